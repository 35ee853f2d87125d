"""Oracle restatement of UpdaterSLAM::update (steps 4-5, update/UpdaterSLAM.cpp:310-470) against independent numpy code."""
import numpy as np
import pytest

from open_vins_b200 import capi, sim


def _cols(order_off, order_sz):
    return np.concatenate([np.arange(o, o + s) for o, s in zip(order_off, order_sz)])


@pytest.mark.parametrize("rep", [capi.REP_GLOBAL_3D, capi.REP_GLOBAL_FULL_INVERSE_DEPTH, capi.REP_ANCHORED_3D,
                                 capi.REP_ANCHORED_FULL_INVERSE_DEPTH, capi.REP_ANCHORED_MSCKF_INVERSE_DEPTH,
                                 capi.REP_ANCHORED_INVERSE_DEPTH_SINGLE])
def test_slam_update_is_the_textbook_update_of_its_stacked_system(oracle, rep):
    single = rep == capi.REP_ANCHORED_INVERSE_DEPTH_SINGLE  # 1-wide landmark, the two bearing columns projected out
    lmw, drop = (1, 2) if single else (3, 0)
    case = sim.make_slam_case(n_landmarks=14, n_clones=8, n_cams=2, seed=40 + rep, rep=rep)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, feat_rep=rep)
    r = oracle.slam_update(case.frame, case.feats, case.landmarks, opts, case.P)
    assert r["status"] == 0 and r["stats"].n_feats_used >= 8
    H, res, Rd = r["H_big"], r["res_big"], r["Rdiag_big"]
    c = _cols(r["order_off"], r["order_sz"])
    assert H.shape == (r["stats"].rows_stacked, len(c)) and len(set(c.tolist())) == len(c)
    # every accepted landmark contributes its own 3 columns, and only its own rows touch them
    used = np.flatnonzero(r["out"].status == 0)
    row = 0
    for f in used:
        m = 2 * int(case.feats.meas_off[f + 1] - case.feats.meas_off[f]) - drop
        j = [int(np.flatnonzero(c == case.lm_off[f] + k)[0]) for k in range(lmw)]
        assert np.abs(H[row:row + m, j]).max() > 0
        mask = np.ones(H.shape[0], dtype=bool)
        mask[row:row + m] = False
        assert not H[mask][:, j].any()
        sig = 1.0 if case.landmarks.sigma_pix is None else case.landmarks.sigma_pix[f]
        assert np.all(Rd[row:row + m] == sig ** 2)
        row += m
    assert row == H.shape[0]
    # textbook EKF update (StateHelper.cpp:116-197) in numpy
    P = case.P
    S = H @ P[np.ix_(c, c)] @ H.T + np.diag(Rd)
    K = P[:, c] @ H.T @ np.linalg.inv(S)
    Pn = P - K @ H @ P[c, :]
    assert np.linalg.norm(r["P"] - Pn) <= 1e-10 * np.linalg.norm(Pn)
    assert np.linalg.norm(r["dx"] - K @ res) <= 1e-10 * np.linalg.norm(K @ res)
    # the gate of every feature: chi2 = res' (H_xf P_marg H_xf' + s2 I)^-1 res on its own rows
    row = 0
    for f in used:
        m = 2 * int(case.feats.meas_off[f + 1] - case.feats.meas_off[f]) - drop
        Sf = S[row:row + m, row:row + m]
        chi2 = res[row:row + m] @ np.linalg.solve(Sf, res[row:row + m])
        assert abs(chi2 - r["out"].chi2[f]) <= 1e-9 * chi2
        row += m


@pytest.mark.parametrize("rep", [capi.REP_GLOBAL_3D, capi.REP_ANCHORED_3D])
def test_landmark_columns_by_finite_differences(oracle, rep):
    """d(residual)/d(landmark) = -H_f: perturb the landmark value (FEJ off, FEJ value = value) and difference the residuals."""
    case = sim.make_slam_case(n_landmarks=6, n_clones=6, n_cams=2, seed=77, rep=rep, two_classes=False)
    lm = case.landmarks
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, feat_rep=rep, do_fej=0, chi2_multipler=1e12)
    mk = lambda v: capi.LandmarkArrays(lm.lm_off, v, v, lm.anchor_cam, lm.anchor_clone)
    r0 = oracle.slam_update(case.frame, case.feats, mk(lm.value), opts, case.P)
    c = _cols(r0["order_off"], r0["order_sz"])
    assert (r0["out"].status == 0).all()
    h = 1e-2  # residuals carry float32 rounding of the distortion (~3e-5 px): the step must dwarf it
    row = 0
    for f in range(6):
        m = 2 * int(case.feats.meas_off[f + 1] - case.feats.meas_off[f])
        for k in range(3):
            vp, vm = lm.value.copy(), lm.value.copy()
            vp[f, k] += h
            vm[f, k] -= h
            rp = oracle.slam_update(case.frame, case.feats, mk(vp), opts, case.P)["res_big"][row:row + m]
            rm = oracle.slam_update(case.frame, case.feats, mk(vm), opts, case.P)["res_big"][row:row + m]
            fd = -(rp - rm) / (2 * h)
            j = int(np.flatnonzero(c == case.lm_off[f] + k)[0])
            an = r0["H_big"][row:row + m, j]
            assert np.abs(fd - an).max() <= 2e-3 * max(np.abs(an).max(), 1.0), (f, k)
        row += m


def _init_case(seed, N=60, r=20, k=3, noise=0.05):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-2 * np.eye(N)
    off, sz = [9, 27, 45, 33], [6, 6, 6, 8]
    n = sum(sz)
    H_R = rng.standard_normal((r, n))
    H_L = rng.standard_normal((r, k))
    res = H_L @ (0.1 * rng.standard_normal(k)) + noise * rng.standard_normal(r)
    return P, off, sz, H_R, H_L, res


@pytest.mark.parametrize("seed,r", [(1, 20), (2, 7), (3, 3), (4, 60)])
def test_initialize_equals_the_joint_posterior_with_no_prior_on_the_new_variable(oracle, seed, r):
    """StateHelper::initialize (Givens split + initialize_invertible + EKFUpdate, StateHelper.cpp:393-577) must equal the
    information-form posterior of the augmented state with zero prior information on the new variable."""
    P, off, sz, H_R, H_L, res = _init_case(seed, r=r)
    s2 = 0.05 ** 2
    st, acc, Po, dxn, dx = oracle.cov_initialize(P, off, sz, H_R, H_L, res, sigma2=s2, chi2_mult=1e9)
    assert st == 0 and acc and Po.shape[0] == P.shape[0] + 3
    N, k = P.shape[0], 3
    cols = np.concatenate([np.arange(o, o + s) for o, s in zip(off, sz)] + [np.arange(N, N + k)])
    Hf = np.zeros((len(res), N + k))
    Hf[:, cols] = np.hstack([H_R, H_L])
    Lam = np.zeros((N + k, N + k))
    Lam[:N, :N] = np.linalg.inv(P)
    Lam += Hf.T @ Hf / s2
    Pn = np.linalg.inv(Lam)
    assert np.linalg.norm(Pn - Po) <= 1e-7 * np.linalg.norm(Po)  # numpy's inverse of the information matrix is the looser side
    dxa = Pn @ Hf.T @ res / s2
    tot = dx.copy()
    tot[N:] += dxn  # the new variable moves by H_Linv res_init and then by the update's cross-covariance term
    assert np.linalg.norm(dxa - tot) <= 1e-7 * np.linalg.norm(dxa)


def test_initialize_gate_rejects_inconsistent_systems(oracle):
    P, off, sz, H_R, H_L, res = _init_case(5, r=30)
    res = res + 3.0  # residuals far beyond what P and the noise explain
    st, acc, Po, dxn, dx = oracle.cov_initialize(P, off, sz, H_R, H_L, res, sigma2=0.05 ** 2, chi2_mult=1.0)
    assert st == 0 and not acc and np.array_equal(Po, P)


ANCHORED = [capi.REP_ANCHORED_3D, capi.REP_ANCHORED_FULL_INVERSE_DEPTH, capi.REP_ANCHORED_MSCKF_INVERSE_DEPTH,
            capi.REP_ANCHORED_INVERSE_DEPTH_SINGLE]


def _global_point(fr, cam, cl, pA, fej=False):
    R = (fr.clone_R_fej if fej else fr.clone_R)[cl]
    p = (fr.clone_p_fej if fej else fr.clone_p)[cl]
    return R.T @ (fr.cam_R[cam].T @ (pA - fr.cam_p[cam])) + p


@pytest.mark.parametrize("rep", ANCHORED)
@pytest.mark.parametrize("ext,fej", [(1, 1), (0, 1), (1, 0)])
def test_anchor_change_matches_oracle_and_keeps_the_point(oracle, rep, ext, fej):
    """ovb_slam_anchor_change (host math of UpdaterSLAM::perform_anchor_change, UpdaterSLAM.cpp:506-647; no GPU involved) against
    the oracle restatement, plus the geometric invariant: the landmark's global position does not move."""
    case = sim.make_slam_case(n_landmarks=6, n_clones=7, n_cams=2, seed=90 + rep, rep=rep)
    fr, lm = case.frame, case.landmarks
    opts = capi.default_opts(do_calib_camera_pose=ext, do_calib_camera_intrinsics=0, feat_rep=rep, do_fej=fej)
    for f, (new_cam, new_clone) in enumerate([(0, 6), (1, 6), (0, 5), (1, 3), (0, 6), (1, 4)]):
        old_cam, old_clone = int(lm.anchor_cam[f]), int(lm.anchor_clone[f])
        got = capi.slam_anchor_change(fr, opts, lm.lm_off[f], lm.value[f], lm.value_fej[f], old_cam, old_clone, new_cam, new_clone)
        ref = oracle.anchor_change(fr, opts, lm.lm_off[f], lm.value[f], lm.value_fej[f], old_cam, old_clone, new_cam, new_clone)
        nv, nvf, off, sz, Phi = got
        assert np.array_equal(off, ref[2]) and np.array_equal(sz, ref[3])
        assert sz[-1] == (1 if rep == capi.REP_ANCHORED_INVERSE_DEPTH_SINGLE else 3) and off[-1] == lm.lm_off[f]
        np.testing.assert_allclose(nv, ref[0], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(nvf, ref[1], rtol=1e-13, atol=1e-15)
        assert np.linalg.norm(Phi - ref[4]) <= 1e-11 * np.linalg.norm(ref[4])
        # same point in the global frame before and after (current estimates, and FEJ estimates for the FEJ value)
        np.testing.assert_allclose(_global_point(fr, new_cam, new_clone, nv), _global_point(fr, old_cam, old_clone, lm.value[f]), atol=1e-12)
        np.testing.assert_allclose(_global_point(fr, new_cam, new_clone, nvf, True), _global_point(fr, old_cam, old_clone, lm.value_fej[f], True),
                                   atol=1e-12)


def test_anchor_change_phi_by_finite_differences(oracle):
    """ANCHORED_3D, FEJ off: Phi's landmark and position blocks are plain derivatives of
    p_new = R_OLDtoNEW p_old + p_OLDinNEW with respect to p_old and the two clone positions."""
    rep = capi.REP_ANCHORED_3D
    case = sim.make_slam_case(n_landmarks=3, n_clones=6, n_cams=2, seed=17, rep=rep)
    fr, lm = case.frame, case.landmarks
    opts = capi.default_opts(do_calib_camera_pose=0, feat_rep=rep, do_fej=0)
    f, new_cam, new_clone = 1, 1, 5
    old_cam, old_clone = int(lm.anchor_cam[f]), int(lm.anchor_clone[f])
    nv, nvf, off, sz, Phi = capi.slam_anchor_change(fr, opts, lm.lm_off[f], lm.value[f], lm.value[f], old_cam, old_clone, new_cam, new_clone)
    cols = np.concatenate([[0], np.cumsum(sz)])
    h = 1e-6

    def new_value(val, dp_old=np.zeros(3), dp_new=np.zeros(3)):
        fr2 = capi.FrameArrays(fr.clone_R, fr.clone_p.copy(), fr.clone_R, fr.clone_p.copy(), fr.clone_off, fr.cam_R, fr.cam_p, fr.cam_intr,
                               fr.cam_model, fr.cam_ext_off, fr.cam_intr_off)
        fr2.clone_p[old_clone] += dp_old
        fr2.clone_p[new_clone] += dp_new
        fr2.clone_p_fej[:] = fr2.clone_p
        return capi.slam_anchor_change(fr2, opts, lm.lm_off[f], val, val, old_cam, old_clone, new_cam, new_clone)[0]

    for k in range(3):
        e = np.zeros(3)
        e[k] = h
        fd_lm = (new_value(lm.value[f] + e) - new_value(lm.value[f] - e)) / (2 * h)
        np.testing.assert_allclose(Phi[:, cols[-2] + k], fd_lm, atol=1e-7)                 # d p_new / d p_old
        i_old = int(np.flatnonzero(off == fr.clone_off[old_clone])[0])
        i_new = int(np.flatnonzero(off == fr.clone_off[new_clone])[0])
        fd_old = (new_value(lm.value[f], dp_old=e) - new_value(lm.value[f], dp_old=-e)) / (2 * h)
        fd_new = (new_value(lm.value[f], dp_new=e) - new_value(lm.value[f], dp_new=-e)) / (2 * h)
        np.testing.assert_allclose(Phi[:, cols[i_old] + 3 + k], fd_old, atol=1e-7)        # position part of the old anchor clone
        np.testing.assert_allclose(Phi[:, cols[i_new] + 3 + k], fd_new, atol=1e-7)        # position part of the new anchor clone
