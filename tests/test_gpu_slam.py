"""ovb_slam_update (UpdaterSLAM::update steps 4-5, update/UpdaterSLAM.cpp:310-470) on the GPU against the oracle."""
import numpy as np
import pytest

from open_vins_b200 import capi, sim

pytestmark = pytest.mark.gpu

REPS = [capi.REP_GLOBAL_3D, capi.REP_GLOBAL_FULL_INVERSE_DEPTH, capi.REP_ANCHORED_3D, capi.REP_ANCHORED_FULL_INVERSE_DEPTH,
        capi.REP_ANCHORED_MSCKF_INVERSE_DEPTH, capi.REP_ANCHORED_INVERSE_DEPTH_SINGLE]


def _check(eng, oracle, case, opts):
    ref = oracle.slam_update(case.frame, case.feats, case.landmarks, opts, case.P)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.slam_update(case.frame, case.feats, case.landmarks, opts)
    assert st == ref["status"] == 0
    assert np.array_equal(out.status, ref["out"].status)
    ok = ref["out"].status == 0
    np.testing.assert_allclose(out.chi2[ok], ref["out"].chi2[ok], rtol=1e-8)
    assert stats.n_feats_used == ref["stats"].n_feats_used and stats.rows_stacked == ref["stats"].rows_stacked
    assert stats.cols_stacked == ref["stats"].cols_stacked
    Pg = eng.cov_get()
    assert np.linalg.norm(Pg - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * max(np.linalg.norm(ref["dx"]), 1e-300)
    assert np.array_equal(Pg, Pg.T)
    return ref, out, stats


@pytest.mark.parametrize("rep", REPS)
@pytest.mark.parametrize("order", [capi.COLS_REFERENCE_FIRST_SEEN, capi.COLS_CANONICAL])
def test_slam_update_parity(oracle, rep, order):
    case = sim.make_slam_case(n_landmarks=14, n_clones=8, n_cams=2, seed=60 + rep, rep=rep)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, feat_rep=rep, col_order=order)
    eng = capi.Engine(max_state=256, max_feats=256, max_meas=4096)
    ref, out, stats = _check(eng, oracle, case, opts)
    assert stats.n_feats_used >= 8
    eng.close()


def test_slam_update_gate_and_no_calibration(oracle):
    """Landmarks far from where the measurements put them are rejected by the chi² gate (per-class multiplier); the rest
    updates the state. Mono, no calibration columns, global representation."""
    case = sim.make_slam_case(n_landmarks=20, n_clones=6, n_cams=1, seed=5, rep=capi.REP_GLOBAL_3D, calib_ext=False, calib_intr=False)
    lm = case.landmarks
    val = lm.value.copy()
    val[[2, 7, 11]] += np.array([0.8, -0.6, 0.9])  # gross landmark errors
    case.landmarks = capi.LandmarkArrays(lm.lm_off, val, lm.value_fej, lm.anchor_cam, lm.anchor_clone, lm.sigma_pix, lm.chi2_multipler)
    opts = capi.default_opts(feat_rep=capi.REP_GLOBAL_3D, col_order=capi.COLS_CANONICAL)
    eng = capi.Engine(max_state=256, max_feats=256, max_meas=4096)
    ref, out, stats = _check(eng, oracle, case, opts)
    assert (out.status[[2, 7, 11]] == capi.FEAT_CHI2).all() and 10 <= stats.n_feats_used <= 17
    eng.close()


def test_msckf_then_slam_on_the_resident_covariance(oracle):
    """VioManager order (core/VioManager.cpp:525-547): MSCKF update, then the SLAM update, P staying on the device."""
    rep = capi.REP_GLOBAL_3D
    sl = sim.make_slam_case(n_landmarks=10, n_clones=8, n_cams=2, seed=9, rep=rep)
    ms = sim.make_update_case(n_feats=40, n_clones=8, n_cams=2, seed=9, calib_ext=True, calib_intr=True)
    assert np.array_equal(ms.frame.clone_R, sl.frame.clone_R)  # same window (same seed): the landmarks extend the state
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, feat_rep=rep, col_order=capi.COLS_CANONICAL)
    eng = capi.Engine(max_state=256, max_feats=256, max_meas=4096)
    eng.cov_set(sl.P)
    st, out, dx, stats = eng.msckf_update(ms.frame, ms.feats, opts)
    r1 = oracle.msckf_update(ms.frame, ms.feats, opts, sl.P, dumps=False)
    assert st == 0 and np.array_equal(out.status, r1["out"].status)
    st, out2, dx2, stats2 = eng.slam_update(sl.frame, sl.feats, sl.landmarks, opts)
    r2 = oracle.slam_update(sl.frame, sl.feats, sl.landmarks, opts, r1["P"])
    assert st == 0 and np.array_equal(out2.status, r2["out"].status)
    assert np.linalg.norm(eng.cov_get() - r2["P"]) <= 1e-9 * np.linalg.norm(r2["P"])
    assert np.linalg.norm(dx2 - r2["dx"]) <= 1e-9 * np.linalg.norm(r2["dx"])
    eng.close()


def test_slam_update_argument_errors():
    case = sim.make_slam_case(n_landmarks=4, n_clones=5, n_cams=1, seed=1, rep=0, calib_ext=False, calib_intr=False)
    eng = capi.Engine(max_state=256, max_feats=64, max_meas=1024)
    eng.cov_set(case.P)
    # an anchored representation without anchors is an argument error, not a silent wrong answer
    opts = capi.default_opts(feat_rep=capi.REP_ANCHORED_3D)
    with pytest.raises(capi.OvbError) as ei:
        eng.slam_update(case.frame, case.feats, case.landmarks, opts)
    assert ei.value.code == capi.OVB_ERR_ARG
    eng.close()


@pytest.mark.parametrize("seed,r,mult", [(1, 20, 1e9), (2, 7, 1e9), (3, 3, 1e9), (4, 60, 1e9), (6, 120, 1e9), (5, 30, 1.0)])
def test_cov_initialize_parity(oracle, seed, r, mult):
    """ovb_cov_initialize (StateHelper::initialize, StateHelper.cpp:393-577): rows < / = / > columns (the last two compress the
    projected part first), plus a rejected system that must leave P untouched."""
    from tests.test_slam_cpu import _init_case
    P, off, sz, H_R, H_L, res = _init_case(seed, r=r)
    if mult == 1.0:
        res = res + 3.0
    s2 = 0.05 ** 2
    st_r, acc_r, P_r, dxn_r, dx_r = oracle.cov_initialize(P, off, sz, H_R, H_L, res, sigma2=s2, chi2_mult=mult)
    eng = capi.Engine(max_state=128, max_feats=16, max_meas=256, max_rows=256)
    eng.cov_set(P)
    st, acc, dxn, dx = eng.cov_initialize(off, sz, H_R, H_L, res, sigma2=s2, chi2_mult=mult)
    assert st == st_r == 0 and acc == acc_r
    Pg = eng.cov_get()
    assert Pg.shape == P_r.shape
    if not acc:
        assert np.array_equal(Pg, P)
    else:
        assert np.linalg.norm(Pg - P_r) <= 1e-9 * np.linalg.norm(P_r)
        assert np.linalg.norm(dxn - dxn_r) <= 1e-9 * np.linalg.norm(dxn_r)
        assert np.linalg.norm(dx - dx_r) <= 1e-9 * max(np.linalg.norm(dx_r), 1e-300)
        assert np.array_equal(Pg, Pg.T)
    eng.close()


def test_delayed_init_composition(oracle):
    """UpdaterSLAM::delayed_init (update/UpdaterSLAM.cpp:68-251) composed from the ABI: triangulate the new tracks, take each
    feature's full (pre-nullspace) Jacobians, StateHelper::initialize them one after the other on the resident covariance.
    The oracle runs the same composition with its own triangulation / Jacobians / initialize."""
    case = sim.make_update_case(n_feats=8, n_clones=8, n_cams=2, seed=21, calib_ext=True, calib_intr=True, outlier_frac=0.0,
                                degenerate_frac=0.0)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1)
    eng = capi.Engine(max_state=256, max_feats=64, max_meas=2048)
    eng.cov_set(case.P)
    tri_g = eng.triangulate(case.frame, case.feats, opts)
    tri_r, _ = oracle.triangulate(case.frame, case.feats, opts)
    assert np.array_equal(tri_g.status, tri_r.status)
    P_r = case.P.copy()
    n_init = 0
    # the state mean is not moved between the initialisations here (the caller's Type::update does that), so the later
    # tracks look inconsistent with the shrinking P: a wide gate keeps most of them, the first one uses the reference's
    mult = 1.0
    for f in np.flatnonzero(tri_r.status == 0)[:5]:
        one = case.feats.subset([f])

        def pick(src):
            o = capi.FeatOut(1)
            o.status[:] = 0
            o.p_FinA[0], o.p_FinG[0] = src.p_FinA[f], src.p_FinG[f]
            o.anchor_cam[0], o.anchor_clone[0] = src.anchor_cam[f], src.anchor_clone[f]
            return o

        Hf_g, Hx_g, res_g, _, col_g = eng.feature_jacobians(case.frame, one, opts, pick(tri_g), stage=0)
        Hf_r, Hx_r, res_r, _ = oracle.feature_jacobians(case.frame, one, opts, pick(tri_r), 0, col_g)
        # H_order = the variables this feature touches, as (offset, size) runs of the dump's columns
        used = np.flatnonzero(np.abs(Hx_r).sum(axis=0) > 0)
        cols = col_g[used]
        starts = [0] + [i for i in range(1, len(cols)) if cols[i] != cols[i - 1] + 1] + [len(cols)]
        off = [int(cols[a]) for a in starts[:-1]]
        sz = [int(b - a) for a, b in zip(starts[:-1], starts[1:])]
        st_r, acc_r, P_r, dxn_r, dx_r = oracle.cov_initialize(P_r, off, sz, Hx_r[:, used], Hf_r, res_r, sigma2=1.0, chi2_mult=mult)
        st_g, acc_g, dxn_g, dx_g = eng.cov_initialize(off, sz, Hx_g[:, used], Hf_g, res_g, sigma2=1.0, chi2_mult=mult)
        assert st_g == st_r == 0 and acc_g == acc_r
        Pg = eng.cov_get()
        assert Pg.shape == P_r.shape
        assert np.linalg.norm(Pg - P_r) <= 1e-9 * np.linalg.norm(P_r)
        mult = 200.0
        if acc_r:
            n_init += 1
            assert np.linalg.norm(dxn_g - dxn_r) <= 1e-8 * max(np.linalg.norm(dxn_r), 1e-12)
            assert np.linalg.norm(dx_g - dx_r) <= 1e-9 * max(np.linalg.norm(dx_r), 1e-300)
    assert n_init >= 3 and eng.cov_dim() == case.P.shape[0] + 3 * n_init
    eng.close()


def test_anchor_change_propagation_shape(oracle):
    """UpdaterSLAM::perform_anchor_change (update/UpdaterSLAM.cpp:574-647) ends in StateHelper::EKFPropagation with a 3-wide
    new block (the landmark itself) and Phi over (landmark, old anchor clone, new anchor clone[, extrinsics]), Q = 0:
    ovb_cov_propagate covers that shape on the resident covariance."""
    rng = np.random.default_rng(3)
    N = 90
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    lm_off, old_clone, new_clone, ext = 84, 30, 66, 15
    old_off, old_sz = [lm_off, old_clone, new_clone, ext], [3, 6, 6, 6]
    Phi = np.hstack([np.eye(3) + 0.1 * rng.standard_normal((3, 3)), 0.3 * rng.standard_normal((3, 18))])
    Q = np.zeros((3, 3))
    st_r, P_r = oracle.cov_propagate(P, lm_off, Phi, Q, old_off, old_sz)
    eng = capi.Engine(max_state=128, max_feats=16, max_meas=256)
    eng.cov_set(P)
    st_g = eng.cov_propagate(lm_off, Phi, Q, old_off, old_sz)
    assert st_g == st_r == 0
    Pg = eng.cov_get()
    assert np.linalg.norm(Pg - P_r) <= 1e-12 * np.linalg.norm(P_r)
    # textbook: P' = J P J' with J = identity except the landmark rows
    J = np.eye(N)
    J[lm_off:lm_off + 3] = 0.0
    cols = np.concatenate([np.arange(o, o + s) for o, s in zip(old_off, old_sz)])
    J[lm_off:lm_off + 3, cols] = Phi
    Pt = J @ P @ J.T
    assert np.linalg.norm(Pg - Pt) <= 1e-12 * np.linalg.norm(Pt)
    eng.close()


def _apply_dx_to_frame(fr, dx):
    """Host side of StateHelper::EKFUpdate's mean update (state/StateHelper.cpp:185-188) for the variables the frame carries:
    JPL left error on rotations (R <- exp(-dtheta) R), additive elsewhere; FEJ values stay. In place (the engine re-reads)."""
    for c, o in enumerate(fr.clone_off):
        fr.clone_R[c] = (sim.exp_so3(-dx[o:o + 3]) @ fr.clone_R[c].reshape(3, 3)).reshape(fr.clone_R[c].shape)
        fr.clone_p[c] += dx[o + 3:o + 6]
    for k in range(fr.n_cams):
        o = fr.cam_ext_off[k]
        if o >= 0:
            fr.cam_R[k] = (sim.exp_so3(-dx[o:o + 3]) @ fr.cam_R[k].reshape(3, 3)).reshape(fr.cam_R[k].shape)
            fr.cam_p[k] += dx[o + 3:o + 6]
        o = fr.cam_intr_off[k]
        if o >= 0:
            fr.cam_intr[k] += dx[o:o + 8]


def test_delayed_init_one_call(oracle):
    """ovb_slam_delayed_init = UpdaterSLAM::delayed_init (update/UpdaterSLAM.cpp:61-251) as ONE ABI call, the state mean moved
    by the caller's callback between the features. Checked against the same sequence composed from the oracle's
    triangulation / Jacobians / StateHelper::initialize with the identical mean update."""
    import copy
    kw = dict(n_feats=10, n_clones=8, n_cams=2, seed=23, calib_ext=True, calib_intr=True, outlier_frac=0.0, degenerate_frac=0.0)
    case_g, case_o = sim.make_update_case(**kw), sim.make_update_case(**kw)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1)
    eng = capi.Engine(max_state=256, max_feats=64, max_meas=2048)
    eng.cov_set(case_g.P)
    log_g = []

    def on_init(f, lm_off, dx_new, dx):
        log_g.append((f, lm_off, dx_new, dx))
        _apply_dx_to_frame(case_g.frame, dx)
    out_g, lm_off = eng.slam_delayed_init(case_g.frame, case_g.feats, opts, on_init)
    # ---- oracle composition
    fr = case_o.frame
    tri, _ = oracle.triangulate(fr, case_o.feats, opts)
    P = case_o.P.copy()
    log_o, status_o = [], tri.status.copy()
    for f in np.flatnonzero(tri.status == 0):
        one = case_o.feats.subset([f])
        o = capi.FeatOut(1)
        o.status[:] = 0
        o.p_FinA[0], o.p_FinG[0] = tri.p_FinA[f], tri.p_FinG[f]
        o.anchor_cam[0], o.anchor_clone[0] = tri.anchor_cam[f], tri.anchor_clone[f]
        # canonical column list = what the engine reports; build it from the layout
        cols = []
        for off, sz in sorted([(int(x), 6) for x in fr.clone_off] + [(int(x), 6) for x in fr.cam_ext_off if x >= 0] + [(int(x), 8) for x in fr.cam_intr_off if x >= 0]):
            cols += list(range(off, off + sz))
        cols = np.array(cols)
        Hf, Hx, res, _ = oracle.feature_jacobians(fr, one, opts, o, 0, cols)
        used = np.flatnonzero(np.abs(Hx).sum(axis=0) > 0)
        cc = cols[used]
        starts = [0] + [i for i in range(1, len(cc)) if cc[i] != cc[i - 1] + 1] + [len(cc)]
        off = [int(cc[a]) for a in starts[:-1]]
        sz = [int(b - a) for a, b in zip(starts[:-1], starts[1:])]
        st, acc, P, dxn, dx = oracle.cov_initialize(P, off, sz, Hx[:, used], Hf, res, sigma2=1.0, chi2_mult=float(opts.chi2_multipler))
        assert st == 0
        if acc:
            log_o.append((int(f), P.shape[0] - 3, dxn, dx))
            _apply_dx_to_frame(fr, dx)
        else:
            status_o[f] = capi.FEAT_CHI2
    assert np.array_equal(out_g.status, status_o)
    assert len(log_g) == len(log_o) >= 4
    for (fg, og, dng, dg), (fo, oo, dno, do) in zip(log_g, log_o):
        assert fg == fo and og == oo == lm_off[fg]
        assert np.linalg.norm(dng - dno) <= 1e-8 * max(np.linalg.norm(dno), 1e-12)
        assert np.linalg.norm(dg - do) <= 1e-8 * max(np.linalg.norm(do), 1e-300)
    Pg = eng.cov_get()
    assert Pg.shape == P.shape and np.linalg.norm(Pg - P) <= 1e-9 * np.linalg.norm(P)
    eng.close()
