"""CPU tests: the C++ oracle against an independent numpy/scipy twin and algebraic invariants.

The reference ships no golden vectors for this path (parity unpinned), so the restatement is pinned three ways:
library linear algebra (numpy/scipy), finite differences of a double-precision measurement model, and invariants.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg

from open_vins_b200 import capi, sim
from tests import np_twin


def _case(**kw):
    d = dict(n_feats=40, n_clones=12, n_cams=1, seed=3)
    d.update(kw)
    return sim.make_update_case(**d)


def test_givens_convention(oracle):
    rng = np.random.default_rng(0)
    cs = np.zeros(2)
    for p, q in list(rng.standard_normal((200, 2))) + [(0.0, 1.5), (2.0, 0.0), (-2.0, 0.0), (0.0, -3.0), (1.0, 1.0), (-1.0, 1.0)]:
        oracle.lib().ovo_make_givens(C.c_double(p), C.c_double(q), cs.ctypes.data_as(capi.c_double_p))
        c, s = cs
        x, y = c * p - s * q, s * p + c * q  # applyOnTheLeft(0,1,G.adjoint())
        assert abs(y) <= 1e-15 * max(1.0, np.hypot(p, q))
        assert x >= 0.0 and abs(x - np.hypot(p, q)) <= 1e-15 * max(1.0, np.hypot(p, q))
        assert abs(c * c + s * s - 1.0) < 1e-15


def test_solve3_and_cond3(oracle):
    rng = np.random.default_rng(1)
    x = np.zeros(3)
    oracle.lib().ovo_cond3.restype = C.c_double
    for _ in range(100):
        B = rng.standard_normal((3, 3))
        A = np.ascontiguousarray(B @ B.T + 0.1 * np.eye(3))
        b = rng.standard_normal(3)
        oracle.lib().ovo_solve3(A.ctypes.data_as(capi.c_double_p), b.ctypes.data_as(capi.c_double_p), x.ctypes.data_as(capi.c_double_p))
        ref = np.linalg.solve(A, b)
        assert np.linalg.norm(x - ref) <= 1e-12 * np.linalg.cond(A) * np.linalg.norm(ref)
        cnd = oracle.lib().ovo_cond3(A.ctypes.data_as(capi.c_double_p))
        assert abs(cnd - np.linalg.cond(A)) <= 1e-10 * np.linalg.cond(A)


@pytest.mark.parametrize("model", [0, 1])
def test_camera_model_against_twin(oracle, model):
    rng = np.random.default_rng(2)
    intr = np.array(sim._INTR[0] if model == 0 else sim._INTR_EQUI, dtype=np.float64)
    uv = np.zeros(2)
    dzn = np.zeros(4)
    dzeta = np.zeros(16)
    for _ in range(50):
        x, y = rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4)
        oracle.lib().ovo_distort_d(C.c_int(model), intr.ctypes.data_as(capi.c_double_p), C.c_double(x), C.c_double(y),
                                   uv.ctypes.data_as(capi.c_double_p))
        ref = np_twin.distort(model, intr, float(np.float32(x)), float(np.float32(y)))
        # distort_d rounds the pixel to float32 (cam/CamBase.h:130-135): half an ulp of a ~500 px value
        assert np.all(np.abs(uv - ref) <= 4e-5)
        assert np.all(uv == uv.astype(np.float32).astype(np.float64))
        oracle.lib().ovo_distort_jacobian(C.c_int(model), intr.ctypes.data_as(capi.c_double_p), C.c_double(x), C.c_double(y),
                                          dzn.ctypes.data_as(capi.c_double_p), dzeta.ctypes.data_as(capi.c_double_p))
        h = 1e-6
        fd = np.zeros((2, 2))
        fd[:, 0] = (np_twin.distort(model, intr, x + h, y) - np_twin.distort(model, intr, x - h, y)) / (2 * h)
        fd[:, 1] = (np_twin.distort(model, intr, x, y + h) - np_twin.distort(model, intr, x, y - h)) / (2 * h)
        assert np.allclose(dzn.reshape(2, 2), fd, rtol=1e-6, atol=1e-5)
        fdz = np.zeros((2, 8))
        for k in range(8):
            dp = intr.copy()
            dm = intr.copy()
            hk = 1e-6 * max(1.0, abs(intr[k]))
            dp[k] += hk
            dm[k] -= hk
            fdz[:, k] = (np_twin.distort(model, dp, x, y) - np_twin.distort(model, dm, x, y)) / (2 * hk)
        assert np.allclose(dzeta.reshape(2, 8), fdz, rtol=1e-6, atol=1e-5)


def test_triangulation_against_twin(oracle):
    case = _case(n_feats=60, n_cams=2, n_clones=10, seed=5)
    opts = capi.default_opts(refine_features=0)
    out, _ = oracle.triangulate(case.frame, case.feats, opts)
    ok = out.status == capi.FEAT_OK
    assert ok.sum() > 30
    for f in np.nonzero(ok)[0]:
        pA, pG, cond = np_twin.triangulate_linear(case.frame, case.feats, f, out.anchor_cam[f], out.anchor_clone[f])
        assert np.linalg.norm(out.p_FinA[f] - pA) <= 1e-11 * cond * np.linalg.norm(pA)
        assert np.linalg.norm(out.p_FinG[f] - pG) <= 1e-11 * cond * np.linalg.norm(pG)
    # anchor rule: camera with most measurements on ties the first visited (descending id here), its newest clone
    for f in range(case.feats.n_feats):
        m0, m1 = case.feats.meas_off[f], case.feats.meas_off[f + 1]
        if m1 - m0 < 2:
            continue
        cams = case.feats.cam[m0:m1]
        counts = {c: int((cams == c).sum()) for c in np.unique(cams)}
        visit = []
        for c in cams:
            if c not in visit:
                visit.append(int(c))
        best = max(visit, key=lambda c: (counts[c], -visit.index(c)))
        assert out.anchor_cam[f] == best
        assert out.anchor_clone[f] == case.feats.clone[m0:m1][cams == best][-1]


def test_gauss_newton_reduces_reprojection_error(oracle):
    case = _case(n_feats=80, n_cams=2, n_clones=12, seed=7, outlier_frac=0.0, degenerate_frac=0.0)
    o0, _ = oracle.triangulate(case.frame, case.feats, capi.default_opts(refine_features=0))
    o1, tr = oracle.triangulate(case.frame, case.feats, capi.default_opts(refine_features=1))
    both = (o0.status == 0) & (o1.status == 0)
    assert both.sum() > 50

    def cost(pG, f):
        c = 0.0
        for i in range(case.feats.meas_off[f], case.feats.meas_off[f + 1]):
            cam, cl = int(case.feats.cam[i]), int(case.feats.clone[i])
            pc = case.frame.cam_R[cam].reshape(3, 3) @ (case.frame.clone_R[cl].reshape(3, 3) @ (pG - case.frame.clone_p[cl])) + case.frame.cam_p[cam]
            c += np.sum((case.feats.uvn[i].astype(np.float64) - pc[:2] / pc[2]) ** 2)
        return c
    worse = 0
    for f in np.nonzero(both)[0]:
        if cost(o1.p_FinG[f], f) > cost(o0.p_FinG[f], f) * (1 + 1e-9):
            worse += 1
    assert worse == 0
    assert tr["runs"][both].max() <= 5 and tr["solves"][both].min() >= 1


@pytest.mark.parametrize("calib", [False, True])
@pytest.mark.parametrize("rep", [capi.REP_GLOBAL_3D, capi.REP_ANCHORED_MSCKF_INVERSE_DEPTH, capi.REP_ANCHORED_3D,
                                 capi.REP_GLOBAL_FULL_INVERSE_DEPTH, capi.REP_ANCHORED_FULL_INVERSE_DEPTH])
def test_feature_jacobian_finite_differences(oracle, calib, rep):
    """Analytic H_x (FEJ off) against central differences of the double-precision twin of the measurement function."""
    case = _case(n_feats=6, n_cams=2, n_clones=6, seed=11, calib_ext=calib, calib_intr=calib, outlier_frac=0.0, degenerate_frac=0.0,
                 full_track_frac=1.0)
    fr = case.frame
    opts = capi.default_opts(do_fej=0, feat_rep=rep, do_calib_camera_pose=int(calib), do_calib_camera_intrinsics=int(calib))
    out, _ = oracle.triangulate(fr, case.feats, opts)
    assert (out.status == 0).all()
    cols = []
    for s_off, s_sz in sorted([(o, 6) for o in case.layout.clone_off] + [(o, 6) for o in case.layout.cam_ext_off if o >= 0] +
                              [(o, 8) for o in case.layout.cam_intr_off if o >= 0]):
        cols += list(range(s_off, s_off + s_sz))
    Hf, Hx, res, row_off = oracle.feature_jacobians(fr, case.feats, opts, out, 0, cols)
    colpos = {c: j for j, c in enumerate(cols)}
    relative = rep in (capi.REP_ANCHORED_3D, capi.REP_ANCHORED_FULL_INVERSE_DEPTH, capi.REP_ANCHORED_MSCKF_INVERSE_DEPTH)
    for f in range(case.feats.n_feats):
        acam, acl = int(out.anchor_cam[f]), int(out.anchor_clone[f])
        pA = out.p_FinA[f].copy()

        def feat_global(frame_R, frame_p, camR, camp):
            if not relative:
                return out.p_FinG[f].copy()
            return frame_R[acl].T @ camR[acam].T @ (pA - camp[acam]) + frame_p[acl]
        R0 = fr.clone_R.reshape(-1, 3, 3).copy()
        p0 = fr.clone_p.copy()
        cR0 = fr.cam_R.reshape(-1, 3, 3).copy()
        cp0 = fr.cam_p.copy()
        in0 = fr.cam_intr.copy()

        def h_all(R, p, cR, cp, intr):
            pG = feat_global(R, p, cR, cp)
            zs = []
            for i in range(case.feats.meas_off[f], case.feats.meas_off[f + 1]):
                cam, cl = int(case.feats.cam[i]), int(case.feats.clone[i])
                zs.append(np_twin.project(fr, cam, cl, pG, R[cl], p[cl], cR[cam], cp[cam], intr[cam]))
            return np.concatenate(zs)
        rows = slice(row_off[f], row_off[f + 1])
        h = 1e-6
        # clone poses: theta (JPL: R_true = (I - [dth x]) R_est  =>  R(dth) = exp(-dth) R) then position
        for cl in range(fr.n_clones):
            for k in range(6):
                def pert(sgn):
                    R, p = R0.copy(), p0.copy()
                    if k < 3:
                        d = np.zeros(3)
                        d[k] = sgn * h
                        R[cl] = np_twin.exp_so3(-d) @ R0[cl]
                    else:
                        p[cl, k - 3] += sgn * h
                    return h_all(R, p, cR0, cp0, in0)
                fd = (pert(+1) - pert(-1)) / (2 * h)
                j = colpos[case.layout.clone_off[cl] + k]
                an = Hx[rows, j]
                assert np.allclose(an, fd, rtol=2e-5, atol=2e-4), (f, cl, k, np.abs(an - fd).max())
        if calib:
            for cam in range(fr.n_cams):
                for k in range(6):
                    def pert(sgn):
                        cR, cp = cR0.copy(), cp0.copy()
                        if k < 3:
                            d = np.zeros(3)
                            d[k] = sgn * h
                            cR[cam] = np_twin.exp_so3(-d) @ cR0[cam]
                        else:
                            cp[cam, k - 3] += sgn * h
                        return h_all(R0, p0, cR, cp, in0)
                    fd = (pert(+1) - pert(-1)) / (2 * h)
                    j = colpos[case.layout.cam_ext_off[cam] + k]
                    assert np.allclose(Hx[rows, j], fd, rtol=2e-5, atol=2e-4), (f, cam, k)
                for k in range(8):
                    hk = 1e-6 * max(1.0, abs(in0[cam, k]))

                    def pert(sgn):
                        intr = in0.copy()
                        intr[cam, k] += sgn * hk
                        return h_all(R0, p0, cR0, cp0, intr)
                    fd = (pert(+1) - pert(-1)) / (2 * hk)
                    j = colpos[case.layout.cam_intr_off[cam] + k]
                    assert np.allclose(Hx[rows, j], fd, rtol=2e-5, atol=2e-4), (f, cam, "intr", k)
        # residual = measured pixel - predicted (float-rounded pixel: 4e-5 slack)
        pred = h_all(R0, p0, cR0, cp0, in0)
        meas = case.feats.uv[case.feats.meas_off[f]:case.feats.meas_off[f + 1]].astype(np.float64).reshape(-1)
        assert np.all(np.abs(res[rows] - (meas - pred)) < 2e-4)


def test_nullspace_projection_invariants(oracle):
    case = _case(n_feats=25, n_cams=2, n_clones=8, seed=13, calib_ext=True, calib_intr=True)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1)
    out, _ = oracle.triangulate(case.frame, case.feats, opts)
    cols = list(range(case.layout.N))
    Hf0, Hx0, r0, ro0 = oracle.feature_jacobians(case.frame, case.feats, opts, out.copy(), 0, cols)
    Hf1, Hx1, r1, ro1 = oracle.feature_jacobians(case.frame, case.feats, opts, out.copy(), 1, cols)
    for f in range(case.feats.n_feats):
        if out.status[f] != 0:
            continue
        A, X, r = Hf0[ro0[f]:ro0[f + 1]], Hx0[ro0[f]:ro0[f + 1]], r0[ro0[f]:ro0[f + 1]]
        Q, _ = scipy.linalg.qr(A)  # full Q
        N = Q[:, 3:]
        Xo, ro = Hx1[ro1[f]:ro1[f + 1]], r1[ro1[f]:ro1[f + 1]]
        assert Xo.shape[0] == A.shape[0] - 3
        ref = X.T @ N @ N.T @ X
        assert np.linalg.norm(Xo.T @ Xo - ref) <= 1e-12 * np.linalg.norm(X.T @ X)
        assert np.linalg.norm(Xo.T @ ro - X.T @ N @ N.T @ r) <= 1e-12 * np.linalg.norm(X.T @ r) + 1e-9


def test_compress_against_scipy_qr(oracle):
    rng = np.random.default_rng(4)
    H = rng.standard_normal((300, 40))
    res = rng.standard_normal(300)
    R, z = oracle.compress(H, res)
    assert R.shape == (40, 40)
    assert np.allclose(np.tril(R, -1), 0.0, atol=1e-13)
    assert (np.diag(R) >= 0).all()  # Givens convention of the reference
    Q2, R2 = scipy.linalg.qr(H, mode="economic")
    sgn = np.sign(np.diag(R2))
    assert np.allclose(R, sgn[:, None] * R2, rtol=0, atol=1e-12 * np.abs(R2).max())
    assert np.allclose(z, sgn * (Q2.T @ res), atol=1e-12 * np.linalg.norm(res))
    assert np.linalg.norm(R.T @ R - H.T @ H) <= 1e-13 * np.linalg.norm(H.T @ H)
    # fat matrix: untouched (UpdaterHelper.cpp:459-460)
    Hf = rng.standard_normal((5, 9))
    Rf, zf = oracle.compress(Hf, res[:5])
    assert np.array_equal(Rf, Hf) and np.array_equal(zf, res[:5])


def test_ekf_update_against_textbook(oracle):
    rng = np.random.default_rng(5)
    N = 60
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    off, sz = [15, 40, 27], [6, 8, 6]
    cols = sum([list(range(o, o + s)) for o, s in zip(off, sz)], [])
    H = rng.standard_normal((12, 20))
    res = rng.standard_normal(12)
    st, Pn, dx = oracle.ekf_update(P, off, sz, H, res, sigma2=0.7)
    assert st == capi.OVB_OK
    Pref, dxref = np_twin.ekf_update(P, cols, H, res, np.full(12, 0.7))
    assert np.linalg.norm(Pn - Pref) <= 1e-12 * np.linalg.norm(Pref)
    assert np.linalg.norm(dx - dxref) <= 1e-12 * np.linalg.norm(dxref)
    assert np.array_equal(Pn, Pn.T)
    Rd = rng.uniform(0.5, 2.0, 12)
    st, Pn2, dx2 = oracle.ekf_update(P, off, sz, H, res, Rdiag=Rd)
    Pref2, dxref2 = np_twin.ekf_update(P, cols, H, res, Rd)
    assert np.linalg.norm(Pn2 - Pref2) <= 1e-12 * np.linalg.norm(Pref2)
    assert np.linalg.norm(dx2 - dxref2) <= 1e-12 * np.linalg.norm(dxref2)


def test_cov_structure_ops(oracle):
    rng = np.random.default_rng(6)
    N = 33
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    # clone (StateHelper.cpp:371-373)
    Pc = oracle.cov_clone(P, 3, 6)
    assert Pc.shape == (N + 6, N + 6)
    assert np.array_equal(Pc[:N, :N], P) and np.array_equal(Pc[N:, N:], P[3:9, 3:9]) and np.array_equal(Pc[:N, N:], P[:, 3:9])
    assert np.array_equal(Pc, Pc.T)
    # clone + time offset Jacobian: J = [I; e_pose rows + dnc e_dt'] so P' = J P J'
    dnc = rng.standard_normal(6)
    dt = 20
    Pd = oracle.cov_clone(P, 3, 6, dnc, dt)
    J = np.zeros((N + 6, N))
    J[:N] = np.eye(N)
    J[N:, 3:9] = np.eye(6)
    J[N:, dt] += dnc
    assert np.linalg.norm(Pd - J @ P @ J.T) <= 1e-13 * np.linalg.norm(P)
    # marginalize
    Pm = oracle.cov_marginalize(P, 9, 6)
    keep = [i for i in range(N) if not 9 <= i < 15]
    assert np.allclose(Pm, P[np.ix_(keep, keep)], rtol=0, atol=0)
    # marginal blocks
    assert np.array_equal(oracle.cov_get_marginal(P, [20, 2], [3, 4]), P[np.ix_([20, 21, 22, 2, 3, 4, 5], [20, 21, 22, 2, 3, 4, 5])])
    # propagation (StateHelper.cpp:80-100)
    p = 15
    Phi = np.eye(p) + 0.01 * rng.standard_normal((p, p))
    Qh = rng.standard_normal((p, p))
    Q = Qh @ Qh.T * 1e-4
    st, Pp = oracle.cov_propagate(P, 0, Phi, Q, [0], [15])
    F = np.eye(N)
    F[:p, :p] = Phi
    Qf = np.zeros((N, N))
    Qf[:p, :p] = Q
    assert st == 0
    assert np.linalg.norm(Pp - (F @ P @ F.T + Qf)) <= 1e-13 * np.linalg.norm(P)


@pytest.mark.parametrize("cfg", [dict(n_feats=30, n_clones=8, n_cams=1), dict(n_feats=30, n_clones=8, n_cams=2, calib_ext=True, calib_intr=True)])
def test_whole_update_is_consistent(oracle, cfg):
    """compression must not change the update; the stacked system must equal the per-feature blocks."""
    case = _case(seed=17, **cfg)
    calib = cfg.get("calib_ext", False)
    opts = capi.default_opts(do_calib_camera_pose=int(calib), do_calib_camera_intrinsics=int(calib))
    r = oracle.msckf_update(case.frame, case.feats, opts, case.P)
    assert r["status"] == capi.OVB_OK
    st = r["stats"]
    assert st.rows_stacked == sum(2 * (case.feats.meas_off[f + 1] - case.feats.meas_off[f]) - 3 for f in range(case.feats.n_feats)
                                  if r["out"].status[f] == 0)
    Hb, rb, Hc, rc = r["H_big"], r["res_big"], r["H_cmp"], r["res_cmp"]
    assert Hc.shape == (min(Hb.shape), Hb.shape[1])
    assert np.linalg.norm(Hc.T @ Hc - Hb.T @ Hb) <= 1e-12 * np.linalg.norm(Hb.T @ Hb)
    assert np.linalg.norm(Hc.T @ rc - Hb.T @ rb) <= 1e-12 * np.linalg.norm(Hb.T @ rb)
    cols = sum([list(range(o, o + s)) for o, s in zip(r["order_off"], r["order_sz"])], [])
    Pref, dxref = np_twin.ekf_update(case.P, cols, Hb, rb, np.ones(len(rb)))
    assert np.linalg.norm(r["P"] - Pref) <= 1e-9 * np.linalg.norm(Pref)
    assert np.linalg.norm(r["dx"] - dxref) <= 1e-9 * np.linalg.norm(dxref)
    # information was gained, covariance stays symmetric positive
    assert np.array_equal(r["P"], r["P"].T)
    assert np.all(np.diag(r["P"]) <= np.diag(case.P) + 1e-18) and np.linalg.eigvalsh(r["P"]).min() > -1e-12
    # rejected features are reported with the reference's reasons
    assert set(np.unique(r["out"].status)).issubset({0, 2, 3, 5, 6, 8})


@pytest.mark.parametrize("model", [0, 1])
def test_camera_model_against_opencv(oracle, model):
    """Third-party pin of CamRadtan / CamEqui (ov_core/src/cam/CamRadtan.h:89-185, CamEqui.h:91-199): OpenCV's own projection
    functions — the library the reference links and whose conventions its camera classes restate — give the distorted pixel and
    its analytic Jacobians with respect to the normalised point and to (fx, fy, cx, cy, d1..d4). cv2.projectPoints is the
    plumb-bob model (k1, k2, p1, p2), cv2.fisheye.projectPoints the equidistant one (k1..k4)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    intr = np.array(sim._INTR[0] if model == 0 else sim._INTR_EQUI, dtype=np.float64)
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
    D = intr[4:8].copy()
    uv, dzn, dzeta = np.zeros(2), np.zeros(4), np.zeros(16)
    zero3 = np.zeros(3)
    for _ in range(40):
        x, y = rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4)
        X = np.array([[[x, y, 1.0]]])
        if model == 0:
            img, jac = cv2.projectPoints(X, zero3, zero3, K, D)
            # columns: rvec(3) tvec(3) f(2) c(2) dist(k1 k2 p1 p2 ...)
            d_point = jac[:, 3:5]  # d uv / d (X, Y) at Z = 1 = d uv / d (xn, yn)
            d_intr = np.hstack([jac[:, 6:8], jac[:, 8:10], jac[:, 10:14]])
        else:
            img, jac = cv2.fisheye.projectPoints(X, zero3, zero3, K, D)
            # columns: f(2) c(2) k(4) rvec(3) tvec(3) alpha(1)
            d_point = jac[:, 11:13]
            d_intr = jac[:, 0:8]
        ref = np.asarray(img, dtype=np.float64).ravel()
        oracle.lib().ovo_distort_d(C.c_int(model), intr.ctypes.data_as(capi.c_double_p), C.c_double(x), C.c_double(y),
                                   uv.ctypes.data_as(capi.c_double_p))
        assert np.all(np.abs(uv - ref) <= 4e-5)  # distort_d rounds the pixel to float32 (cam/CamBase.h:130-135)
        oracle.lib().ovo_distort_jacobian(C.c_int(model), intr.ctypes.data_as(capi.c_double_p), C.c_double(x), C.c_double(y),
                                          dzn.ctypes.data_as(capi.c_double_p), dzeta.ctypes.data_as(capi.c_double_p))
        assert np.allclose(dzn.reshape(2, 2), d_point, rtol=1e-9, atol=1e-9)
        assert np.allclose(dzeta.reshape(2, 8), d_intr, rtol=1e-9, atol=1e-9)


def test_gauss_newton_against_scipy_least_squares(oracle):
    """Third-party pin of single_gaussnewton (feat/FeatureInitializer.cpp:197-335): the refined point minimises the
    reprojection error in the anchor frame over (alpha, beta, rho); scipy.optimize.least_squares on the same cost (double
    precision, no float casts) lands on the same minimiser up to the reference's own stopping rule (min_dx = 1e-6 on the
    inverse-depth parameters) and float32 residuals."""
    from scipy.optimize import least_squares
    case = _case(n_feats=60, n_cams=2, n_clones=12, seed=9, outlier_frac=0.0, degenerate_frac=0.0)
    out, _ = oracle.triangulate(case.frame, case.feats, capi.default_opts(refine_features=1))
    ok = np.nonzero(out.status == capi.FEAT_OK)[0]
    assert len(ok) > 40
    fr, fb = case.frame, case.feats
    worst = 0.0
    for f in ok[:25]:
        acam, acl = int(out.anchor_cam[f]), int(out.anchor_clone[f])
        R_GtoA = fr.cam_R[acam].reshape(3, 3) @ fr.clone_R[acl].reshape(3, 3)
        p_AinG = fr.clone_p[acl] - R_GtoA.T @ fr.cam_p[acam]
        rows = range(fb.meas_off[f], fb.meas_off[f + 1])
        poses = []
        for i in rows:
            cam, cl = int(fb.cam[i]), int(fb.clone[i])
            R_GtoC = fr.cam_R[cam].reshape(3, 3) @ fr.clone_R[cl].reshape(3, 3)
            p_CinG = fr.clone_p[cl] - R_GtoC.T @ fr.cam_p[cam]
            poses.append((R_GtoC @ R_GtoA.T, R_GtoA @ (p_CinG - p_AinG), fb.uvn[i].astype(np.float64)))

        def resid(x):
            a, b, rho = x
            r = []
            for R_AtoC, p_CinA, z in poses:
                h = R_AtoC @ (np.array([a, b, 1.0]) - rho * p_CinA)
                r += [z[0] - h[0] / h[2], z[1] - h[1] / h[2]]
            return np.array(r)
        pA = out.p_FinA[f]
        x0 = np.array([pA[0] / pA[2], pA[1] / pA[2], 1.0 / pA[2]])
        sol = least_squares(resid, x0, method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14)
        # the oracle's point is (numerically) a stationary point already: scipy moves it by less than the LM stopping tolerance
        worst = max(worst, float(np.abs(sol.x - x0).max()))
        assert np.sum(resid(x0) ** 2) <= np.sum(sol.fun ** 2) * (1 + 1e-6) + 1e-12
    assert worst <= 5e-6  # measured 2.1e-7


def test_residual_and_feature_jacobian_against_opencv(oracle):
    """Third-party pin of the measurement model in get_feature_jacobian_full (update/UpdaterHelper.cpp:313-393): for every
    measurement, the residual is the tracked pixel minus OpenCV's projection of p_FinG through the clone / extrinsic pose, and
    H_f (GLOBAL_3D) is OpenCV's Jacobian with respect to the camera-frame point times R_GtoC."""
    cv2 = pytest.importorskip("cv2")
    case = _case(n_feats=12, n_cams=2, n_clones=8, seed=21, outlier_frac=0.0, degenerate_frac=0.0)
    opts = capi.default_opts(feat_rep=capi.REP_GLOBAL_3D, do_fej=0)
    fr, fb = case.frame, case.feats
    tri, _ = oracle.triangulate(fr, fb, opts)
    cols = np.concatenate([np.arange(o, o + 6) for o in fr.clone_off])
    Hf, Hx, res, row_off = oracle.feature_jacobians(fr, fb, opts, tri, 0, cols)
    checked = 0
    for f in np.nonzero(tri.status == capi.FEAT_OK)[0]:
        r0 = int(row_off[f])
        for k, i in enumerate(range(fb.meas_off[f], fb.meas_off[f + 1])):
            cam, cl = int(fb.cam[i]), int(fb.clone[i])
            R_GtoC = fr.cam_R[cam].reshape(3, 3) @ fr.clone_R[cl].reshape(3, 3)
            t = fr.cam_R[cam].reshape(3, 3) @ (-fr.clone_R[cl].reshape(3, 3) @ fr.clone_p[cl]) + fr.cam_p[cam]
            intr = fr.cam_intr[cam]
            K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
            img, jac = cv2.projectPoints(tri.p_FinG[f].reshape(1, 1, 3), cv2.Rodrigues(R_GtoC)[0], t, K, intr[4:8])
            ref_res = fb.uv[i].astype(np.float64) - img.ravel()
            assert np.all(np.abs(res[r0 + 2 * k:r0 + 2 * k + 2] - ref_res) <= 6e-5)  # distort_d rounds the pixel to float32
            ref_Hf = jac[:, 3:6] @ R_GtoC
            got = Hf[r0 + 2 * k:r0 + 2 * k + 2]
            assert np.allclose(got, ref_Hf, rtol=1e-9, atol=1e-9 * np.abs(ref_Hf).max())
            checked += 1
    assert checked > 60


@pytest.mark.parametrize("calib", [False, True])
def test_msckf_gate_against_scipy_nullspace(oracle, calib):
    """Third-party pin of nullspace_project_inplace + the chi² gate (update/UpdaterHelper.cpp:426-454, UpdaterMSCKF.cpp:209-234):
    chi² does not depend on the basis of the left nullspace of H_f, so scipy.linalg.null_space (LAPACK SVD) in place of the
    reference's Givens sweep, numpy's solve in place of its LLT, and the pre-nullspace Jacobians must reproduce the value the
    oracle reports for every feature that reaches the gate; and the gate decision follows scipy's 0.95 quantile."""
    from scipy.linalg import null_space
    from scipy.stats import chi2 as chi2_dist
    case = _case(n_feats=40, n_cams=2, n_clones=9, seed=31, calib_ext=calib, calib_intr=calib)
    opts = capi.default_opts(do_calib_camera_pose=int(calib), do_calib_camera_intrinsics=int(calib))
    r = oracle.msckf_update(case.frame, case.feats, opts, case.P)
    out = r["out"]
    N = case.P.shape[0]
    cols = np.arange(N)
    pre = capi.FeatOut(case.feats.n_feats)
    for k in ("status", "p_FinA", "p_FinG", "anchor_cam", "anchor_clone"):
        getattr(pre, k)[...] = getattr(out, k)
    reached = np.isin(out.status, (capi.FEAT_OK, capi.FEAT_CHI2))
    pre.status[reached] = capi.FEAT_OK  # rebuild the Jacobians of everything that was triangulated, gated or not
    Hf, Hx, res, row_off = oracle.feature_jacobians(case.frame, case.feats, opts, pre, 0, cols)
    s2 = float(opts.sigma_pix) ** 2
    n = 0
    for f in np.nonzero(reached)[0]:
        a, b = int(row_off[f]), int(row_off[f + 1])
        if b - a < 4:
            continue
        Nl = null_space(Hf[a:b].T)  # (2M) x (2M-3)
        Ho, ro = Nl.T @ Hx[a:b], Nl.T @ res[a:b]
        S = Ho @ case.P @ Ho.T + s2 * np.eye(Ho.shape[0])
        chi2 = float(ro @ np.linalg.solve(S, ro))
        assert abs(chi2 - out.chi2[f]) <= 1e-9 * chi2
        accept = chi2 <= float(opts.chi2_multipler) * chi2_dist.ppf(0.95, Ho.shape[0])
        assert (out.status[f] == capi.FEAT_OK) == accept
        n += 1
    assert n > 20
