"""BASELINE.json configs 3 and 5 at full size on the GPU, checked through size-independent properties and an independent
numpy restatement of the update (the oracle's Givens compression would need minutes at these sizes)."""
import numpy as np
import pytest

from open_vins_b200 import capi, sim

pytestmark = pytest.mark.gpu


def _info_form_update(P, H, res, cols, sigma2):
    """P+ = (I + P A)^-1 P and dx = P+ b with A = H'H/s2, b = H'r/s2 scattered to the state: the textbook EKF update
    P - P H'(H P H' + s2 I)^-1 H P in information form (valid for singular P, never forms the m x m innovation)."""
    N = P.shape[0]
    A = np.zeros((N, N))
    A[np.ix_(cols, cols)] = (H.T @ H) / sigma2
    b = np.zeros(N)
    b[cols] = (H.T @ res) / sigma2
    Pp = np.linalg.solve(np.eye(N) + P @ A, P)
    return 0.5 * (Pp + Pp.T), Pp @ b


@pytest.mark.parametrize("F,n_cams,n_clones", [(1024, 2, 21), (4096, 1, 11), (4096, 2, 21)])
def test_config3_feature_sweep(F, n_cams, n_clones):
    """Config 3: synthetic batches of up to 4096 features. The engine's own projected Jacobians (parity-tested against the
    oracle at small sizes) are pulled back through the staged API and the update is rebuilt with numpy."""
    case = sim.make_update_case(n_feats=F, n_clones=n_clones, n_cams=n_cams, seed=300 + F, calib_ext=True, calib_intr=True)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, col_order=capi.COLS_CANONICAL)
    eng = capi.Engine(max_state=256, max_feats=4096, max_meas=4096 * 2 * 21)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.msckf_update(case.frame, case.feats, opts)
    P = eng.cov_get()
    assert st == 0 and stats.n_feats_in == F
    assert stats.n_feats_used > F // 2
    assert np.array_equal(P, P.T) and np.isfinite(P).all() and np.isfinite(dx).all()
    assert np.linalg.eigvalsh(P).min() > -1e-11 * np.abs(P).max()
    assert np.all(np.diag(P) <= np.diag(case.P) * (1 + 1e-12))
    # independent rebuild: triangulate -> projected Jacobians (gated features come back as zero rows) -> numpy update
    eng.cov_set(case.P)
    tri = eng.triangulate(case.frame, case.feats, opts)
    assert np.array_equal(tri.status == 0, (out.status == 0) | (out.status == capi.FEAT_CHI2))
    ok = tri.status == 0
    np.testing.assert_array_equal(tri.p_FinG[ok], out.p_FinG[ok])
    Hf, Hx, res, row_off, col_index = eng.feature_jacobians(case.frame, case.feats, opts, tri, stage=1)
    assert stats.rows_stacked == sum(int(row_off[f + 1] - row_off[f]) for f in range(F) if out.status[f] == 0)
    assert not np.any(Hx[np.repeat(out.status != 0, np.diff(row_off))])  # rejected features contribute nothing
    Pn, dxn = _info_form_update(case.P, Hx, res, col_index, float(opts.sigma_pix) ** 2)
    assert np.linalg.norm(P - Pn) <= 1e-8 * np.linalg.norm(Pn)
    assert np.linalg.norm(dx - dxn) <= 1e-7 * np.linalg.norm(dxn)
    eng.close()


@pytest.mark.parametrize("structured", [False, True])
def test_config5_tsqr_ekf_microbench(structured):
    """Config 5: H in R^{8000 x 500} (dense i.i.d. / MSCKF-structured block-sparse), N = n = 500, P = A A'/500 + 1e-4 I."""
    m, n = 8000, 500
    rng = np.random.default_rng(0)
    if structured:
        H, res, _ = sim.make_compress_case(m=m, n=n, seed=0, structured=True)
    else:
        H = rng.standard_normal((m, n))
        res = rng.standard_normal(m)
    A = rng.standard_normal((n, n))
    P = A @ A.T / n + 1e-4 * np.eye(n)
    eng = capi.Engine(max_state=512, max_feats=64, max_meas=4096, max_rows=8192)
    R, z = eng.compress(H, res)
    G = H.T @ H
    assert np.allclose(np.tril(R, -1), 0.0, atol=0) and (np.diag(R) >= 0).all()
    assert np.linalg.norm(R.T @ R - G) <= 1e-12 * np.linalg.norm(G)
    assert np.linalg.norm(R.T @ z - H.T @ res) <= 1e-12 * np.linalg.norm(H) * np.linalg.norm(res)
    if not structured:  # full column rank: R is unique, compare with LAPACK's Householder QR (sign-normalised)
        Rq = np.linalg.qr(np.column_stack([H, res]), mode="r")
        sgn = np.sign(np.diag(Rq)[:n])
        assert np.abs(R - sgn[:, None] * Rq[:n, :n]).max() <= 1e-10 * np.abs(Rq).max()
        assert np.abs(z - sgn * Rq[:n, n]).max() <= 1e-10 * np.abs(Rq).max()
    # EKF update from the raw 8000-row system (compress-then-update inside the call) and from the compressed one
    off, sz = [0], [n]
    eng.cov_set(P)
    st, dx = eng.ekf_update(off, sz, H, res, sigma2=1.0)
    Pg = eng.cov_get()
    assert st == 0
    Pn, dxn = _info_form_update(P, H, res, np.arange(n), 1.0)
    assert np.linalg.norm(Pg - Pn) <= 1e-9 * np.linalg.norm(Pn)
    assert np.linalg.norm(dx - dxn) <= 1e-9 * np.linalg.norm(dxn)
    eng.cov_set(P)
    st2, dx2 = eng.ekf_update(off, sz, R, z, sigma2=1.0)
    assert st2 == 0
    assert np.linalg.norm(eng.cov_get() - Pn) <= 1e-9 * np.linalg.norm(Pn)
    assert np.linalg.norm(dx2 - dxn) <= 1e-9 * np.linalg.norm(dxn)
    eng.close()


def test_config4_window_msckf_and_slam(oracle):
    """Config 4's window (4 cameras, 31 clone poses, tracks of up to 124 measurements: the per-feature innovation is
    245 x 245 and lives in the kernel's global-memory scratch instead of shared memory) with online calibration, then a
    SLAM update of 25 landmarks: 31 + 8 + 25 = 64 state variables, the engine's per-call limit. Oracle parity at a
    feature count the CPU restatement finishes in about a second."""
    ms = sim.make_update_case(n_feats=40, n_clones=31, n_cams=4, seed=4, calib_ext=True, calib_intr=True)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, col_order=capi.COLS_CANONICAL)
    assert int((ms.feats.meas_off[1:] - ms.feats.meas_off[:-1]).max()) == 124
    eng = capi.Engine(max_state=384, max_feats=256, max_meas=8192)
    eng.cov_set(ms.P)
    st, out, dx, stats = eng.msckf_update(ms.frame, ms.feats, opts)
    ref = oracle.msckf_update(ms.frame, ms.feats, opts, ms.P, dumps=False)
    assert st == ref["status"] == 0 and np.array_equal(out.status, ref["out"].status) and stats.n_feats_used > 30
    ok = ref["out"].status == 0
    np.testing.assert_allclose(out.chi2[ok], ref["out"].chi2[ok], rtol=1e-8)
    assert np.linalg.norm(eng.cov_get() - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])
    sl = sim.make_slam_case(n_landmarks=25, n_clones=31, n_cams=4, seed=4, rep=capi.REP_GLOBAL_3D)
    eng.cov_set(sl.P)
    st, out, dx, stats = eng.slam_update(sl.frame, sl.feats, sl.landmarks, opts)
    ref = oracle.slam_update(sl.frame, sl.feats, sl.landmarks, opts, sl.P)
    assert st == ref["status"] == 0 and np.array_equal(out.status, ref["out"].status) and stats.n_feats_used > 15
    assert stats.cols_stacked == ref["stats"].cols_stacked
    assert np.linalg.norm(eng.cov_get() - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])
    # one landmark more than the limit: a clear capacity error, not a wrong answer
    sl2 = sim.make_slam_case(n_landmarks=26, n_clones=31, n_cams=4, seed=4, rep=capi.REP_GLOBAL_3D)
    eng2 = capi.Engine(max_state=384, max_feats=256, max_meas=8192)
    eng2.cov_set(sl2.P)
    with pytest.raises(capi.OvbError) as ei:
        eng2.slam_update(sl2.frame, sl2.feats, sl2.landmarks, opts)
    assert ei.value.code == capi.OVB_ERR_CAPACITY
    eng.close()
    eng2.close()
