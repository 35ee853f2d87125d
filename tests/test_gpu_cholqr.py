"""GPU tests of the shifted CholeskyQR2 compression (OVB_COMPRESS_CHOLQR2, csrc/k_cholqr.cu) and of the DMMA Cholesky it
shares with the EKF update. Reference for measurement_compress_inplace: ov_msckf/src/update/UpdaterHelper.cpp:456-487."""
import numpy as np
import pytest

from open_vins_b200 import capi, sim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = capi.Engine(max_state=256, max_feats=1024, max_meas=1024 * 48)
    yield e
    e.close()


def _posterior(R, z, P):
    """P+ and dx of an EKF update with unit noise from the compressed system, information form (float64 numpy)."""
    Pi = np.linalg.inv(P)
    Pp = np.linalg.inv(Pi + R.T @ R)
    return Pp, Pp @ (R.T @ z)


@pytest.mark.parametrize("shape", [(300, 40), (1000, 86), (5000, 154), (25000, 154), (22487, 154), (2500, 159), (130, 126), (60, 90), (17, 16),
                                   (9, 159), (4, 3), (40000, 7)])
def test_compress_cholqr2_parity(eng, oracle, shape):
    m, n = shape
    rng = np.random.default_rng(m + 7 * n)
    H = rng.standard_normal((m, n))
    res = rng.standard_normal(m)
    R, z = eng.compress(H, res, mode=capi.COMPRESS_CHOLQR2)
    assert np.array_equal(np.tril(R, -1), np.zeros_like(R)) and (np.diag(R) >= 0).all() and np.isfinite(R).all() and np.isfinite(z).all()
    G = H.T @ H
    # BASELINE.json: compressed H within 1e-12 rel, evaluated on the invariants
    assert np.linalg.norm(R.T @ R - G) <= 1e-12 * np.linalg.norm(G)
    assert np.linalg.norm(R.T @ z - H.T @ res) <= 1e-12 * np.linalg.norm(H) * np.linalg.norm(res)
    if m > 2 * n and m * n * n < 2e8:
        # well conditioned, full column rank: the factor IS the reference's Givens R (diag >= 0), row for row
        Rr, zr = oracle.compress(H, res)
        assert np.abs(R - Rr).max() <= 1e-11 * np.abs(Rr).max()
        assert np.abs(z - zr).max() <= 1e-11 * np.abs(zr).max()


@pytest.mark.parametrize("kappa", [1e3, 1e5, 1e7])
@pytest.mark.parametrize("n", [40, 154])
def test_compress_cholqr2_ill_conditioned(eng, kappa, n):
    """Columns with a wide range of scales AND a nearly dependent column set: the case that sank the one-pass Gram path
    (kappa^2 eps). The posterior built from [R z] must match the one from a Householder QR (numpy/LAPACK) at 1e-9."""
    m = 6000
    rng = np.random.default_rng(int(np.log10(kappa)) + n)
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = np.logspace(0, -np.log10(kappa), n)
    H = (U * s) @ V.T
    H *= np.logspace(0, -3, n)[rng.permutation(n)]  # badly scaled state variables on top
    H *= 300.0
    res = rng.standard_normal(m)
    R, z = eng.compress(H, res, mode=capi.COMPRESS_CHOLQR2)
    assert np.isfinite(R).all() and np.isfinite(z).all()
    Q, Rq = np.linalg.qr(np.column_stack([H, res]))
    Rh, zh = Rq[:n, :n], Rq[:n, n]
    A = rng.standard_normal((n, n))
    P = A @ A.T / n + 1e-2 * np.eye(n)
    Pp, dx = _posterior(R, z, P)
    Ph, dxh = _posterior(Rh, zh, P)
    assert np.linalg.norm(Pp - Ph) <= 1e-9 * np.linalg.norm(Ph)
    assert np.linalg.norm(dx - dxh) <= 1e-9 * np.linalg.norm(dxh)
    # column-scaled Gram error: no condition-number factor
    d = np.sqrt(np.diag(H.T @ H))
    E = (R.T @ R - H.T @ H) / np.outer(d, d)
    assert np.abs(E).max() <= 1e-11


def test_compress_cholqr2_rank_deficient(eng):
    H, res, _ = sim.make_compress_case(m=3000, n=120, seed=3, structured=True)
    H[:, 7] = 0.0
    H[:, 30] = H[:, 31]
    R, z = eng.compress(H, res, mode=capi.COMPRESS_CHOLQR2)
    G = H.T @ H
    assert np.isfinite(R).all() and np.isfinite(z).all()
    assert np.linalg.norm(R.T @ R - G) <= 1e-12 * np.linalg.norm(G)
    assert np.linalg.norm(R.T @ z - H.T @ res) <= 1e-11 * np.linalg.norm(H) * np.linalg.norm(res)
    # the shifts leave at most round-off level information in the empty / duplicated directions
    assert np.abs(R[7]).max() <= 1e-9 * np.abs(R).max()


def test_compress_cholqr2_zero_and_capacity(eng):
    H = np.zeros((50, 12))
    R, z = eng.compress(H, np.zeros(50), mode=capi.COMPRESS_CHOLQR2)
    assert not R.any() and not z.any()
    with pytest.raises(capi.OvbError):  # wider than the wide path as well (n + 1 > 513)
        eng.compress(np.ones((600, 520)), np.ones(600), mode=capi.COMPRESS_CHOLQR2)


CASES = [
    dict(n_feats=50, n_clones=12, n_cams=1, seed=1),
    dict(n_feats=50, n_clones=12, n_cams=1, seed=2, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),
    dict(n_feats=120, n_clones=21, n_cams=2, seed=3, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),
    dict(n_feats=400, n_clones=21, n_cams=2, seed=42, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),
    dict(n_feats=400, n_clones=21, n_cams=2, seed=0, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),
    dict(n_feats=6, n_clones=8, n_cams=1, seed=9),  # fewer rows than columns
]


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("order", [capi.COLS_CANONICAL, capi.COLS_REFERENCE_FIRST_SEEN])
def test_update_parity_cholqr2(eng, oracle, cfg, order):
    """Full MSCKF update with the CholeskyQR2 compression against the oracle (Givens compression): the 1e-9 bar of
    BASELINE.json on P and dx, with weakly observable calibration columns in the update."""
    case = sim.make_update_case(**cfg)
    opts = capi.default_opts(do_calib_camera_pose=int(case.meta["calib_ext"]), do_calib_camera_intrinsics=int(case.meta["calib_intr"]),
                             compress=capi.COMPRESS_CHOLQR2, col_order=order)
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P, dumps=False)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.msckf_update(case.frame, case.feats, opts)
    P = eng.cov_get()
    assert st == ref["status"] == 0
    assert np.array_equal(out.status, ref["out"].status)
    assert np.linalg.norm(P - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])
    assert np.array_equal(P, P.T)


@pytest.mark.parametrize("r", [1, 7, 8, 9, 16, 63, 64, 65, 120, 154, 159, 160])
def test_ekf_chol_dmma_sizes(eng, oracle, r):
    """StateHelper::EKFUpdate (state/StateHelper.cpp:116-197) through the DMMA Cholesky for every block-edge size."""
    N = 200
    rng = np.random.default_rng(r)
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    off, sz = [3], [max(r, 6)]
    n = sz[0]
    H = rng.standard_normal((r, n))
    res = rng.standard_normal(r)
    st_r, P_r, dx_r = oracle.ekf_update(P, off, sz, H, res, sigma2=0.25)
    eng.cov_set(P)
    st_g, dx_g = eng.ekf_update(off, sz, H, res, sigma2=0.25)
    P_g = eng.cov_get()
    assert st_g == st_r == capi.OVB_OK
    assert np.linalg.norm(P_g - P_r) <= 1e-11 * np.linalg.norm(P_r)
    assert np.linalg.norm(dx_g - dx_r) <= 1e-10 * np.linalg.norm(dx_r)
    assert np.array_equal(P_g, P_g.T)


def test_ekf_not_spd_leaves_P(eng):
    """A failed innovation-covariance factorisation must not touch the resident covariance (status is recoverable)."""
    N = 60
    rng = np.random.default_rng(5)
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    P[10:16, 10:16] -= 50.0 * np.eye(6)  # indefinite "covariance": S = H P H' + R is not SPD
    eng.cov_set(P)
    H = np.eye(6)
    st, dx = eng.ekf_update([10], [6], H, np.ones(6), sigma2=1.0, allow=(capi.OVB_ERR_NOT_SPD,))
    assert st == capi.OVB_ERR_NOT_SPD
    assert np.array_equal(eng.cov_get(), P)
    assert not dx.any()


# ---------------------------------------------------------------------------------------------------------------- wide systems
@pytest.fixture(scope="module")
def eng_wide():
    e = capi.Engine(max_state=640, max_feats=256, max_meas=8192, max_rows=16384)
    yield e
    e.close()


@pytest.mark.parametrize("shape", [(700, 160), (3000, 200), (8000, 243), (6000, 317), (8000, 500), (1200, 511), (300, 400)])
def test_compress_cholqr2_wide(eng_wide, shape):
    """More columns than one CTA's Cholesky takes (configs 4 and 5): blocked DMMA factorisation / solve. Same invariants."""
    m, n = shape
    rng = np.random.default_rng(m + 3 * n)
    H = rng.standard_normal((m, n))
    res = rng.standard_normal(m)
    R, z = eng_wide.compress(H, res, mode=capi.COMPRESS_CHOLQR2)
    assert np.array_equal(np.tril(R, -1), np.zeros_like(R)) and (np.diag(R) >= 0).all() and np.isfinite(R).all() and np.isfinite(z).all()
    G = H.T @ H
    assert np.linalg.norm(R.T @ R - G) <= 1e-12 * np.linalg.norm(G)
    assert np.linalg.norm(R.T @ z - H.T @ res) <= 1e-12 * np.linalg.norm(H) * np.linalg.norm(res)
    if m > 2 * n:
        Rq = np.linalg.qr(np.column_stack([H, res]), mode="r")
        sgn = np.sign(np.diag(Rq)[:n])
        assert np.abs(R - sgn[:, None] * Rq[:n, :n]).max() <= 1e-10 * np.abs(Rq).max()
        assert np.abs(z - sgn * Rq[:n, n]).max() <= 1e-10 * np.abs(Rq).max()


def test_compress_cholqr2_wide_ill_conditioned(eng_wide):
    m, n, kappa = 5000, 300, 1e6
    rng = np.random.default_rng(11)
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    H = (U * np.logspace(0, -np.log10(kappa), n)) @ V.T
    H *= np.logspace(0, -3, n)[rng.permutation(n)] * 300.0
    res = rng.standard_normal(m)
    R, z = eng_wide.compress(H, res, mode=capi.COMPRESS_CHOLQR2)
    Rq = np.linalg.qr(np.column_stack([H, res]), mode="r")
    A = rng.standard_normal((n, n))
    P = A @ A.T / n + 1e-2 * np.eye(n)
    Pp, dx = _posterior(R, z, P)
    Ph, dxh = _posterior(Rq[:n, :n], Rq[:n, n], P)
    assert np.linalg.norm(Pp - Ph) <= 1e-9 * np.linalg.norm(Ph)
    assert np.linalg.norm(dx - dxh) <= 1e-9 * np.linalg.norm(dxh)


@pytest.mark.parametrize("r", [161, 200, 256, 257, 384, 500])
def test_ekf_wide_sizes(eng_wide, oracle, r):
    """EKFUpdate with an innovation covariance wider than 160: blocked Cholesky + blocked gain solve."""
    N = 520
    rng = np.random.default_rng(r)
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    off, sz, n = [4], [r + 3], r + 3
    H = rng.standard_normal((r, n))
    res = rng.standard_normal(r)
    st_r, P_r, dx_r = oracle.ekf_update(P, off, sz, H, res, sigma2=0.5)
    eng_wide.cov_set(P)
    st_g, dx_g = eng_wide.ekf_update(off, sz, H, res, sigma2=0.5)
    P_g = eng_wide.cov_get()
    assert st_g == st_r == capi.OVB_OK
    assert np.linalg.norm(P_g - P_r) <= 1e-10 * np.linalg.norm(P_r)
    assert np.linalg.norm(dx_g - dx_r) <= 1e-9 * np.linalg.norm(dx_r)
    assert np.array_equal(P_g, P_g.T)


def test_config4_msckf_cholqr2(eng_wide, oracle):
    """Config-4 window (4 cameras, 31 clone poses, full calibration: 242 stacked columns) through the wide CholeskyQR2."""
    ms = sim.make_update_case(n_feats=60, n_clones=31, n_cams=4, seed=4, calib_ext=True, calib_intr=True)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, col_order=capi.COLS_CANONICAL, compress=capi.COMPRESS_CHOLQR2)
    eng_wide.cov_set(ms.P)
    st, out, dx, stats = eng_wide.msckf_update(ms.frame, ms.feats, opts)
    ref = oracle.msckf_update(ms.frame, ms.feats, opts, ms.P, dumps=False)
    assert st == ref["status"] == 0 and np.array_equal(out.status, ref["out"].status) and stats.cols_stacked == 242
    assert np.linalg.norm(eng_wide.cov_get() - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])
