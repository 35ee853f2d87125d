"""GPU tests of the normal-equations compression (OVB_COMPRESS_NORMAL_EQUATIONS) and of both compression modes end to end."""
import numpy as np
import pytest

from open_vins_b200 import capi, sim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = capi.Engine(max_state=256, max_feats=1024, max_meas=1024 * 48)
    yield e
    e.close()


@pytest.mark.parametrize("shape", [(300, 40), (1000, 86), (5000, 154), (25000, 154), (2500, 194), (130, 126), (60, 90), (17, 16)])
def test_compress_gram_parity(eng, oracle, shape):
    m, n = shape
    rng = np.random.default_rng(m + 7 * n)
    H = rng.standard_normal((m, n))
    res = rng.standard_normal(m)
    R, z = eng.compress(H, res, mode=capi.COMPRESS_NORMAL_EQUATIONS)
    assert np.allclose(np.tril(R, -1), 0.0, atol=0) and (np.diag(R) >= 0).all() and np.isfinite(R).all()
    if m <= n:
        return  # rank deficient by shape: only finiteness and zero rows are promised (see the rank-deficient test)
    G = H.T @ H
    assert np.linalg.norm(R.T @ R - G) <= 1e-12 * np.linalg.norm(G)
    assert np.linalg.norm(R.T @ z - H.T @ res) <= 1e-12 * np.linalg.norm(H) * np.linalg.norm(res)
    if m > 2 * n:
        # well conditioned full column rank: the Cholesky factor of H'H IS the reference's Givens R (diag >= 0)
        Rr, zr = oracle.compress(H, res)
        assert np.abs(R - Rr).max() <= 1e-11 * np.abs(Rr).max()
        assert np.abs(z - zr).max() <= 1e-11 * np.abs(zr).max()


def test_compress_gram_rank_deficient(eng):
    H, res, _ = sim.make_compress_case(m=3000, n=120, seed=3, structured=True)
    H[:, 7] = 0.0
    H[:, 30] = H[:, 31]
    R, z = eng.compress(H, res, mode=capi.COMPRESS_NORMAL_EQUATIONS)
    G = H.T @ H
    assert np.isfinite(R).all() and np.isfinite(z).all()
    assert np.linalg.norm(R.T @ R - G) <= 1e-12 * np.linalg.norm(G)
    assert np.linalg.norm(R.T @ z - H.T @ res) <= 1e-11 * np.linalg.norm(H) * np.linalg.norm(res)
    assert not R[7].any()  # the unused variable contributes a zero row, not noise


CASES = [
    dict(n_feats=50, n_clones=12, n_cams=1, seed=1),
    dict(n_feats=50, n_clones=12, n_cams=1, seed=2, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),
    dict(n_feats=120, n_clones=21, n_cams=2, seed=3, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),
    dict(n_feats=400, n_clones=21, n_cams=2, seed=42, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),
    dict(n_feats=6, n_clones=8, n_cams=1, seed=9),  # fewer rows than columns
]


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("mode", [capi.COMPRESS_HOUSEHOLDER_TSQR, capi.COMPRESS_NORMAL_EQUATIONS])
@pytest.mark.parametrize("order", [capi.COLS_CANONICAL, capi.COLS_REFERENCE_FIRST_SEEN])
def test_update_parity_both_compressions(eng, oracle, cfg, mode, order):
    case = sim.make_update_case(**cfg)
    opts = capi.default_opts(do_calib_camera_pose=int(case.meta["calib_ext"]), do_calib_camera_intrinsics=int(case.meta["calib_intr"]),
                             compress=mode, col_order=order)
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P, dumps=False)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.msckf_update(case.frame, case.feats, opts)
    P = eng.cov_get()
    assert st == ref["status"] == 0
    assert np.array_equal(out.status, ref["out"].status)
    # Householder TSQR: the 1e-9 bar of BASELINE.json. Normal equations: the squared condition number costs accuracy on
    # weakly observable calibration states (include/ovb200.h, ovb_compress_mode): 1e-9 without calibration columns, 1e-5 with
    calib = bool(case.meta["calib_ext"] or case.meta["calib_intr"])
    tol = 1e-9 if (mode == capi.COMPRESS_HOUSEHOLDER_TSQR or not calib) else 1e-5
    assert np.linalg.norm(P - ref["P"]) <= tol * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= tol * np.linalg.norm(ref["dx"])
    assert np.array_equal(P, P.T)
