"""Third pin of the oracle: 50-digit mpmath recomputation of the two dense steps whose rounding the
double-precision numpy twins share with the oracle — the Givens compression invariants and the EKF update — on small cases."""
import mpmath as mp
import numpy as np
import pytest

mp.mp.dps = 50


def _mp(A):
    return mp.matrix(A.tolist())


def _np(M, shape):
    return np.array([[float(M[i, j]) for j in range(shape[1])] for i in range(shape[0])])


@pytest.mark.parametrize("seed,N,n,r", [(0, 18, 12, 12), (1, 24, 12, 7), (2, 26, 18, 18)])
def test_ekf_update_against_50_digit_arithmetic(oracle, seed, N, n, r):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    off = list(range(3, 3 + n, 6))
    sz = [6] * len(off)
    n = 6 * len(off)
    cols = np.concatenate([np.arange(o, o + 6) for o in off])
    H = rng.standard_normal((r, n))
    res = rng.standard_normal(r)
    s2 = 0.37
    st, P_o, dx_o = oracle.ekf_update(P, off, sz, H, res, sigma2=s2)
    assert st == 0
    Pm, Hm, rm = _mp(P), mp.zeros(r, N), _mp(res.reshape(-1, 1))
    for i in range(r):
        for j, c in enumerate(cols):
            Hm[i, int(c)] = mp.mpf(float(H[i, j]))
    S = Hm * Pm * Hm.T + mp.mpf(s2) * mp.eye(r)
    K = Pm * Hm.T * mp.inverse(S)
    Pn = _np(Pm - K * Hm * Pm, (N, N))
    dxn = _np(K * rm, (N, 1)).ravel()
    assert np.linalg.norm(P_o - Pn) <= 1e-12 * np.linalg.norm(Pn)
    assert np.linalg.norm(dx_o - dxn) <= 1e-12 * np.linalg.norm(dxn)


@pytest.mark.parametrize("seed,m,n", [(3, 40, 9), (4, 25, 12)])
def test_compression_invariants_against_50_digit_arithmetic(oracle, seed, m, n):
    """R'R = H'H and R'z = H'r hold to double rounding when H'H / H'r are formed in 50-digit arithmetic (the Givens sweep of
    UpdaterHelper.cpp:456-487 is orthogonal; numpy's own H'H carries its own rounding, this does not)."""
    rng = np.random.default_rng(seed)
    H = rng.standard_normal((m, n))
    res = rng.standard_normal(m)
    R, z = oracle.compress(H, res)
    Hm, rm = _mp(H), _mp(res.reshape(-1, 1))
    G = _np(Hm.T * Hm, (n, n))
    g = _np(Hm.T * rm, (n, 1)).ravel()
    assert np.linalg.norm(R.T @ R - G) <= 5e-15 * np.linalg.norm(G) * n
    assert np.linalg.norm(R.T @ z - g) <= 5e-15 * np.linalg.norm(H) * np.linalg.norm(res) * n
    # the rotated-away entries are computed, not set: they are zero to rounding, exactly as in the reference's Givens sweep
    assert (np.diag(R) >= 0).all() and np.abs(np.tril(R, -1)).max() <= 1e-14 * np.abs(R).max()
