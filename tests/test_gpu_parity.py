"""GPU parity tests: the CUDA path (through the C ABI of include/ovb200.h) against the CPU oracle on identical inputs.

Tolerances are BASELINE.json's: triangulated points and the compressed system 1e-12 relative (the compressed H on
the invariants of SURVEY.md App. A.6), post-update state/covariance 1e-9 relative Frobenius.
"""
import numpy as np
import pytest

from open_vins_b200 import capi, sim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = capi.Engine(max_state=256, max_feats=1024, max_meas=1024 * 48)
    yield e
    e.close()


def _opts(case, **kw):
    return capi.default_opts(do_calib_camera_pose=int(case.meta["calib_ext"]), do_calib_camera_intrinsics=int(case.meta["calib_intr"]), **kw)


CASES = [
    dict(n_feats=50, n_clones=12, n_cams=1, seed=1),                                            # BASELINE config 1 (no calib, N=87)
    dict(n_feats=50, n_clones=12, n_cams=1, seed=2, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),  # N=126
    dict(n_feats=120, n_clones=21, n_cams=2, seed=3, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True),  # config 2 layout
    dict(n_feats=64, n_clones=9, n_cams=2, seed=4, cam_model=1),                                 # equidistant cameras
    dict(n_feats=40, n_clones=8, n_cams=4, seed=5, calib_ext=True, mono_frac=0.3),               # 4 cameras, ragged tracks
]


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("tri1d", [0, 1])
def test_triangulation_parity(eng, oracle, cfg, tri1d):
    case = sim.make_update_case(**cfg)
    opts = _opts(case, triangulate_1d=tri1d)
    ref, _ = oracle.triangulate(case.frame, case.feats, opts)
    got = eng.triangulate(case.frame, case.feats, opts)
    assert np.array_equal(got.status, ref.status)
    assert np.array_equal(got.anchor_cam, ref.anchor_cam) and np.array_equal(got.anchor_clone, ref.anchor_clone)
    ok = ref.status == capi.FEAT_OK
    assert ok.sum() >= 0.5 * len(ok)
    rel = np.linalg.norm(got.p_FinG[ok] - ref.p_FinG[ok], axis=1) / np.linalg.norm(ref.p_FinG[ok], axis=1)
    relA = np.linalg.norm(got.p_FinA[ok] - ref.p_FinA[ok], axis=1) / np.linalg.norm(ref.p_FinA[ok], axis=1)
    # BASELINE.json: triangulated points within 1e-12 rel, for EVERY feature. The kernel accumulates in the reference's
    # measurement order (seq_add, csrc/k_triangulate.cu), so the float32 casts inside the LM loop see the same doubles as the
    # oracle and no rounding flip can separate the two (round 1 tolerated 1 % outliers up to 1e-6 from butterfly sums)
    assert rel.max() <= 1e-12 and relA.max() <= 1e-12, f"max rel {rel.max():.3e} / {relA.max():.3e}"
    exact = np.all(got.p_FinG[ok] == ref.p_FinG[ok], axis=1).mean()
    assert exact >= 0.99, f"only {exact:.3f} of the features bit-identical" 


@pytest.mark.parametrize("cfg", CASES[:4])
def test_jacobian_parity_pre_nullspace(eng, oracle, cfg):
    case = sim.make_update_case(**cfg)
    opts = _opts(case)
    eng.cov_set(case.P)
    tri, _ = oracle.triangulate(case.frame, case.feats, opts)
    Hf, Hx, res, row_off, cols = eng.feature_jacobians(case.frame, case.feats, opts, tri.copy(), 0)
    Hf_r, Hx_r, res_r, row_off_r = oracle.feature_jacobians(case.frame, case.feats, opts, tri.copy(), 0, cols)
    assert np.array_equal(row_off, row_off_r)
    assert np.array_equal(res, res_r)  # float-rounded pixels: exact
    sc = max(np.abs(Hx_r).max(), 1.0)
    assert np.abs(Hx - Hx_r).max() <= 1e-12 * sc
    assert np.abs(Hf - Hf_r).max() <= 1e-12 * max(np.abs(Hf_r).max(), 1.0)


@pytest.mark.parametrize("rep", [capi.REP_ANCHORED_MSCKF_INVERSE_DEPTH, capi.REP_ANCHORED_3D, capi.REP_GLOBAL_FULL_INVERSE_DEPTH,
                                 capi.REP_ANCHORED_FULL_INVERSE_DEPTH, capi.REP_ANCHORED_INVERSE_DEPTH_SINGLE])
@pytest.mark.parametrize("calib", [False, True])
def test_jacobian_parity_representations(eng, oracle, rep, calib):
    case = sim.make_update_case(n_feats=40, n_clones=10, n_cams=2, seed=21, calib_ext=calib, calib_intr=calib)
    opts = _opts(case, feat_rep=rep)
    eng.cov_set(case.P)
    tri, _ = oracle.triangulate(case.frame, case.feats, opts)
    Hf, Hx, res, row_off, cols = eng.feature_jacobians(case.frame, case.feats, opts, tri.copy(), 0)
    Hf_r, Hx_r, res_r, _ = oracle.feature_jacobians(case.frame, case.feats, opts, tri.copy(), 0, cols)
    assert np.abs(res - res_r).max() <= 1e-9  # anchored reps recompute p_FinG; float rounding of the pixel may flip
    assert np.abs(Hx - Hx_r).max() <= 1e-11 * max(np.abs(Hx_r).max(), 1.0)
    assert np.abs(Hf - Hf_r).max() <= 1e-11 * max(np.abs(Hf_r).max(), 1.0)


@pytest.mark.parametrize("cfg", CASES[:4])
def test_nullspace_and_gate_parity(eng, oracle, cfg):
    case = sim.make_update_case(**cfg)
    opts = _opts(case)
    eng.cov_set(case.P)
    tri, _ = oracle.triangulate(case.frame, case.feats, opts)
    out_g = tri.copy()
    _, Hx, res, row_off, cols = eng.feature_jacobians(case.frame, case.feats, opts, out_g, 1)
    out_r = tri.copy()
    _, Hx_r, res_r, row_off_r = oracle.feature_jacobians(case.frame, case.feats, opts, out_r, 1, cols, P=case.P)
    assert np.array_equal(row_off, row_off_r)
    assert np.array_equal(out_g.status, out_r.status), "chi² gate decisions differ"
    seen = np.isfinite(out_r.chi2)
    assert np.allclose(out_g.chi2[seen], out_r.chi2[seen], rtol=1e-9, atol=0)
    for f in range(case.feats.n_feats):
        a, b = row_off[f], row_off[f + 1]
        if out_r.status[f] != 0:
            assert not Hx[a:b].any() and not res[a:b].any()  # rejected rows are zero rows
            continue
        G, Gr = Hx[a:b].T @ Hx[a:b], Hx_r[a:b].T @ Hx_r[a:b]
        assert np.linalg.norm(G - Gr) <= 1e-12 * np.linalg.norm(Gr)
        g, gr = Hx[a:b].T @ res[a:b], Hx_r[a:b].T @ res_r[a:b]
        assert np.linalg.norm(g - gr) <= 1e-11 * (np.linalg.norm(Hx_r[a:b]) * np.linalg.norm(res_r[a:b]))
        assert abs(res[a:b] @ res[a:b] - res_r[a:b] @ res_r[a:b]) <= 1e-12 * (res_r[a:b] @ res_r[a:b])


@pytest.mark.parametrize("shape", [(300, 40), (1000, 86), (5000, 154), (2500, 194), (130, 126), (60, 90), (17, 16), (4000, 33)])
def test_compress_parity(eng, oracle, shape):
    m, n = shape
    rng = np.random.default_rng(m + n)
    H = rng.standard_normal((m, n))
    res = rng.standard_normal(m)
    R, z = eng.compress(H, res)
    assert np.allclose(np.tril(R, -1), 0.0, atol=0)
    G = H.T @ H
    assert np.linalg.norm(R.T @ R - G) <= 1e-12 * np.linalg.norm(G)
    assert np.linalg.norm(R.T @ z - H.T @ res) <= 1e-12 * np.linalg.norm(H) * np.linalg.norm(res)
    if m > n:
        Rr, zr = oracle.compress(H, res)
        assert (np.diag(R) >= 0).all()
        # full column rank, modest condition number: the Givens R of the reference is unique, compare element-wise
        assert np.abs(R - Rr).max() <= 1e-11 * np.abs(Rr).max()
        assert np.abs(z - zr).max() <= 1e-11 * np.abs(zr).max()


def test_compress_rank_deficient_and_structured(eng, oracle):
    H, res, _ = sim.make_compress_case(m=3000, n=120, seed=3, structured=True)
    H[:, 7] = 0.0            # an unused variable: zero column
    H[:, 30] = H[:, 31]      # exact dependency
    R, z = eng.compress(H, res)
    G = H.T @ H
    assert np.linalg.norm(R.T @ R - G) <= 1e-12 * np.linalg.norm(G)
    assert np.linalg.norm(R.T @ z - H.T @ res) <= 1e-12 * np.linalg.norm(H) * np.linalg.norm(res)


@pytest.mark.parametrize("N,n,r", [(87, 72, 72), (194, 154, 154), (141, 60, 25), (230, 126, 400)])
def test_ekf_update_parity(eng, oracle, N, n, r):
    rng = np.random.default_rng(N + n + r)
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    # variables: 6-wide blocks scattered through the state, not in ascending order
    nblk = n // 6
    starts = rng.permutation(np.arange(15, N - 6, 6))[:nblk]
    off, sz = list(map(int, starts)), [6] * nblk
    n = 6 * nblk
    H = rng.standard_normal((r, n))
    res = rng.standard_normal(r)
    st_r, P_r, dx_r = oracle.ekf_update(P, off, sz, H, res, sigma2=0.5)
    eng.cov_set(P)
    st_g, dx_g = eng.ekf_update(off, sz, H, res, sigma2=0.5)
    P_g = eng.cov_get()
    assert st_g == st_r == capi.OVB_OK
    assert np.linalg.norm(P_g - P_r) <= 1e-9 * np.linalg.norm(P_r)
    assert np.linalg.norm(dx_g - dx_r) <= 1e-9 * np.linalg.norm(dx_r)
    assert np.array_equal(P_g, P_g.T)
    Rd = rng.uniform(0.5, 2.0, r)
    st_r, P_r, dx_r = oracle.ekf_update(P, off, sz, H, res, Rdiag=Rd)
    eng.cov_set(P)
    st_g, dx_g = eng.ekf_update(off, sz, H, res, Rdiag=Rd)
    P_g = eng.cov_get()
    assert np.linalg.norm(P_g - P_r) <= 1e-9 * np.linalg.norm(P_r)
    assert np.linalg.norm(dx_g - dx_r) <= 1e-9 * np.linalg.norm(dx_r)


def test_cov_structure_ops_parity(eng, oracle):
    rng = np.random.default_rng(9)
    N = 101
    A = rng.standard_normal((N, N))
    P = A @ A.T / N + 1e-3 * np.eye(N)
    eng.cov_set(P)
    assert np.array_equal(eng.cov_get(), P)
    assert np.array_equal(eng.cov_get_marginal([40, 3], [6, 8]), oracle.cov_get_marginal(P, [40, 3], [6, 8]))
    eng.cov_clone(0, 6)
    Pc = oracle.cov_clone(P, 0, 6)
    assert np.array_equal(eng.cov_get(), Pc)
    dnc = rng.standard_normal(6)
    eng.cov_set(P)
    eng.cov_clone(0, 6, dnc, 39)
    Pd = oracle.cov_clone(P, 0, 6, dnc, 39)
    assert np.abs(eng.cov_get() - Pd).max() <= 1e-15 * np.abs(Pd).max()
    eng.cov_marginalize(60, 6)
    Pm = oracle.cov_marginalize(Pd, 60, 6)
    assert np.array_equal(eng.cov_get(), Pm) or np.abs(eng.cov_get() - Pm).max() <= 1e-15 * np.abs(Pm).max()
    for p, old in [(15, ([0], [15])), (39, ([0, 15, 21, 27, 36], [15, 6, 6, 9, 3])), (3, ([50], [3]))]:
        Phi = np.eye(p) + 0.01 * rng.standard_normal((p, p))
        Qh = rng.standard_normal((p, p))
        Q = Qh @ Qh.T * 1e-5
        new_off = old[0][0]
        Pin = eng.cov_get()
        st_r, Pp = oracle.cov_propagate(Pin, new_off, Phi, Q, old[0], old[1])
        st_g = eng.cov_propagate(new_off, Phi, Q, old[0], old[1])
        assert st_g == st_r == 0
        assert np.linalg.norm(eng.cov_get() - Pp) <= 1e-13 * np.linalg.norm(Pp)


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("order", [capi.COLS_REFERENCE_FIRST_SEEN, capi.COLS_CANONICAL])
def test_full_update_parity(eng, oracle, cfg, order):
    case = sim.make_update_case(**cfg)
    opts = _opts(case, col_order=order)
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.msckf_update(case.frame, case.feats, opts)
    P = eng.cov_get()
    assert st == ref["status"] == capi.OVB_OK
    assert np.array_equal(out.status, ref["out"].status)
    ok = ref["out"].status == 0
    rel = np.linalg.norm(out.p_FinG[ok] - ref["out"].p_FinG[ok], axis=1) / np.linalg.norm(ref["out"].p_FinG[ok], axis=1)
    assert rel.max() <= 1e-12
    seen = np.isfinite(ref["out"].chi2)
    assert np.allclose(out.chi2[seen], ref["out"].chi2[seen], rtol=1e-8)
    assert stats.n_feats_used == ref["stats"].n_feats_used and stats.rows_stacked == ref["stats"].rows_stacked
    assert stats.cols_stacked == ref["stats"].cols_stacked
    assert np.linalg.norm(P - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])
    assert np.array_equal(P, P.T)


def test_empty_and_degenerate_batches(eng, oracle):
    case = sim.make_update_case(n_feats=10, n_clones=6, n_cams=1, seed=31)
    opts = _opts(case)
    eng.cov_set(case.P)
    # no features: nothing happens (UpdaterMSCKF.cpp:61-62)
    empty = case.feats.subset([])
    st, out, dx, stats = eng.msckf_update(case.frame, empty, opts)
    assert st == 0 and not dx.any() and np.array_equal(eng.cov_get(), case.P)
    # single-measurement features are dropped (UpdaterMSCKF.cpp:88); an all-rejected batch leaves P untouched
    one = capi.FeatArrays([0, 1, 2], case.feats.cam[:2], case.feats.clone[:2], case.feats.uv[:2], case.feats.uvn[:2])
    st, out, dx, stats = eng.msckf_update(case.frame, one, opts)
    assert st == 0 and (out.status == capi.FEAT_FEW_MEAS).all() and not dx.any()
    assert np.array_equal(eng.cov_get(), case.P)
    # fewer rows than columns: no compression in the reference, same update here
    few = case.feats.subset([0, 1])
    ref = oracle.msckf_update(case.frame, few, opts, case.P)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.msckf_update(case.frame, few, opts)
    assert np.array_equal(out.status, ref["out"].status)
    assert np.linalg.norm(eng.cov_get() - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * max(np.linalg.norm(ref["dx"]), 1e-300)


def test_config2_size_properties(eng, oracle):
    """BASELINE config 2 at full size (stereo, 21 clone poses, 400 features): size-independent properties + oracle parity."""
    case = sim.make_update_case(n_feats=400, n_clones=21, n_cams=2, seed=42, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True)
    opts = _opts(case)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.msckf_update(case.frame, case.feats, opts)
    P = eng.cov_get()
    assert st == 0
    assert np.array_equal(P, P.T)
    assert np.all(np.diag(P) <= np.diag(case.P) * (1 + 1e-12)) and np.linalg.eigvalsh(P).min() > -1e-12 * np.abs(P).max()
    assert np.isfinite(dx).all() and stats.n_feats_used > 250
    # idempotence of the gate: re-running on the same prior gives the same decisions and the same answer bit for bit
    eng.cov_set(case.P)
    st2, out2, dx2, _ = eng.msckf_update(case.frame, case.feats, opts)
    assert np.array_equal(out.status, out2.status) and np.array_equal(dx, dx2) and np.array_equal(P, eng.cov_get())
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P, dumps=False)
    assert np.array_equal(out.status, ref["out"].status)
    assert np.linalg.norm(P - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])


def test_sharded_update_matches_single(oracle):
    """The N>1 data path on one GPU: two engine contexts stand in for two ranks; their compressed blocks are stacked
    (what the all-gather produces) and each finishes the update. Result = the single-context update = the oracle."""
    import torch
    from open_vins_b200 import multigpu
    case = sim.make_update_case(n_feats=120, n_clones=21, n_cams=2, seed=3, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True)
    opts = _opts(case, col_order=capi.COLS_CANONICAL)
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P, dumps=False)
    world = 2
    dev = torch.device("cuda", 0)
    engs = [capi.Engine(max_state=256, max_feats=512, max_meas=512 * 48) for _ in range(world)]
    parts = multigpu.partition_features(case.feats.meas_off, world)
    blocks, shards = [], []
    cap = 512 * 520
    for r, e in enumerate(engs):
        e.cov_set(case.P)
        shard = case.feats.subset(np.arange(*parts[r]))
        shards.append(shard)
        buf = torch.zeros(cap, dtype=torch.float64, device=dev)
        n, ld = e.shard_compress(case.frame, shard, opts, buf.data_ptr(), cap)
        torch.cuda.synchronize()
        blocks.append(buf[: n * ld].clone())
    status = []
    for r, e in enumerate(engs):
        stacked = torch.cat(blocks).contiguous()
        st, out, dx, stats = e.shard_finish(stacked.data_ptr(), world, shards[r].n_feats)
        assert st == 0
        status.append(out.status)
        P = e.cov_get()
        assert np.linalg.norm(P - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"])
        assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"])
        if r == 0:
            P0, dx0 = P, dx
        else:
            assert np.array_equal(P, P0) and np.array_equal(dx, dx0)  # replicas stay bitwise identical
    assert np.array_equal(np.concatenate(status), ref["out"].status)
    for e in engs:
        e.close()
