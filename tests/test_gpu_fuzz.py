"""Seeded random sweep over small configurations of the MSCKF and SLAM updates (camera count, window size, batch size,
representation, calibration flags, camera model, column order, 1-D triangulation, gate multiplier): engine vs oracle at the
parity bars. Complements the hand-picked cases of test_gpu_parity.py / test_gpu_slam.py."""
import numpy as np
import pytest

from open_vins_b200 import capi, sim

pytestmark = pytest.mark.gpu


def _draw(seed):
    rng = np.random.default_rng(1000 + seed)
    n_cams = int(rng.integers(1, 4))
    calib_ext, calib_intr = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    return dict(
        rng=rng, n_cams=n_cams, n_clones=int(rng.integers(3, 13)), n_feats=int(rng.integers(1, 61)), rep=int(rng.integers(0, 6)),
        calib_ext=calib_ext, calib_intr=calib_intr, cam_model=int(rng.integers(0, 2)), order=int(rng.integers(0, 2)),
        tri1d=int(rng.integers(0, 4) == 0), mult=float(rng.choice([0.5, 1.0, 5.0])), fej=int(rng.integers(0, 4) != 0),
        mono_frac=float(rng.choice([0.0, 0.5])))


@pytest.mark.parametrize("seed", range(24))
def test_msckf_update_random_configuration(oracle, seed):
    c = _draw(seed)
    case = sim.make_update_case(n_feats=c["n_feats"], n_clones=c["n_clones"], n_cams=c["n_cams"], seed=5000 + seed, calib_ext=c["calib_ext"],
                                calib_intr=c["calib_intr"], calib_dt=bool(seed % 2), cam_model=c["cam_model"], mono_frac=c["mono_frac"],
                                min_track=min(5, c["n_clones"]), cam_order="descending" if seed % 3 else "ascending")
    opts = capi.default_opts(do_calib_camera_pose=int(c["calib_ext"]), do_calib_camera_intrinsics=int(c["calib_intr"]), feat_rep=c["rep"],
                             col_order=c["order"], triangulate_1d=c["tri1d"], chi2_multipler=c["mult"], do_fej=c["fej"])
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P, dumps=False)
    eng = capi.Engine(max_state=256, max_feats=128, max_meas=128 * 48)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.msckf_update(case.frame, case.feats, opts)
    assert st == ref["status"], (st, ref["status"], c)
    assert np.array_equal(out.status, ref["out"].status), c
    ok = ~np.isnan(ref["out"].p_FinG[:, 0])
    if ok.any():
        assert (np.linalg.norm(out.p_FinG[ok] - ref["out"].p_FinG[ok], axis=1) <= 1e-12 * np.linalg.norm(ref["out"].p_FinG[ok], axis=1)).all()
    assert stats.n_feats_used == ref["stats"].n_feats_used and stats.rows_stacked == ref["stats"].rows_stacked
    if st == 0:
        assert np.linalg.norm(eng.cov_get() - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"]), c
        assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * max(np.linalg.norm(ref["dx"]), 1e-300), c
    eng.close()


@pytest.mark.parametrize("seed", range(12))
def test_slam_update_random_configuration(oracle, seed):
    c = _draw(100 + seed)
    n_clones = max(c["n_clones"], 4)
    case = sim.make_slam_case(n_landmarks=max(2, c["n_feats"] // 3), n_clones=n_clones, n_cams=c["n_cams"], seed=7000 + seed, rep=c["rep"],
                              calib_ext=c["calib_ext"], calib_intr=c["calib_intr"], track_len=(1, min(4, n_clones)), two_classes=bool(seed % 2))
    opts = capi.default_opts(do_calib_camera_pose=int(c["calib_ext"]), do_calib_camera_intrinsics=int(c["calib_intr"]), feat_rep=c["rep"],
                             col_order=c["order"], chi2_multipler=c["mult"], do_fej=c["fej"])
    ref = oracle.slam_update(case.frame, case.feats, case.landmarks, opts, case.P)
    eng = capi.Engine(max_state=256, max_feats=128, max_meas=128 * 48)
    eng.cov_set(case.P)
    st, out, dx, stats = eng.slam_update(case.frame, case.feats, case.landmarks, opts)
    assert st == ref["status"] == 0, c
    assert np.array_equal(out.status, ref["out"].status), c
    assert np.linalg.norm(eng.cov_get() - ref["P"]) <= 1e-9 * np.linalg.norm(ref["P"]), c
    assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * max(np.linalg.norm(ref["dx"]), 1e-300), c
    eng.close()
