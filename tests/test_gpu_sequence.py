"""Sequence parity: the per-frame covariance pipeline of VioManager::do_feature_propagate_update
(ov_msckf/src/core/VioManager.cpp:341-596) — EKFPropagation, clone (+ time-offset term), UpdaterMSCKF::update,
marginalize of the oldest clone — run for many frames with P resident on the GPU, against the oracle in lock step.
The rpng_sim simulator is not restated yet, so poses/tracks of every frame come from the synthetic
generator; what is carried from frame to frame is the covariance, exactly the quantity that never leaves the device."""
import numpy as np
import pytest

from open_vins_b200 import capi, sim

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chi2_mult,min_used", [(1.0, 20), (1e12, 400)])
def test_twenty_frames_lockstep(oracle, chi2_mult, min_used):
    """chi2_mult = 1: the reference's gate (few tracks survive: every frame's synthetic poses are drawn around a fresh
    truth while P keeps shrinking). chi2_mult = 1e12: gate open, nearly every track enters the update."""
    mode = capi.COMPRESS_HOUSEHOLDER_TSQR
    C = 8               # max_clones: window holds C+1 poses during the update
    n_cams = 2
    lay = sim.StateLayout(n_cams, C + 1, calib_ext=True, calib_intr=True, calib_imu=False, calib_dt=True)
    N_full = lay.N
    rng = np.random.default_rng(123)
    # prior before the first frame: window of C clones (the newest clone slot is appended by cov_clone each frame)
    case0 = sim.make_update_case(n_feats=4, n_clones=C + 1, n_cams=n_cams, seed=1000, calib_ext=True, calib_intr=True, calib_dt=True)
    assert case0.layout.N == N_full
    P = case0.P[: N_full - 6, : N_full - 6].copy()
    eng = capi.Engine(max_state=256, max_feats=256, max_meas=256 * 24)
    eng.cov_set(P)
    opts = capi.default_opts(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, compress=mode, col_order=capi.COLS_CANONICAL,
                             chi2_multipler=chi2_mult)
    used_total = 0
    for frame in range(20):
        # ---- propagate the IMU block (Propagator::propagate_and_clone -> StateHelper::EKFPropagation, p = 15)
        Phi = np.eye(15) + 0.02 * rng.standard_normal((15, 15))
        Qh = rng.standard_normal((15, 15))
        Q = Qh @ Qh.T * 1e-8
        st_r, P = oracle.cov_propagate(P, 0, Phi, Q, [0], [15])
        st_g = eng.cov_propagate(0, Phi, Q, [0], [15])
        assert st_r == st_g == 0
        # ---- clone the IMU pose with the camera time-offset Jacobian (StateHelper::augment_clone)
        dnc = np.concatenate([0.3 * rng.standard_normal(3), rng.standard_normal(3)])
        P = oracle.cov_clone(P, 0, 6, dnc, lay.dt_off)
        eng.cov_clone(0, 6, dnc, lay.dt_off)
        assert P.shape[0] == N_full == eng.cov_dim()
        # ---- MSCKF update on this frame's tracks (fresh poses; the carried quantity is P)
        case = sim.make_update_case(n_feats=60, n_clones=C + 1, n_cams=n_cams, seed=2000 + frame, calib_ext=True, calib_intr=True,
                                    calib_dt=True, t0=3.0 + 0.1 * frame)
        ref = oracle.msckf_update(case.frame, case.feats, opts, P, dumps=False)
        st, out, dx, stats = eng.msckf_update(case.frame, case.feats, opts)
        assert st == ref["status"] == 0, frame
        assert np.array_equal(out.status, ref["out"].status), f"gate decisions differ in frame {frame}"
        P = ref["P"]
        Pg = eng.cov_get()
        assert np.linalg.norm(Pg - P) <= 1e-9 * np.linalg.norm(P), (frame, np.linalg.norm(Pg - P) / np.linalg.norm(P))
        assert np.linalg.norm(dx - ref["dx"]) <= 1e-9 * np.linalg.norm(ref["dx"]), frame
        used_total += stats.n_feats_used
        # ---- marginalize the oldest clone (StateHelper::marginalize_old_clone)
        P = oracle.cov_marginalize(P, lay.clone_off[0], 6)
        eng.cov_marginalize(lay.clone_off[0], 6)
        # keep the oracle on ITS OWN trajectory (no re-sync): the device copy must track it for all 20 frames
    Pg = eng.cov_get()
    assert np.linalg.norm(Pg - P) <= 1e-9 * np.linalg.norm(P)
    assert used_total >= min_used, used_total
    assert np.linalg.eigvalsh(Pg).min() > -1e-12 * np.abs(Pg).max()
    eng.close()
