"""Closed-loop sanity of the update path's conventions with the oracle backend (tests/mini_filter.py): the estimate must stay
inside its own uncertainty for a whole run — a wrong Jacobian sign, FEJ rule, column map or dx convention makes it drift."""
import numpy as np
import pytest

from tests import mini_filter as mf


@pytest.mark.parametrize("calib", [False, True])
def test_closed_loop_stays_consistent(oracle, calib):
    r = mf.run(lambda P: mf.OracleBackend(oracle, P), n_frames=40, window=8, n_cams=2, feats_per_frame=25, calib=calib, seed=1)
    err = np.linalg.norm(r["p_est"] - r["p_true"], axis=1)
    assert sum(r["used"]) > 300                      # the gate accepts the bulk of the tracks: estimate and P agree
    assert (err < 4.0 * np.array(r["sigma_p"])).all()  # inside the 4-sigma band at every frame
    assert err[-1] < 0.05 and err.max() < 0.06       # metres, after 4 s of motion at ~1 m/s
    P = r["P_final"]
    assert np.allclose(P, P.T, atol=0) and np.linalg.eigvalsh(P).min() > -1e-12 * np.abs(P).max()


def test_closed_loop_detects_a_wrong_correction_sign(oracle):
    """Negative control: applying the correction with the wrong sign must break the consistency the test above relies on."""
    orig = mf._apply_pose
    try:
        mf._apply_pose = lambda R, p, d: orig(R, p, -d)
        r = mf.run(lambda P: mf.OracleBackend(oracle, P), n_frames=40, window=8, n_cams=2, feats_per_frame=25, calib=False, seed=1)
    finally:
        mf._apply_pose = orig
    err = np.linalg.norm(r["p_est"] - r["p_true"], axis=1)
    assert (err > 4.0 * np.array(r["sigma_p"])).any()  # the band the correct run never leaves
