#!/usr/bin/env python
"""Full-size oracle fixtures for BASELINE.json configs 3 and 4 (tests/golden/full_*.npz).

The oracle (CPU restatement of the reference's Eigen arithmetic) needs minutes at these sizes — the column-major Givens
compression of a 230k x 155 or 196k x 243 system — so it is run ONCE here and its outputs are committed; the GPU tests
(tests/test_gpu_fullsize.py) regenerate the inputs from the seeded generators of open_vins_b200/sim.py (a hash of the
inputs is stored to detect a generator or numpy change) and compare the engine against these numbers.
  full_config3_f4096   : synthetic 4096-feature stereo batch, 21 clone poses, full calibration, N = 194
  full_config4_msckf800: 4 cameras, 31 clone poses, 800 MSCKF features, full calibration (N = 282), 242 stacked columns
  full_config4_slam100 : the same window with 100 SLAM landmarks in the state (N = 542 + ...), UpdaterSLAM::update in
                         4 sequential batches of 25 (max_slam_in_update), each ONE EKFUpdate like the reference
Outputs: per-feature status / chi2 / p_FinG, dx, and the posterior covariance (upper triangle, float64).
Regenerate:  python tests/golden/make_fullsize.py [config3|msckf800|slam100 ...]
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from open_vins_b200 import capi, sim  # noqa: E402
from oracle import ovo_py as oracle  # noqa: E402


def input_hash(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def triu(P):
    return P[np.triu_indices(P.shape[0])]


CONFIG3 = dict(n_feats=4096, n_clones=21, n_cams=2, seed=0, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True)
CONFIG4 = dict(n_feats=800, n_clones=31, n_cams=4, seed=0, calib_ext=True, calib_intr=True, calib_imu=True, calib_dt=True)
SLAM4 = dict(n_landmarks=100, n_clones=31, n_cams=4, seed=4, rep=capi.REP_GLOBAL_3D)
OPTS = dict(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, col_order=capi.COLS_CANONICAL)


def msckf(name, kw):
    case = sim.make_update_case(**kw)
    opts = capi.default_opts(**OPTS)
    t = time.time()
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P, dumps=False)
    print(name, "oracle %.1f s" % (time.time() - t), "used", ref["stats"].n_feats_used, "rows", ref["stats"].rows_stacked, "cols", ref["stats"].cols_stacked)
    assert ref["status"] == 0
    out = ref["out"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), input_hash=input_hash(case.P, case.feats.uv, case.feats.uvn, case.frame.clone_R),
                        status=out.status.astype(np.int8), chi2=out.chi2, p_FinG=out.p_FinG, dx=ref["dx"], P_triu=triu(ref["P"]),
                        n_used=ref["stats"].n_feats_used, rows=ref["stats"].rows_stacked, cols=ref["stats"].cols_stacked, stage_s=ref["times"])


def slam_batches(sl, batch=25):
    """feature index ranges of the sequential UpdaterSLAM::update calls (core/VioManager.cpp:533-544)"""
    F = sl.feats.n_feats
    return [(a, min(a + batch, F)) for a in range(0, F, batch)]


def slam_subset(sl, a, b):
    from open_vins_b200.capi import LandmarkArrays
    idx = np.arange(a, b)
    lm = sl.landmarks
    sub = LandmarkArrays(lm.lm_off[idx], lm.value[idx], lm.value_fej[idx], lm.anchor_cam[idx], lm.anchor_clone[idx],
                         None if lm.sigma_pix is None else lm.sigma_pix[idx], None if lm.chi2_multipler is None else lm.chi2_multipler[idx])
    return sl.feats.subset(idx), sub


def slam(name, kw):
    sl = sim.make_slam_case(**kw)
    opts = capi.default_opts(feat_rep=kw["rep"], **OPTS)
    P = sl.P.copy()
    dxs, status, chi2 = [], [], []
    t = time.time()
    for a, b in slam_batches(sl):
        feats, lms = slam_subset(sl, a, b)
        ref = oracle.slam_update(sl.frame, feats, lms, opts, P)
        assert ref["status"] == 0
        P = ref["P"]
        dxs.append(ref["dx"])
        status.append(ref["out"].status)
        chi2.append(ref["out"].chi2)
    print(name, "oracle %.1f s" % (time.time() - t), "N", P.shape[0], "accepted", int((np.concatenate(status) == 0).sum()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), input_hash=input_hash(sl.P, sl.feats.uv, sl.feats.uvn, sl.landmarks.value),
                        status=np.concatenate(status).astype(np.int8), chi2=np.concatenate(chi2), dx=np.array(dxs), P_triu=triu(P))


if __name__ == "__main__":
    which = sys.argv[1:] or ["config3", "msckf800", "slam100"]
    if "msckf800" in which:
        msckf("full_config4_msckf800", CONFIG4)
    if "slam100" in which:
        slam("full_config4_slam100", SLAM4)
    if "config3" in which:
        msckf("full_config3_f4096", CONFIG3)
