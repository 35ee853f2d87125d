#!/usr/bin/env python
"""Generate the golden "update case" fixtures of tests/golden/*.npz.

The reference ships no golden vectors for this path and cannot be built in this image (PARITY UNPINNED),
so these fixtures are produced by the restated oracle (oracle/libovoracle.so) on inputs of open_vins_b200/sim.py. They
pin the ORACLE against drift (tests/test_golden_cpu.py) and let the GPU parity tests run against committed numbers
without the oracle in the loop (tests/test_gpu_golden.py). Regenerate with:  python tests/golden/make_golden.py
Each file: inputs (frame, feature SoA, P, options) and outputs (per-feature status / p_FinG / chi2, dx, posterior P,
compressed-system invariants R'R and R'z).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from open_vins_b200 import capi, sim  # noqa: E402
import ovo_py as oracle  # noqa: E402

CASES = {
    # name: (sim kwargs, option kwargs)
    "mono_11_nocalib": (dict(n_feats=24, n_clones=11, n_cams=1, seed=11), dict()),
    "stereo_8_calib": (dict(n_feats=30, n_clones=8, n_cams=2, seed=12, calib_ext=True, calib_intr=True, calib_dt=True),
                       dict(do_calib_camera_pose=1, do_calib_camera_intrinsics=1)),
    "stereo_8_anchored": (dict(n_feats=20, n_clones=8, n_cams=2, seed=13, calib_ext=True, calib_intr=True),
                          dict(do_calib_camera_pose=1, do_calib_camera_intrinsics=1,
                               feat_rep=capi.REP_ANCHORED_MSCKF_INVERSE_DEPTH)),
}
FRAME_FIELDS = ["clone_R", "clone_p", "clone_R_fej", "clone_p_fej", "clone_off", "cam_R", "cam_p", "cam_intr", "cam_model",
                "cam_ext_off", "cam_intr_off"]
FEAT_FIELDS = ["meas_off", "cam", "clone", "uv", "uvn"]


def build(name):
    skw, okw = CASES[name]
    case = sim.make_update_case(**skw)
    opts = capi.default_opts(**okw)
    ref = oracle.msckf_update(case.frame, case.feats, opts, case.P, dumps=True)
    assert ref["status"] == 0
    d = {"P": case.P, "opt_keys": np.array(sorted(okw.keys())), "opt_vals": np.array([okw[k] for k in sorted(okw.keys())], dtype=np.int64)}
    for k in FRAME_FIELDS:
        d["frame_" + k] = getattr(case.frame, k)
    for k in FEAT_FIELDS:
        d["feat_" + k] = getattr(case.feats, k)
    out = ref["out"]
    d.update(out_status=out.status, out_p_FinG=out.p_FinG, out_chi2=out.chi2, dx=ref["dx"], P_post=ref["P"])
    R, z = ref["H_cmp"], ref["res_cmp"]  # compressed system, columns in the reference's first-seen order
    d.update(RtR=R.T @ R, Rtz=R.T @ z, order_off=ref["order_off"], order_sz=ref["order_sz"], n_used=np.int64(ref["stats"].n_feats_used))
    return d


SLAM_CASES = {
    # name: (sim.make_slam_case kwargs, option kwargs)
    "slam_global3d": (dict(n_landmarks=12, n_clones=8, n_cams=2, seed=31, rep=capi.REP_GLOBAL_3D),
                      dict(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, feat_rep=capi.REP_GLOBAL_3D)),
    "slam_single_depth": (dict(n_landmarks=12, n_clones=8, n_cams=2, seed=32, rep=capi.REP_ANCHORED_INVERSE_DEPTH_SINGLE),
                          dict(do_calib_camera_pose=1, do_calib_camera_intrinsics=1, feat_rep=capi.REP_ANCHORED_INVERSE_DEPTH_SINGLE)),
}
LM_FIELDS = ["lm_off", "value", "value_fej", "anchor_cam", "anchor_clone", "sigma_pix", "chi2_multipler"]


def build_slam(name):
    skw, okw = SLAM_CASES[name]
    case = sim.make_slam_case(**skw)
    opts = capi.default_opts(**okw)
    ref = oracle.slam_update(case.frame, case.feats, case.landmarks, opts, case.P)
    assert ref["status"] == 0
    d = {"P": case.P, "opt_keys": np.array(sorted(okw.keys())), "opt_vals": np.array([okw[k] for k in sorted(okw.keys())], dtype=np.int64)}
    for k in FRAME_FIELDS:
        d["frame_" + k] = getattr(case.frame, k)
    for k in FEAT_FIELDS:
        d["feat_" + k] = getattr(case.feats, k)
    for k in LM_FIELDS:
        d["lm_" + k] = getattr(case.landmarks, k)
    d.update(out_status=ref["out"].status, out_chi2=ref["out"].chi2, dx=ref["dx"], P_post=ref["P"], H_big=ref["H_big"], res_big=ref["res_big"],
             Rdiag_big=ref["Rdiag_big"], order_off=ref["order_off"], order_sz=ref["order_sz"], n_used=np.int64(ref["stats"].n_feats_used))
    return d


def main():
    for name in SLAM_CASES:
        d = build_slam(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, os.path.getsize(path), "bytes; used", int(d["n_used"]), "of", len(d["out_status"]))
    for name in CASES:
        d = build(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, os.path.getsize(path), "bytes; used", int(d["n_used"]), "of", len(d["out_status"]))


if __name__ == "__main__":
    main()
