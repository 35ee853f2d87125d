"""Load a tests/golden/*.npz update case (written by tests/golden/make_golden.py) back into capi structures."""
import glob
import os

import numpy as np

from open_vins_b200 import capi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAME_FIELDS = ["clone_R", "clone_p", "clone_R_fej", "clone_p_fej", "clone_off", "cam_R", "cam_p", "cam_intr", "cam_model",
                "cam_ext_off", "cam_intr_off"]
FEAT_FIELDS = ["meas_off", "cam", "clone", "uv", "uvn"]


LM_FIELDS = ["lm_off", "value", "value_fej", "anchor_cam", "anchor_clone", "sigma_pix", "chi2_multipler"]


def _all():
    # full_*.npz are the output-only full-size fixtures of make_fullsize.py (inputs are regenerated): tests/test_gpu_fullsize.py
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))) if not n.startswith("full_"))


def names():
    """MSCKF update cases."""
    return [n for n in _all() if not n.startswith("slam_")]


def slam_names():
    return [n for n in _all() if n.startswith("slam_")]


def load_slam(name):
    d, frame, feats, opts = load(name)
    lms = capi.LandmarkArrays(*[d["lm_" + k] for k in LM_FIELDS])
    return d, frame, feats, lms, opts


def load(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    frame = capi.FrameArrays(*[d["frame_" + k] for k in FRAME_FIELDS])
    feats = capi.FeatArrays(*[d["feat_" + k] for k in FEAT_FIELDS])
    okw = {str(k): int(v) for k, v in zip(d["opt_keys"], d["opt_vals"])}
    return d, frame, feats, capi.default_opts(**okw)
