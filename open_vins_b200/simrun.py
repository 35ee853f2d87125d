"""Python side of the rpng_sim runner (tools/run_simulation.cpp over include/ovb200_vio.hpp): launch the product executable
open_vins_b200/ovb_run_simulation (CUDA engine) and read what it writes — the JSON summary, the estimate/ground-truth
trajectory file, the timing CSV (columns of ov_msckf/src/core/VioManager.cpp:117-121) and captured update cases."""
from __future__ import annotations

import ctypes as C
import json
import os
import re
import subprocess

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_EXE = os.path.join(HERE, "ovb_run_simulation")
TRAJ_FIXTURE = os.path.join(os.path.dirname(HERE), "tests", "golden", "traj_tum_corridor1_head.bin")
# update cases captured from rpng_sim runs (tests/golden/make_rpng_sim_cases.py): BASELINE.json configs 1 and 2
CASE_CONFIG1 = os.path.join(os.path.dirname(HERE), "tests", "golden", "rpng_sim_mono11_f50.case.gz")
CASE_CONFIG2 = os.path.join(os.path.dirname(HERE), "tests", "golden", "rpng_sim_stereo20_f400.case.gz")


def run(exe=None, traj=None, cams=2, clones=11, msckf=10, pts=250, frames=0, calib=1, est=None, timing=None, capture=None, integration="rk4",
        compress="cholqr2", timeout=1800):
    """Runs the simulation; returns the parsed JSON summary. capture = (frame_index, path_prefix) dumps that update's inputs."""
    cmd = [exe or ENGINE_EXE, "--traj", traj or TRAJ_FIXTURE, "--cams", str(cams), "--clones", str(clones), "--msckf", str(msckf), "--pts", str(pts),
           "--frames", str(frames), "--calib", str(int(calib)), "--integration", integration, "--compress", compress]
    if est:
        cmd += ["--est", est]
    if timing:
        cmd += ["--timing", timing]
    if capture:
        cmd += ["--capture", str(capture[0]), capture[1]]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=timeout)
    return json.loads(out.stdout.strip().splitlines()[-1])


def load_estimate(path):
    """rows: t, est p(3) q(4), gt p(3) q(4)"""
    a = np.loadtxt(path, comments="#")
    return a[:, 0], a[:, 1:4], a[:, 4:8], a[:, 8:11], a[:, 11:15]


def ate_rmse(p_est, p_gt):
    """ov_eval/src/calc/ResultTrajectory.cpp:82-109 (position part, alignment 'none'): sqrt(mean |p_gt - p_est|^2)"""
    return float(np.sqrt(np.mean(np.sum((p_gt - p_est) ** 2, axis=1))))


def load_case(path):
    """One captured MSCKF update (written by the runner's --capture): returns (FrameArrays, FeatArrays, ovb_opts, P)."""
    import gzip
    with (gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")) as f:
        hdr = f.readline().decode()
        assert hdr.startswith("OVBCASE1"), hdr
        kv = {k: int(v) for k, v in re.findall(r"(\w+)=(\d+)", hdr)}
        C_, K, F, M, NK, N = kv["n_clones"], kv["n_cams"], kv["n_feats"], kv["n_meas"], kv["n_keys"], kv["N"]

        def rd(dt, n):
            return np.frombuffer(f.read(np.dtype(dt).itemsize * n), dtype=dt).copy()
        clone_R, clone_p, clone_Rf, clone_pf = rd("<f8", 9 * C_), rd("<f8", 3 * C_), rd("<f8", 9 * C_), rd("<f8", 3 * C_)
        clone_off = rd("<i4", C_)
        cam_R, cam_p, cam_intr = rd("<f8", 9 * K), rd("<f8", 3 * K), rd("<f8", 8 * K)
        cam_model, cam_ext, cam_intr_off = rd("<i4", K), rd("<i4", K), rd("<i4", K)
        meas_off, cam, clone = rd("<i4", F + 1), rd("u1", M), rd("<u2", M)
        uv, uvn = rd("<f4", 2 * M), rd("<f4", 2 * M)
        keys_off, keys = rd("<i4", F + 1), rd("u1", NK)
        opts = capi.ovb_opts.from_buffer_copy(f.read(kv["opts"]))
        P = rd("<f8", N * N).reshape(N, N)
    frame = capi.FrameArrays(clone_R.reshape(C_, 9), clone_p.reshape(C_, 3), clone_Rf.reshape(C_, 9), clone_pf.reshape(C_, 3), clone_off,
                             cam_R.reshape(K, 9), cam_p.reshape(K, 3), cam_intr.reshape(K, 8), cam_model, cam_ext, cam_intr_off)
    feats = capi.FeatArrays(meas_off, cam, clone, uv.reshape(M, 2), uvn.reshape(M, 2), keys_off, keys)
    return frame, feats, opts, P
