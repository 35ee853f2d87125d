#!/usr/bin/env python
"""Regenerates the captured rpng_sim update cases of tests/golden/ (BASELINE.json configs 1 and 2).

Each case is ONE MSCKF update's marshalled inputs (frame, feature batch, options, prior covariance) dumped by the host
runner (tools/run_simulation.cpp --capture) in the middle of a simulation on the rpng_sim trajectory head
(tests/golden/traj_tum_corridor1_head.bin, seeds 0, rpng_sim calibration/noise, full online calibration). The runs use
the CPU oracle as backend so that the fixture does not depend on any GPU arithmetic; bench.py and the tests feed the
very same bytes to the engine and to the oracle.
  config 1: mono,   max_clones 11, max_msckf_in_update  50, num_pts  400, update of frame 25
  config 2: stereo, max_clones 20, max_msckf_in_update 400, num_pts 6000, update of frame 38 (window full: 21 clone poses, N = 194)
"""
import gzip
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from open_vins_b200 import build as b  # noqa: E402

from oracle import ovo_py  # noqa: E402
exe = ovo_py.build_sim_runner()
traj = os.path.join(ROOT, "tests", "golden", "traj_tum_corridor1_head.bin")
for name, args, frame in [("rpng_sim_mono11_f50", ["--cams", "1", "--clones", "11", "--msckf", "50", "--pts", "400", "--frames", "30"], 25),
                          ("rpng_sim_stereo20_f400", ["--cams", "2", "--clones", "20", "--msckf", "400", "--pts", "6000", "--frames", "40"], 38)]:
    prefix = os.path.join("/tmp", name)
    subprocess.check_call([exe, "--traj", traj, "--capture", str(frame), prefix] + args)
    with open(prefix + ".case", "rb") as f, gzip.GzipFile(os.path.join(ROOT, "tests", "golden", name + ".case.gz"), "wb", mtime=0) as g:
        g.write(f.read())
    print("wrote", name)
