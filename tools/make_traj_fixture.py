#!/usr/bin/env python
"""Writes tests/golden/traj_tum_corridor1_head.bin: the first N poses of the rpng_sim trajectory
(ov_data/sim/tum_corridor1_512_16_okvis.txt of the reference, config/rpng_sim/estimator_config.yaml:124) as raw
little-endian float64 rows [t x y z qx qy qz qw]. Input fixture only (the simulator of include/ovb200_sim.hpp builds its
B-spline from it); the reference tree does not exist on the GPU box, hence the committed copy of the numbers.
Usage: python tools/make_traj_fixture.py [/root/reference] [rows]"""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
src = os.path.join(ref, "ov_data", "sim", "tum_corridor1_512_16_okvis.txt")
a = np.loadtxt(src, comments="#")[:rows]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "traj_tum_corridor1_head.bin")
a.astype("<f8").tofile(out)
print(out, a.shape, f"{a[-1, 0] - a[0, 0]:.1f} s")
