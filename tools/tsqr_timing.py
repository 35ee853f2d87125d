"""Debug helper: per-phase cycle counts of k_tsqr_level (needs tools/libovb200_timing.so built with -DOVB_TSQR_TIMING)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_vins_b200 import capi
lib = capi.load_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libovb200_timing.so"))
eng = capi.Engine(max_state=256, max_feats=64, max_meas=16384, lib=lib)
rng = np.random.default_rng(0)
m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 25000, 154
H = rng.standard_normal((m, n)); res = rng.standard_normal(m)
for _ in range(2):
    R, z = eng.compress(H, res)
print("ok", np.linalg.norm(R.T @ R - H.T @ H) / np.linalg.norm(H.T @ H))
