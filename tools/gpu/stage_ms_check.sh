cd $GRAFT_REPO_ROOT
timeout 100 python - <<'PY'
import numpy as np
from open_vins_b200 import capi, simrun
fr, fe, op, P = simrun.load_case(simrun.CASE_CONFIG2)
eng = capi.Engine(max_state=256, max_feats=1024, max_meas=65536)
for _ in range(3):
    eng.cov_set(P); st, out, dx, stats = eng.msckf_update(fr, fe, op)
ms = eng.last_stage_ms()
print('stage ms', [round(float(x), 4) for x in ms], 'sum5', round(float(np.sum(ms[:5])), 4), 'total', round(float(stats.ms_total), 4))
ms2 = eng.last_stage_ms()
assert np.array_equal(ms, ms2) and 0.3 < float(np.sum(ms[:5])) < 1.0 and abs(float(np.sum(ms[:5])) - float(ms[5])) < 0.2
print('lazy stage times ok')
PY
