cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
A="--traj tests/golden/traj_tum_corridor1_head.bin --cams 1 --clones 11 --msckf 50 --pts 200 --frames 300"
timeout 200 tests/cpp/run_simulation_oracle $A --est gpurun_out/est_o.txt | cut -c1-330
for c in cholqr2 tsqr; do timeout 200 open_vins_b200/ovb_run_simulation $A --compress $c --est gpurun_out/est_$c.txt | cut -c1-330; done
python - <<'PY'
import numpy as np
o=np.loadtxt('gpurun_out/est_o.txt',comments='#')
for c in ('cholqr2','tsqr'):
    g=np.loadtxt(f'gpurun_out/est_{c}.txt',comments='#')
    d=np.abs(g[:,1:4]-o[:,1:4]).max(axis=1)
    print(c, 'max', d.max(), 'at', [float(d[i]) for i in (10,50,100,150,200,250,299)])
PY
timeout 300 python bench.py --config 4 --steps 20 --no-cpu-baseline > gpurun_out/bench_r02_c4.json 2>gpurun_out/c4.err; tail -c 300 gpurun_out/c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c4.json')); print('c4', d['value'], d['e2e']['value'], d['stage_ms']); print(d['kernels_us'])"
