cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sim.py -x -q 2>&1 | tail -12
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -c 300 gpurun_out/bench_r02_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_n1.json'))
print(d['value'], d['e2e']['value'], d['stage_ms'], d['gpu_launches_per_step'], d.get('cpu_baseline',{}).get('value'))
for k in d['kernels_us']: print(k)
print(d['roofline']['frac'], [ (r['kernel'][:20], round(r['frac'],3)) for r in d['rooflines']])
PY
timeout 200 python bench.py --config 1 --steps 100 > gpurun_out/bench_r02_c1.json 2>gpurun_out/c1.err; tail -c 600 gpurun_out/bench_r02_c1.json | head -c 300; echo
timeout 300 python bench.py --config 3 --steps 30 > gpurun_out/bench_r02_c3.json 2>gpurun_out/c3.err; tail -c 300 gpurun_out/c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c3.json')); print('c3', d['value'], d['e2e']['value'], d['stage_ms'])"
timeout 300 python bench.py --config 5 --steps 20 > gpurun_out/bench_r02_c5.json 2>gpurun_out/c5.err; tail -c 300 gpurun_out/c5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c5.json')); print('c5', d['value'], d['kernels_us'])"
