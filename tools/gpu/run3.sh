cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
tools/ubench/cholqr_bench 2>&1 | tail -14
tools/ubench/cholqr_bench_np 2>&1 | tail -3
python -m pytest tests/test_gpu_cholqr.py -x -q 2>&1 | tail -8
python bench.py --steps 50 --warmup 5 --compress cholqr2 --no-cpu-baseline > gpurun_out/bench_cholqr2.json 2> gpurun_out/bench_cholqr2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cholqr2.json'))
print(d['value'], d['e2e']['value'], d['stage_ms'], d['gpu_launches_per_step'])
PY
ncu --set full --import-source on --clock-control none -k regex:"k_cq_gram|k_cq_trsm" -c 2 -f -o gpurun_out/cq_r02a tools/ubench/cholqr_bench_np > /dev/null 2>&1
ls -la gpurun_out/
