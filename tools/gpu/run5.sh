cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 60 tools/ubench/cholqr_bench_np 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_cholqr.py tests/test_gpu_parity.py tests/test_gpu_slam.py -x -q 2>&1 | tail -5
timeout 120 python bench.py --steps 50 --warmup 5 --compress cholqr2 --no-cpu-baseline > gpurun_out/bench_cholqr2.json 2> gpurun_out/bench_cholqr2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cholqr2.json'))
print(d['value'], d['e2e']['value'], d['stage_ms'], d['gpu_launches_per_step'])
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cq.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -22 gpurun_out/launches_cq.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    print(r[4][:30], r[-1])
"
