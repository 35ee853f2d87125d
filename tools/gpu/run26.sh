cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q 2>&1 | tail -4
for t in np t256 t512 t640; do echo "== $t"; timeout 100 tools/ubench/cholqr_bench_$t 2>&1 | grep "^rep 3"; done
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/b26.json 2>gpurun_out/b26.err; tail -c 300 gpurun_out/b26.err; python -c "
import json; d=json.load(open('gpurun_out/b26.json')); print('cur', d['value'], d['e2e']['value'], d['stage_ms'], d['e2e'].get('host_us_inside_call'))"
