cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_r02_c2_n1.json 2> gpurun_out/bench_r02_c2_n1.err; tail -c 300 gpurun_out/bench_r02_c2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c2_n1.json')); print(d['value'], d['e2e']['value'], d['stage_ms'], d['e2e']['host_us_inside_call'], d['cpu_baseline']['value'])"
