cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cholqr.py tests/test_gpu_configs.py tests/test_gpu_sim.py -x -q 2>&1 | tail -15
timeout 300 python bench.py --config 5 --steps 20 > gpurun_out/bench_r02_c5.json 2>gpurun_out/c5.err; tail -c 300 gpurun_out/c5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c5.json')); print('c5', d['value'], d['kernels_us'])"
timeout 300 python bench.py --config 4 --steps 20 --no-cpu-baseline > gpurun_out/bench_r02_c4.json 2>gpurun_out/c4.err; tail -c 300 gpurun_out/c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c4.json')); print('c4', d['value'], d['e2e']['value'], d['stage_ms']); print(d['kernels_us'])"
