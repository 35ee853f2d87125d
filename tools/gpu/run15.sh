cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_slam.py tests/test_gpu_configs.py tests/test_gpu_cholqr.py -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/b15.json 2>gpurun_out/b15.err; tail -c 300 gpurun_out/b15.err; python -c "
import json; d=json.load(open('gpurun_out/b15.json')); print(d['value'], d['e2e']['value'], d['stage_ms']); print([(k['kernel'],round(k['us_per_step'],1)) for k in d['kernels_us']])"
