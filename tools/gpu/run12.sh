cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 60 tools/ubench/cholqr_bench_np 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_cholqr.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -8
for c in 1 0; do OVB_GRAM_CLUSTER=$c timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/b$c.json 2>gpurun_out/b$c.err; tail -c 300 gpurun_out/b$c.err; python -c "
import json; d=json.load(open('gpurun_out/b$c.json')); print('cluster=$c', d['value'], d['e2e']['value'], d['stage_ms']); print([(k['kernel'],round(k['us_per_step'],1)) for k in d['kernels_us']])"; done
