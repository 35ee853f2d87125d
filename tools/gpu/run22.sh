cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_slam.py tests/test_gpu_configs.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/b22.json 2>gpurun_out/b22.err; tail -c 300 gpurun_out/b22.err; python -c "
import json; d=json.load(open('gpurun_out/b22.json')); print(d['value'], d['e2e']['value'], d['stage_ms']); print([(k['kernel'],round(k['us_per_step'],1)) for k in d['kernels_us']])"
cp open_vins_b200/libovb200.so /tmp/lib_keep.so
cp open_vins_b200/libovb200_probe.so open_vins_b200/libovb200.so
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/probe23.txt 2>&1
grep "^feat" gpurun_out/probe23.txt | tail -3
cp /tmp/lib_keep.so open_vins_b200/libovb200.so
