cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r02.txt
timeout 400 python bench.py > gpurun_out/bench_r02_c2_n1.json 2> gpurun_out/bench_r02_c2_n1.err; tail -c 400 gpurun_out/bench_r02_c2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c2_n1.json')); print(d['value'], d['e2e']['value'], d['stage_ms'], [r.get('traffic') for r in d['rooflines']], d['roofline'].get('traffic'))"
