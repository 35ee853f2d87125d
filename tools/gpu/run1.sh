set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_cholqr.py -x -q 2>&1 | tail -30
python -m pytest tests/test_gpu_gram.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python bench.py --steps 50 --warmup 5 --compress cholqr2 --no-cpu-baseline > gpurun_out/bench_cholqr2.json 2> gpurun_out/bench_cholqr2.err; tail -c 1500 gpurun_out/bench_cholqr2.json
python bench.py --steps 50 --warmup 5 --compress tsqr --no-cpu-baseline > gpurun_out/bench_tsqr.json 2>/dev/null; tail -c 600 gpurun_out/bench_tsqr.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cq.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -60 gpurun_out/launches_cq.csv | cut -d, -f5,12- 
