cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_slam.py -x -q 2>&1 | tail -5
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 210 -c 210 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_ncu_list.log 2>&1
tail -3 gpurun_out/launches_r02.csv
# full-set capture of the step's kernels (one step's worth after warm-up)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_' -s 105 -c 21 \
    -o gpurun_out/prof_r02 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02.ncu-rep
