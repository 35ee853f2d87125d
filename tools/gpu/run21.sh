cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_feature_system|k_cq_chol' -s 30 -c 6 \
    -o gpurun_out/prof_r02b python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_full2.log 2>&1
ls -la gpurun_out/prof_r02b.ncu-rep
