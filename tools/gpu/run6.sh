cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 ncu --set full --import-source on --clock-control none -k regex:"k_cq_trsm" -c 3 -f -o gpurun_out/cq_r02b tools/ubench/cholqr_bench_np > /dev/null 2>&1
ls -la gpurun_out
