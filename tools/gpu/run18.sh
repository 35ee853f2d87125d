cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 60 tools/ubench/diag8_bench 2>&1 | tee gpurun_out/diag8_bench.txt
