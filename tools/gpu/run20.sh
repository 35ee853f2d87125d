cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cp open_vins_b200/libovb200.so /tmp/lib_keep.so
cp open_vins_b200/libovb200_probe.so open_vins_b200/libovb200.so
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/probe22.txt 2>&1
grep "^feat" gpurun_out/probe22.txt | tail -3; grep "^chol" gpurun_out/probe22.txt | tail -12
cp /tmp/lib_keep.so open_vins_b200/libovb200.so
