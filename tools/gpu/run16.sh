cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cp open_vins_b200/libovb200.so /tmp/lib_keep.so
cp open_vins_b200/libovb200_probe.so open_vins_b200/libovb200.so
timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>&1 | grep "^feat" | tail -9
cp /tmp/lib_keep.so open_vins_b200/libovb200.so
