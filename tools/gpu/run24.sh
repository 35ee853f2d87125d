cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cp open_vins_b200/libovb200.so /tmp/lib_keep.so
for v in nb3 nb4; do
cp open_vins_b200/libovb200_$v.so open_vins_b200/libovb200.so
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/b_$v.json 2>gpurun_out/b_$v.err; tail -c 300 gpurun_out/b_$v.err; python -c "
import json; d=json.load(open('gpurun_out/b_$v.json')); print('$v', d['value'], d['e2e']['value'], d['stage_ms'])"
done
cp /tmp/lib_keep.so open_vins_b200/libovb200.so
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/b24.json 2>gpurun_out/b24.err; tail -c 300 gpurun_out/b24.err; python -c "
import json; d=json.load(open('gpurun_out/b24.json')); print('cur', d['value'], d['e2e']['value'], d['stage_ms'])"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
