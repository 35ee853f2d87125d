cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02_n2.json 2> gpurun_out/n2.err; echo rc=$?; tail -c 1500 gpurun_out/n2.err | tail -15
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_n2.json'))
print(d['value'], d['e2e']['value'], d['n_gpus'], d['config']['multi_gpu'][:60])
print(d.get('sweep_4096'))
PY
