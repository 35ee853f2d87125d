# bench.py under torchrun on N GPUs of one box (N = $1): config 2 (replicated below the row threshold) + the 4096-feature sharded point
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline \
    > gpurun_out/bench_r02_c2_n$N.json 2> gpurun_out/bench_r02_c2_n$N.err
tail -c 600 gpurun_out/bench_r02_c2_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02_c2_n$N.json'))
print('N=$N', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['config'].get('multi_gpu','')[:60])
print(json.dumps(d.get('sweep_4096'), indent=0)[:900])
PY
