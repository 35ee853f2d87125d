# refresh of the round-2 artefacts on the final build (everything except the default bench line and the parity suite, which tools/gpu/suite_and_bench.sh produced)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_r02.txt
timeout 400 python bench.py --impl reference --steps 60 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; tail -c 300 gpurun_out/bench_r02_reference.err
for c in 1 3 4 5; do
  timeout 500 python bench.py --config $c --steps 100 > gpurun_out/bench_r02_c${c}_n1.json 2> gpurun_out/bench_r02_c${c}_n1.err; tail -c 300 gpurun_out/bench_r02_c${c}_n1.err
done
python - <<'PY'
import json
for c in (1,3,4,5):
    try:
        d=json.load(open(f'gpurun_out/bench_r02_c{c}_n1.json'))
        print(c, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('stage_ms'))
    except Exception as e: print(c, 'ERR', e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 210 -c 210 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_' -s 105 -c 21 \
    -o gpurun_out/prof_r02_final python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02_final.ncu-rep
( timeout 60 tools/ubench/fp64_rate; timeout 60 tools/ubench/diag8_bench; timeout 120 tools/ubench/cholqr_bench_np ) > gpurun_out/ubench_r02.txt 2>&1; tail -2 gpurun_out/ubench_r02.txt
du -sh gpurun_out
