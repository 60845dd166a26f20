mkdir -p gpurun_out
./scripts/micro/mma_rate > gpurun_out/r2_mma_rate.txt 2>&1; cat gpurun_out/r2_mma_rate.txt
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
timeout 900 python bench.py --sharded-entry --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_sharded_entry_n2.json 2> gpurun_out/r2_bench_sharded_entry_n2.err; echo "sharded entry rc=$?"; tail -c 400 gpurun_out/r2_bench_sharded_entry_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_sharded_entry_n2.json',):
    try:
        d=json.load(open(f))
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d['details'])
    except Exception as e: print(f, 'ERR', e)
PY
