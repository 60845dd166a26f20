mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2_tests_sharded_n2.log 2>&1; echo "sharded tests rc=$?"; tail -12 gpurun_out/r2_tests_sharded_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_v10.json 2> gpurun_out/r2_bench_n2_v10.err; echo "bench n2 rc=$?"; tail -c 600 gpurun_out/r2_bench_n2_v10.err
timeout 900 python bench.py --sharded-entry --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_sharded_entry_n2.json 2> gpurun_out/r2_bench_sharded_entry_n2.err; echo "sharded entry rc=$?"; tail -c 400 gpurun_out/r2_bench_sharded_entry_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n2_v10.json','gpurun_out/r2_bench_sharded_entry_n2.json'):
    try:
        d=json.load(open(f)); s=d.pop('secondary',None)
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d.get('roofline',{}).get('stage_ms'))
        if s: print('  C4', s['value'], s['ms_per_step'], s.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
