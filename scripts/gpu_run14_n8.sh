mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_v14.json 2> gpurun_out/r2_bench_n8_v14.err; echo "bench n8 rc=$?"; tail -c 500 gpurun_out/r2_bench_n8_v14.err
timeout 900 python bench.py --sharded-entry --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_sharded_entry_n8.json 2> gpurun_out/r2_bench_sharded_entry_n8.err; echo "sharded entry rc=$?"; tail -c 300 gpurun_out/r2_bench_sharded_entry_n8.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n8_v14.json','gpurun_out/r2_bench_sharded_entry_n8.json'):
    try:
        d=json.load(open(f)); s=d.pop('secondary',None)
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d.get('roofline',{}).get('stage_ms'), d.get('roofline',{}).get('kernel_ms'))
        if s: print('  C4', s['value'], s['ms_per_step'], s.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
