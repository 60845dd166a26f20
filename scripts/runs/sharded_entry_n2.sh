# two GPUs, ONE process: the pyrope_sharded_* entry (tests + bench line)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_peer_exchange.py -x -q 2>&1 | tail -3
timeout 1200 python bench.py --sharded-entry --gpus 2 --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/bench_sharded_entry_n2.json 2> gpurun_out/bench_sharded_entry_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_sharded_entry_n2.json'))
print('sharded entry N=2', d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('parity') or {}).get('mismatch'))
PY
tail -2 gpurun_out/bench_sharded_entry_n2.err
