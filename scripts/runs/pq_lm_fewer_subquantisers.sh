mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lm.py -x -q 2>&1 | tail -2
for W in c5m8 c5m4; do
timeout 900 python bench.py --workload $W --scale 0.1 --secondary none --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v24_${W}_scale01.json 2> gpurun_out/r2_bench_v24_${W}.err; echo "bench rc=$?"
PYROPE_PQ_LM=0 timeout 900 python bench.py --workload $W --scale 0.1 --secondary none --steps 3 --warmup 3 --recall-queries 0 --no-cpu > gpurun_out/r2_bench_v24_${W}_scale01_query_major.json 2> gpurun_out/r2_bench_v24_${W}_qm.err; echo "bench rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v24_*.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatch'), d['roofline'].get('kernel'), d['roofline'].get('kernel_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
