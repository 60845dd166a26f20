mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coarse.py tests/test_gpu_fullshape.py -x -q 2>&1 | tail -8
PYROPE_COARSE_DEBUG=1 timeout 300 python scripts/bench_coarse.py 2>&1 | grep -E "coarse" | tail -3
PYROPE_COARSE_TF32=1 PYROPE_COARSE_DEBUG=1 timeout 300 python scripts/bench_coarse.py 2>&1 | grep -E "coarse" | tail -3
timeout 900 python bench.py --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v29.json 2> gpurun_out/r2_bench_v29.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v29.json'))
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['gpu_launches'])
PY
