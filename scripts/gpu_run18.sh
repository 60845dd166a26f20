mkdir -p gpurun_out
PYROPE_COARSE_DEBUG=1 timeout 300 python scripts/bench_coarse.py 2>&1 | grep -E "coarse" | tail -4
PYROPE_COARSE_TF32_COPY=1 PYROPE_COARSE_DEBUG=1 timeout 300 python scripts/bench_coarse.py 2>&1 | grep -E "coarse" | tail -4
timeout 300 python scripts/bench_coarse.py 2>&1 | grep "coarse probe"
timeout 300 python scripts/bench_coarse.py 65536 1250 2>&1 | grep "coarse probe"
timeout 600 python -m pytest tests/test_gpu_coarse.py tests/test_gpu_fullshape.py -x -q 2>&1 | tail -2
timeout 900 python bench.py --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v18.json 2> gpurun_out/r2_bench_v18.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v18.json'))
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['gpu_launches'])
PY
