mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lm.py tests/test_gpu_fullshape.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
PYROPE_LM_STAGES=1 timeout 600 python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 2>&1 >/dev/null | grep "lm stages" | tail -1
timeout 900 python bench.py --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v26.json 2> gpurun_out/r2_bench_v26.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v26.json'))
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['gpu_launches'])
PY
