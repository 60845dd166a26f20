mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_fullshape.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_tests_v8a.log 2>&1; echo "lm tests rc=$?"; tail -4 gpurun_out/r2_tests_v6a.log
PYROPE_LM_STAGES=1 timeout 600 python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 > gpurun_out/r2_stages_v8.json 2> gpurun_out/r2_stages_v8.err; tail -2 gpurun_out/r2_stages_v8.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v8.json 2> gpurun_out/r2_bench_v8.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_v8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v8.json'))
s=d.pop('secondary')
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'])
print('C4', s['value'], s['ms_per_step'], s['parity']['mismatch'], s['parity'].get('scores_bit_exact_vs_oracle'), s['roofline']['frac'], s['roofline']['kernel_ms'])
PY
