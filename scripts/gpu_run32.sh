mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8
timeout 900 python bench.py --workload c4 --secondary none --steps 3 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v32_c4.json 2> gpurun_out/r2_bench_v32_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v32_c4.json'))
print('C4', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['parity']['scores_bit_exact_vs_oracle'], d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['frac'])
PY
tail -3 gpurun_out/r2_bench_v32_c4.err
