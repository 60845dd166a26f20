mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_vector_index.py -x -q 2>&1 | tail -12
timeout 900 python bench.py --workload c4 --secondary none --steps 3 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v30_c4.json 2> gpurun_out/r2_bench_v30_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v30_c4.json'))
print('C4', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'], d['roofline'])
PY
PYROPE_FLAT_TF32=1 timeout 900 python bench.py --workload c4 --secondary none --steps 3 --warmup 3 --recall-queries 0 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4 tf32 4-stage', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms'])"
tail -3 gpurun_out/r2_bench_v30_c4.err
