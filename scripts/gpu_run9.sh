mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_fullshape.py tests/test_gpu_parity.py tests/test_gpu_vector_index.py -x -q > gpurun_out/r2_tests_v9a.log 2>&1; echo "lm tests rc=$?"; tail -3 gpurun_out/r2_tests_v9a.log
PYROPE_LM_STAGES=1 timeout 600 python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 > gpurun_out/r2_stages_v9.json 2> gpurun_out/r2_stages_v9.err; tail -1 gpurun_out/r2_stages_v9.err
CMD="python bench.py --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 --profile-step"
timeout 600 $CMD > gpurun_out/r2_bench_v9.json 2> gpurun_out/r2_bench_v9.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v9.json'))
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'])
PY
CMD2="python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 --profile-step"
timeout 600 $CMD2 > gpurun_out/plain.json 2> gpurun_out/plain.err && timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivfpq_lm_scan_kernel -o gpurun_out/r2_scan_v9 $CMD2 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
