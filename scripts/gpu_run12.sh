mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_v12.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_v12.log
PYROPE_LM_STAGES=1 timeout 600 python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 > gpurun_out/r2_stages_v12.json 2> gpurun_out/r2_stages_v12.err; tail -1 gpurun_out/r2_stages_v12.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v12.json 2> gpurun_out/r2_bench_v12.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v12.json'))
s=d.pop('secondary')
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['gpu_launches'])
print('C4', s['value'], s['ms_per_step'], s['parity']['mismatch'], s['roofline']['frac'])
PY
CMD="python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 --profile-step"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_v12_c5_step.csv $CMD > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
