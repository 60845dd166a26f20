mkdir -p gpurun_out
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c5_c4.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5_c4.json'))
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('step_level_frac'), d['gpu_launches'], d['cpu_baseline']['value'])
s=d['secondary']; print('C4', s['value'], s['ms_per_step'], s['e2e']['value'], s['parity']['mismatch'], s['roofline']['kernel'], s['roofline']['kernel_ms'], s['roofline']['frac'], s['clocks'])
PY
CMD="python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 --profile-step"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_c5_step.csv $CMD > gpurun_out/ncu.log 2>&1; echo "ncu list rc=$?"
