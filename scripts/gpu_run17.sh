mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_coarse.py tests/test_gpu_fullshape.py -x -q 2>&1 | tail -2
timeout 300 python scripts/bench_coarse.py 2>&1 | grep "coarse probe"
timeout 300 python scripts/bench_coarse.py 65536 1250 2>&1 | grep "coarse probe"
timeout 900 python bench.py --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v17.json 2> gpurun_out/r2_bench_v17.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v17.json'))
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['gpu_launches'])
PY
CMD="python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 --profile-step"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2_launches_v17_c5_step.csv $CMD > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
grep -E "coarse_tc|coarse_rank|coarse_tau" gpurun_out/r2_launches_v17_c5_step.csv | cut -d, -f5,12- | head -12
