mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_v13.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_v13.log
timeout 300 python scripts/bench_coarse.py 2>&1 | grep "coarse probe"
timeout 300 python scripts/bench_coarse.py 65536 1250 2>&1 | grep "coarse probe"
timeout 900 python bench.py --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v13.json 2> gpurun_out/r2_bench_v13.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v13.json'))
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['gpu_launches'])
PY
