mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v38_c5_c4.json 2> gpurun_out/r2_bench_v38.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v38_c5_c4.json'))
print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('step_level_frac'), d['roofline'].get('traffic'), d['roofline'].get('traffic_note'), d['gpu_launches'])
s=d['secondary']; print('C4', s['value'], s['ms_per_step'], s['e2e']['value'], s['parity']['mismatch'], s['roofline']['kernel'], s['roofline']['kernel_ms'], s['roofline']['frac'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
