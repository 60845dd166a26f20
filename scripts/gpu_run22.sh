timeout 900 python -m pytest tests/test_gpu_lm.py tests/test_gpu_fullshape.py -x -q 2>&1 | tail -2
timeout 600 python bench.py --workload c5 --secondary none --steps 10 --warmup 3 --no-cpu --recall-queries 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['parity']['mismatch'] if d.get('parity') else None)"
