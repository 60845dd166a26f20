timeout 600 python bench.py --workload c5 --secondary none --steps 10 --warmup 3 --no-cpu --recall-queries 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'])"
