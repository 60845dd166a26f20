timeout 900 python bench.py --workload c4 --secondary none --steps 3 --warmup 3 --recall-queries 0 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4 prefetch', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms'])"
PYROPE_FLAT_NOPREFETCH=1 timeout 900 python bench.py --workload c4 --secondary none --steps 3 --warmup 3 --recall-queries 0 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4 no prefetch', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms'])"
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
