mkdir -p gpurun_out
CMD="python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0"
PYROPE_COARSE_DEBUG=1 timeout 600 $CMD 2>&1 >/dev/null | grep "\[coarse\]" | tail -4
PYROPE_COARSE_TF32_COPY=1 PYROPE_COARSE_DEBUG=1 timeout 600 $CMD 2>&1 >/dev/null | grep "\[coarse\]" | tail -4
PYROPE_COARSE_TF32_COPY=1 timeout 600 $CMD 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('copy', d['ms_per_step'], d['roofline']['stage_ms'])"
timeout 600 $CMD 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('direct', d['ms_per_step'], d['roofline']['stage_ms'])"
