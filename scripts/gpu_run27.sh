mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lm.py -x -q 2>&1 | tail -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v27_n2.json 2> gpurun_out/r2_bench_v27_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v27_n2.json'))
print('C5 N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['stage_ms'], d['roofline']['kernel_ms'], d['roofline']['step_phases_ms'], d['gpu_launches'])
PY
