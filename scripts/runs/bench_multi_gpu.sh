mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --recall-queries 0 > gpurun_out/r2_bench_v36_c5_c4_n$N.json 2> gpurun_out/r2_bench_v36_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_v36_c5_c4_n$N.json'))
print('C5 N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['kernel_ms'], d['roofline']['step_phases_ms'], d['gpu_launches'], d['details']['collective'])
s=d.get('secondary')
if s: print('C4 N=$N', s['value'], s['ms_per_step'], s['e2e']['value'], s['parity']['mismatch'], s['roofline']['kernel'], s['roofline']['kernel_ms'], s['roofline'].get('step_phases_ms'))
PY
tail -2 gpurun_out/r2_bench_v36_n$N.err
