mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for MODE in peer nccl; do
FLAG=""; [ $MODE = nccl ] && FLAG="--nccl-exchange"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c5 --secondary none --steps 10 --warmup 3 --recall-queries 0 $FLAG > gpurun_out/r2_bench_v28_n2_$MODE.json 2> gpurun_out/r2_bench_v28_n2_$MODE.err; echo "bench $MODE rc=$?"
done
python - <<'PY'
import json
for m in ('peer','nccl'):
    try:
        d=json.load(open(f'gpurun_out/r2_bench_v28_n2_{m}.json'))
        print(m, d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['mismatch'], d['roofline']['kernel_ms'], d['roofline']['step_phases_ms'], d['gpu_launches'], d['details']['collective'])
    except Exception as e: print(m,'ERR',e)
PY
tail -3 gpurun_out/r2_bench_v28_n2_peer.err
