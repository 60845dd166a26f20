# IVF_FLAT with rows of 768 floats (2 M rows, nlist 1024, nprobe 16, 10 k queries): list-major wide kernels vs query-major
mkdir -p gpurun_out
timeout 900 python bench.py --workload c2w --secondary none --steps 5 --warmup 3 --recall-queries 0 > gpurun_out/bench_c2w_list_major.json 2> gpurun_out/bench_c2w.err; echo "bench rc=$?"
PYROPE_PQ_LM=0 timeout 900 python bench.py --workload c2w --secondary none --steps 3 --warmup 3 --recall-queries 0 --no-cpu > gpurun_out/bench_c2w_query_major.json 2> gpurun_out/bench_c2w_qm.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('list_major','query_major'):
    d=json.load(open(f'gpurun_out/bench_c2w_{f}.json'))
    print(f, d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatch'), d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['frac'])
PY
tail -2 gpurun_out/bench_c2w.err
