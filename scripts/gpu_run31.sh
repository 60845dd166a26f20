mkdir -p gpurun_out
CMD="python bench.py --workload c4 --scale 0.05 --secondary none --steps 2 --warmup 2 --no-cpu --recall-queries 0 --profile-step"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:flat_tc_kernel -o gpurun_out/r2_flat_fp16_v31 $CMD > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/plain.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
