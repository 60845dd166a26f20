mkdir -p gpurun_out
CMD="python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 --profile-step"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && timeout 1200 ncu --profile-from-start off --set full --clock-control none -k "regex:coarse|ivfpq_lm|lm_prepare" -o gpurun_out/r2_step_v37 $CMD > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_step_v37.ncu-rep
