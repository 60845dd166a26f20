mkdir -p gpurun_out
python profiles/measure_tf32_peak.py > gpurun_out/r2_tf32.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_v1.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests_v1.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_v1.err
CMD="python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 --profile-step"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && \
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:flat_tc_kernel|flat_rescore|lm_prepare|lm_final|lm_seed|tc_gmax|coarse_rerank' -o gpurun_out/r2_coarse_v0 $CMD > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
