mkdir -p gpurun_out
PYROPE_COARSE_DEBUG=1 python scripts/bench_coarse.py 2>&1 | grep -E "coarse" | tail -3
python scripts/bench_coarse.py 65536 1250 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests_v4.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2_tests_v4.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v4.json 2> gpurun_out/r2_bench_v4.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_v4.err
CMD="python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 --profile-step"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && \
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2_step_v4 $CMD > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log
