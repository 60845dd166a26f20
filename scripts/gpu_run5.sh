mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_coarse.py tests/test_gpu_fullshape.py -x -q > gpurun_out/r2_tests_v5a.log 2>&1; echo "lm/coarse tests rc=$?"; tail -6 gpurun_out/r2_tests_v5a.log
PYROPE_COARSE_DEBUG=1 timeout 300 python scripts/bench_coarse.py 2>&1 | grep -E "coarse" | tail -2
timeout 300 python scripts/bench_coarse.py 65536 1250 2>&1 | grep "coarse probe"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_v5.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2_tests_v5.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v5.json 2> gpurun_out/r2_bench_v5.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_v5.err
PYROPE_LM_STAGES=1 timeout 600 python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 > gpurun_out/r2_stages_v5.json 2> gpurun_out/r2_stages_v5.err; tail -2 gpurun_out/r2_stages_v5.err
