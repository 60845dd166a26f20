mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coarse.py tests/test_gpu_fullshape.py tests/test_gpu_tc.py -x -q > gpurun_out/r2_tests_v2a.log 2>&1; echo "coarse tests rc=$?"; tail -15 gpurun_out/r2_tests_v2a.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_v2.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2_tests_v2.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2_bench_v2.err
PYROPE_LM_STAGES=1 timeout 600 python bench.py --workload c5 --secondary none --steps 3 --warmup 3 --no-cpu --recall-queries 0 > gpurun_out/r2_stages_v2.json 2> gpurun_out/r2_stages_v2.err; tail -4 gpurun_out/r2_stages_v2.err
