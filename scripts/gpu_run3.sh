mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_coarse.py -x -q 2>&1 | tail -3
python scripts/bench_coarse.py 2>&1 | tail -1
PYROPE_COARSE_STREAMING=1 python scripts/bench_coarse.py 2>&1 | tail -1
python scripts/bench_coarse.py 65536 1250 2>&1 | tail -1
PROFILE_ONE=1 python scripts/bench_coarse.py > gpurun_out/plain_coarse.log 2>&1 && PROFILE_ONE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2_coarse_v1 python scripts/bench_coarse.py > gpurun_out/ncu_coarse.log 2>&1; echo ncu rc=$?
