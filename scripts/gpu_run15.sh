./scripts/micro/mma_rate > gpurun_out/r2_mma_rate_v2.txt 2>&1; cat gpurun_out/r2_mma_rate_v2.txt
