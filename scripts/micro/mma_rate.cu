// tcgen05.mma issue / execution rate on one SM: cycles per instruction for kind::tf32 and kind::f16 (bf16), N = 128 / 256,
// with 1, 2 or 4 accumulators used round-robin (dependent-accumulate distance).  Operands are whatever shared memory holds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int KIND>  // 0 tf32, 1 f16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND, int N, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_s;
    const uint32_t idesc = (1u << 4) | ((KIND == 0 ? 2u : 1u) << 7) | ((KIND == 0 ? 2u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1 && elect_one()) {
        const uint64_t a0 = make_desc(smem_u32(smem)), b0 = make_desc(smem_u32(smem + 65536));
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                // a new 32-byte K slice every instruction (4 per 128-byte row), operands wander over 64 KiB each
                const uint64_t adv = (uint64_t)((u & 3) * 2 + ((u >> 2) * 1024));
                mma<KIND>(tm + (uint32_t)((u % NACC) * N), a0 + adv, b0 + adv, idesc, (it | (u / NACC)) != 0);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    }
}
// MMA stream on accumulator columns [0, 256) while four other warps read columns [256, 512) with tcgen05.ld (the
// epilogue of a double-buffered GEMM): do the two slow each other down, and what is the TMEM read rate?
__global__ void __launch_bounds__(192, 1) overlap_kernel(long long* out, int iters, int with_mma, int with_ld) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    __shared__ int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1) {
        if (elect_one()) {
            const uint64_t a0 = make_desc(smem_u32(smem)), b0 = make_desc(smem_u32(smem + 65536));
            const long long t0 = clock64();
            if (with_mma) {
                for (int it = 0; it < iters; ++it) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint64_t adv = (uint64_t)((u & 3) * 2 + ((u >> 2) * 1024));
                        mma<0>(tm + (uint32_t)((u & 1) * 128), a0 + adv, b0 + adv, idesc, (it | (u >> 1)) != 0);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t ok = 0;
                while (!ok) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
                }
            } else {
                while (clock64() - t0 < 400000) {}
            }
            out[0] = clock64() - t0;
            *(volatile int*)&stop = 1;
        }
    } else if (warp >= 2 && with_ld) {
        const int ew = (warp - 2) & 3;
        long long n = 0;
        float acc = 0.f;
        const long long t0 = clock64();
        while (!*(volatile int*)&stop) {
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(tm + ((uint32_t)(ew * 32) << 16) + 256 + c * 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
                ++n;
            }
        }
        const long long t1 = clock64();
        if (lane == 0) { out[2 + ew * 2] = n; out[3 + ew * 2] = t1 - t0; }
        if (acc == 123.456f) out[15] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    }
}
void run_overlap(const char* name, int with_mma, int with_ld) {
    long long* d;
    cudaMalloc(&d, 128);
    cudaMemset(d, 0, 128);
    const int iters = 1024;
    cudaFuncSetAttribute(overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    overlap_kernel<<<1, 192, 160 * 1024>>>(d, iters, with_mma, with_ld);
    long long h[16] = {0};
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    double ldbytes = 0, ldcyc = 1;
    for (int w = 0; w < 4; ++w) { ldbytes += (double)h[2 + 2 * w] * 32 * 32 * 4; ldcyc = h[3 + 2 * w] > ldcyc ? h[3 + 2 * w] : ldcyc; }
    printf("%-44s", name);
    if (with_mma) printf(" %.1f cyc/mma (N=128)", h[0] / (16.0 * iters));
    if (with_ld) printf("  tcgen05.ld by 4 warps: %.1f B/clk", ldbytes / ldcyc);
    printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

template <int KIND, int N, int NACC>
void run(const char* name) {
    long long* d;
    cudaMalloc(&d, 16);
    const int iters = 256;
    cudaFuncSetAttribute(rate_kernel<KIND, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int rep = 0; rep < 2; ++rep) rate_kernel<KIND, N, NACC><<<1, 128, 160 * 1024>>>(d, iters);
    long long h[2] = {0, 0};
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double n = 16.0 * iters;
    printf("%-28s issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d)  %s\n", name, h[0] / n, h[1] / n, N / 2 * (KIND == 0 ? 1 : 1),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    run<0, 128, 1>("tf32 N=128 1 accumulator");
    run<0, 128, 2>("tf32 N=128 2 accumulators");
    run<0, 128, 4>("tf32 N=128 4 accumulators");
    run<0, 256, 1>("tf32 N=256 1 accumulator");
    run<0, 256, 2>("tf32 N=256 2 accumulators");
    run<1, 128, 1>("bf16 N=128 1 accumulator");
    run<1, 128, 2>("bf16 N=128 2 accumulators");
    run<1, 128, 4>("bf16 N=128 4 accumulators");
    run<1, 256, 1>("bf16 N=256 1 accumulator");
    run<1, 256, 2>("bf16 N=256 2 accumulators");
    run_overlap("MMA alone (2 accumulators)", 1, 0);
    run_overlap("tcgen05.ld alone (4 warps, other columns)", 0, 1);
    run_overlap("MMA + tcgen05.ld together", 1, 1);
    return 0;
}
