// common.cuh — device helpers shared by the scan kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pyrope {

constexpr int kMaxTopK = 1024;

// ---------------------------------------------------------------------------------------------
// Candidate keys.  Every top-k structure in the library works on one 64-bit key per candidate:
//   high 32 bits = order-preserving image of the fp32 score (larger = better),
//   low  32 bits = ~position, so that among equal scores the LOWER position wins when keys are
//                  compared as unsigned integers.  key == 0 means "empty".
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t score_to_ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_score(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t pos) {
    return ((uint64_t)score_to_ord(score) << 32) | (uint64_t)(~pos);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ord_to_score((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_pos(uint64_t k) { return ~(uint32_t)k; }

__host__ __device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------------------------
// Bitonic sort, descending, of P (power of two) keys in shared memory by `nthr` cooperating
// threads (a warp with WARP=true, else the whole CTA).  Pads are zeros and sink to the end.
// ---------------------------------------------------------------------------------------------
template <bool WARP>
__device__ __forceinline__ void group_sync() {
    if (WARP) __syncwarp(); else __syncthreads();
}

template <bool WARP>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* s, int P, int tid, int nthr) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (P >> 1); i += nthr) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = s[lo], b = s[hi];
                if ((a < b) == desc) { s[lo] = b; s[hi] = a; }
            }
            group_sync<WARP>();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-wide candidate queue in shared memory (used by the list-scan kernels): threads push keys
// above the running threshold; when the queue may overflow the CTA sorts it, keeps the best k
// and raises the threshold to the k-th key.
// ---------------------------------------------------------------------------------------------
struct CtaQueue {
    uint64_t* keys;   // [cap] shared
    int* cnt;         // shared
    uint64_t* thr;    // shared: accept only keys > *thr
    int cap, k;

    __device__ __forceinline__ void reset(int tid) {
        if (tid == 0) { *cnt = 0; *thr = 0; }
    }
    __device__ __forceinline__ void push(uint64_t key) {
        if (key > *thr) {
            int pos = atomicAdd(cnt, 1);
            if (pos < cap) keys[pos] = key;
        }
    }
    // All threads of the CTA must call.  Leaves the best min(cnt,k) keys sorted descending.
    __device__ __forceinline__ void prune(int tid, int nthr) {
        __syncthreads();
        int n = min(*cnt, cap);
        int P = next_pow2(max(n, 2));
        for (int i = n + tid; i < P; i += nthr) keys[i] = 0;
        __syncthreads();
        bitonic_sort_desc<false>(keys, P, tid, nthr);
        if (tid == 0) {
            int keep = min(n, k);
            *cnt = keep;
            *thr = (keep == k) ? keys[k - 1] : 0;
        }
        __syncthreads();
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pyrope
