// exact_arith.cuh — the reference's fp32 evaluation orders, restated with __fmul_rn/__fadd_rn/__fsub_rn
// so nvcc can never contract them into FMAs (conventions in DESIGN.md: 8-lane Vector<float>
// accumulators, separate multiply and add, pairwise horizontal sum).
//   a2_eval    : VectorMath.L2Squared / DotProduct        (VectorMath.cs:8-70)
//   a1_l2_eval : VectorMath.L2SquaredUnsafe                (VectorMath.cs:188-253)
//   norm_eval  : VectorMath.ComputeNorm                    (VectorMath.cs:72-100)
#pragma once
#include <cuda_runtime.h>

namespace pyrope {
namespace exact {

__device__ __forceinline__ float hsum8(const float (&v)[8]) {
    float lo = __fadd_rn(__fadd_rn(v[0], v[1]), __fadd_rn(v[2], v[3]));
    float hi = __fadd_rn(__fadd_rn(v[4], v[5]), __fadd_rn(v[6], v[7]));
    return __fadd_rn(lo, hi);
}

enum Arith { A2_L2 = 0, A2_DOT = 1, A1_L2 = 2 };

template <int OP>  // 0 l2, 1 dot
__device__ __forceinline__ float term(float a, float b) {
    if (OP == 0) { float d = __fsub_rn(a, b); return __fmul_rn(d, d); }
    return __fmul_rn(a, b);
}

// VectorMath.L2Squared / DotProduct: one 8-lane accumulator, scalar tail.  a: shared/global, b: global
template <int OP>
static __device__ float a2_eval(const float* a, const float* b, int n) {
    int i = 0;
    float sum = 0.f;
    if (n >= 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (; i <= n - 8; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], term<OP>(a[i + j], __ldg(b + i + j)));
        }
        sum = __fadd_rn(sum, hsum8(acc));
    }
    for (; i < n; ++i) sum = __fadd_rn(sum, term<OP>(a[i], __ldg(b + i)));
    return sum;
}

// VectorMath.L2SquaredUnsafe: 4 accumulators for len >= 32, remainder accumulator, scalar tail
static __device__ float a1_l2_eval(const float* a, const float* b, int n) {
    int i = 0;
    float sum = 0.f;
    if (n >= 32) {
        float a1[8], a2[8], a3[8], a4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a1[j] = a2[j] = a3[j] = a4[j] = 0.f;
        for (; i <= n - 32; i += 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a1[j] = __fadd_rn(a1[j], term<0>(a[i + j], __ldg(b + i + j)));
                a2[j] = __fadd_rn(a2[j], term<0>(a[i + 8 + j], __ldg(b + i + 8 + j)));
                a3[j] = __fadd_rn(a3[j], term<0>(a[i + 16 + j], __ldg(b + i + 16 + j)));
                a4[j] = __fadd_rn(a4[j], term<0>(a[i + 24 + j], __ldg(b + i + 24 + j)));
            }
        }
        float fin[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) fin[j] = __fadd_rn(__fadd_rn(__fadd_rn(a1[j], a2[j]), a3[j]), a4[j]);
        sum = __fadd_rn(sum, hsum8(fin));
    }
    if (i <= n - 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (; i <= n - 8; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], term<0>(a[i + j], __ldg(b + i + j)));
        }
        sum = __fadd_rn(sum, hsum8(acc));
    }
    for (; i < n; ++i) sum = __fadd_rn(sum, term<0>(a[i], __ldg(b + i)));
    return sum;
}

// ComputeNorm (VectorMath.cs:72-100)
static __device__ float norm_eval(const float* v, int n) {
    int i = 0;
    float sum = 0.f;
    if (n >= 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (; i <= n - 8; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], __fmul_rn(v[i + j], v[i + j]));
        }
        sum = __fadd_rn(sum, hsum8(acc));
    }
    for (; i < n; ++i) sum = __fadd_rn(sum, __fmul_rn(v[i], v[i]));
    return __fsqrt_rn(sum);
}


// L2SquaredUnsafe on a compile-time length with `a` in registers (N < 32: no unrolled main loop)
template <int N>
__device__ __forceinline__ float a1_l2_fixed(const float (&a)[N], const float* __restrict__ b) {
    static_assert(N < 32, "short sub-vectors only");
    int i = 0;
    float sum = 0.f;
    if (N >= 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (; i <= N - 8; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], term<0>(a[i + j], __ldg(b + i + j)));
        }
        sum = __fadd_rn(sum, hsum8(acc));
    }
#pragma unroll
    for (; i < N; ++i) sum = __fadd_rn(sum, term<0>(a[i], __ldg(b + i)));
    return sum;
}

}  // namespace exact
}  // namespace pyrope
