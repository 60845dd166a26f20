// pq.cu — K5 IVF_PQ: per-probe ADC lookup-table build + PQ-code scan with fused top-k.
//
// Replaces ProductQuantizer.ComputeDistanceTable (ProductQuantizer.cs:98-120) and the ADC loop of
// IvfPqVectorIndex.Search (IvfPqVectorIndex.cs:152-199): for every probed list the residual query
// r = q - centroid is formed, LUT[m][k] = ||r_m - codeword_mk||^2 is built in shared memory, and
// every code of the list is scored as sum_m LUT[m][code_m] (score = -dist for every metric, :194).
//
// Fast kernel (m in {4,8,16,32}, k <= 256):
//   * the PQ codebook (dim*k fp32, 128 KiB for d=128) lives in shared memory for the CTA's
//     lifetime, 16-byte chunks XOR-swizzled so that 8 consecutive codewords hit 8 distinct
//     bank groups;
//   * the LUT is stored bank-per-sub-quantiser: word address = code*32 + bank, bank = table index
//     (+ replica offset when m < 32).  During the scan lane l looks up table (l+s) mod m at step
//     s, so the 32 lanes of a warp always touch 32 distinct banks whatever the code bytes are —
//     the random 4-byte lookups are conflict-free by construction;
//   * the code row (m bytes) is one 32/64/128-bit load per lane (coalesced across the warp) and
//     is byte-rotated once per lane so step s reads a compile-time byte position;
//   * candidates go through the CTA queue of common.cuh (threshold filter, rare bitonic prune).
// Generic kernel: any m, plain [m][k] table — the fallback for other shapes and the on-GPU
// cross-check of the fast kernel in tests.
#include "common.cuh"
#include "kernels.h"

namespace pyrope {
namespace {

constexpr int QCAP = 2048;

// ---------------------------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------------------------
constexpr int GNT = 256;

__global__ void __launch_bounds__(GNT) ivfpq_scan_generic_kernel(IvfPqScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);   // [QCAP]
    float* lut = reinterpret_cast<float*>(keys + QCAP);        // [m][ksub]
    float* res = lut + (size_t)p.m * p.ksub;                   // [dim]
    __shared__ int s_cnt;
    __shared__ uint64_t s_thr;

    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const int y = blockIdx.y;
    const int dim = p.dim, m = p.m, K = p.ksub, sub = dim / m;

    CtaQueue Qu{keys, &s_cnt, &s_thr, QCAP, p.k};
    Qu.reset(tid);
    __syncthreads();
    uint64_t thr_reg = 0;
    int cnt_known = 0;

    for (int pr = y; pr < p.nprobe; pr += p.groups) {
        const int64_t l = p.probes[q * p.nprobe + pr];
        if (l < 0) continue;
        const int64_t beg = p.list_off[l], end = p.list_off[l + 1];
        if (beg >= end) continue;  // IvfPqVectorIndex.cs:155 skip empty lists
        __syncthreads();           // previous LUT fully consumed
        for (int i = tid; i < dim; i += GNT) res[i] = p.Q[q * dim + i] - p.centroids[l * dim + i];
        __syncthreads();
        for (int e = tid; e < m * K; e += GNT) {
            int mi = e / K, kk = e - mi * K;
            const float* cw = p.codebook + ((size_t)mi * K + kk) * sub;
            const float* r = res + mi * sub;
            float a = 0.f;
            for (int j = 0; j < sub; ++j) { float d = r[j] - cw[j]; a = fmaf(d, d, a); }
            lut[e] = a;
        }
        __syncthreads();
        for (int64_t c0 = beg; c0 < end; c0 += GNT) {
            if (cnt_known + GNT > QCAP || (thr_reg == 0 && cnt_known >= 2 * p.k && cnt_known >= 64)) {
                __syncthreads();
                int actual = s_cnt;
                if (actual + GNT > QCAP || (s_thr == 0 && actual >= 2 * p.k && actual >= 64)) {
                    Qu.prune(tid, GNT);
                    actual = s_cnt;
                }
                __syncthreads();
                cnt_known = actual;
                thr_reg = s_thr;
            }
            const int64_t r = c0 + tid;
            if (r < end && !(p.dead && p.dead[r])) {
                const uint8_t* code = p.codes + r * m;
                float dist = 0.f;
                for (int mi = 0; mi < m; ++mi) dist += lut[mi * K + code[mi]];
                uint64_t key = make_key(-dist, (uint32_t)r);
                if (key > thr_reg) {
                    int pos = atomicAdd(&s_cnt, 1);
                    if (pos < QCAP) keys[pos] = key;
                }
            }
            cnt_known += (int)min((int64_t)GNT, end - c0);
        }
    }
    Qu.prune(tid, GNT);
    const int keep = s_cnt;
    const int64_t ob = (q * p.out.parts_total + p.out.part_base + y) * (int64_t)p.k;
    for (int i = tid; i < p.k; i += GNT) {
        if (i < keep) {
            uint64_t key = keys[i];
            p.out.scores[ob + i] = key_score(key);
            p.out.labels[ob + i] = p.labels[key_pos(key)];
        } else {
            p.out.scores[ob + i] = 0.f;
            p.out.labels[ob + i] = -1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fast kernel
// ---------------------------------------------------------------------------------------------
constexpr int FNT = 512;      // 16 warps
constexpr int LUT_ROWS = 256; // code values
constexpr int LUT_WORDS = LUT_ROWS * 32;

// swizzle of 16-byte chunk c of codeword kk within its row of C chunks (C = sub/4)
__device__ __forceinline__ int cw_swz(int c, int kk, int C) {
    if ((C & (C - 1)) != 0) return c;  // non power of two: no swizzle
    if (C >= 8) return c ^ (kk & 7);
    if (C == 4) return c ^ ((kk >> 1) & 3);
    if (C == 2) return c ^ ((kk >> 2) & 1);
    return c;
}

template <int M>
struct LutMap {
    static constexpr int MB = M;                         // M <= 32
    static constexpr int R = 32 / MB;                    // replicas
    static constexpr int P = (MB < 8) ? MB : 8;          // steps per 8-codeword block
    static constexpr int NB = (MB >= 8) ? (32 / MB) : 4; // 8-blocks per warp unit
    static constexpr int UNITS = LUT_ROWS / (8 * NB);
    __device__ static __forceinline__ int mi(int lane) { return lane % MB; }
    __device__ static __forceinline__ int blk(int lane) { return (MB >= 8) ? ((lane >> 3) / (MB / 8 > 0 ? MB / 8 : 1)) : (lane >> 3); }
    __device__ static __forceinline__ int rep(int lane) { return lane / MB; }
};

template <int W>
__device__ __forceinline__ void rotate_bytes(uint32_t (&w)[W], int rb) {
    // new byte s = old byte (s + rb) mod 4W
    int a = rb >> 2;
#pragma unroll
    for (int sft = 1; sft < W; sft <<= 1) {
        if (W > sft) {
            uint32_t t[W];
            bool on = (a & sft) != 0;
#pragma unroll
            for (int i = 0; i < W; ++i) t[i] = on ? w[(i + sft) % W] : w[i];
#pragma unroll
            for (int i = 0; i < W; ++i) w[i] = t[i];
        }
    }
    int bits = (rb & 3) * 8;
    uint32_t t[W];
#pragma unroll
    for (int i = 0; i < W; ++i) t[i] = __funnelshift_r(w[i], w[(i + 1) % W], bits);
#pragma unroll
    for (int i = 0; i < W; ++i) w[i] = t[i];
}

template <int W>
__device__ __forceinline__ void load_code(const uint8_t* ptr, uint32_t (&w)[W]) {
    if (W == 1) {
        w[0] = __ldg(reinterpret_cast<const uint32_t*>(ptr));
    } else if (W == 2) {
        uint2 v = __ldg(reinterpret_cast<const uint2*>(ptr));
        w[0] = v.x; w[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(ptr) + i);
            w[4 * i + 0] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
}

template <int M, bool CB_SMEM>
__global__ void __launch_bounds__(FNT, 1) ivfpq_scan_fast_kernel(IvfPqScanParams p) {
    using Map = LutMap<M>;
    constexpr int W = M / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);                       // [256][32]
    uint64_t* keys = reinterpret_cast<uint64_t*>(lut + LUT_WORDS);          // [QCAP]
    float* res = reinterpret_cast<float*>(keys + QCAP);                     // [dim]
    float* cbs = res + ((p.dim + 3) / 4) * 4;                               // [m][K][sub] swizzled (CB_SMEM)
    __shared__ int s_cnt;
    __shared__ uint64_t s_thr;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const int y = blockIdx.y;
    const int dim = p.dim, K = p.ksub, sub = dim / M, C = sub / 4;

    if (CB_SMEM) {
        // stage the codebook once per CTA: chunk (row, c) -> row*C + swz(c)
        const int rows = M * K;
        const float4* src = reinterpret_cast<const float4*>(p.codebook);
        float4* dst = reinterpret_cast<float4*>(cbs);
        for (int i = tid; i < rows * C; i += FNT) {
            int row = i / C, c = i - row * C;
            int kk = row % K;
            dst[row * C + cw_swz(c, kk, C)] = __ldg(src + i);
        }
    }
    CtaQueue Qu{keys, &s_cnt, &s_thr, QCAP, p.k};
    Qu.reset(tid);
    __syncthreads();
    uint64_t thr_reg = 0;
    int cnt_known = 0;

    // per-lane constants
    const int my_mi = Map::mi(lane);
    const int my_blk = Map::blk(lane);
    const int my_rep = Map::rep(lane);
    const int q8 = lane & 7;
    const int rot = lane % M;                       // byte rotation of this lane's codes
    const int bank_hi = lane & ~(M - 1);            // replica base bank during the scan

    for (int pr = y; pr < p.nprobe; pr += p.groups) {
        const int64_t l = p.probes[q * p.nprobe + pr];
        if (l < 0) continue;
        const int64_t beg = p.list_off[l], end = p.list_off[l + 1];
        if (beg >= end) continue;
        __syncthreads();  // previous LUT fully consumed
        for (int i = tid; i < dim; i += FNT) res[i] = p.Q[q * dim + i] - p.centroids[l * dim + i];
        __syncthreads();

        // ---- LUT build: lane <-> sub-quantiser, conflict-free stores into bank = table index
        for (int u = warp; u < Map::UNITS; u += FNT / 32) {
#pragma unroll
            for (int j = 0; j < Map::P; ++j) {
                const int kk = 8 * (Map::NB * u + my_blk) + ((q8 + j) & 7);
                float a = 0.f;
                if (kk < K) {
                    const float* r = res + my_mi * sub;
                    if (CB_SMEM) {
                        const float4* row = reinterpret_cast<const float4*>(cbs) + ((size_t)my_mi * K + kk) * C;
                        for (int c = 0; c < C; ++c) {
                            float4 cv = row[cw_swz(c, kk, C)];
                            float4 rv = *reinterpret_cast<const float4*>(r + 4 * c);
                            float d0 = rv.x - cv.x, d1 = rv.y - cv.y, d2 = rv.z - cv.z, d3 = rv.w - cv.w;
                            a = fmaf(d0, d0, a); a = fmaf(d1, d1, a); a = fmaf(d2, d2, a); a = fmaf(d3, d3, a);
                        }
                    } else {
                        const float4* row = reinterpret_cast<const float4*>(p.codebook) + ((size_t)my_mi * K + kk) * C;
                        for (int c = 0; c < C; ++c) {
                            float4 cv = __ldg(row + c);
                            float4 rv = *reinterpret_cast<const float4*>(r + 4 * c);
                            float d0 = rv.x - cv.x, d1 = rv.y - cv.y, d2 = rv.z - cv.z, d3 = rv.w - cv.w;
                            a = fmaf(d0, d0, a); a = fmaf(d1, d1, a); a = fmaf(d2, d2, a); a = fmaf(d3, d3, a);
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < Map::R; ++t)
                    lut[kk * 32 + my_mi + M * ((my_rep + t) % Map::R)] = a;
            }
        }
        __syncthreads();

        // ---- scan
        for (int64_t c0 = beg; c0 < end; c0 += FNT) {
            if (cnt_known + FNT > QCAP || (thr_reg == 0 && cnt_known >= 2 * p.k && cnt_known >= 64)) {
                __syncthreads();
                int actual = s_cnt;
                if (actual + FNT > QCAP || (s_thr == 0 && actual >= 2 * p.k && actual >= 64)) {
                    Qu.prune(tid, FNT);
                    actual = s_cnt;
                }
                __syncthreads();
                cnt_known = actual;
                thr_reg = s_thr;
            }
            const int64_t r = c0 + tid;
            if (r < end) {
                uint32_t w[W];
                load_code<W>(p.codes + r * M, w);
                const bool live = !(p.dead && p.dead[r]);
                rotate_bytes<W>(w, rot);
                float dist = 0.f;
#pragma unroll
                for (int s = 0; s < M; ++s) {
                    const uint32_t byte = (w[s >> 2] >> (8 * (s & 3))) & 0xffu;
                    const int bank = ((lane + s) & (M - 1)) | bank_hi;
                    dist += lut[byte * 32 + bank];
                }
                if (live) {
                    uint64_t key = make_key(-dist, (uint32_t)r);
                    if (key > thr_reg) {
                        int pos = atomicAdd(&s_cnt, 1);
                        if (pos < QCAP) keys[pos] = key;
                    }
                }
            }
            cnt_known += (int)min((int64_t)FNT, end - c0);
        }
    }
    Qu.prune(tid, FNT);
    const int keep = s_cnt;
    const int64_t ob = (q * p.out.parts_total + p.out.part_base + y) * (int64_t)p.k;
    for (int i = tid; i < p.k; i += FNT) {
        if (i < keep) {
            uint64_t key = keys[i];
            p.out.scores[ob + i] = key_score(key);
            p.out.labels[ob + i] = p.labels[key_pos(key)];
        } else {
            p.out.scores[ob + i] = 0.f;
            p.out.labels[ob + i] = -1;
        }
    }
}

// ProductQuantizer.ComputeDistanceTable for parity tests: table[q][m][k]
__global__ void pq_table_kernel(const float* Q, int64_t nq, int dim, const float* cb, int m, int K,
                                float* table) {
    int64_t q = blockIdx.x;
    int sub = dim / m;
    for (int e = threadIdx.x; e < m * K; e += blockDim.x) {
        int mi = e / K, kk = e - mi * K;
        const float* cw = cb + ((size_t)mi * K + kk) * sub;
        const float* r = Q + q * dim + mi * sub;
        float a = 0.f;
        for (int j = 0; j < sub; ++j) { float d = r[j] - cw[j]; a = fmaf(d, d, a); }
        table[(q * m + mi) * K + kk] = a;
    }
}

template <int M>
cudaError_t launch_fast(const IvfPqScanParams& p, cudaStream_t st) {
    size_t base = sizeof(float) * LUT_WORDS + sizeof(uint64_t) * QCAP + sizeof(float) * (size_t)((p.dim + 3) / 4 * 4);
    size_t cb_bytes = sizeof(float) * (size_t)p.dim * p.ksub;
    bool cb_smem = base + cb_bytes <= 220 * 1024;
    size_t smem = base + (cb_smem ? cb_bytes : 0);
    dim3 grid((unsigned)p.nq, (unsigned)p.groups);
    cudaError_t e;
    if (cb_smem) {
        e = cudaFuncSetAttribute(ivfpq_scan_fast_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ivfpq_scan_fast_kernel<M, true><<<grid, FNT, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(ivfpq_scan_fast_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ivfpq_scan_fast_kernel<M, false><<<grid, FNT, smem, st>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_ivfpq_scan(const IvfPqScanParams& p, cudaStream_t st) {
    if (p.nq <= 0) return cudaSuccess;
    const int sub = p.dim / p.m;
    const bool fast_ok = !p.force_generic && p.ksub <= 256 && (sub % 4 == 0) &&
                         (p.m == 4 || p.m == 8 || p.m == 16 || p.m == 32);
    if (fast_ok) {
        switch (p.m) {
            case 4: return launch_fast<4>(p, st);
            case 8: return launch_fast<8>(p, st);
            case 16: return launch_fast<16>(p, st);
            default: return launch_fast<32>(p, st);
        }
    }
    size_t smem = sizeof(uint64_t) * QCAP + sizeof(float) * ((size_t)p.m * p.ksub + p.dim);
    if (smem > 220 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(ivfpq_scan_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)p.nq, (unsigned)p.groups);
    ivfpq_scan_generic_kernel<<<grid, GNT, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_pq_distance_table(const float* Q, int64_t nq, int dim, const float* codebook, int m,
                                     int k, float* table, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    pq_table_kernel<<<(unsigned)nq, 256, 0, st>>>(Q, nq, dim, codebook, m, k, table);
    return cudaGetLastError();
}

}  // namespace pyrope
