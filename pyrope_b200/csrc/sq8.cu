// sq8.cu — the FLAT index's 8-bit scalar-quantised scan (BruteForceVectorIndex.EnableQuantization).
//
// Replaces ScalarQuantizer.Quantize (Vector/ScalarQuantizer.cs:22-62: per-VECTOR min / max, (v - min) * (255 / range),
// Math.Round to even, clamp) and the quantised branch of BruteForceVectorIndex.Search (BruteForceVectorIndex.cs:297-336):
// the query is quantised with ITS OWN min / max and rows are ranked by the integer distance between the byte
// vectors — score = -(float)L2Squared8Bit (VectorMath.cs:447-560) for L2, (float)DotProduct8Bit (:565-) for inner
// product AND cosine.  Everything is integer or correctly rounded, so bytes and scores are bit-exact.
//
// Layout: X8[cap][dpad] u8 (dpad = dim rounded up to 16, zero padded: padding contributes 0 to both distances),
// qvalid[cap] u8 (rows written while quantisation was off have no quantised form: counted by MaxScans, never
// returned — :312-322).  HBM-bound byte scan: a warp reads rows as 16-byte words (one row per lane, rows of a
// warp are contiguous so every 128-byte line fetched is used in full), eight queries of the tile sit in shared
// memory, |a-b| by __vabsdiffu4 and the sums by DP4A.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace pyrope {
namespace {

constexpr int SQ_TQ = 8;       // queries per CTA
constexpr int SQ_THREADS = 256;

// one warp per row: min / max, then bytes
__global__ void sq8_quantize_kernel(const float* __restrict__ X, int64_t n, int dim, int64_t ldx, uint8_t* __restrict__ out,
                                    int dpad, uint8_t* __restrict__ qvalid) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const float* x = X + r * ldx;
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;  // float.MaxValue / float.MinValue (:36-37)
    for (int d = lane; d < dim; d += 32) {
        const float v = __ldg(x + d);
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const float range = __fsub_rn(mx, mn);
    const float scale = range == 0.f ? 0.f : __fdiv_rn(255.0f, range);
    uint8_t* o8 = out + r * dpad;
    for (int d = lane; d < dpad; d += 32) {
        int b = 0;
        if (d < dim && range != 0.f) {
            const float nz = __fmul_rn(__fsub_rn(__ldg(x + d), mn), scale);
            b = __float2int_rn(nz);  // Math.Round: to nearest, ties to even
            b = min(max(b, 0), 255);
        }
        o8[d] = (uint8_t)b;
    }
    if (lane == 0 && qvalid) qvalid[r] = 1;
}

struct Sq8Params {
    const uint8_t* Q8; int64_t nq; int dpad;
    const uint8_t* X8; int64_t n_scan; const uint8_t* dead; const uint8_t* qvalid; const int64_t* labels;
    int metric, k, cap, splits;
    PairOut out;
};

__global__ void __launch_bounds__(SQ_THREADS) sq8_scan_kernel(Sq8Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);                          // [TQ][cap]
    uint32_t* q8 = reinterpret_cast<uint32_t*>(keys + (size_t)SQ_TQ * p.cap);         // [TQ][dpad/4]
    __shared__ int s_cnt[SQ_TQ];
    __shared__ uint64_t s_thr[SQ_TQ];
    const int tid = threadIdx.x;
    const int64_t qtiles = (p.nq + SQ_TQ - 1) / SQ_TQ;
    const int64_t qt = blockIdx.x % qtiles, sp = blockIdx.x / qtiles;
    const int64_t q0 = qt * SQ_TQ;
    const int nqt = (int)min((int64_t)SQ_TQ, p.nq - q0);
    const int dw = p.dpad / 4;
    for (int i = tid; i < SQ_TQ * dw; i += SQ_THREADS) {
        const int j = i / dw, w = i - j * dw;
        q8[i] = j < nqt ? reinterpret_cast<const uint32_t*>(p.Q8 + (q0 + j) * p.dpad)[w] : 0u;
    }
    if (tid < SQ_TQ) { s_cnt[tid] = 0; s_thr[tid] = 0ull; }
    __syncthreads();
    const int64_t per = (p.n_scan + p.splits - 1) / p.splits;
    const int64_t r_begin = sp * per, r_end = min(p.n_scan, r_begin + per);
    for (int64_t base = r_begin; base < r_end; base += SQ_THREADS) {
        // queues: a batch adds at most SQ_THREADS keys per query
        bool need = false;
        for (int j = 0; j < nqt; ++j) need |= s_cnt[j] + SQ_THREADS > p.cap;
        if (need) {
            for (int j = 0; j < nqt; ++j) {
                CtaQueue Qu{keys + (size_t)j * p.cap, &s_cnt[j], &s_thr[j], p.cap, p.k};
                Qu.prune(tid, SQ_THREADS);
            }
        }
        const int64_t r = base + tid;
        if (r < r_end && !(p.dead && p.dead[r]) && (!p.qvalid || p.qvalid[r])) {
            unsigned acc[SQ_TQ];
#pragma unroll
            for (int j = 0; j < SQ_TQ; ++j) acc[j] = 0u;
            const uint4* xr = reinterpret_cast<const uint4*>(p.X8 + r * p.dpad);
            for (int w4 = 0; w4 < dw / 4; ++w4) {
                const uint4 xv = __ldg(xr + w4);
                const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int j = 0; j < SQ_TQ; ++j) {
                        const uint32_t qw = q8[j * dw + w4 * 4 + u];
                        if (p.metric == kL2) {
                            const uint32_t df = __vabsdiffu4(xw[u], qw);
                            acc[j] = __dp4a(df, df, acc[j]);
                        } else {
                            acc[j] = __dp4a(xw[u], qw, acc[j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < SQ_TQ; ++j) {
                if (j < nqt) {
                    // (float)long: exact below 2^24, round-to-nearest above, like the reference's conversion
                    const float score = p.metric == kL2 ? -(float)acc[j] : (float)acc[j];
                    const uint64_t key = make_key(score, (uint32_t)r);
                    if (key > s_thr[j]) {
                        const int pos = atomicAdd(&s_cnt[j], 1);
                        if (pos < p.cap) keys[(size_t)j * p.cap + pos] = key;
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int j = 0; j < nqt; ++j) {
        CtaQueue Qu{keys + (size_t)j * p.cap, &s_cnt[j], &s_thr[j], p.cap, p.k};
        Qu.prune(tid, SQ_THREADS);
        const int keep = s_cnt[j];
        const int64_t ob = ((q0 + j) * p.out.parts_total + p.out.part_base + sp) * (int64_t)p.k;
        for (int i = tid; i < p.k; i += SQ_THREADS) {
            if (i < keep) {
                const uint64_t key = keys[(size_t)j * p.cap + i];
                const uint32_t pos = key_pos(key);
                p.out.scores[ob + i] = key_score(key);
                p.out.labels[ob + i] = p.labels ? p.labels[pos] : (int64_t)pos;
            } else {
                p.out.scores[ob + i] = 0.f;
                p.out.labels[ob + i] = -1;
            }
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_sq8_quantize(const float* X, int64_t n, int dim, int64_t ldx, uint8_t* out, int dpad, uint8_t* qvalid,
                                cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    sq8_quantize_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(X, n, dim, ldx, out, dpad, qvalid);
    return cudaGetLastError();
}

int sq8_scan_cap(int k) { return next_pow2(k + SQ_THREADS); }

int sq8_pick_splits(int64_t nq, int64_t n_scan, int k, int num_sms) {
    const int64_t qtiles = (nq + SQ_TQ - 1) / SQ_TQ;
    int64_t want = (4 * (int64_t)num_sms + qtiles - 1) / qtiles;
    const int64_t by_rows = std::max<int64_t>(1, n_scan / (4 * SQ_THREADS));
    const int64_t lim = kMergeMaxCandidates / (k > 0 ? k : 1);
    want = std::min(std::min(want, by_rows), lim);
    return (int)std::max<int64_t>(1, want);
}

// writes `splits` parts starting at out.part_base
cudaError_t launch_sq8_scan(const uint8_t* Q8, int64_t nq, int dpad, const uint8_t* X8, int64_t n_scan, const uint8_t* dead,
                            const uint8_t* qvalid, const int64_t* labels, int metric, int k, int splits, PairOut out,
                            cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    Sq8Params p{};
    p.Q8 = Q8; p.nq = nq; p.dpad = dpad; p.X8 = X8; p.n_scan = n_scan; p.dead = dead; p.qvalid = qvalid; p.labels = labels;
    p.metric = metric; p.k = k; p.cap = sq8_scan_cap(k); p.splits = splits; p.out = out;
    const size_t smem = sizeof(uint64_t) * (size_t)SQ_TQ * p.cap + (size_t)SQ_TQ * dpad;
    cudaError_t e = cudaFuncSetAttribute(sq8_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t qtiles = (nq + SQ_TQ - 1) / SQ_TQ;
    sq8_scan_kernel<<<(unsigned)(qtiles * splits), SQ_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace pyrope
