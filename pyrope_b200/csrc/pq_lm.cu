// pq_lm.cu — K5 list-major IVF_PQ ADC scan (the batched hot path of BASELINE config 5).
//
// Replaces, for a whole query batch, ProductQuantizer.ComputeDistanceTable (ProductQuantizer.cs:98-120)
// and the ADC loop of IvfPqVectorIndex.Search (IvfPqVectorIndex.cs:152-199).  The reference walks
// query -> probed list -> code; a batch of 10^4 queries x 64 probes hits every inverted list ~10 times,
// so this kernel inverts the loop: (query, probe) pairs are grouped BY LIST and each work item is one
// list x up to four of the queries that probe it.
//
//   * the four queries' lookup tables are interleaved as float4 — LUT[e][m] = {q0,q1,q2,q3} at byte
//     e*256 + m*16 — so one 128-bit shared-memory load serves four (query, code) lookups;
//   * during the scan lane l reads table (l+t) mod 16 at step t: the eight lanes of every quarter warp
//     touch eight distinct 16-byte bank groups whatever the code bytes are — conflict-free by
//     construction.  The 16 code bytes are rotated once per lane so step t uses a compile-time byte
//     and the address (code<<8 | table<<4) is a single PRMT;
//   * the PQ codebook (m*k*sub fp32 = 128 KiB at d=128) stays in REGISTERS for the life of the
//     persistent CTA (8 codewords per thread) and the table is built with packed FFMA2 as
//     |p|^2 + |r_m|^2 - 2 r_m.p for the four residual queries at once;
//   * each list's codes are staged into shared memory with TMA bulk copies (cp.async.bulk +
//     mbarrier), double buffered, the next segment in flight while the current one is scanned;
//   * candidates pass a per-query global threshold (the k-th best of any finished item), go to a
//     per-slot queue, and at most k per item are appended to the query's pool in HBM;
//   * ivfpq_lm_final_kernel selects the best k of the pool and RE-SCORES them in the reference's exact
//     fp32 order (L2SquaredUnsafe per sub-vector, sequential sum over m), so reported distances never
//     come from the fused-multiply-add path.
// HBM traffic is one pass over the probed lists' codes (shared by all queries of the batch) instead
// of one pass per (query, probe); the kernel is bound by the 128 B/clk/SM shared-memory crossbar.
#include <cub/cub.cuh>

#include "common.cuh"
#include "exact_arith.cuh"
#include "kernels.h"

namespace pyrope {
namespace {

constexpr int LM_THREADS = 512;
constexpr int LM_QS = 4;            // query slots per work item
constexpr int LM_CODE_CAP = 1024;   // vectors per staged segment (16 KiB)
constexpr int LM_QC = 2048;         // candidate queue entries per slot (>= LM_CODE_CAP + kMaxTopK)
constexpr int LM_LUT_BYTES = 256 * 256;
constexpr int LM_SMEM = LM_LUT_BYTES + 2 * LM_CODE_CAP * 16 + LM_QS * LM_QC * 8;

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();  // a broken pipeline must fault, not hang the GPU
    }
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

struct LmParams {
    const float* Q; int64_t nq; int dim;
    const float* centroids; const float* codebook; int ksub;
    const uint8_t* codes; const uint8_t* dead; const int64_t* list_off; int nlist;
    const int2* items; const int32_t* n_items; const int32_t* pair_off; const int32_t* pairq;
    unsigned long long* pool; int32_t* pool_cnt; uint32_t* pool_thr; int pool_cap; int k;
    int32_t* work_ctr;
};

// ---- grouping (query, probe) pairs by list ---------------------------------------------------------
__global__ void lm_count_kernel(const int64_t* __restrict__ probes, int64_t npairs, const int64_t* __restrict__ list_off,
                                int32_t* lcnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int64_t l = probes[i];
    if (l >= 0 && list_off[l + 1] > list_off[l]) atomicAdd(&lcnt[l], 1);  // IvfPqVectorIndex.cs:155 skips empty lists
}
__global__ void lm_items_per_list_kernel(const int32_t* __restrict__ lcnt, int n, int32_t* nit) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nit[i] = (lcnt[i] + LM_QS - 1) / LM_QS;
}
__global__ void lm_fill_pairs_kernel(const int64_t* __restrict__ probes, int64_t npairs, int P,
                                     const int64_t* __restrict__ list_off, const int32_t* __restrict__ loff, int32_t* lcur,
                                     int32_t* pairq) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int64_t l = probes[i];
    if (l >= 0 && list_off[l + 1] > list_off[l]) pairq[loff[l] + atomicAdd(&lcur[l], 1)] = (int32_t)(i / P);
}
__global__ void lm_fill_items_kernel(const int32_t* __restrict__ nit, const int32_t* __restrict__ ioff, int nlist, int2* items) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    const int n = nit[l], o = ioff[l];
    for (int g = 0; g < n; ++g) items[o + g] = make_int2(l, g);
}

// ---- the scan --------------------------------------------------------------------------------------
template <int SUB>
__global__ void __launch_bounds__(LM_THREADS, 1) ivfpq_lm_scan_kernel(LmParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* lut = smem;                                                         // [256][16] float4
    unsigned char* cbuf = smem + LM_LUT_BYTES;                                         // [2][CODE_CAP] uint4
    uint64_t* qkeys = reinterpret_cast<uint64_t*>(cbuf + 2 * LM_CODE_CAP * 16);        // [QS][QC]
    __shared__ __align__(8) uint64_t s_mbar[2];
    __shared__ int s_qcnt[LM_QS];
    __shared__ int s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.ksub, dim = p.dim;
    const int m = lane & 15;                 // build: this thread's sub-quantiser
    const int eb = (lane >> 4) + 2 * warp;   // build: its codewords are eb + 32 j, j < 8
    const int n_items = *p.n_items;
    const uint32_t bar0 = smem_u32(&s_mbar[0]), bar1 = smem_u32(&s_mbar[1]);

    // codebook slice in registers for the CTA's lifetime
    float cb[8][SUB], pn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int e = eb + 32 * j;
        float s = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < SUB / 4; ++d4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < K) v = __ldg(reinterpret_cast<const float4*>(p.codebook + ((size_t)m * K + e) * SUB) + d4);
            cb[j][4 * d4 + 0] = v.x; cb[j][4 * d4 + 1] = v.y; cb[j][4 * d4 + 2] = v.z; cb[j][4 * d4 + 3] = v.w;
            s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
        }
        pn[j] = s;
    }
    // scan: lane reads table (lane + t) & 15 at step t; op[i] packs the table byte offsets of steps 2i, 2i+1
    uint32_t op[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        op[i] = (uint32_t)(((lane + 2 * i) & 15) << 4) | ((uint32_t)(((lane + 2 * i + 1) & 15) << 4) << 8);
    const int rot = lane & 15;

    auto seg_issue = [&](int item, int seg, int b) {  // thread 0 only
        const int2 it = __ldg(&p.items[item]);
        const int64_t beg = __ldg(p.list_off + it.x) + (int64_t)seg * LM_CODE_CAP;
        const int64_t end = __ldg(p.list_off + it.x + 1);
        const uint32_t bytes = (uint32_t)min((int64_t)LM_CODE_CAP, end - beg) * 16u;
        const uint32_t bar = b ? bar1 : bar0;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(cbuf + b * LM_CODE_CAP * 16), p.codes + beg * 16, bytes, bar);
    };
    // CTA-wide: sort slot j's queue, keep the best k; returns the kept count (all threads call)
    auto prune_slot = [&](int j) -> int {
        uint64_t* kq = qkeys + j * LM_QC;
        const int n = min(s_qcnt[j], LM_QC);
        const int P2 = next_pow2(max(n, 2));
        for (int i = n + tid; i < P2; i += LM_THREADS) kq[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc<false>(kq, P2, tid, LM_THREADS);
        const int keep = min(n, p.k);
        if (tid == 0) s_qcnt[j] = keep;
        __syncthreads();
        return keep;
    };

    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_next = atomicAdd(p.work_ctr, 1);
    }
    __syncthreads();
    int cur = s_next;
    __syncthreads();
    uint32_t ph0 = 0, ph1 = 0;
    int buf = 0;
    if (tid == 0 && cur < n_items) seg_issue(cur, 0, 0);

    while (cur < n_items) {
        const int2 it = __ldg(&p.items[cur]);
        const int l = it.x, g = it.y;
        const int64_t beg = __ldg(p.list_off + l), end = __ldg(p.list_off + l + 1);
        const int pbeg = __ldg(p.pair_off + l), pend = __ldg(p.pair_off + l + 1);
        int qid[LM_QS];
        float thrd[LM_QS];
#pragma unroll
        for (int j = 0; j < LM_QS; ++j) {
            const int idx = pbeg + LM_QS * g + j;
            qid[j] = idx < pend ? __ldg(p.pairq + idx) : -1;
            thrd[j] = -INFINITY;
            if (qid[j] >= 0) {
                const uint32_t u = __ldcg(p.pool_thr + qid[j]);
                thrd[j] = u ? -ord_to_score(u) : INFINITY;
            }
        }
        int nxt_local = 0;
        if (tid == 0) nxt_local = atomicAdd(p.work_ctr, 1);
        if (tid < LM_QS) s_qcnt[tid] = 0;

        // ---- lookup tables of the (up to) four residual queries: |p|^2 + |r_m|^2 - 2 r_m.p
        {
            float cen[SUB];
#pragma unroll
            for (int d4 = 0; d4 < SUB / 4; ++d4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.centroids + (size_t)l * dim + m * SUB) + d4);
                cen[4 * d4 + 0] = v.x; cen[4 * d4 + 1] = v.y; cen[4 * d4 + 2] = v.z; cen[4 * d4 + 3] = v.w;
            }
            float tt[LM_QS][SUB], rr[LM_QS];
#pragma unroll
            for (int j = 0; j < LM_QS; ++j) {
                rr[j] = 0.f;
#pragma unroll
                for (int d4 = 0; d4 < SUB / 4; ++d4) {
                    float4 v = make_float4(cen[4 * d4], cen[4 * d4 + 1], cen[4 * d4 + 2], cen[4 * d4 + 3]);
                    if (qid[j] >= 0) v = __ldg(reinterpret_cast<const float4*>(p.Q + (size_t)qid[j] * dim + m * SUB) + d4);
                    const float r0 = v.x - cen[4 * d4], r1 = v.y - cen[4 * d4 + 1], r2 = v.z - cen[4 * d4 + 2], r3 = v.w - cen[4 * d4 + 3];
                    rr[j] = fmaf(r0, r0, rr[j]); rr[j] = fmaf(r1, r1, rr[j]); rr[j] = fmaf(r2, r2, rr[j]); rr[j] = fmaf(r3, r3, rr[j]);
                    tt[j][4 * d4 + 0] = -2.f * r0; tt[j][4 * d4 + 1] = -2.f * r1; tt[j][4 * d4 + 2] = -2.f * r2; tt[j][4 * d4 + 3] = -2.f * r3;
                }
            }
            unsigned long long t01[SUB], t23[SUB];
#pragma unroll
            for (int d = 0; d < SUB; ++d) { t01[d] = pack2(tt[0][d], tt[1][d]); t23[d] = pack2(tt[2][d], tt[3][d]); }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                unsigned long long a01 = pack2(pn[j] + rr[0], pn[j] + rr[1]);
                unsigned long long a23 = pack2(pn[j] + rr[2], pn[j] + rr[3]);
#pragma unroll
                for (int d = 0; d < SUB; ++d) {
                    const unsigned long long c2 = pack2(cb[j][d], cb[j][d]);
                    a01 = ffma2(c2, t01[d], a01);
                    a23 = ffma2(c2, t23[d], a23);
                }
                *reinterpret_cast<ulonglong2*>(lut + (eb + 32 * j) * 256 + m * 16) = make_ulonglong2(a01, a23);
            }
        }
        if (tid == 0) s_next = nxt_local;
        __syncthreads();  // tables, counters and the next item index are visible
        const int nxt = s_next;

        const int nseg = (int)((end - beg + LM_CODE_CAP - 1) / LM_CODE_CAP);
        for (int sg = 0; sg < nseg; ++sg) {
            if (tid == 0) {  // keep the other buffer in flight: next segment of this list, else the next item
                if (sg + 1 < nseg) seg_issue(cur, sg + 1, buf ^ 1);
                else if (nxt < n_items) seg_issue(nxt, 0, buf ^ 1);
            }
            const int64_t sbeg = beg + (int64_t)sg * LM_CODE_CAP;
            const int nvec = (int)min((int64_t)LM_CODE_CAP, end - sbeg);
            if (sg > 0) {  // make room: a cold (no threshold yet) slot can take every vector of a segment
#pragma unroll
                for (int j = 0; j < LM_QS; ++j) {
                    if (s_qcnt[j] + nvec > LM_QC) {
                        const int keep = prune_slot(j);
                        if (keep == p.k) thrd[j] = fminf(thrd[j], -key_score(qkeys[j * LM_QC + p.k - 1]));
                    }
                }
            }
            if (buf == 0) { mbar_wait(bar0, ph0); ph0 ^= 1; } else { mbar_wait(bar1, ph1); ph1 ^= 1; }
            const unsigned char* cseg = cbuf + buf * LM_CODE_CAP * 16;
            for (int v = tid; v < nvec; v += LM_THREADS) {
                const uint4 cw = *reinterpret_cast<const uint4*>(cseg + v * 16);
                uint32_t w[4] = {cw.x, cw.y, cw.z, cw.w};
                {   // rotate the 16 code bytes left by `rot` positions: new byte t = old byte (t + rot) & 15
                    const bool r8 = rot & 8, r4 = rot & 4;
                    uint32_t a0 = r8 ? w[2] : w[0], a1 = r8 ? w[3] : w[1], a2 = r8 ? w[0] : w[2], a3 = r8 ? w[1] : w[3];
                    uint32_t b0 = r4 ? a1 : a0, b1 = r4 ? a2 : a1, b2 = r4 ? a3 : a2, b3 = r4 ? a0 : a3;
                    const int sh = (rot & 3) * 8;
                    w[0] = __funnelshift_r(b0, b1, sh); w[1] = __funnelshift_r(b1, b2, sh);
                    w[2] = __funnelshift_r(b2, b3, sh); w[3] = __funnelshift_r(b3, b0, sh);
                }
                unsigned long long acc01 = 0ull, acc23 = 0ull;
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    // byte0 = table offset, byte1 = code byte: address = code * 256 + table * 16
                    const uint32_t sel = 0x7600u | (uint32_t)((t & 3) << 4) | (uint32_t)(4 + (t & 1));
                    const uint32_t a = __byte_perm(w[t >> 2], op[t >> 1], sel);
                    const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(lut + a);
                    acc01 = fadd2(acc01, e.x);
                    acc23 = fadd2(acc23, e.y);
                }
                float d0, d1, d2, d3;
                unpack2(acc01, d0, d1);
                unpack2(acc23, d2, d3);
                if ((d0 < thrd[0]) | (d1 < thrd[1]) | (d2 < thrd[2]) | (d3 < thrd[3])) {
                    const int64_t gpos = sbeg + v;
                    if (!(p.dead && p.dead[gpos])) {
                        const float dd[LM_QS] = {d0, d1, d2, d3};
#pragma unroll
                        for (int j = 0; j < LM_QS; ++j) {
                            if (dd[j] < thrd[j]) {
                                const int pos = atomicAdd(&s_qcnt[j], 1);
                                if (pos < LM_QC) qkeys[j * LM_QC + pos] = make_key(-dd[j], (uint32_t)gpos);
                            }
                        }
                    }
                }
            }
            __syncthreads();  // the segment buffer may be refilled; queue counts are visible
            buf ^= 1;
        }

        // ---- hand at most k candidates per slot to the queries' pools
#pragma unroll
        for (int j = 0; j < LM_QS; ++j)
            if (s_qcnt[j] > p.k && s_qcnt[j] > 64) prune_slot(j);
        if (warp < LM_QS) {
            const int j = warp;
            const int n = min(s_qcnt[j], LM_QC);
            const int q = j == 0 ? qid[0] : j == 1 ? qid[1] : j == 2 ? qid[2] : qid[3];
            if (n > 0 && q >= 0) {
                const uint64_t* kq = qkeys + j * LM_QC;
                unsigned long long* dst = p.pool + (size_t)q * p.pool_cap;
                uint64_t mink = ~0ull;
                int kept;
                if (n > p.k) {  // k < n <= 64: select by rank counting inside the warp
                    const uint64_t a = lane < n ? kq[lane] : 0ull, b = lane + 32 < n ? kq[lane + 32] : 0ull;
                    int ra = 0, rb = 0;
                    for (int i = 0; i < n; ++i) {
                        const uint64_t x = kq[i];
                        ra += x > a;
                        rb += x > b;
                    }
                    const bool ka = lane < n && ra < p.k, kb = lane + 32 < n && rb < p.k;
                    const unsigned ma = __ballot_sync(0xffffffffu, ka), mb = __ballot_sync(0xffffffffu, kb);
                    kept = __popc(ma) + __popc(mb);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&p.pool_cnt[q], kept);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const unsigned below = (1u << lane) - 1u;
                    const int ia = base + __popc(ma & below), ib = base + __popc(ma) + __popc(mb & below);
                    if (ka) { if (ia < p.pool_cap) dst[ia] = a; mink = a; }
                    if (kb) { if (ib < p.pool_cap) dst[ib] = b; mink = b < mink ? b : mink; }
                } else {
                    kept = n;
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&p.pool_cnt[q], kept);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (int i = lane; i < n; i += 32) {
                        const uint64_t x = kq[i];
                        if (base + i < p.pool_cap) dst[base + i] = x;
                        mink = x < mink ? x : mink;
                    }
                }
                if (kept >= p.k) {  // this item alone proves k candidates at or above mink
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const uint64_t x = __shfl_xor_sync(0xffffffffu, mink, o);
                        mink = x < mink ? x : mink;
                    }
                    if (lane == 0) atomicMax(p.pool_thr + q, (uint32_t)(mink >> 32));
                }
            }
        }
        __syncthreads();
        cur = nxt;
    }
}

// ---- pool -> best k, exact re-score, final order ----------------------------------------------------
struct LmFinalParams {
    const float* Q; int dim;
    const float* centroids; const float* codebook; int ksub;
    const uint8_t* codes; const int64_t* list_off; int nlist; const int64_t* labels;
    const unsigned long long* pool; const int32_t* pool_cnt; int pool_cap; int k;
    PairOut out;
};

template <int SUB>
__global__ void __launch_bounds__(256) ivfpq_lm_final_kernel(LmFinalParams p, int P2max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [P2max]
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(p.pool_cnt[q], p.pool_cap);
    const int P2 = next_pow2(max(n, 2));
    const unsigned long long* src = p.pool + (size_t)q * p.pool_cap;
    for (int i = tid; i < P2; i += blockDim.x) keys[i] = i < n ? src[i] : 0ull;
    __syncthreads();
    bitonic_sort_desc<false>(keys, P2, tid, blockDim.x);
    const int kk = min(n, p.k);
    // exact re-score of the survivors: IvfPqVectorIndex.cs:161-166,182-186 in the reference's order
    for (int i = warp; i < kk; i += blockDim.x / 32) {
        const uint32_t pos = key_pos(keys[i]);
        int lo = 0, hi = p.nlist;  // list with list_off[l] <= pos < list_off[l+1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(p.list_off + mid) <= (int64_t)pos) lo = mid; else hi = mid;
        }
        float dm = 0.f;
        if (lane < 16) {
            float r[SUB];
#pragma unroll
            for (int d = 0; d < SUB; ++d)
                r[d] = __fsub_rn(__ldg(p.Q + q * p.dim + lane * SUB + d), __ldg(p.centroids + (size_t)lo * p.dim + lane * SUB + d));
            const int code = p.codes[(size_t)pos * 16 + lane];
            dm = exact::a1_l2_fixed<SUB>(r, p.codebook + ((size_t)lane * p.ksub + code) * SUB);
        }
        float dist = 0.f;
#pragma unroll
        for (int mi = 0; mi < 16; ++mi) dist = __fadd_rn(dist, __shfl_sync(0xffffffffu, dm, mi));
        __syncwarp();
        if (lane == 0) keys[i] = make_key(-dist, pos);
    }
    __syncthreads();
    const int P3 = next_pow2(max(kk, 2));
    for (int i = kk + tid; i < P3; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc<false>(keys, P3, tid, blockDim.x);
    const int64_t ob = (q * p.out.parts_total + p.out.part_base) * (int64_t)p.k;
    for (int i = tid; i < p.k; i += blockDim.x) {
        if (i < kk) {
            const uint64_t key = keys[i];
            p.out.scores[ob + i] = key_score(key);
            p.out.labels[ob + i] = p.labels[key_pos(key)];
        } else {
            p.out.scores[ob + i] = 0.f;
            p.out.labels[ob + i] = -1;
        }
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct LmLayout {
    size_t zero_bytes;  // leading region cleared per search
    size_t lcnt, lcur, pool_cnt, pool_thr, ctr, loff, nit, ioff, pairq, items, pool, temp, total;
    size_t temp_bytes;
    int64_t max_items;
};

LmLayout lm_layout(int64_t nq, int P, int k, int nlist) {
    LmLayout L{};
    const int64_t npairs = nq * P;
    size_t o = 0;
    L.lcnt = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.lcur = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.pool_cnt = o; o += align_up(sizeof(int32_t) * (size_t)nq, 256);
    L.pool_thr = o; o += align_up(sizeof(uint32_t) * (size_t)nq, 256);
    L.ctr = o; o += 256;
    L.zero_bytes = o;
    L.loff = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.nit = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.ioff = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.pairq = o; o += align_up(sizeof(int32_t) * (size_t)npairs, 256);
    L.max_items = npairs / LM_QS + std::min<int64_t>(npairs, nlist) + 1;
    L.items = o; o += align_up(sizeof(int2) * (size_t)L.max_items, 256);
    L.pool = o; o += align_up(sizeof(unsigned long long) * (size_t)nq * P * k, 256);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr, nlist + 1);
    L.temp_bytes = tb + 256;
    L.temp = o; o += align_up(L.temp_bytes, 256);
    L.total = o;
    return L;
}

template <int SUB>
cudaError_t launch_lm(const IvfPqScanParams& p, void* scratch, int num_sms, cudaStream_t st) {
    const int P = p.nprobe;
    const int64_t npairs = p.nq * P;
    const LmLayout L = lm_layout(p.nq, P, p.k, p.nlist);
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    int32_t* lcnt = reinterpret_cast<int32_t*>(base + L.lcnt);
    int32_t* lcur = reinterpret_cast<int32_t*>(base + L.lcur);
    int32_t* pool_cnt = reinterpret_cast<int32_t*>(base + L.pool_cnt);
    uint32_t* pool_thr = reinterpret_cast<uint32_t*>(base + L.pool_thr);
    int32_t* ctr = reinterpret_cast<int32_t*>(base + L.ctr);
    int32_t* loff = reinterpret_cast<int32_t*>(base + L.loff);
    int32_t* nit = reinterpret_cast<int32_t*>(base + L.nit);
    int32_t* ioff = reinterpret_cast<int32_t*>(base + L.ioff);
    int32_t* pairq = reinterpret_cast<int32_t*>(base + L.pairq);
    int2* items = reinterpret_cast<int2*>(base + L.items);
    unsigned long long* pool = reinterpret_cast<unsigned long long*>(base + L.pool);
    void* temp = base + L.temp;
    size_t tb = L.temp_bytes;

    cudaError_t e = cudaMemsetAsync(base, 0, L.zero_bytes, st);
    if (e != cudaSuccess) return e;
    const unsigned gb = (unsigned)((npairs + 255) / 256), lb = (unsigned)((p.nlist + 1 + 255) / 256);
    lm_count_kernel<<<gb, 256, 0, st>>>(p.probes, npairs, p.list_off, lcnt);
    lm_items_per_list_kernel<<<lb, 256, 0, st>>>(lcnt, p.nlist + 1, nit);
    e = cub::DeviceScan::ExclusiveSum(temp, tb, lcnt, loff, p.nlist + 1, st);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(temp, tb, nit, ioff, p.nlist + 1, st);
    if (e != cudaSuccess) return e;
    lm_fill_pairs_kernel<<<gb, 256, 0, st>>>(p.probes, npairs, P, p.list_off, loff, lcur, pairq);
    lm_fill_items_kernel<<<lb, 256, 0, st>>>(nit, ioff, p.nlist, items);

    LmParams sp{};
    sp.Q = p.Q; sp.nq = p.nq; sp.dim = p.dim; sp.centroids = p.centroids; sp.codebook = p.codebook; sp.ksub = p.ksub;
    sp.codes = p.codes; sp.dead = p.dead; sp.list_off = p.list_off; sp.nlist = p.nlist;
    sp.items = items; sp.n_items = ioff + p.nlist; sp.pair_off = loff; sp.pairq = pairq;
    sp.pool = pool; sp.pool_cnt = pool_cnt; sp.pool_thr = pool_thr; sp.pool_cap = P * p.k; sp.k = p.k;
    sp.work_ctr = ctr;
    e = cudaFuncSetAttribute(ivfpq_lm_scan_kernel<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM);
    if (e != cudaSuccess) return e;
    const int64_t grid = std::min<int64_t>(num_sms, L.max_items);
    ivfpq_lm_scan_kernel<SUB><<<(unsigned)grid, LM_THREADS, LM_SMEM, st>>>(sp);

    LmFinalParams fp{};
    fp.Q = p.Q; fp.dim = p.dim; fp.centroids = p.centroids; fp.codebook = p.codebook; fp.ksub = p.ksub;
    fp.codes = p.codes; fp.list_off = p.list_off; fp.nlist = p.nlist; fp.labels = p.labels;
    fp.pool = pool; fp.pool_cnt = pool_cnt; fp.pool_cap = P * p.k; fp.k = p.k; fp.out = p.out;
    const int P2max = next_pow2(std::max(2, P * p.k));
    const size_t fsm = sizeof(uint64_t) * (size_t)P2max;
    e = cudaFuncSetAttribute(ivfpq_lm_final_kernel<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    if (e != cudaSuccess) return e;
    ivfpq_lm_final_kernel<SUB><<<(unsigned)p.nq, 256, fsm, st>>>(fp, P2max);
    return cudaGetLastError();
}

}  // namespace

bool ivfpq_lm_supported(int dim, int m, int ksub, int nprobe, int k, int64_t nq, int64_t list_total) {
    if (m != 16 || ksub > 256 || ksub < 1) return false;
    const int sub = dim / m;
    if (sub != 4 && sub != 8) return false;
    if (k < 1 || k > kMaxTopK || (int64_t)nprobe * k > 8192) return false;
    if (nq * nprobe >= ((int64_t)1 << 31) || list_total >= ((int64_t)1 << 32)) return false;
    return true;
}

size_t ivfpq_lm_scratch_bytes(int64_t nq, int nprobe, int k, int nlist) { return lm_layout(nq, nprobe, k, nlist).total; }

int ivfpq_lm_launches() { return 8; }  // count, items-per-list, 2 scans, pair fill, item fill, scan, final

cudaError_t launch_ivfpq_scan_lm(const IvfPqScanParams& p, void* scratch, int num_sms, cudaStream_t st) {
    if (p.nq <= 0) return cudaSuccess;
    return (p.dim / p.m == 8) ? launch_lm<8>(p, scratch, num_sms, st) : launch_lm<4>(p, scratch, num_sms, st);
}

}  // namespace pyrope
