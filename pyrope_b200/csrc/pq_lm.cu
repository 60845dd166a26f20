// pq_lm.cu — K5 list-major IVF_PQ ADC scan (the batched hot path of BASELINE config 5).
//
// Replaces, for a whole query batch, ProductQuantizer.ComputeDistanceTable (ProductQuantizer.cs:98-120)
// and the ADC loop of IvfPqVectorIndex.Search (IvfPqVectorIndex.cs:152-199).  The reference walks
// query -> probed list -> code; a batch of 10^4 queries x 64 probes hits every inverted list ~10 times,
// so this path inverts the loop: (query, probe) pairs are grouped BY LIST and one work item is one
// inverted list x up to four of the queries that probe it.
//
// Pipeline (all on one stream, no host synchronisation):
//   lm_count / scans / lm_fill_pairs   group the pairs by list (counting sort on device);
//   lm_prepare_kernel                  one warp per item writes the item block: header (list, code range,
//                                      four query ids, their pool slots) + the four residual queries
//                                      -2(q - c) interleaved as float4 per dimension (slot d*16 + m);
//   ivfpq_lm_seed_kernel               per query, an upper bound of its k-th best ADC distance from (a
//                                      sample of) its nearest list, so no item starts without a threshold;
//   ivfpq_lm_scan_kernel               two 256-thread persistent CTAs per SM, static item striding.  Per item:
//       - the item block arrives by a TMA bulk copy (cp.async.bulk + mbarrier), one item ahead;
//       - the PQ codebook (m*k*sub fp32 = 128 KiB at d=128) is parked in TENSOR MEMORY for the CTA's lifetime
//         (tcgen05.st once, tcgen05.ld per build; 2 CTAs x 256 columns = the whole TMEM), which leaves the
//         register file free for two resident CTAs: one builds tables (FMA pipe) while the other scans
//         (shared-memory crossbar);
//       - the four lookup tables are built with packed FFMA2 as |p|^2 + |r_m|^2 - 2 r_m.p and stored
//         interleaved, LUT[e][m] = {q0,q1,q2,q3} at byte e*256 + m*16;
//       - scan: codes stream from L2/HBM as one coalesced 16-byte row per lane, prefetched a chunk ahead;
//         lane l reads table (l+t) mod 16 at step t, so the eight lanes of every quarter warp touch eight
//         distinct 16-byte bank groups whatever the code bytes are (conflict-free by construction); the 16
//         code bytes are rotated once per lane so step t uses a compile-time byte and the address
//         (code<<8 | table<<4) is one PRMT; one LDS.128 = four (query, code) lookups, summed with FADD2;
//       - candidates below the query's global threshold go to a small per-slot queue; after the item, one
//         warp per slot hands at most k of them to the pair's private pool region in HBM and tightens the
//         threshold (atomicMax).  A queue that overflows sends the (query, item) to the redo list;
//   ivfpq_lm_redo_kernel               plain per-(query, item) scan for the rare overflows;
//   ivfpq_lm_final_kernel              best k of the pool, RE-SCORED in the reference's exact fp32 order
//                                      (L2SquaredUnsafe per sub-vector, sequential sum over m), so
//                                      reported distances never come from the fused-multiply-add path.
// HBM traffic is one pass over the probed lists' codes (shared by the batch through L2) instead of one pass
// per (query, probe); the scan is bound by the 128 B/clk/SM shared-memory crossbar (ncu: 4 wavefronts per
// LDS.128, zero excess).
#include <cub/cub.cuh>

#include <cstdio>

#include "common.cuh"
#include "exact_arith.cuh"
#include "kernels.h"

namespace pyrope {
namespace {

constexpr int LM_THREADS = 256;       // two CTAs per SM
constexpr int LM_QS = 4;            // query slots per work item
constexpr int LM_QC = 512;          // candidate queue entries per slot
constexpr int LM_HDR = 64;          // item-block header bytes
constexpr int LM_MAX_DIM = 128;     // m = 16, sub <= 8
constexpr int LM_LUT_BYTES = 256 * 256;
constexpr int LM_BLK_MAX = LM_HDR + LM_MAX_DIM * 16;
constexpr int LM_SMEM = LM_LUT_BYTES + 2 * LM_BLK_MAX + LM_QS * LM_QC * 8;
constexpr int SEED_NQ = 2;          // queries per seed CTA (share the codebook reads; small tables -> 5 CTAs/SM)
constexpr int SEED_CAP = 1024;      // sampled distances per query
constexpr int REDO_QCAP = 2048;

struct __align__(16) LmHeader {
    int list, nvec;
    long long vbeg;
    int qid[LM_QS];
    int pslot[LM_QS];  // (probe rank * maxseg + segment): the pair's private slot in the query's pool
    int pad[4];
};
static_assert(sizeof(LmHeader) == LM_HDR, "header size");

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


#ifdef PYROPE_LM_TIMING
#define LM_T(i) do { long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } while (0)
#else
#define LM_T(i) do { } while (0)
#endif

// ---- TMEM as a constant table: the PQ codebook lives in tensor memory for the CTA's lifetime -------------
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[N]) {
    static_assert(N == 4 || N == 8, "4 or 8 columns");
    if (N == 8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4 % N]), "r"(r[5 % N]), "r"(r[6 % N]), "r"(r[7 % N])
                     : "memory");
    else
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                     : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]) {  // no wait: pair with tmem_ld_wait()
    if (N == 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4 % N]), "=r"(r[5 % N]), "=r"(r[6 % N]), "=r"(r[7 % N])
                     : "r"(taddr)
                     : "memory");
    else
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(taddr)
                     : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct LmParams {
    long long* timing;  // [grid][16 warps][8] cycle sums per phase (PYROPE_LM_TIMING builds only)
    int dim, ksub, k;
    const float* codebook; const uint8_t* codes; const uint8_t* dead;
    const unsigned char* iblk; const int32_t* n_items;
    unsigned long long* pool; int32_t* pool_cnt; uint32_t* pool_thr; int pslots;  // pool [nq][pslots][k], counts [nq][pslots]
    int2* redo; int32_t* redo_cnt;
};

// ---- grouping (query, probe) pairs by list ---------------------------------------------------------
__global__ void lm_count_kernel(const int64_t* __restrict__ probes, int64_t npairs, const int64_t* __restrict__ list_off,
                                int32_t* lcnt, unsigned long long* scanned) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long len = 0;
    if (i < npairs) {
        const int64_t l = probes[i];
        if (l >= 0) {
            len = (unsigned long long)(list_off[l + 1] - list_off[l]);
            if (len) atomicAdd(&lcnt[l], 1);  // IvfPqVectorIndex.cs:155 skips empty lists
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
    if ((threadIdx.x & 31) == 0 && len) atomicAdd(scanned, len);
}
__global__ void lm_items_per_list_kernel(const int32_t* __restrict__ lcnt, const int64_t* __restrict__ list_off, int nlist,
                                         int32_t* nit) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nlist) return;
    int v = 0;
    if (i < nlist) {
        v = (lcnt[i] + LM_QS - 1) / LM_QS;
    }
    nit[i] = v;
}
__global__ void lm_fill_pairs_kernel(const int64_t* __restrict__ probes, int64_t npairs, int P,
                                     const int64_t* __restrict__ list_off, const int32_t* __restrict__ loff, int32_t* lcur,
                                     int32_t* pairq, int32_t* pairp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int64_t l = probes[i];
    if (l >= 0 && list_off[l + 1] > list_off[l]) {
        const int slot = loff[l] + atomicAdd(&lcur[l], 1);
        pairq[slot] = (int32_t)(i / P);
        pairp[slot] = (int32_t)(i % P);
    }
}

// item -> list map: thread per list writes its (few) items
__global__ void lm_fill_items_kernel(const int32_t* __restrict__ nit, const int32_t* __restrict__ ioff, int nlist, int32_t* item_list) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    const int n = nit[l], o = ioff[l];
    for (int g = 0; g < n; ++g) item_list[o + g] = l;
}

// one warp per item: header + the four residual queries t = -2 (q - c), interleaved per dimension
struct LmPrep {
    const int32_t* ioff; const int32_t* item_list; const int32_t* loff; const int32_t* pairq; const int32_t* pairp;
    const int64_t* list_off; int nlist;
    int maxseg;
    const float* Q; const float* centroids; int dim;
    unsigned char* iblk; int blk;
};
__global__ void __launch_bounds__(256) lm_prepare_kernel(LmPrep a) {
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= a.ioff[a.nlist]) return;
    const int l = a.item_list[w], rel = w - a.ioff[l];
    const int64_t beg = a.list_off[l], len = a.list_off[l + 1] - beg;
    const int g = rel;
    const int pbeg = a.loff[l], pend = a.loff[l + 1];
    int qid[LM_QS], psl[LM_QS];
#pragma unroll
    for (int j = 0; j < LM_QS; ++j) {
        const int idx = pbeg + LM_QS * g + j;
        qid[j] = idx < pend ? a.pairq[idx] : -1;
        psl[j] = idx < pend ? a.pairp[idx] : 0;
    }
    unsigned char* blkp = a.iblk + (size_t)w * a.blk;
    if (lane == 0) {
        LmHeader h{};
        h.list = l;
        h.vbeg = beg;
        h.nvec = (int)len;
#pragma unroll
        for (int j = 0; j < LM_QS; ++j) { h.qid[j] = qid[j]; h.pslot[j] = psl[j]; }
        *reinterpret_cast<LmHeader*>(blkp) = h;
    }
    if (lane * 4 < a.dim) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(a.centroids + (size_t)l * a.dim) + lane);
        float4 t[LM_QS];
#pragma unroll
        for (int j = 0; j < LM_QS; ++j) {
            float4 q = c;
            if (qid[j] >= 0) q = __ldg(reinterpret_cast<const float4*>(a.Q + (size_t)qid[j] * a.dim) + lane);
            t[j] = make_float4(-2.f * (q.x - c.x), -2.f * (q.y - c.y), -2.f * (q.z - c.z), -2.f * (q.w - c.w));
        }
        // dimension D = mi*sub + d is stored at slot d*16 + mi, so the 16 sub-quantiser lanes of the
        // table build read 256 contiguous bytes per d (no bank conflicts)
        float4* dst = reinterpret_cast<float4*>(blkp + LM_HDR);
        const int sub = a.dim >> 4, D0 = lane * 4;
        dst[((D0 + 0) % sub) * 16 + (D0 + 0) / sub] = make_float4(t[0].x, t[1].x, t[2].x, t[3].x);
        dst[((D0 + 1) % sub) * 16 + (D0 + 1) / sub] = make_float4(t[0].y, t[1].y, t[2].y, t[3].y);
        dst[((D0 + 2) % sub) * 16 + (D0 + 2) / sub] = make_float4(t[0].z, t[1].z, t[2].z, t[3].z);
        dst[((D0 + 3) % sub) * 16 + (D0 + 3) / sub] = make_float4(t[0].w, t[1].w, t[2].w, t[3].w);
    }
}

// ---- plain ADC pieces shared by the seed and redo kernels ------------------------------------------------
// lut[j][m*256 + e] = |r_j,m - codeword(m,e)|^2 for NQ residual queries res[j][dim]; thread e <-> codeword
template <int NQ>
__device__ __forceinline__ void lut_direct(const float* __restrict__ codebook, int K, int sub, const float* res, int dim,
                                           float* lut, int tid, int nthr) {
    for (int e = tid; e < 256; e += nthr) {
#pragma unroll 4
        for (int mi = 0; mi < 16; ++mi) {
            float a[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) a[j] = 0.f;
            if (e < K) {
                const float4* cw = reinterpret_cast<const float4*>(codebook + ((size_t)mi * K + e) * sub);
                for (int d4 = 0; d4 < sub / 4; ++d4) {  // sub is 4 or 8 on this path
                    const float4 c = __ldg(cw + d4);
                    const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int j = 0; j < NQ; ++j) {
                            const float df = res[j * dim + mi * sub + d4 * 4 + u] - cc[u];
                            a[j] = fmaf(df, df, a[j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NQ; ++j) lut[j * 4096 + mi * 256 + e] = a[j];
        }
    }
}
__device__ __forceinline__ float adc_plain(const uint8_t* __restrict__ code16, const float* lut) {
    const uint4 cw = __ldg(reinterpret_cast<const uint4*>(code16));
    const uint32_t w[4] = {cw.x, cw.y, cw.z, cw.w};
    float d = 0.f;
#pragma unroll
    for (int mi = 0; mi < 16; ++mi) d += lut[mi * 256 + ((w[mi >> 2] >> (8 * (mi & 3))) & 0xffu)];
    return d;
}

// ---- seed: an upper bound of every query's k-th best ADC distance ---------------------------------------
struct LmSeed {
    const float* Q; int64_t nq; int dim; const int64_t* probes; int P;
    const float* centroids; const float* codebook; int ksub;
    const uint8_t* codes; const uint8_t* dead; const int64_t* list_off;
    uint32_t* pool_thr; int k; int sample;
};
__global__ void __launch_bounds__(256) ivfpq_lm_seed_kernel(LmSeed a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);      // [SEED_NQ][16*256]
    float* dist = lut + SEED_NQ * 4096;                   // [SEED_NQ][SEED_CAP]
    float* res = dist + SEED_NQ * SEED_CAP;               // [SEED_NQ][dim]
    __shared__ int s_cnt[SEED_NQ], s_pr[SEED_NQ];
    __shared__ int64_t s_list[SEED_NQ];
    __shared__ float s_rr[SEED_NQ];
    __shared__ int s_more;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * SEED_NQ;
    const int dim = a.dim, sub = dim / 16;
    if (tid < SEED_NQ) { s_cnt[tid] = 0; s_pr[tid] = 0; s_rr[tid] = 0.f; }
    __syncthreads();
    for (int round = 0; round < a.P; ++round) {
        // every query still short of k distances takes its next non-empty probed list
        if (tid < SEED_NQ) {
            const int64_t q = q0 + tid;
            int64_t l = -1;
            if (q < a.nq && s_cnt[tid] < a.k) {
                int pr = s_pr[tid];
                while (pr < a.P) {
                    const int64_t c = a.probes[q * a.P + pr++];
                    if (c >= 0 && a.list_off[c + 1] > a.list_off[c]) { l = c; break; }
                }
                s_pr[tid] = pr;
            }
            s_list[tid] = l;
        }
        if (tid == 0) s_more = 0;
        __syncthreads();
        bool any_list = false;
#pragma unroll
        for (int j = 0; j < SEED_NQ; ++j) any_list |= s_list[j] >= 0;
        if (!any_list) break;
        for (int i = tid; i < SEED_NQ * dim; i += 256) {
            const int j = i / dim, d = i - j * dim;
            const int64_t l = s_list[j];
            res[i] = l >= 0 ? a.Q[(q0 + j) * dim + d] - a.centroids[l * dim + d] : 0.f;
        }
        __syncthreads();
        if (warp < SEED_NQ && s_list[warp] >= 0) {  // |q - c|^2 scales the rounding margin below
            float s = 0.f;
            for (int d = lane; d < dim; d += 32) s = fmaf(res[warp * dim + d], res[warp * dim + d], s);
            s = warp_sum(s);
            if (lane == 0) s_rr[warp] = fmaxf(s_rr[warp], s);
        }
        lut_direct<SEED_NQ>(a.codebook, a.ksub, sub, res, dim, lut, tid, 256);
        __syncthreads();
        for (int j = 0; j < SEED_NQ; ++j) {
            const int64_t l = s_list[j];
            if (l < 0) continue;
            const int64_t beg = a.list_off[l];
            const int room = min(a.sample, SEED_CAP) - s_cnt[j];
            const int nv = (int)min((int64_t)max(room, 0), a.list_off[l + 1] - beg);
            __syncthreads();  // s_cnt[j] read by everyone before it moves
            for (int v = tid; v < nv; v += 256) {
                const int64_t pos = beg + v;
                if (a.dead && a.dead[pos]) continue;
                const float d = adc_plain(a.codes + pos * 16, lut + j * 4096);
                dist[j * SEED_CAP + atomicAdd(&s_cnt[j], 1)] = d;
            }
        }
        __syncthreads();
        if (tid < SEED_NQ && q0 + tid < a.nq && s_cnt[tid] < a.k && s_pr[tid] < a.P) s_more = 1;
        __syncthreads();
        if (!s_more) break;
    }
    __syncthreads();
    // one warp per query: smallest t (within a factor-2 count window) with count(dist <= t) >= k
    if (warp < SEED_NQ && q0 + warp < a.nq) {
        const int n = s_cnt[warp];
        if (n >= a.k) {
            const float* dj = dist + warp * SEED_CAP;
            uint32_t hi = 0;
            for (int i = lane; i < n; i += 32) hi = max(hi, __float_as_uint(fmaxf(dj[i], 0.f)));
            hi = __reduce_max_sync(0xffffffffu, hi);
            uint32_t lo = 0;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                int c = 0;
                for (int i = lane; i < n; i += 32) c += __float_as_uint(fmaxf(dj[i], 0.f)) <= mid;
                c = __reduce_add_sync(0xffffffffu, c);
                if (c >= a.k) { hi = mid; if (c <= 2 * a.k) break; } else lo = mid + 1;
            }
            if (lane == 0) {
                const float t = __uint_as_float(hi);
                // the scan kernel evaluates the same distances in |p|^2+|r|^2-2r.p form: cover its rounding
                const float tp = t + 2e-5f * t + 1e-5f * s_rr[warp] + 1e-12f;
                a.pool_thr[q0 + warp] = score_to_ord(-tp) - 1u;  // accept iff dist <= tp
            }
        }
    }
}

// ---- the scan --------------------------------------------------------------------------------------
// Two 256-thread CTAs per SM, each a plain build -> scan -> hand-over loop over its own items: while one
// CTA builds lookup tables (FMA pipe) the other scans (shared-memory crossbar), so the two phases overlap
// without any intra-CTA software pipelining.
template <int SUB>
__global__ void __launch_bounds__(LM_THREADS, 2) ivfpq_lm_scan_kernel(LmParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* lut = smem;                                                     // [256][16] float4
    unsigned char* rbuf = lut + LM_LUT_BYTES;                                       // [2] item blocks
    uint64_t* qkeys = reinterpret_cast<uint64_t*>(rbuf + 2 * LM_BLK_MAX);          // [QS][QC]
    __shared__ __align__(8) uint64_t s_mbar[2];
    __shared__ int s_qcnt[LM_QS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.ksub;
    const int blk = LM_HDR + p.dim * 16;
    const int m = lane & 15;                 // build: this thread's sub-quantiser
    const int eb = (lane >> 4) + 2 * warp;   // build: its codewords are eb + 16 j, j < 16
    const int n_items = *p.n_items;
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_n = first < n_items ? (n_items - first + stride - 1) / stride : 0;

    // The PQ codebook (m*k*sub fp32 = 128 KiB at d=128) is parked in TENSOR MEMORY for the CTA's lifetime:
    // 16 codewords per thread at the thread's own TMEM lane, columns (warp/4)*16*SUB + j*SUB.  Two CTAs per
    // SM x 256 columns fill the 512-column TMEM exactly; the register file stays free for two resident CTAs.
    constexpr int EPT = 256 * 16 / LM_THREADS;  // codewords per thread
    constexpr int TCOLS = (LM_THREADS / 128) * EPT * SUB;  // 256 (SUB = 8) or 128 (SUB = 4)
    __shared__ uint32_t s_tmem;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tcb = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * EPT * SUB);
    float pn[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        const int e = eb + (LM_THREADS / 16) * j;
        float s = 0.f;
        uint32_t r[SUB];
#pragma unroll
        for (int d4 = 0; d4 < SUB / 4; ++d4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < K) v = __ldg(reinterpret_cast<const float4*>(p.codebook + ((size_t)m * K + e) * SUB) + d4);
            r[4 * d4 + 0] = __float_as_uint(v.x); r[4 * d4 + 1] = __float_as_uint(v.y);
            r[4 * d4 + 2] = __float_as_uint(v.z); r[4 * d4 + 3] = __float_as_uint(v.w);
            s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
        }
        tmem_st<SUB>(tcb + j * SUB, r);
        pn[j] = s;
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // scan: lane reads table (lane + t) & 15 at step t; op[i] packs the table byte offsets of steps 2i, 2i+1
    uint32_t op[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        op[i] = (uint32_t)(((lane + 2 * i) & 15) << 4) | ((uint32_t)(((lane + 2 * i + 1) & 15) << 4) << 8);
        asm volatile("" : "+r"(op[i]));  // keep the eight offset words in registers (no rematerialisation in the loop)
    }
    const int rot = lane & 15;

    const uint32_t bar_r = smem_u32(&s_mbar[0]);  // +8*s: item-block stage s
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) mbar_init(bar_r + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < LM_QS) s_qcnt[tid] = 0;
    __syncthreads();

    auto hdr_ptr = [&](int i) { return p.iblk + (size_t)(first + (size_t)i * stride) * blk; };
    auto issue_block = [&](int i) {  // thread 0: TMA of item i's block (header + residual queries)
        const uint32_t br = bar_r + 8 * (i & 1);
        mbar_expect_tx(br, (uint32_t)blk);
        bulk_g2s(smem_u32(rbuf + (i & 1) * LM_BLK_MAX), hdr_ptr(i), (uint32_t)blk, br);
    };
    if (tid == 0 && my_n > 0) issue_block(0);

    // one warp per slot: hand at most k of the slot's candidates to the pair's private region of the query's
    // pool (plain stores: no returning atomics on this path) and tighten the query's threshold
    auto finalize = [&](const LmHeader* hd, int gitem) {
        const int j = warp;
        int* cntp = &s_qcnt[j];
        const int n = *cntp;
        const int q = hd->qid[j];
        if (n > 0 && q >= 0) {
            uint64_t* kq = qkeys + j * LM_QC;
            if (n > LM_QC) {  // candidates were dropped: the plain kernel redoes this (query, item)
                if (lane == 0) p.redo[atomicAdd(p.redo_cnt, 1)] = make_int2(q, gitem * LM_QS + j);
            } else {
                const size_t ps = (size_t)q * p.pslots + hd->pslot[j];
                unsigned long long* dst = p.pool + ps * p.k;
                uint64_t mink = ~0ull;
                int kept;
                if (n > p.k && n <= 64) {  // select by rank counting inside the warp
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
                    const unsigned below = (1u << lane) - 1u;
                    if (ka) { dst[__popc(ma & below)] = a; mink = a; }
                    if (kb) { dst[__popc(ma) + __popc(mb & below)] = b; mink = b < mink ? b : mink; }
                } else {
                    if (n > p.k) {  // 64 < n <= QC: warp-level sort, best first
                        const int P2 = next_pow2(n);
                        for (int i = n + lane; i < P2; i += 32) kq[i] = 0ull;
                        __syncwarp();
                        bitonic_sort_desc<true>(kq, P2, lane, 32);
                    }
                    kept = min(n, p.k);
                    for (int i = lane; i < kept; i += 32) {
                        const uint64_t x = kq[i];
                        dst[i] = x;
                        mink = x < mink ? x : mink;
                    }
                }
                if (lane == 0) p.pool_cnt[ps] = kept;
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
        __syncwarp();
        if (lane == 0) *cntp = 0;
    };

#ifdef PYROPE_LM_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    for (int i = 0; i < my_n; ++i) {
        const int rs = i & 1;
        // ---- build the four lookup tables of item i: |p|^2 + |r_m|^2 - 2 r_m.p
        mbar_wait(bar_r + 8 * rs, (uint32_t)(i >> 1) & 1u);
        LM_T(0);
        const unsigned char* blkp = rbuf + rs * LM_BLK_MAX;
        const LmHeader* hd = reinterpret_cast<const LmHeader*>(blkp);
        const int4 qv = *reinterpret_cast<const int4*>(hd->qid);
        const int nvec = hd->nvec;
        const long long vbeg = hd->vbeg;
        // codes stream straight from L2/HBM, one coalesced 16-byte row per lane, one chunk ahead; the first
        // chunk's load is in flight during the table build
        const uint4* cp = reinterpret_cast<const uint4*>(p.codes) + vbeg;
        uint4 cnext = make_uint4(0u, 0u, 0u, 0u);
        if (warp * 32 + lane < nvec) cnext = __ldg(cp + warp * 32 + lane);
        uint32_t tu[LM_QS];
        tu[0] = qv.x >= 0 ? __ldcg(p.pool_thr + qv.x) : 0xffffffffu;  // in flight during the build
        tu[1] = qv.y >= 0 ? __ldcg(p.pool_thr + qv.y) : 0xffffffffu;
        tu[2] = qv.z >= 0 ? __ldcg(p.pool_thr + qv.z) : 0xffffffffu;
        tu[3] = qv.w >= 0 ? __ldcg(p.pool_thr + qv.w) : 0xffffffffu;
        {
            const ulonglong2* rt = reinterpret_cast<const ulonglong2*>(blkp + LM_HDR) + m;  // slot d*16 + m
            unsigned long long t01[SUB], t23[SUB], rr01 = 0ull, rr23 = 0ull;
#pragma unroll
            for (int d = 0; d < SUB; ++d) {
                const ulonglong2 v = rt[d * 16];
                t01[d] = v.x; t23[d] = v.y;
                rr01 = ffma2(v.x, v.x, rr01);
                rr23 = ffma2(v.y, v.y, rr23);
            }
            const unsigned long long quarter = pack2(0.25f, 0.25f);  // t = -2 r  =>  |r|^2 = sum t^2 / 4
            unsigned char* lw = lut + eb * 256 + m * 16;
            uint32_t cw[2][SUB];  // codeword j+1 streams in from TMEM while codeword j is multiplied
            tmem_ld<SUB>(tcb, cw[0]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                if (j + 1 < EPT) tmem_ld<SUB>(tcb + (j + 1) * SUB, cw[(j + 1) & 1]);
                const unsigned long long pj = pack2(pn[j], pn[j]);
                unsigned long long a01 = ffma2(rr01, quarter, pj), a23 = ffma2(rr23, quarter, pj);
#pragma unroll
                for (int d = 0; d < SUB; ++d) {
                    const float c = __uint_as_float(cw[j & 1][d]);
                    const unsigned long long c2 = pack2(c, c);
                    a01 = ffma2(c2, t01[d], a01);
                    a23 = ffma2(c2, t23[d], a23);
                }
                *reinterpret_cast<ulonglong2*>(lw + j * ((LM_THREADS / 16) * 256)) = make_ulonglong2(a01, a23);
                if (j + 1 < EPT) tmem_ld_wait();
            }
        }
        LM_T(1);
        __syncthreads();  // tables complete; the previous item's hand-over is done
        LM_T(2);
        if (tid == 0 && i + 1 < my_n) issue_block(i + 1);

        // ---- scan item i
        float thrd[LM_QS];
#pragma unroll
        for (int j = 0; j < LM_QS; ++j)
            thrd[j] = tu[j] == 0xffffffffu ? -INFINITY : (tu[j] ? -ord_to_score(tu[j]) : INFINITY);
        LM_T(3);
        for (int c = warp; c * 32 < nvec; c += LM_THREADS / 32) {
            const int v = c * 32 + lane;
            const uint4 cw = cnext;
            if (v + LM_THREADS < nvec) cnext = __ldg(cp + v + LM_THREADS);
            if (v < nvec) {
                uint32_t w[4];
                {   // rotate the 16 code bytes: new byte t = old byte (t + rot) & 15
                    const bool r8 = rot & 8, r4 = rot & 4;
                    const uint32_t a0 = r8 ? cw.z : cw.x, a1 = r8 ? cw.w : cw.y, a2 = r8 ? cw.x : cw.z, a3 = r8 ? cw.y : cw.w;
                    const uint32_t b0 = r4 ? a1 : a0, b1 = r4 ? a2 : a1, b2 = r4 ? a3 : a2, b3 = r4 ? a0 : a3;
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
                    const long long gpos = vbeg + v;
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
        }
        LM_T(4);
        __syncthreads();  // tables fully consumed, queue counts visible
        LM_T(5);
        if (warp < LM_QS) finalize(hd, first + i * stride);
        LM_T(6);
    }
#ifdef PYROPE_LM_TIMING
    if (lane == 0 && p.timing)
        for (int u = 0; u < 8; ++u) p.timing[((size_t)blockIdx.x * (LM_THREADS / 32) + warp) * 8 + u] = tacc[u];
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TCOLS) : "memory");
    }
}

// ---- redo: plain scan of one (query, item) whose queue overflowed -------------------------------------------
struct LmRedo {
    const float* Q; int dim; const float* centroids; const float* codebook; int ksub;
    const uint8_t* codes; const uint8_t* dead;
    const unsigned char* iblk; int blk; const int2* redo; const int32_t* redo_cnt;
    unsigned long long* pool; int32_t* pool_cnt; int pslots; int k;
    const uint32_t* pool_thr;  // the query's current threshold: a valid bound, so the redo queue starts warm
};
__global__ void __launch_bounds__(256) ivfpq_lm_redo_kernel(LmRedo a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);   // [REDO_QCAP]
    float* lut = reinterpret_cast<float*>(keys + REDO_QCAP);  // [16*256]
    float* res = lut + 4096;                                  // [dim]
    __shared__ int s_cnt;
    __shared__ uint64_t s_thr;
    const int tid = threadIdx.x;
    const int n_redo = *a.redo_cnt;
    for (int e = blockIdx.x; e < n_redo; e += gridDim.x) {
        const int2 en = a.redo[e];
        const LmHeader* hd = reinterpret_cast<const LmHeader*>(a.iblk + (size_t)(en.y / LM_QS) * a.blk);
        const size_t ps = (size_t)en.x * a.pslots + hd->pslot[en.y % LM_QS];
        const int l = hd->list, nvec = hd->nvec;
        const long long vbeg = hd->vbeg;
        __syncthreads();
        for (int d = tid; d < a.dim; d += 256) res[d] = a.Q[(size_t)en.x * a.dim + d] - a.centroids[(size_t)l * a.dim + d];
        CtaQueue Qu{keys, &s_cnt, &s_thr, REDO_QCAP, a.k};
        Qu.reset(tid);
        __syncthreads();
        if (tid == 0) s_thr = (uint64_t)__ldcg(a.pool_thr + en.x) << 32;  // keys at or below it cannot be in the top k
        __syncthreads();
        lut_direct<1>(a.codebook, a.ksub, a.dim / 16, res, a.dim, lut, tid, 256);
        __syncthreads();
        for (int c0 = 0; c0 < nvec; c0 += 256) {
            if (s_cnt + 256 > REDO_QCAP) Qu.prune(tid, 256);  // s_cnt is stable here: every push is behind a barrier
            const int v = c0 + tid;
            if (v < nvec && !(a.dead && a.dead[vbeg + v])) {
                const float d = adc_plain(a.codes + (size_t)(vbeg + v) * 16, lut);
                Qu.push(make_key(-d, (uint32_t)(vbeg + v)));
            }
            __syncthreads();
        }
        Qu.prune(tid, 256);
        const int keep = s_cnt;
        if (tid == 0) a.pool_cnt[ps] = keep;
        for (int i = tid; i < keep; i += 256) a.pool[ps * a.k + i] = keys[i];
    }
}

// ---- pool -> best k, exact re-score, final order ----------------------------------------------------
struct LmFinalParams {
    const float* Q; int dim;
    const float* centroids; const float* codebook; int ksub;
    const uint8_t* codes; const int64_t* list_off; int nlist; const int64_t* labels;
    const int64_t* probes; int P;
    const unsigned long long* pool; const int32_t* pool_cnt; int pslots; int k;
    PairOut out;
};

template <int SUB>
__global__ void __launch_bounds__(256) ivfpq_lm_final_kernel(LmFinalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [next_pow2(pool_cap)]
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_n;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int sl = tid; sl < p.pslots; sl += blockDim.x) {  // gather the pairs' private regions
        const size_t ps = (size_t)q * p.pslots + sl;
        const int c = min(p.pool_cnt[ps], p.k);
        if (c > 0) {
            const int base = atomicAdd(&s_n, c);
            for (int i = 0; i < c; ++i) keys[base + i] = p.pool[ps * p.k + i];
        }
    }
    __syncthreads();
    const int n = s_n;
    const int kk = min(n, p.k);
    if (n > kk) {
        const int P2 = next_pow2(max(n, 2));
        for (int i = n + tid; i < P2; i += blockDim.x) keys[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc<false>(keys, P2, tid, blockDim.x);
    }
    // exact re-score of the survivors: IvfPqVectorIndex.cs:161-166,182-186 in the reference's order
    for (int i = warp; i < kk; i += blockDim.x / 32) {
        const uint32_t pos = key_pos(keys[i]);
        int lo = 0;  // the list holding pos is one of this query's probed lists: test them in parallel
        for (int p0 = 0; p0 < p.P; p0 += 32) {
            const int64_t l = p0 + lane < p.P ? __ldg(p.probes + q * p.P + p0 + lane) : -1;
            const bool hit = l >= 0 && __ldg(p.list_off + l) <= (int64_t)pos && (int64_t)pos < __ldg(p.list_off + l + 1);
            const unsigned mh = __ballot_sync(0xffffffffu, hit);
            if (mh) { lo = (int)__shfl_sync(0xffffffffu, l, __ffs(mh) - 1); break; }
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
inline int lm_maxseg(int64_t) { return 1; }  // one item covers a whole list (codes stream from L2/HBM, no staging buffer)

struct LmLayout {
    size_t zero_bytes;  // leading region cleared per search
    size_t lcnt, lcur, pool_cnt, pool_thr, redo_cnt, scanned, loff, nit, ioff, pairq, pairp, item_list, redo, iblk, pool, temp, total;
    size_t temp_bytes;
    int64_t max_items;
    int pool_cap, pslots, blk;
};

LmLayout lm_layout(int64_t nq, int P, int k, int nlist, int dim, int64_t max_list_len) {
    LmLayout L{};
    const int64_t npairs = nq * P;
    const int maxseg = lm_maxseg(max_list_len);
    size_t o = 0;
    L.lcnt = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.lcur = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.pool_cnt = o; o += align_up(sizeof(int32_t) * (size_t)nq * P * maxseg, 256);
    L.pool_thr = o; o += align_up(sizeof(uint32_t) * (size_t)nq, 256);
    L.redo_cnt = o; o += 256;
    L.scanned = o; o += 256;
    L.zero_bytes = o;
    L.loff = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.nit = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.ioff = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.pairq = o; o += align_up(sizeof(int32_t) * (size_t)npairs, 256);
    L.pairp = o; o += align_up(sizeof(int32_t) * (size_t)npairs, 256);
    L.max_items = (npairs / LM_QS + std::min<int64_t>(npairs, nlist) + 1) * maxseg;
    L.item_list = o; o += align_up(sizeof(int32_t) * (size_t)L.max_items, 256);
    L.redo = o; o += align_up(sizeof(int2) * (size_t)L.max_items * LM_QS, 256);
    L.blk = LM_HDR + dim * 16;
    L.iblk = o; o += align_up((size_t)L.blk * (size_t)L.max_items, 256);
    L.pslots = P * maxseg;
    L.pool_cap = L.pslots * k;
    L.pool = o; o += align_up(sizeof(unsigned long long) * (size_t)nq * L.pool_cap, 256);
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
    const LmLayout L = lm_layout(p.nq, P, p.k, p.nlist, p.dim, p.max_list_len);
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    int32_t* lcnt = reinterpret_cast<int32_t*>(base + L.lcnt);
    int32_t* lcur = reinterpret_cast<int32_t*>(base + L.lcur);
    int32_t* pool_cnt = reinterpret_cast<int32_t*>(base + L.pool_cnt);
    uint32_t* pool_thr = reinterpret_cast<uint32_t*>(base + L.pool_thr);
    int32_t* redo_cnt = reinterpret_cast<int32_t*>(base + L.redo_cnt);
    unsigned long long* scanned = reinterpret_cast<unsigned long long*>(base + L.scanned);
    int32_t* loff = reinterpret_cast<int32_t*>(base + L.loff);
    int32_t* nit = reinterpret_cast<int32_t*>(base + L.nit);
    int32_t* ioff = reinterpret_cast<int32_t*>(base + L.ioff);
    int32_t* pairq = reinterpret_cast<int32_t*>(base + L.pairq);
    int32_t* pairp = reinterpret_cast<int32_t*>(base + L.pairp);
    int32_t* item_list = reinterpret_cast<int32_t*>(base + L.item_list);
    int2* redo = reinterpret_cast<int2*>(base + L.redo);
    unsigned char* iblk = base + L.iblk;
    unsigned long long* pool = reinterpret_cast<unsigned long long*>(base + L.pool);
    void* temp = base + L.temp;
    size_t tb = L.temp_bytes;

    // PYROPE_LM_STAGES=1: print the duration of every kernel of this pipeline (debug aid, synchronises)
    static const bool stage_dbg = getenv("PYROPE_LM_STAGES") != nullptr;
    cudaEvent_t sev[8];
    int nsev = 0;
    auto mark = [&]() { if (stage_dbg && nsev < 8) { cudaEventCreate(&sev[nsev]); cudaEventRecord(sev[nsev], st); ++nsev; } };
    cudaError_t e = cudaMemsetAsync(base, 0, L.zero_bytes, st);
    if (e != cudaSuccess) return e;
    mark();
    // the seed kernel only needs the probe lists: it runs on the caller-provided side stream while the
    // pairs are grouped and the item blocks are written
    const bool fork = p.aux_stream && p.ev_fork && p.ev_join && !stage_dbg;
    cudaStream_t sst = fork ? p.aux_stream : st;
    if (fork) {
        cudaEventRecord(p.ev_fork, st);
        cudaStreamWaitEvent(p.aux_stream, p.ev_fork, 0);
        LmSeed sd{};
        sd.Q = p.Q; sd.nq = p.nq; sd.dim = p.dim; sd.probes = p.probes; sd.P = P; sd.centroids = p.centroids;
        sd.codebook = p.codebook; sd.ksub = p.ksub; sd.codes = p.codes; sd.dead = p.dead; sd.list_off = p.list_off;
        sd.pool_thr = pool_thr; sd.k = p.k; sd.sample = std::max(512, 16 * p.k);
        const size_t seed_smem = sizeof(float) * ((size_t)SEED_NQ * 4096 + (size_t)SEED_NQ * SEED_CAP + (size_t)SEED_NQ * p.dim);
        e = cudaFuncSetAttribute(ivfpq_lm_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem);
        if (e != cudaSuccess) return e;
        ivfpq_lm_seed_kernel<<<(unsigned)((p.nq + SEED_NQ - 1) / SEED_NQ), 256, seed_smem, sst>>>(sd);
        cudaEventRecord(p.ev_join, p.aux_stream);
    }
    const unsigned gb = (unsigned)((npairs + 255) / 256), lb = (unsigned)((p.nlist + 1 + 255) / 256);
    lm_count_kernel<<<gb, 256, 0, st>>>(p.probes, npairs, p.list_off, lcnt, scanned);
    lm_items_per_list_kernel<<<lb, 256, 0, st>>>(lcnt, p.list_off, p.nlist, nit);
    e = cub::DeviceScan::ExclusiveSum(temp, tb, lcnt, loff, p.nlist + 1, st);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(temp, tb, nit, ioff, p.nlist + 1, st);
    if (e != cudaSuccess) return e;
    lm_fill_pairs_kernel<<<gb, 256, 0, st>>>(p.probes, npairs, P, p.list_off, loff, lcur, pairq, pairp);

    LmPrep pa{};
    pa.maxseg = lm_maxseg(p.max_list_len); pa.pairp = pairp;
    lm_fill_items_kernel<<<lb, 256, 0, st>>>(nit, ioff, p.nlist, item_list);
    pa.item_list = item_list;
    pa.ioff = ioff; pa.loff = loff; pa.pairq = pairq; pa.list_off = p.list_off; pa.nlist = p.nlist;
    pa.Q = p.Q; pa.centroids = p.centroids; pa.dim = p.dim; pa.iblk = iblk; pa.blk = L.blk;
    mark();
    lm_prepare_kernel<<<(unsigned)((L.max_items * 32 + 255) / 256), 256, 0, st>>>(pa);
    mark();

    if (!fork) {
        LmSeed sd{};
        sd.Q = p.Q; sd.nq = p.nq; sd.dim = p.dim; sd.probes = p.probes; sd.P = P; sd.centroids = p.centroids;
        sd.codebook = p.codebook; sd.ksub = p.ksub; sd.codes = p.codes; sd.dead = p.dead; sd.list_off = p.list_off;
        sd.pool_thr = pool_thr; sd.k = p.k; sd.sample = std::max(512, 16 * p.k);
        const size_t seed_smem = sizeof(float) * ((size_t)SEED_NQ * 4096 + (size_t)SEED_NQ * SEED_CAP + (size_t)SEED_NQ * p.dim);
        e = cudaFuncSetAttribute(ivfpq_lm_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem);
        if (e != cudaSuccess) return e;
        ivfpq_lm_seed_kernel<<<(unsigned)((p.nq + SEED_NQ - 1) / SEED_NQ), 256, seed_smem, st>>>(sd);
    } else {
        cudaStreamWaitEvent(st, p.ev_join, 0);
    }
    mark();

    LmParams sp{};
#ifdef PYROPE_LM_TIMING
    static long long* d_timing = nullptr;
    if (!d_timing) cudaMalloc(&d_timing, sizeof(long long) * 512 * 16 * 8);
    sp.timing = d_timing;
#else
    sp.timing = nullptr;
#endif
    sp.dim = p.dim; sp.ksub = p.ksub; sp.k = p.k; sp.codebook = p.codebook; sp.codes = p.codes; sp.dead = p.dead;
    sp.iblk = iblk; sp.n_items = ioff + p.nlist;
    sp.pool = pool; sp.pool_cnt = pool_cnt; sp.pool_thr = pool_thr; sp.pslots = L.pslots;
    sp.redo = redo; sp.redo_cnt = redo_cnt;
    e = cudaFuncSetAttribute(ivfpq_lm_scan_kernel<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM);
    if (e != cudaSuccess) return e;
    const int64_t grid = std::min<int64_t>(2 * num_sms, L.max_items);
    if (p.ev_k0) cudaEventRecord(p.ev_k0, st);
    ivfpq_lm_scan_kernel<SUB><<<(unsigned)grid, LM_THREADS, LM_SMEM, st>>>(sp);
    if (p.ev_k1) cudaEventRecord(p.ev_k1, st);
#ifdef PYROPE_LM_TIMING
    {
        cudaStreamSynchronize(st);
        static long long h[512 * 16 * 8];
        cudaMemcpy(h, d_timing, sizeof(long long) * (size_t)grid * (LM_THREADS / 32) * 8, cudaMemcpyDeviceToHost);
        const char* names[8] = {"wait_blk", "build", "sync1", "wait_codes", "scan", "sync2", "finalize", "-"};
        for (int w = 0; w < LM_THREADS / 32; w += 1) {
            fprintf(stderr, "[lm timing] warp %2d:", w);
            for (int u = 0; u < 8; ++u) {
                double sum = 0;
                for (int b = 0; b < grid; ++b) sum += (double)h[((size_t)b * (LM_THREADS / 32) + w) * 8 + u];
                fprintf(stderr, " %s=%.0fk", names[u], sum / (double)grid / 1e3);
            }
            fprintf(stderr, "\n");
        }
    }
#endif

    LmRedo rd{};
    rd.Q = p.Q; rd.dim = p.dim; rd.centroids = p.centroids; rd.codebook = p.codebook; rd.ksub = p.ksub;
    rd.codes = p.codes; rd.dead = p.dead; rd.iblk = iblk; rd.blk = L.blk; rd.redo = redo; rd.redo_cnt = redo_cnt;
    rd.pool = pool; rd.pool_cnt = pool_cnt; rd.pslots = L.pslots; rd.k = p.k; rd.pool_thr = pool_thr;
    const size_t redo_smem = sizeof(uint64_t) * REDO_QCAP + sizeof(float) * (4096 + (size_t)p.dim);
    mark();
    ivfpq_lm_redo_kernel<<<(unsigned)(2 * num_sms), 256, redo_smem, st>>>(rd);
    mark();

    LmFinalParams fp{};
    fp.Q = p.Q; fp.dim = p.dim; fp.centroids = p.centroids; fp.codebook = p.codebook; fp.ksub = p.ksub;
    fp.codes = p.codes; fp.list_off = p.list_off; fp.nlist = p.nlist; fp.labels = p.labels;
    fp.probes = p.probes; fp.P = P;
    fp.pool = pool; fp.pool_cnt = pool_cnt; fp.pslots = L.pslots; fp.k = p.k; fp.out = p.out;
    const size_t fsm = sizeof(uint64_t) * (size_t)next_pow2(std::max(2, L.pool_cap));
    e = cudaFuncSetAttribute(ivfpq_lm_final_kernel<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    if (e != cudaSuccess) return e;
    ivfpq_lm_final_kernel<SUB><<<(unsigned)p.nq, (P * p.k <= 1024 ? 128 : 256), fsm, st>>>(fp);
    mark();
    if (stage_dbg && nsev == 7) {
        cudaEventSynchronize(sev[6]);
        const char* names[6] = {"group", "prepare", "seed", "scan", "redo", "final"};
        fprintf(stderr, "[lm stages]");
        for (int i = 0; i < 6; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, sev[i], sev[i + 1]);
            fprintf(stderr, " %s=%.3fms", names[i], ms);
        }
        int nredo = 0;
        cudaMemcpy(&nredo, redo_cnt, sizeof(int), cudaMemcpyDeviceToHost);
        int nit = 0;
        cudaMemcpy(&nit, ioff + p.nlist, sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, " items=%d redo=%d\n", nit, nredo);
        for (int i = 0; i < nsev; ++i) cudaEventDestroy(sev[i]);
    }
    return cudaGetLastError();
}

}  // namespace

bool ivfpq_lm_supported(int dim, int m, int ksub, int nprobe, int k, int64_t nq, int64_t list_total, int64_t max_list_len) {
    if (m != 16 || ksub > 256 || ksub < 1) return false;
    const int sub = dim / m;
    if (sub != 4 && sub != 8) return false;
    if (k < 1 || k > kMaxTopK || (int64_t)nprobe * k * lm_maxseg(max_list_len) > 16384) return false;
    if (nq * nprobe >= ((int64_t)1 << 29) || list_total >= ((int64_t)1 << 32)) return false;
    return true;
}

size_t ivfpq_lm_scratch_bytes(int64_t nq, int nprobe, int k, int nlist, int dim, int64_t max_list_len) {
    return lm_layout(nq, nprobe, k, nlist, dim, max_list_len).total;
}

// count, items-per-list, 2 scans, pair fill, prepare, seed, scan, redo, final
int ivfpq_lm_launches() { return 10; }

// codes scored by the most recent list-major search that used `scratch` (sum of probed list lengths)
cudaError_t ivfpq_lm_scanned_codes(const void* scratch, int64_t nq, int nprobe, int k, int nlist, int dim,
                                   int64_t max_list_len, unsigned long long* out, cudaStream_t st) {
    const LmLayout L = lm_layout(nq, nprobe, k, nlist, dim, max_list_len);
    cudaError_t e = cudaMemcpyAsync(out, reinterpret_cast<const unsigned char*>(scratch) + L.scanned, sizeof(*out),
                                    cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

cudaError_t launch_ivfpq_scan_lm(const IvfPqScanParams& p, void* scratch, int num_sms, cudaStream_t st) {
    if (p.nq <= 0) return cudaSuccess;
    return (p.dim / p.m == 8) ? launch_lm<8>(p, scratch, num_sms, st) : launch_lm<4>(p, scratch, num_sms, st);
}

}  // namespace pyrope
