// pq_lm.cu — K5 list-major IVF_PQ ADC scan (the batched hot path of BASELINE config 5).
//
// Replaces, for a whole query batch, ProductQuantizer.ComputeDistanceTable (ProductQuantizer.cs:98-120)
// and the ADC loop of IvfPqVectorIndex.Search (IvfPqVectorIndex.cs:152-199).  The reference walks
// query -> probed list -> code; a batch of 10^4 queries x 64 probes hits every inverted list ~10 times,
// so this path inverts the loop: (query, probe) pairs are grouped BY LIST and one work item is one
// inverted list x up to EIGHT of the queries that probe it.
//
// Pipeline (all on one stream, no host synchronisation):
//   lm_cmax_kernel                     max codeword norm per sub-quantiser (bound of every table value);
//   lm_count / scans / lm_fill_pairs   group the pairs by list (counting sort on device);
//   lm_prepare_kernel                  one warp per item writes the item block: header (list, code range,
//                                      eight query ids, their pool slots, their fixed-point scales) + the eight
//                                      residual queries -2(q - c), two float4-interleaved halves (slot d*16 + m);
//   ivfpq_lm_seed_kernel               per query, an upper bound of its k-th best ADC distance from (a
//                                      sample of) its nearest list, so no item starts without a threshold;
//   ivfpq_lm_scan_kernel               one 512-thread persistent CTA per SM, warp-specialised (see the kernel):
//       - builder warps: item block by TMA bulk copy (cp.async.bulk + mbarrier), the item's codes pulled into L2
//         by a bulk prefetch, PQ codebook parked in TENSOR MEMORY (tcgen05.st once, tcgen05.ld per build), tables
//         |p|^2 + |r_m|^2 - 2 r_m.p built with packed FFMA2 and stored as 16-bit fixed point, eight queries per
//         16 bytes: LUT[e][m] = {q0..q7} at byte e*256 + m*16, double buffered;
//       - scan warps: codes stream from L2 as one coalesced 16-byte row per lane, four chunks in flight; lane l
//         reads table (l+t) mod 16 at step t, so the eight lanes of every quarter warp touch eight distinct
//         16-byte bank groups whatever the code bytes are (conflict-free by construction); the 16 code bytes are
//         rotated once per lane so step t uses a compile-time byte and the address (code<<8 | table<<4) is one
//         PRMT; one LDS.128 = eight (query, code) lookups, summed as exact integers, two queries per IADD;
//       - candidates within the query's threshold go to a per-slot queue; the builder warps hand each finished
//         item's queues to the pairs' private pool regions in HBM and tighten the thresholds (per pair, and
//         across pairs through a per-query histogram of candidate distances).  A queue or region that overflows
//         sends the (query, item) to the redo list;
//   ivfpq_lm_redo_kernel               plain per-(query, item) scan for the rare overflows;
//   ivfpq_lm_final_kernel              everything in the pool within the rounding band of the k-th best is
//                                      RE-SCORED in the reference's exact fp32 order (L2SquaredUnsafe per
//                                      sub-vector, sequential sum over m), then the best k are taken: reported
//                                      distances never come from the fixed-point path.
// HBM traffic is one pass over the probed lists' codes (shared by the batch through L2) instead of one pass
// per (query, probe); the scan itself is bound by the 128 B/clk/SM shared-memory crossbar (4 wavefronts per
// LDS.128, zero excess) and the builders by the FMA pipe.
#include <cub/cub.cuh>

#include <cstdio>
#include <type_traits>

#include "common.cuh"
#include "exact_arith.cuh"
#include "kernels.h"

namespace pyrope {
namespace {

constexpr int LM_SCAN_WARPS = 8;    // warps 0-7 scan
constexpr int LM_BUILD_WARPS = 8;   // the last 8 warps build the next item's tables and hand finished items over (one per query slot)
constexpr int LM_THREADS = 32 * (LM_SCAN_WARPS + LM_BUILD_WARPS);  // one persistent CTA per SM
constexpr int LM_BLK_STAGES = 6;    // item blocks in flight: arriving (i+1), built from (i), scanned (i-1), handed over (i-2) — and two more,
                                    // because a stage may only be reused once EVERY builder warp has left its hand-over, which the
                                    // pipeline proves two rounds later (wait_done(i+1) at the top of round i+3, which claims block i+4)
constexpr int LM_QS = 8;            // query slots per work item (two halves of four)
constexpr int LM_QC = 256;          // candidate queue entries per slot (two sets: items alternate)
constexpr int LM_PF = 4;            // code chunks (256 rows each) a scan warp keeps in flight (even: two per iteration)
constexpr int LM_HDR = 96;          // item-block header bytes
constexpr int LM_MAX_DIM = 128;     // m = 16, sub <= 8
constexpr int LM_LUT_BYTES = 256 * 256;  // [256 codes][16 tables][8 queries] u16
constexpr int LM_BLK_MAX = LM_HDR + LM_MAX_DIM * 32;
constexpr int LM_QSETS = 3;         // candidate-queue sets: item i pushes into set i % 3 while item i-2's set is still being handed over
constexpr int LM_SMEM = 2 * LM_LUT_BYTES + LM_BLK_STAGES * LM_BLK_MAX + LM_QSETS * LM_QS * LM_QC * 8;
// Fixed-point lookup tables: entry = round(T * s) with s = LM_QMAX / B, B >= every table value of that (query,
// item); 16 entries sum to < 2^15, so two queries share one 32-bit add and bit 15 is free for the threshold test.
constexpr float LM_QMAX = 2046.f;
// |sum of 16 rounded entries - s * sum of the fp32 entries| <= 16 * 0.5 (+ the product's own rounding)
constexpr float LM_QERR = 8.25f;
// Threshold tightening: every query keeps a histogram of its candidates' distances over [thr0/2, thr0] (thr0 = the
// seed bound); the upper edge of the bucket where the running count reaches k bounds the k-th best distance.
constexpr int LM_HB = 64;
constexpr int SEED_NQ = 2;          // queries per seed CTA (share the codebook reads; small tables -> 5 CTAs/SM)
constexpr int SEED_CAP = 1024;      // sampled distances per query
constexpr int REDO_QCAP = 2048;

struct __align__(16) LmHeader {
    int list, nvec;
    long long vbeg;
    int qid[LM_QS];
    short pslot[LM_QS];  // (probe rank * maxseg + segment): the pair's private slot in the query's pool
    float s[LM_QS];      // fixed-point scale of the slot's lookup tables (0: slot unused)
};
static_assert(LM_BUILD_WARPS == LM_QS && LM_SCAN_WARPS % 4 == 0, "one builder warp per query slot; builders start a TMEM lane quadrant");
static_assert(sizeof(LmHeader) == LM_HDR, "header size");

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of a converged warp (see flat_tc.cu: under `lane == 0` ptxas serialises every uniform-datapath instruction)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// ask L2 to fetch a byte range from HBM (no destination, no completion: SASS UBLKPF)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}



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
    int dim, ksub, k;
    const float* codebook; const uint8_t* codes; const uint8_t* dead;
    const unsigned char* iblk; const int32_t* n_items;
    unsigned long long* pool; int32_t* pool_cnt; uint32_t* pool_thr; int pslots;  // pool [nq][pslots][kc], counts [nq][pslots]
    int kc;              // pool entries per (query, probe) pair: k plus room for candidates tied within the rounding band
    uint32_t* hist;            // [nq][LM_HB] candidates per distance bucket (zero-initialised)
    const float* thr0;         // [nq] seed bound the buckets are laid over (0: query was not seeded)
    const uint32_t* sinv_max;  // [nq] max 1/scale over the query's items (float bits)
    int2* redo; int32_t* redo_cnt;
    int32_t* item_ctr;   // next unclaimed work item (zero-initialised): CTAs claim items as they go
    // multi-GPU: bounds published by the peer ranks (nullable) and the peers' arrays (NVLink peer memory).  A word is
    // (batch epoch << 32 | ordered bound): a reader only believes a word of ITS batch, so a slow peer's late write for
    // an earlier batch is inert whenever it lands, and atomicMax lets a newer batch's word replace an older one
    const unsigned long long* thr_pub;
    unsigned long long* peer_thr[7];
    int n_peers;
    uint32_t epoch;
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
// (items, pairs) of a list packed into one 64-bit word, so ONE exclusive scan yields both offsets: pair offset in the
// low half, item offset in the high half (sums stay far below 2^32: at most nq * nprobe pairs)
struct LmPackOp {
    const int32_t* lcnt; int nlist;
    __host__ __device__ unsigned long long operator()(int i) const {
        const unsigned c = i < nlist ? (unsigned)lcnt[i] : 0u;
        return ((unsigned long long)((c + LM_QS - 1) / LM_QS) << 32) | c;
    }
};
// one launch after the scan: unpack the offsets (loff / ioff, [nlist + 1]), write the item -> list map, scatter the pairs
__global__ void lm_fill_kernel(const int64_t* __restrict__ probes, int64_t npairs, int P, const int64_t* __restrict__ list_off,
                               const unsigned long long* __restrict__ poff, int nlist, int32_t* loff, int32_t* ioff, int32_t* lcur,
                               int32_t* pairq, int32_t* pairp, int32_t* item_list) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nlist) {
        const unsigned long long w = poff[i];
        loff[i] = (int32_t)(uint32_t)w;
        ioff[i] = (int32_t)(w >> 32);
        if (i < nlist) {
            const int n = (int)(poff[i + 1] >> 32) - (int)(w >> 32), o = (int)(w >> 32);
            for (int g = 0; g < n; ++g) item_list[o + g] = (int)i;
        }
    }
    if (i < npairs) {
        const int64_t l = probes[i];
        if (l >= 0 && list_off[l + 1] > list_off[l]) {
            const int slot = (int)(uint32_t)poff[l] + atomicAdd(&lcur[l], 1);
            pairq[slot] = (int32_t)(i / P);
            pairp[slot] = (int32_t)(i % P);
        }
    }
}

// one warp per item: header + the four residual queries t = -2 (q - c), interleaved per dimension
struct LmPrep {
    const int32_t* ioff; const int32_t* item_list; const int32_t* loff; const int32_t* pairq; const int32_t* pairp;
    const int64_t* list_off; int nlist;
    int maxseg;
    const float* Q; const float* centroids; int dim;
    unsigned char* iblk; int blk;
    const float* cmax;     // [16] max codeword norm per sub-quantiser
    uint32_t* sinv_max;    // [nq] max over the query's items of 1 / scale, as float bits (zero-initialised)
};
// max_e |codeword(m, e)| per sub-quantiser (slightly rounded up): the table bound of lm_prepare_kernel
__global__ void __launch_bounds__(256) lm_cmax_kernel(const float* __restrict__ codebook, int K, int sub, float* cmax) {
    __shared__ uint32_t s_mx[16];
    const int e = threadIdx.x;
    if (e < 16) s_mx[e] = 0u;
    __syncthreads();
    if (e < K) {
        for (int mi = 0; mi < 16; ++mi) {
            float a = 0.f;
            for (int d = 0; d < sub; ++d) { const float c = codebook[((size_t)mi * K + e) * sub + d]; a = fmaf(c, c, a); }
            atomicMax(&s_mx[mi], __float_as_uint(sqrtf(a) * 1.000001f));  // non-negative floats order like their bits
        }
    }
    __syncthreads();
    if (e < 16) cmax[e] = __uint_as_float(s_mx[e]);
}
__global__ void __launch_bounds__(256) lm_prepare_kernel(LmPrep a) {
    const int lane = threadIdx.x & 31;
    const int n_items = a.ioff[a.nlist];
    const int wstride = (int)((gridDim.x * blockDim.x) >> 5);
    // a resident grid walks the items (one warp per item at a time): 100 k short-lived blocks kept the SMs a third full
    for (int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < n_items; w += wstride) {
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
        const int sub = a.dim >> 4, D0 = lane * 4;
        const bool on = lane * 4 < a.dim;
        const float cm = on ? __ldg(a.cmax + D0 / sub) : 0.f;  // the lane's four dimensions lie in one sub-vector
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) c = __ldg(reinterpret_cast<const float4*>(a.centroids + (size_t)l * a.dim) + lane);
        float scale[LM_QS];
    #pragma unroll
        for (int h = 0; h < LM_QS / 4; ++h) {  // queries 4h .. 4h+3 form one float4-interleaved half
            float4 t[4];
    #pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 q = c;
                if (on && qid[4 * h + j] >= 0) q = __ldg(reinterpret_cast<const float4*>(a.Q + (size_t)qid[4 * h + j] * a.dim) + lane);
                t[j] = make_float4(-2.f * (q.x - c.x), -2.f * (q.y - c.y), -2.f * (q.z - c.z), -2.f * (q.w - c.w));
                // fixed-point scale: every table value |r_m - p|^2 is at most B = max_m (|r_m| + max_e |p_m,e|)^2
                float r2 = 0.25f * (t[j].x * t[j].x + t[j].y * t[j].y + t[j].z * t[j].z + t[j].w * t[j].w);
                if (sub == 8) r2 += __shfl_xor_sync(0xffffffffu, r2, 1);
                float b = on ? sqrtf(r2) * 1.000001f + cm : 0.f;
                b *= b;
    #pragma unroll
                for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
                scale[4 * h + j] = qid[4 * h + j] >= 0 ? LM_QMAX / (b * 1.00001f + 1e-30f) : 0.f;
            }
            if (on) {
                // dimension D = mi*sub + d is stored at slot d*16 + mi, so the 16 sub-quantiser lanes of the
                // table build read 256 contiguous bytes per d (no bank conflicts)
                float4* dst = reinterpret_cast<float4*>(blkp + LM_HDR + (size_t)h * a.dim * 16);
                dst[((D0 + 0) % sub) * 16 + (D0 + 0) / sub] = make_float4(t[0].x, t[1].x, t[2].x, t[3].x);
                dst[((D0 + 1) % sub) * 16 + (D0 + 1) / sub] = make_float4(t[0].y, t[1].y, t[2].y, t[3].y);
                dst[((D0 + 2) % sub) * 16 + (D0 + 2) / sub] = make_float4(t[0].z, t[1].z, t[2].z, t[3].z);
                dst[((D0 + 3) % sub) * 16 + (D0 + 3) / sub] = make_float4(t[0].w, t[1].w, t[2].w, t[3].w);
            }
        }
        if (lane == 0) {
            LmHeader h{};
            h.list = l;
            h.vbeg = beg;
            h.nvec = (int)len;
    #pragma unroll
            for (int j = 0; j < LM_QS; ++j) { h.qid[j] = qid[j]; h.pslot[j] = (short)psl[j]; h.s[j] = scale[j]; }
            *reinterpret_cast<LmHeader*>(blkp) = h;
        }
        if (lane < LM_QS) {
            float mys = 0.f;
            int myq = -1;
    #pragma unroll
            for (int j = 0; j < LM_QS; ++j) { mys = lane == j ? scale[j] : mys; myq = lane == j ? qid[j] : myq; }
            if (myq >= 0) atomicMax(a.sinv_max + myq, __float_as_uint(1.f / mys));  // positive floats order like their bits
        }
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
    float* thr0;  // [nq] out: the bound as a distance (histogram range of the scan's threshold tightening)
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
                a.thr0[q0 + warp] = tp;
            }
        }
    }
}

// ---- the scan --------------------------------------------------------------------------------------
// One 512-thread persistent CTA per SM, warp-specialised: warps 8-15 BUILD the lookup tables of item i+1 (FMA pipe,
// codewords from tensor memory) into one half of a double-buffered table while warps 0-7 SCAN item i from the
// other half (shared-memory crossbar).  Hand-over is by mbarriers only: block-arrived (TMA), table-full (builders
// -> scanners), item-done (scanners -> builders); the two groups never meet at a CTA-wide barrier inside the loop.
//
// Tables are FIXED POINT: the fp32 value T = |p|^2 + |r_m|^2 - 2 r_m.p of (query j, table m, code e) is stored as
// the 16-bit integer round(T * s_j), s_j = LM_QMAX / B_j with B_j = max_m (|r_j,m| + max_e |p_m,e|)^2 >= T (triangle
// inequality; lm_prepare_kernel puts s_j in the item header), so one LDS.128 serves EIGHT queries and a (code, query)
// sum is an exact integer below 2^15: two queries share one 32-bit IADD, and `(0x8000 | threshold) - sum` keeps bit 15
// set exactly when sum <= threshold.  The integer sum differs from s_j x (the fp32 sum) by at most LM_QERR, which
// every threshold below allows for; reported distances never come from this path (the final kernel re-scores in
// the reference's order).
template <int SUB>
__global__ void __launch_bounds__(LM_THREADS, 1) ivfpq_lm_scan_kernel(LmParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* lut0 = smem;                                                    // [2][256][16] x 8 u16
    unsigned char* rbuf = lut0 + 2 * LM_LUT_BYTES;                                  // [3] item blocks
    uint64_t* qkeys = reinterpret_cast<uint64_t*>(rbuf + LM_BLK_STAGES * LM_BLK_MAX);  // [QSETS][QS][QC]
    __shared__ __align__(8) uint64_t s_mbar[LM_BLK_STAGES + 4];
    __shared__ int s_qcnt[LM_QSETS * LM_QS];
    __shared__ int s_ti[LM_SCAN_WARPS * LM_QS];     // per scan warp: integer thresholds of its current item (-1: slot unused)
    __shared__ float s_inv[LM_SCAN_WARPS * LM_QS];  // per scan warp: 1 / s_j
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool builder = warp >= LM_SCAN_WARPS;
    const int K = p.ksub;
    const int blk = LM_HDR + p.dim * 32;
    const int n_items = *p.n_items;
    // Items are claimed dynamically (one atomicAdd per item, two items ahead of the scan): lists differ a lot in
    // length, and with few items per CTA a static assignment leaves SMs idle at the end.  s_item[i % stages] holds
    // the global index of this CTA's i-th item, -1 once the counter ran past the end.
    __shared__ int s_item[LM_BLK_STAGES];

    const uint32_t bar_blk = smem_u32(&s_mbar[0]);                      // +8*s: item block stage s arrived (TMA)
    const uint32_t bar_full = smem_u32(&s_mbar[LM_BLK_STAGES]);         // +8*b: table half b built
    const uint32_t bar_done = smem_u32(&s_mbar[LM_BLK_STAGES + 2]);     // +8*b: every scan warp has left the item in half b
    if (tid == 0) {
        for (int i = 0; i < LM_BLK_STAGES; ++i) mbar_init(bar_blk + 8 * i, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_full + 8 * i, LM_BUILD_WARPS); mbar_init(bar_done + 8 * i, LM_SCAN_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < LM_QSETS * LM_QS) s_qcnt[tid] = 0;
    for (int i = tid; i < 2 * LM_LUT_BYTES / 16; i += LM_THREADS) reinterpret_cast<uint4*>(lut0)[i] = make_uint4(0u, 0u, 0u, 0u);

    // The PQ codebook (m*k*sub fp32 = 128 KiB at d=128) is parked in TENSOR MEMORY for the CTA's lifetime:
    // 16 codewords per builder thread at the thread's own TMEM lane, columns (bw/4)*16*SUB + j*SUB.
    constexpr int BT = LM_BUILD_WARPS * 32;     // builder threads
    constexpr int EPT = 256 * 16 / BT;          // codewords per builder thread
    constexpr int TCOLS = (BT / 128) * EPT * SUB;  // 256 (SUB = 8) or 128 (SUB = 4)
    if (warp == LM_SCAN_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // one thread: claim this CTA's i-th item and start the TMA of its block (header + residual queries); past the
    // end, publish -1 and complete the barrier phase by hand so that every waiter wakes up and leaves
    auto claim_block = [&](int i) -> bool {
        const uint32_t br = bar_blk + 8 * (i % LM_BLK_STAGES);
        const int g = atomicAdd(p.item_ctr, 1);
        if (g < n_items) {
            s_item[i % LM_BLK_STAGES] = g;
            mbar_expect_tx(br, (uint32_t)blk);  // release: the index above is visible to whoever sees the phase complete
            bulk_g2s(smem_u32(rbuf + (i % LM_BLK_STAGES) * LM_BLK_MAX), p.iblk + (size_t)g * blk, (uint32_t)blk, br);
            return true;
        }
        s_item[i % LM_BLK_STAGES] = -1;
        mbar_arrive(br);
        return false;
    };

    if (builder) {
        // =================================== table builders ===================================
        const int bw = warp - LM_SCAN_WARPS;
        const int m = lane & 15;                 // this thread's sub-quantiser
        const int eb = (lane >> 4) + 2 * bw;     // its codewords are eb + 16 j, j < 16
        const uint32_t tcb = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((bw >> 2) * EPT * SUB);
        float pn[EPT];
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int e = eb + (BT / 16) * j;
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
        bool more = false;  // the claiming lane: items may be left
        // one elected lane of builder warp 0 (elect.sync keeps the bulk-copy operands in uniform registers)
        const bool claimer = bw == 0 && elect_one();
        if (claimer) more = claim_block(0);
        // Hand-over runs on the BUILDER warps (they have the slack), one warp per slot, two items behind the build:
        // hand the slot's candidates to the pair's private region of the query's pool (plain
        // stores: no returning atomics on this path) and tighten the query's threshold.  Candidate distances are
        // approximate (fixed point), so beyond the k best the region also keeps whatever lies within twice the
        // rounding bound of the k-th: one of those may be the better one once re-scored.
        // a tighter bound for query q: locally, and (multi-GPU) into every peer's published array — k candidates at or
        // below it exist in the whole base, whichever shard found them
        auto tighten = [&](int q, uint32_t ord) {
            if (p.n_peers == 0) {
                atomicMax(p.pool_thr + q, ord);  // result unused: a reduction, nothing to wait for
            } else if (atomicMax(p.pool_thr + q, ord) < ord) {
                const unsigned long long w = ((unsigned long long)p.epoch << 32) | ord;
                for (int r = 0; r < p.n_peers; ++r) atomicMax(p.peer_thr[r] + q, w);
            }
        };
        auto finalize = [&](const LmHeader* hd, int gitem, uint64_t* qk, int* qcnt) {
            const int j = bw;
            int* cntp = &qcnt[j];
            const int n = *cntp;
            const int q = hd->qid[j];
            if (n > 0 && q >= 0) {
                uint64_t* kq = qk + j * LM_QC;
                bool redo = n > LM_QC;  // candidates were dropped: the plain kernel redoes this (query, item)
                if (!redo) {
                    const size_t ps = (size_t)q * p.pslots + hd->pslot[j];
                    unsigned long long* dst = p.pool + ps * p.kc;
                    const float err = LM_QERR / hd->s[j];
                    int kept = n;
                    float dk = 0.f;  // k-th best approximate distance of this pair
                    if (n > p.k) {   // warp-level sort, best first
                        const int P2 = next_pow2(n);
                        for (int i = n + lane; i < P2; i += 32) kq[i] = 0ull;
                        __syncwarp();
                        bitonic_sort_desc<true>(kq, P2, lane, 32);
                        dk = -key_score(kq[p.k - 1]);
                        const float lim = dk + 2.f * err;
                        int extra = 0;
                        for (int i = p.k + lane; i < n; i += 32) extra += (-key_score(kq[i]) <= lim);
                        extra = __reduce_add_sync(0xffffffffu, extra);  // sorted: the band is a prefix of the tail
                        kept = p.k + extra;
                        if (kept > p.kc) redo = true;
                    } else if (n == p.k) {
                        float mx = 0.f;
                        for (int i = lane; i < n; i += 32) mx = fmaxf(mx, -key_score(kq[i]));
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                        dk = mx;
                    }
                    if (!redo) {
                        for (int i = lane; i < kept; i += 32) dst[i] = kq[i];
                        if (lane == 0) {
                            p.pool_cnt[ps] = kept;
                            // k candidates whose true distance is at most dk + err each: a bound of the k-th best
                            if (n >= p.k) tighten(q, score_to_ord(-(dk + err)) - 1u);
                        }
                        // Across pairs: count the candidates per distance bucket; once the running count reaches k
                        // the bucket's upper edge (+ the rounding bound) is a bound of the query's k-th best.
                        // Counts read here may lag other CTAs' additions: a lagging count only loosens the bound.
                        const float t0 = __ldg(p.thr0 + q);
                        if (t0 > 0.f) {
                            uint32_t* hq = p.hist + (size_t)q * LM_HB;
                            const float lo = 0.5f * t0, wid = t0 * (0.5f / LM_HB);
                            for (int i = lane; i < kept; i += 32) {
                                const int bk = (int)fminf(fmaxf((-key_score(kq[i]) - lo) / wid, 0.f), (float)(LM_HB - 1));
                                atomicAdd(hq + bk, 1u);
                            }
                            __syncwarp();
                            const uint2 cc = __ldcg(reinterpret_cast<const uint2*>(hq) + lane);  // LM_HB = 64: two buckets per lane
                            uint32_t cum = cc.x + cc.y;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const uint32_t up = __shfl_up_sync(0xffffffffu, cum, o);
                                if (lane >= o) cum += up;
                            }
                            const unsigned reach = __ballot_sync(0xffffffffu, cum >= (uint32_t)p.k);
                            if (reach) {
                                const int fl = __ffs(reach) - 1;
                                if (lane == fl) {
                                    const int bk = 2 * lane + ((cum - cc.y >= (uint32_t)p.k) ? 0 : 1);
                                    // a bucket holds distances up to its upper edge (and anything clamped into bucket 0 / the last
                                    // one is counted as if at that bucket's edge: bucket 0 is conservative, the last is never tighter
                                    // than the seed bound)
                                    const float edge = bk == LM_HB - 1 ? t0 * 1.01f + 2.f * err : lo + wid * (float)(bk + 1);
                                    const float errq = LM_QERR * 1.001f * __uint_as_float(__ldg(p.sinv_max + q));
                                    tighten(q, score_to_ord(-(edge * 1.00001f + errq)) - 1u);
                                }
                            }
                        }
                    }
                }
                if (redo && lane == 0) p.redo[atomicAdd(p.redo_cnt, 1)] = make_int2(q, gitem * LM_QS + j);
            }
            __syncwarp();
            if (lane == 0) *cntp = 0;
        };

        // item i: wait until every scan warp has left it (that frees its table half) ...
        auto wait_done = [&](int i) { mbar_wait(bar_done + 8 * (i & 1), (uint32_t)(i >> 1) & 1u); };
        // ... and empty its queues.  The hand-over runs AFTER the next table is built and published, so the scanners never
        // wait for it: queue sets rotate over three items (item i pushes into set i % 3 while set (i-2) % 3 is emptied).
        auto hand_over = [&](int i) {
            finalize(reinterpret_cast<const LmHeader*>(rbuf + (i % LM_BLK_STAGES) * LM_BLK_MAX), s_item[i % LM_BLK_STAGES],
                     qkeys + (i % LM_QSETS) * (LM_QS * LM_QC), s_qcnt + (i % LM_QSETS) * LM_QS);
        };
        const unsigned long long magic2 = pack2(8388608.f, 8388608.f);  // 2^23: the sum's low mantissa bits are the integer
        const unsigned long long quarter = pack2(0.25f, 0.25f);          // t = -2 r  =>  |r|^2 = sum t^2 / 4
        int i = 0;
        for (;; ++i) {
            const int b = i & 1;
            if (i >= 2) wait_done(i - 2);  // table half b is free
            if (more) more = claim_block(i + 1);
            mbar_wait(bar_blk + 8 * (i % LM_BLK_STAGES), (uint32_t)(i / LM_BLK_STAGES) & 1u);
            if (s_item[i % LM_BLK_STAGES] < 0) break;  // no item i: items 0 .. i-1 were this CTA's share
            const unsigned char* blkp = rbuf + (i % LM_BLK_STAGES) * LM_BLK_MAX;
            const LmHeader* hd = reinterpret_cast<const LmHeader*>(blkp);
            // the builders run one item ahead of the scanners: pull this item's codes from HBM into L2 now, so the
            // scanners' loads (a few hundred rows ahead at most) find them there
            if (claimer && hd->nvec > 0)
                bulk_prefetch_l2(p.codes + (size_t)hd->vbeg * 16, (uint32_t)hd->nvec * 16u);
            unsigned char* lut = lut0 + b * LM_LUT_BYTES;
            // four queries at a time: |p|^2 + |r_m|^2 - 2 r_m.p, x s_j, rounded
#pragma unroll 1
            for (int h = 0; h < LM_QS / 4; ++h) {
                // slots fill in order: a half whose first slot is unused has no query at all.  Its table entries keep
                // whatever an earlier item left there (or the initial zeros): at most LM_QMAX each, so the sums of the
                // unused lanes stay below 2^15 and never disturb their neighbours
                if (hd->qid[4 * h] < 0) continue;
                const ulonglong2* rt = reinterpret_cast<const ulonglong2*>(blkp + LM_HDR) + h * p.dim + m;  // slot d*16 + m
                unsigned long long t01[SUB], t23[SUB], rr01 = 0ull, rr23 = 0ull;
#pragma unroll
                for (int d = 0; d < SUB; ++d) {
                    const ulonglong2 v = rt[d * 16];
                    t01[d] = v.x; t23[d] = v.y;
                    rr01 = ffma2(v.x, v.x, rr01);
                    rr23 = ffma2(v.y, v.y, rr23);
                }
                const float4 sh4 = *reinterpret_cast<const float4*>(hd->s + 4 * h);  // 0 for an unused slot: its entries are 0
                const unsigned long long s01 = pack2(sh4.x, sh4.y), s23 = pack2(sh4.z, sh4.w);
                // slots fill in order: no fifth query = a compact table (8-byte entries, the scanners read it with LDS.64)
                const bool compact = hd->qid[LM_QS / 2] < 0;
                unsigned char* lw = compact ? lut + eb * 128 + m * 8 : lut + eb * 256 + m * 16 + h * 8;
                const int jstride = compact ? (BT / 16) * 128 : (BT / 16) * 256;
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
                    a01 = ffma2(a01, s01, magic2);
                    a23 = ffma2(a23, s23, magic2);
                    const uint32_t w0 = __byte_perm((uint32_t)a01, (uint32_t)(a01 >> 32), 0x5410);
                    const uint32_t w1 = __byte_perm((uint32_t)a23, (uint32_t)(a23 >> 32), 0x5410);
                    *reinterpret_cast<uint2*>(lw + j * jstride) = make_uint2(w0, w1);
                    if (j + 1 < EPT) tmem_ld_wait();
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * b);  // release: this warp's table stores are visible to the waiters
            if (i >= 2) hand_over(i - 2);
        }
        // no item i: items i-2 (scan finished: waited for above) and i-1 are still to be handed over
        if (i >= 2) hand_over(i - 2);
        if (i >= 1) { wait_done(i - 1); hand_over(i - 1); }
    } else {
        // =================================== scanners ===================================
        // lane reads table (lane + t) & 15 at step t; op[i] packs the table byte offsets of steps 2i, 2i+1
        uint32_t op[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            op[i] = (uint32_t)(((lane + 2 * i) & 15) << 4) | ((uint32_t)(((lane + 2 * i + 1) & 15) << 4) << 8);
            asm volatile("" : "+r"(op[i]));  // keep the eight offset words in registers (no rematerialisation in the loop)
        }
        const int rot = lane & 15;

        // codes stream straight from L2/HBM, one coalesced 16-byte row per lane, LM_PF chunks of 256 rows ahead; the
        // first chunks of item i+1 are requested before item i is handed over, so they arrive behind the barrier
        uint4 cq[LM_PF];
        auto prefetch_item = [&](int i) -> bool {  // false: this CTA has no i-th item
            mbar_wait(bar_blk + 8 * (i % LM_BLK_STAGES), (uint32_t)(i / LM_BLK_STAGES) & 1u);
            if (s_item[i % LM_BLK_STAGES] < 0) return false;
            const LmHeader* h = reinterpret_cast<const LmHeader*>(rbuf + (i % LM_BLK_STAGES) * LM_BLK_MAX);
            const int nv = h->nvec;
            const uint4* cp = reinterpret_cast<const uint4*>(p.codes) + h->vbeg;
#pragma unroll
            for (int u = 0; u < LM_PF; ++u) {
                const int v = (warp + u * LM_SCAN_WARPS) * 32 + lane;
                cq[u] = make_uint4(0u, 0u, 0u, 0u);
                if (v < nv) cq[u] = __ldg(cp + v);
            }
            return true;
        };
        bool have = prefetch_item(0);
        for (int i = 0; have; ++i) {
            const int b = i & 1;
            const unsigned char* blkp = rbuf + (i % LM_BLK_STAGES) * LM_BLK_MAX;  // arrived: prefetch_item(i) waited for it
            const LmHeader* hd = reinterpret_cast<const LmHeader*>(blkp);
            const int nvec = hd->nvec;
            const long long vbeg = hd->vbeg;
            const uint4* cp = reinterpret_cast<const uint4*>(p.codes) + vbeg;
            uint64_t* qk = qkeys + (i % LM_QSETS) * (LM_QS * LM_QC);  // queue sets rotate over three items: this item's pushes
            int* qcnt = s_qcnt + (i % LM_QSETS) * LM_QS;              // never meet the hand-over of item i-2 or i-1
            // every warp keeps its own copy of the slots' scalars (lanes 0-7 compute one slot each): the current
            // threshold becomes an integer bound, accept sum <= floor(s * tau) + QERR (rounded up)
            int* ti_w = s_ti + warp * LM_QS;
            float* inv_w = s_inv + warp * LM_QS;
            {
                int ti = -1;
                float inv = 0.f;
                if (lane < LM_QS) {
                    const int myq = hd->qid[lane];
                    if (myq >= 0) {
                        const float mys = hd->s[lane];
                        uint32_t mytu = __ldcg(p.pool_thr + myq);
                        if (p.thr_pub) {  // a bound a peer rank proved for THIS batch
                            const unsigned long long w = __ldcg(p.thr_pub + myq);
                            if ((uint32_t)(w >> 32) == p.epoch) mytu = max(mytu, (uint32_t)w);
                        }
                        const float tau = mytu ? -ord_to_score(mytu) : INFINITY;
                        ti = (int)fminf(fmaxf(tau, 0.f) * mys + (LM_QERR + 1.f), 32767.f);
                        inv = 1.f / mys;
                    }
                    ti_w[lane] = ti;
                    inv_w[lane] = inv;
                }
                __syncwarp();
            }
            uint32_t th[LM_QS / 2];  // per query pair: (0x8000 | threshold) halves; an unused slot never passes (0x7fff, no guard bit)
#pragma unroll
            for (int u = 0; u < LM_QS / 2; ++u) {
                const int t0 = ti_w[2 * u], t1 = ti_w[2 * u + 1];
                th[u] = (t0 >= 0 ? (0x8000u | (uint32_t)t0) : 0x7fffu) | ((t1 >= 0 ? (0x8000u | (uint32_t)t1) : 0x7fffu) << 16);
            }
            mbar_wait(bar_full + 8 * b, (uint32_t)(i >> 1) & 1u);  // acquire: the builders' table stores
            const unsigned char* lut = lut0 + b * LM_LUT_BYTES;
            // two chunks (rows v0 and v1 = v0 + 256) per iteration: 32 independent table reads in flight per lane.
            // HALF: an item with at most four queries uses the compact table (LUT[e][m] = {q0..q3}, 8 bytes at e*128 + m*8):
            // one LDS.64 — two crossbar wavefronts instead of four — per 32 (code, table) lookups.
            auto scan_rows = [&](auto half_tag) {
            constexpr bool HALF = decltype(half_tag)::value;
            for (int c = warp; c * 32 < nvec; c += 2 * LM_SCAN_WARPS) {
                const int v0 = c * 32 + lane, v1 = v0 + LM_SCAN_WARPS * 32;
                const uint4 cw0 = cq[0], cw1 = cq[1];
#pragma unroll
                for (int u = 0; u + 2 < LM_PF; ++u) cq[u] = cq[u + 2];
                if (v0 + LM_PF * LM_SCAN_WARPS * 32 < nvec) cq[LM_PF - 2] = __ldg(cp + v0 + LM_PF * LM_SCAN_WARPS * 32);
                if (v1 + LM_PF * LM_SCAN_WARPS * 32 < nvec) cq[LM_PF - 1] = __ldg(cp + v1 + LM_PF * LM_SCAN_WARPS * 32);
                uint32_t wa[4], wb[4];
                {   // rotate the 16 code bytes: new byte t = old byte (t + rot) & 15
                    const bool r8 = rot & 8, r4 = rot & 4;
                    const int sh = (rot & 3) * 8;
                    {
                        const uint32_t a0 = r8 ? cw0.z : cw0.x, a1 = r8 ? cw0.w : cw0.y, a2 = r8 ? cw0.x : cw0.z, a3 = r8 ? cw0.y : cw0.w;
                        const uint32_t b0 = r4 ? a1 : a0, b1 = r4 ? a2 : a1, b2 = r4 ? a3 : a2, b3 = r4 ? a0 : a3;
                        wa[0] = __funnelshift_r(b0, b1, sh); wa[1] = __funnelshift_r(b1, b2, sh);
                        wa[2] = __funnelshift_r(b2, b3, sh); wa[3] = __funnelshift_r(b3, b0, sh);
                    }
                    {
                        const uint32_t a0 = r8 ? cw1.z : cw1.x, a1 = r8 ? cw1.w : cw1.y, a2 = r8 ? cw1.x : cw1.z, a3 = r8 ? cw1.y : cw1.w;
                        const uint32_t b0 = r4 ? a1 : a0, b1 = r4 ? a2 : a1, b2 = r4 ? a3 : a2, b3 = r4 ? a0 : a3;
                        wb[0] = __funnelshift_r(b0, b1, sh); wb[1] = __funnelshift_r(b1, b2, sh);
                        wb[2] = __funnelshift_r(b2, b3, sh); wb[3] = __funnelshift_r(b3, b0, sh);
                    }
                }
                // rows past the end read table entries of code 0 (their code words are zero) and are discarded below
                uint32_t A0 = 0u, A1 = 0u, A2 = 0u, A3 = 0u, B0 = 0u, B1 = 0u, B2 = 0u, B3 = 0u;
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    // byte0 = table offset, byte1 = code byte: address = code * 256 + table * 16
                    const uint32_t sel = 0x7600u | (uint32_t)((t & 3) << 4) | (uint32_t)(4 + (t & 1));
                    const uint32_t aa = __byte_perm(wa[t >> 2], op[t >> 1], sel);
                    const uint32_t ab = __byte_perm(wb[t >> 2], op[t >> 1], sel);
                    if (HALF) {
                        const uint2 ea = *reinterpret_cast<const uint2*>(lut + (aa >> 1));
                        const uint2 eb = *reinterpret_cast<const uint2*>(lut + (ab >> 1));
                        A0 += ea.x; A1 += ea.y;
                        B0 += eb.x; B1 += eb.y;
                    } else {
                        const uint4 ea = *reinterpret_cast<const uint4*>(lut + aa);
                        const uint4 eb = *reinterpret_cast<const uint4*>(lut + ab);
                        A0 += ea.x; A1 += ea.y; A2 += ea.z; A3 += ea.w;
                        B0 += eb.x; B1 += eb.y; B2 += eb.z; B3 += eb.w;
                    }
                }
                // bit 15 / 31 of (guarded threshold - sum) survives exactly where sum <= threshold
                const uint32_t ha = HALF ? ((th[0] - A0) | (th[1] - A1)) : ((th[0] - A0) | (th[1] - A1) | (th[2] - A2) | (th[3] - A3));
                const uint32_t hb = HALF ? ((th[0] - B0) | (th[1] - B1)) : ((th[0] - B0) | (th[1] - B1) | (th[2] - B2) | (th[3] - B3));
                const uint32_t hita = v0 < nvec ? ha & 0x80008000u : 0u;
                const uint32_t hitb = v1 < nvec ? hb & 0x80008000u : 0u;
                if (hita | hitb) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (r ? hitb : hita) {
                            const long long gpos = vbeg + (r ? v1 : v0);
                            if (!(p.dead && p.dead[gpos])) {
                                const uint32_t acc[4] = {r ? B0 : A0, r ? B1 : A1, r ? B2 : A2, r ? B3 : A3};
#pragma unroll
                                for (int j = 0; j < (HALF ? LM_QS / 2 : LM_QS); ++j) {
                                    const int sum = (int)((acc[j >> 1] >> (16 * (j & 1))) & 0xffffu);
                                    if (sum <= ti_w[j]) {
                                        const int pos = atomicAdd(&qcnt[j], 1);
                                        if (pos < LM_QC) qk[j * LM_QC + pos] = make_key(-((float)sum * inv_w[j]), (uint32_t)gpos);
                                    }
                                }
                            }
                        }
                    }
                }
            }
            };
            if (hd->qid[LM_QS / 2] < 0) scan_rows(std::true_type{}); else scan_rows(std::false_type{});
            have = prefetch_item(i + 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_done + 8 * b);  // release: this warp's pushes; it no longer reads table half b
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == LM_SCAN_WARPS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TCOLS) : "memory");
    }
}

// ---- redo: plain scan of one (query, item) whose queue overflowed -------------------------------------------
struct LmRedo {
    const float* Q; int dim; const float* centroids; const float* codebook; int ksub;
    const uint8_t* codes; const uint8_t* dead;
    const unsigned char* iblk; int blk; const int2* redo; const int32_t* redo_cnt;
    unsigned long long* pool; int32_t* pool_cnt; int pslots; int k; int kc;
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
        for (int i = tid; i < keep; i += 256) a.pool[ps * a.kc + i] = keys[i];
    }
}

// ---- pool -> candidates within the rounding band of the k-th, exact re-score, final order -----------------
struct LmFinalParams {
    const float* Q; int dim;
    const float* centroids; const float* codebook; int ksub; int m;  // the index's own quantiser: [m][ksub][dim/m]
    const uint8_t* codes; const int64_t* list_off; int nlist; const int64_t* labels;  // codes [total][m]
    const int64_t* probes; int P; int64_t nq;
    const unsigned long long* pool; const int32_t* pool_cnt; int pslots; int k; int kc;
    const uint32_t* sinv_max;  // [nq] max 1/scale over the query's items (float bits): bounds every pool entry's error
    int keys_cap;              // keys in the sort window (the list ranges sit behind it in shared memory)
    PairOut out;
    int32_t* counts;           // nullable: results per query (when this kernel writes the search's final output)
};

// One CTA per query.  The k-th best APPROXIMATE distance is bounded with a 256-bucket histogram over the pool's range
// (a bucket's upper edge, not a full sort: the pool holds a few hundred entries, the band a few dozen), everything within
// the rounding band of it is compacted to the front and re-scored - one half warp per survivor - and only those
// survivors are sorted.
template <int SUB>
__global__ void __launch_bounds__(256) ivfpq_lm_final_kernel(LmFinalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [next_pow2(pool_cap)]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_n, s_m;
    __shared__ uint32_t s_lohi[2];
    __shared__ int s_hist[256];
    __shared__ float s_lim;
    for (int64_t q = blockIdx.x; q < p.nq; q += gridDim.x) {  // launched with one CTA per query
        if (tid == 0) { s_n = 0; s_m = 0; s_lohi[0] = 0xffffffffu; s_lohi[1] = 0u; }
        for (int i = tid; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        // the probed lists and their code ranges, once per query: a survivor's list is found by a shared-memory search below
        // instead of dependent global loads per survivor
        long long* s_lo = reinterpret_cast<long long*>(keys + p.keys_cap);
        long long* s_hi = s_lo + p.P;
        int* s_list = reinterpret_cast<int*>(s_hi + p.P);
        float* s_res = reinterpret_cast<float*>(s_list + p.P) + (tid >> 4) * p.dim;  // m < 16: this half warp's residual query
        for (int sl = tid; sl < p.P; sl += blockDim.x) {
            const int64_t l = __ldg(p.probes + q * p.P + sl);
            s_list[sl] = (int)l;
            s_lo[sl] = l >= 0 ? __ldg(p.list_off + l) : 0;
            s_hi[sl] = l >= 0 ? __ldg(p.list_off + l + 1) : 0;
        }
        for (int sl = tid; sl < p.pslots; sl += blockDim.x) {  // gather the pairs' private regions
            const size_t ps = (size_t)q * p.pslots + sl;
            const int c = min(p.pool_cnt[ps], p.kc);
            if (c > 0) {
                const int base = atomicAdd(&s_n, c);
                uint32_t lo = 0xffffffffu, hi = 0u;
                for (int i = 0; i < c; ++i) {
                    const uint64_t key = p.pool[ps * p.kc + i];
                    keys[base + i] = key;
                    lo = min(lo, (uint32_t)(key >> 32)); hi = max(hi, (uint32_t)(key >> 32));
                }
                atomicMin(&s_lohi[0], lo);
                atomicMax(&s_lohi[1], hi);
            }
        }
        __syncthreads();
        const int n = s_n;
        const int kk = min(n, p.k);
        // Pool distances are approximate (fixed-point tables): |d - true| <= err.  Everything within 2 err of the k-th
        // best approximate distance can still belong to the true top k, so all of it is re-scored.
        int nres = n;
        if (n > kk) {
            // distances run from dmin (the best score) to dmax; bucket b holds dmin + [b, b+1) / scale
            const float dmin = -ord_to_score(s_lohi[1]), dmax = -ord_to_score(s_lohi[0]);
            const float scale = dmax > dmin ? 255.f / (dmax - dmin) : 0.f;
            for (int i = tid; i < n; i += blockDim.x)
                atomicAdd(&s_hist[min(255, (int)((-key_score(keys[i]) - dmin) * scale))], 1);
            __syncthreads();
            if (warp == 0) {  // the first bucket where the running count reaches k: its upper edge bounds the k-th best
                int c[8], tot = 0;
    #pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = s_hist[lane * 8 + j]; tot += c[j]; }
                int cum = tot;
    #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, cum, o);
                    if (lane >= o) cum += up;
                }
                const unsigned reach = __ballot_sync(0xffffffffu, cum >= kk);
                if (lane == __ffs(reach) - 1) {
                    int run = cum - tot, b = lane * 8;
    #pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        run += c[j];
                        if (run >= kk) break;
                        ++b;
                    }
                    const float edge = scale > 0.f ? fminf(dmax, dmin + (float)(b + 1) / scale * 1.00001f) : dmax;
                    const float err = LM_QERR * 1.001f * __uint_as_float(__ldg(p.sinv_max + q));
                    s_lim = edge + 2.f * err + 2e-5f * edge;
                }
            }
            __syncthreads();
            const float lim = s_lim;
            // compact the band to the front, a block of entries at a time (writes never pass the entries still to be read)
            for (int i0 = 0; i0 < n; i0 += blockDim.x) {
                const int i = i0 + tid;
                const uint64_t key = i < n ? keys[i] : 0ull;
                const bool keep = i < n && -key_score(key) <= lim;
                const unsigned mk = __ballot_sync(0xffffffffu, keep);
                int base = 0;
                __syncthreads();
                if (lane == 0 && mk) base = atomicAdd(&s_m, __popc(mk));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) keys[base + __popc(mk & ((1u << lane) - 1u))] = key;
            }
            __syncthreads();
            nres = s_m;
        }
        // exact re-score of the survivors: IvfPqVectorIndex.cs:161-166,182-186 in the reference's order; one half warp each
        const int hl = tid & 15, hw = tid >> 4, nhw = blockDim.x >> 4;
        const unsigned hmask = 0xffffu << (lane & 16);
        for (int i0 = 0; i0 < nres; i0 += nhw) {  // uniform trip count: the shuffles below need the whole warp
            const int i = i0 + hw;
            const bool on = i < nres;
            const uint32_t pos = key_pos(keys[on ? i : i0]);
            int lo = 0;  // the list holding pos is one of this query's probed lists: test them sixteen at a time
            for (int p0 = 0; p0 < p.P; p0 += 16) {
                const bool hit = p0 + hl < p.P && s_lo[p0 + hl] <= (long long)pos && (long long)pos < s_hi[p0 + hl];
                const unsigned mh = __ballot_sync(0xffffffffu, hit) & hmask;
                if (mh) lo = s_list[p0 + ((__ffs(mh) - 1) & 15)];
                if (__all_sync(0xffffffffu, mh != 0 || lo != 0)) break;  // both halves found theirs (list 0 is re-checked, harmlessly)
            }
            float dm = 0.f, dist = 0.f;
            if (p.m == 16) {
                float r[SUB];
    #pragma unroll
                for (int d = 0; d < SUB; ++d)
                    r[d] = __fsub_rn(__ldg(p.Q + q * p.dim + hl * SUB + d), __ldg(p.centroids + (size_t)lo * p.dim + hl * SUB + d));
                const int code = p.codes[(size_t)pos * 16 + hl];
                dm = exact::a1_l2_fixed<SUB>(r, p.codebook + ((size_t)hl * p.ksub + code) * SUB);
    #pragma unroll
                for (int mi = 0; mi < 16; ++mi) dist = __fadd_rn(dist, __shfl_sync(0xffffffffu, dm, (lane & 16) + mi));
            } else {
                // fewer, longer sub-vectors (dim/m = 16 .. 128): L2SquaredUnsafe's four-accumulator form applies from 32 on
                const int subr = p.dim / p.m;
                for (int d = hl; d < p.dim; d += 16)
                    s_res[d] = __fsub_rn(__ldg(p.Q + q * p.dim + d), __ldg(p.centroids + (size_t)lo * p.dim + d));
                __syncwarp();
                if (hl < p.m) {
                    const int code = p.codes[(size_t)pos * p.m + hl];
                    dm = exact::a1_l2_eval(s_res + hl * subr, p.codebook + ((size_t)hl * p.ksub + code) * subr, subr);
                }
                for (int mi = 0; mi < p.m; ++mi) dist = __fadd_rn(dist, __shfl_sync(0xffffffffu, dm, (lane & 16) + mi));
                __syncwarp();
            }
            if (hl == 0 && on) keys[i] = make_key(-dist, pos);
        }
        __syncthreads();
        const int P3 = next_pow2(max(nres, 2));
        for (int i = nres + tid; i < P3; i += blockDim.x) keys[i] = 0ull;
        __syncthreads();
        if (P3 <= 64) {  // the usual case: one warp sorts without block barriers
            if (warp == 0) bitonic_sort_desc<true>(keys, P3, lane, 32);
            __syncthreads();
        } else {
            bitonic_sort_desc<false>(keys, P3, tid, blockDim.x);
        }
        const int64_t ob = (q * p.out.parts_total + p.out.part_base) * (int64_t)p.k;
        if (tid == 0 && p.counts) p.counts[q] = kk;
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
        __syncthreads();  // the next query re-uses every shared array
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int lm_maxseg(int64_t) { return 1; }  // one item covers a whole list (codes stream from L2/HBM, no staging buffer)

struct LmLayout {
    size_t zero_bytes;  // leading region cleared per search
    size_t lcnt, lcur, pool_cnt, pool_thr, sinv_max, hist, thr0, redo_cnt, item_ctr, scanned, cmax, loff, nit, ioff, pairq, pairp, item_list, redo, iblk, pool, temp, total;
    size_t temp_bytes;
    int64_t max_items;
    int pool_cap, pslots, blk, kc;
};
// pool entries per (query, probe) pair: the k best plus room for candidates within the rounding band of the k-th
inline int lm_kc(int k) { return k + 6 + k / 8; }

LmLayout lm_layout(int64_t nq, int P, int k, int nlist, int dim, int64_t max_list_len) {
    LmLayout L{};
    const int64_t npairs = nq * P;
    const int maxseg = lm_maxseg(max_list_len);
    size_t o = 0;
    L.lcnt = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.lcur = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.pool_cnt = o; o += align_up(sizeof(int32_t) * (size_t)nq * P * maxseg, 256);
    L.pool_thr = o; o += align_up(sizeof(uint32_t) * (size_t)nq, 256);
    L.sinv_max = o; o += align_up(sizeof(uint32_t) * (size_t)nq, 256);
    L.hist = o; o += align_up(sizeof(uint32_t) * (size_t)nq * LM_HB, 256);
    L.thr0 = o; o += align_up(sizeof(float) * (size_t)nq, 256);
    L.redo_cnt = o; o += 256;
    L.item_ctr = o; o += 256;
    L.scanned = o; o += 256;
    L.zero_bytes = o;
    L.cmax = o; o += 256;
    L.loff = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.nit = o; o += align_up(sizeof(unsigned long long) * ((size_t)nlist + 2), 256);  // packed (item, pair) offsets
    L.ioff = o; o += align_up(sizeof(int32_t) * ((size_t)nlist + 1), 256);
    L.pairq = o; o += align_up(sizeof(int32_t) * (size_t)npairs, 256);
    L.pairp = o; o += align_up(sizeof(int32_t) * (size_t)npairs, 256);
    L.max_items = (npairs / LM_QS + std::min<int64_t>(npairs, nlist) + 1) * maxseg;
    L.item_list = o; o += align_up(sizeof(int32_t) * (size_t)L.max_items, 256);
    L.redo = o; o += align_up(sizeof(int2) * (size_t)L.max_items * LM_QS, 256);
    L.blk = LM_HDR + dim * 32;
    L.iblk = o; o += align_up((size_t)L.blk * (size_t)L.max_items, 256);
    L.pslots = P * maxseg;
    L.kc = lm_kc(k);
    L.pool_cap = L.pslots * L.kc;
    L.pool = o; o += align_up(sizeof(unsigned long long) * (size_t)nq * L.pool_cap, 256);
    size_t tb = 0;
    {
        cub::TransformInputIterator<unsigned long long, LmPackOp, cub::CountingInputIterator<int>> it(cub::CountingInputIterator<int>(0),
                                                                                                     LmPackOp{nullptr, nlist});
        cub::DeviceScan::ExclusiveSum(nullptr, tb, it, (unsigned long long*)nullptr, nlist + 1);
    }
    L.temp_bytes = tb + 256;
    L.temp = o; o += align_up(L.temp_bytes, 256);
    L.total = o;
    return L;
}

template <int SUB>
cudaError_t launch_lm(const IvfPqScanParams& p, void* scratch, int num_sms, cudaStream_t st) {
    const int P = p.nprobe;
    const int64_t npairs = p.nq * P;
    // the approximate stages (seed, scan, redo) run on 16 tables: the index's own when m = 16, else the equivalent split
    const float* cb16 = p.lm_codebook ? p.lm_codebook : p.codebook;
    const uint8_t* codes16 = p.lm_codes ? p.lm_codes : p.codes;
    if (p.m != 16 && (!p.lm_codebook || !p.lm_codes)) return cudaErrorInvalidValue;
    const LmLayout L = lm_layout(p.nq, P, p.k, p.nlist, p.dim, p.max_list_len);
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    int32_t* lcnt = reinterpret_cast<int32_t*>(base + L.lcnt);
    int32_t* lcur = reinterpret_cast<int32_t*>(base + L.lcur);
    int32_t* pool_cnt = reinterpret_cast<int32_t*>(base + L.pool_cnt);
    uint32_t* pool_thr = reinterpret_cast<uint32_t*>(base + L.pool_thr);
    uint32_t* sinv_max = reinterpret_cast<uint32_t*>(base + L.sinv_max);
    float* cmax = reinterpret_cast<float*>(base + L.cmax);
    uint32_t* hist = reinterpret_cast<uint32_t*>(base + L.hist);
    float* thr0 = reinterpret_cast<float*>(base + L.thr0);
    int32_t* redo_cnt = reinterpret_cast<int32_t*>(base + L.redo_cnt);
    unsigned long long* scanned = reinterpret_cast<unsigned long long*>(base + L.scanned);
    int32_t* loff = reinterpret_cast<int32_t*>(base + L.loff);
    unsigned long long* poff = reinterpret_cast<unsigned long long*>(base + L.nit);
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
        sd.codebook = cb16; sd.ksub = p.ksub; sd.codes = codes16; sd.dead = p.dead; sd.list_off = p.list_off;
        sd.pool_thr = pool_thr; sd.k = p.k; sd.sample = std::max(512, 16 * p.k); sd.thr0 = thr0;
        const size_t seed_smem = sizeof(float) * ((size_t)SEED_NQ * 4096 + (size_t)SEED_NQ * SEED_CAP + (size_t)SEED_NQ * p.dim);
        e = cudaFuncSetAttribute(ivfpq_lm_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem);
        if (e != cudaSuccess) return e;
        ivfpq_lm_seed_kernel<<<(unsigned)((p.nq + SEED_NQ - 1) / SEED_NQ), 256, seed_smem, sst>>>(sd);
        cudaEventRecord(p.ev_join, p.aux_stream);
    }
    const unsigned gb = (unsigned)((npairs + 255) / 256);
    if (p.cmax) cmax = const_cast<float*>(p.cmax);  // cached per index: depends on the codebook only
    else lm_cmax_kernel<<<1, 256, 0, st>>>(cb16, p.ksub, p.dim / 16, cmax);
    lm_count_kernel<<<gb, 256, 0, st>>>(p.probes, npairs, p.list_off, lcnt, scanned);
    {
        cub::TransformInputIterator<unsigned long long, LmPackOp, cub::CountingInputIterator<int>> it(cub::CountingInputIterator<int>(0),
                                                                                                     LmPackOp{lcnt, p.nlist});
        e = cub::DeviceScan::ExclusiveSum(temp, tb, it, poff, p.nlist + 1, st);
        if (e != cudaSuccess) return e;
    }
    const unsigned fb = (unsigned)((std::max<int64_t>(npairs, p.nlist + 1) + 255) / 256);
    lm_fill_kernel<<<fb, 256, 0, st>>>(p.probes, npairs, P, p.list_off, poff, p.nlist, loff, ioff, lcur, pairq, pairp, item_list);

    LmPrep pa{};
    pa.maxseg = lm_maxseg(p.max_list_len); pa.pairp = pairp;
    pa.item_list = item_list;
    pa.ioff = ioff; pa.loff = loff; pa.pairq = pairq; pa.list_off = p.list_off; pa.nlist = p.nlist;
    pa.Q = p.Q; pa.centroids = p.centroids; pa.dim = p.dim; pa.iblk = iblk; pa.blk = L.blk;
    pa.cmax = cmax; pa.sinv_max = sinv_max;
    mark();
    lm_prepare_kernel<<<(unsigned)std::min<int64_t>((L.max_items * 32 + 255) / 256, (int64_t)num_sms * 8), 256, 0, st>>>(pa);
    mark();

    if (!fork) {
        LmSeed sd{};
        sd.Q = p.Q; sd.nq = p.nq; sd.dim = p.dim; sd.probes = p.probes; sd.P = P; sd.centroids = p.centroids;
        sd.codebook = cb16; sd.ksub = p.ksub; sd.codes = codes16; sd.dead = p.dead; sd.list_off = p.list_off;
        sd.pool_thr = pool_thr; sd.k = p.k; sd.sample = std::max(512, 16 * p.k); sd.thr0 = thr0;
        const size_t seed_smem = sizeof(float) * ((size_t)SEED_NQ * 4096 + (size_t)SEED_NQ * SEED_CAP + (size_t)SEED_NQ * p.dim);
        e = cudaFuncSetAttribute(ivfpq_lm_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem);
        if (e != cudaSuccess) return e;
        ivfpq_lm_seed_kernel<<<(unsigned)((p.nq + SEED_NQ - 1) / SEED_NQ), 256, seed_smem, st>>>(sd);
    } else {
        cudaStreamWaitEvent(st, p.ev_join, 0);
    }
    mark();

    LmParams sp{};
    sp.dim = p.dim; sp.ksub = p.ksub; sp.k = p.k; sp.codebook = cb16; sp.codes = codes16; sp.dead = p.dead;
    sp.iblk = iblk; sp.n_items = ioff + p.nlist;
    sp.pool = pool; sp.pool_cnt = pool_cnt; sp.pool_thr = pool_thr; sp.pslots = L.pslots; sp.kc = L.kc;
    sp.hist = hist; sp.thr0 = thr0; sp.sinv_max = sinv_max;
    sp.redo = redo; sp.redo_cnt = redo_cnt; sp.item_ctr = reinterpret_cast<int32_t*>(base + L.item_ctr);
    sp.thr_pub = p.thr_pub; sp.n_peers = p.n_peers; sp.epoch = p.epoch;
    for (int r = 0; r < 7; ++r) sp.peer_thr[r] = r < p.n_peers ? p.peer_thr[r] : nullptr;
    e = cudaFuncSetAttribute(ivfpq_lm_scan_kernel<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM);
    if (e != cudaSuccess) return e;
    const int64_t grid = std::min<int64_t>(num_sms, L.max_items);
    if (p.ev_k0) cudaEventRecord(p.ev_k0, st);
    ivfpq_lm_scan_kernel<SUB><<<(unsigned)grid, LM_THREADS, LM_SMEM, st>>>(sp);
    if (p.ev_k1) cudaEventRecord(p.ev_k1, st);

    LmRedo rd{};
    rd.Q = p.Q; rd.dim = p.dim; rd.centroids = p.centroids; rd.codebook = cb16; rd.ksub = p.ksub;
    rd.codes = codes16; rd.dead = p.dead; rd.iblk = iblk; rd.blk = L.blk; rd.redo = redo; rd.redo_cnt = redo_cnt;
    rd.pool = pool; rd.pool_cnt = pool_cnt; rd.pslots = L.pslots; rd.k = p.k; rd.kc = L.kc; rd.pool_thr = pool_thr;
    const size_t redo_smem = sizeof(uint64_t) * REDO_QCAP + sizeof(float) * (4096 + (size_t)p.dim);
    mark();
    ivfpq_lm_redo_kernel<<<(unsigned)(2 * num_sms), 256, redo_smem, st>>>(rd);
    mark();

    LmFinalParams fp{};
    fp.Q = p.Q; fp.dim = p.dim; fp.centroids = p.centroids; fp.codebook = p.codebook; fp.ksub = p.ksub; fp.m = p.m;
    fp.codes = p.codes; fp.list_off = p.list_off; fp.nlist = p.nlist; fp.labels = p.labels;
    fp.probes = p.probes; fp.P = P; fp.nq = p.nq;
    fp.pool = pool; fp.pool_cnt = pool_cnt; fp.pslots = L.pslots; fp.k = p.k; fp.kc = L.kc; fp.sinv_max = sinv_max; fp.out = p.out; fp.counts = p.out_counts;
    fp.keys_cap = next_pow2(std::max(2, L.pool_cap));
    const int fthreads = L.pool_cap <= 2048 ? 128 : 256;
    const size_t fsm = sizeof(uint64_t) * (size_t)fp.keys_cap + (2 * sizeof(long long) + sizeof(int)) * (size_t)P +
                       (p.m == 16 ? 0 : sizeof(float) * (size_t)(fthreads / 16) * p.dim);
    e = cudaFuncSetAttribute(ivfpq_lm_final_kernel<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    if (e != cudaSuccess) return e;
    // one CTA per query (a resident grid walking the queries measured slower: 0.164 against 0.133 ms on C5)
    ivfpq_lm_final_kernel<SUB><<<(unsigned)p.nq, fthreads, fsm, st>>>(fp);
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
        int nitems = 0;
        cudaMemcpy(&nitems, ioff + p.nlist, sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, " items=%d redo=%d\n", nitems, nredo);
        for (int i = 0; i < nsev; ++i) cudaEventDestroy(sev[i]);
    }
    return cudaGetLastError();
}

// ---- the equivalent 16-table quantiser of an m < 16 index ------------------------------------------------------
// |r_m - p|^2 over a sub-vector of dim/m dimensions is the sum of the same expression over its 16/m pieces of dim/16
// dimensions, all addressed by the one code byte: piece v = m * (16/m) + j of codeword e is p[j * dim/16 ...].
__global__ void lm_expand_codebook_kernel(const float* __restrict__ cb, int m, int ksub, int dim, float* __restrict__ cb16) {
    const int sub16 = dim / 16, subr = dim / m, per = 16 / m;
    const int64_t n = (int64_t)16 * ksub * sub16;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % sub16), e = (int)((i / sub16) % ksub), v = (int)(i / ((int64_t)sub16 * ksub));
        cb16[i] = cb[((size_t)(v / per) * ksub + e) * subr + (v % per) * sub16 + d];
    }
}
__global__ void lm_expand_codes_kernel(const uint8_t* __restrict__ codes, int64_t n, int m, uint4* __restrict__ codes16) {
    const int per = 16 / m;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int v = 0; v < 16; ++v) w[v >> 2] |= (uint32_t)codes[r * m + v / per] << (8 * (v & 3));
        codes16[r] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace

cudaError_t launch_pq_lm_expand_codebook(const float* codebook, int m, int ksub, int dim, float* cb16, cudaStream_t st) {
    lm_expand_codebook_kernel<<<64, 256, 0, st>>>(codebook, m, ksub, dim, cb16);
    return cudaGetLastError();
}
cudaError_t launch_pq_lm_expand_codes(const uint8_t* codes, int64_t n, int m, uint8_t* codes16, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    lm_expand_codes_kernel<<<grid, 256, 0, st>>>(codes, n, m, reinterpret_cast<uint4*>(codes16));
    return cudaGetLastError();
}

bool ivfpq_lm_supported(int dim, int m, int ksub, int nprobe, int k, int64_t nq, int64_t list_total, int64_t max_list_len) {
    // m < 16 runs on the equivalent 16-table quantiser (IvfPqScanParams::lm_codebook)
    if (m < 1 || m > 16 || 16 % m != 0 || dim % 16 != 0 || ksub > 256 || ksub < 1) return false;
    const int sub = dim / 16;
    if (sub != 4 && sub != 8) return false;
    if (k < 1 || k > kMaxTopK || (int64_t)nprobe * lm_kc(k) * lm_maxseg(max_list_len) > 16384 || nprobe > 32767) return false;
    if (nq * nprobe >= ((int64_t)1 << 29) || list_total >= ((int64_t)1 << 32)) return false;
    return true;
}

size_t ivfpq_lm_scratch_bytes(int64_t nq, int nprobe, int k, int nlist, int dim, int64_t max_list_len) {
    return lm_layout(nq, nprobe, k, nlist, dim, max_list_len).total;
}

// count, scan (2), fill, prepare, seed, scan, redo, final (+ the codeword bound when it is not cached)
int ivfpq_lm_launches() { return 9; }

// codes scored by the most recent list-major search that used `scratch` (sum of probed list lengths)
cudaError_t ivfpq_lm_scanned_codes(const void* scratch, int64_t nq, int nprobe, int k, int nlist, int dim,
                                   int64_t max_list_len, unsigned long long* out, cudaStream_t st) {
    const LmLayout L = lm_layout(nq, nprobe, k, nlist, dim, max_list_len);
    cudaError_t e = cudaMemcpyAsync(out, reinterpret_cast<const unsigned char*>(scratch) + L.scanned, sizeof(*out),
                                    cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

cudaError_t launch_pq_cmax(const float* codebook, int ksub, int sub, float* cmax16, cudaStream_t st) {
    lm_cmax_kernel<<<1, 256, 0, st>>>(codebook, ksub, sub, cmax16);
    return cudaGetLastError();
}

cudaError_t launch_ivfpq_scan_lm(const IvfPqScanParams& p, void* scratch, int num_sms, cudaStream_t st) {
    if (p.nq <= 0) return cudaSuccess;
    return (p.dim / 16 == 8) ? launch_lm<8>(p, scratch, num_sms, st) : launch_lm<4>(p, scratch, num_sms, st);
}

}  // namespace pyrope
