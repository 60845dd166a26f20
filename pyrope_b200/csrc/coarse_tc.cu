// coarse_tc.cu — the coarse probe of a batched IVF search on the 5th-gen tensor cores, query-tile stationary.
//
// Replaces, for a whole query batch, the centroid ranking of IvfFlatVectorIndex.Search (IvfFlatVectorIndex.cs:186-198)
// and IvfPqVectorIndex.Search (IvfPqVectorIndex.cs:141-150): score ALL centroids, sort descending, take nprobe.
// At BASELINE config 5 that is 10,000 x 65,536 scores of 128 dimensions per batch; the reported probe lists are
// ranked in the reference's own arithmetic (VectorMath.L2Squared / DotProduct / Cosine, one 8-lane accumulator),
// so they equal the oracle's bit for bit — the tensor cores only decide WHICH few centroids get that treatment.
//
// Why a second tcgen05 kernel beside flat_tc.cu: at d = 128 a tile has only 16 MMA k-steps, and the streaming
// kernel re-fetches its 128-query tile with every 256-row tile (192 KiB of L2 -> shared-memory traffic per 2,048
// MMA cycles = 26 TB/s chip-wide, twice what L2 delivers), so it was bound by operand traffic, and its pass B by a
// divergent per-lane filter.  Here:
//   * one CTA keeps a 256-query tile (two UMMA M = 128 halves) RESIDENT in shared memory and streams centroid tiles
//     of 128 rows (16 KiB per 128-byte K chunk, 5 TMA stages).  Operands are fp16 copies (kind::f16: the mantissa of
//     tf32 at twice the rate and half the bytes; L2 / IP, values within the fp16 range - launch_coarse_tc) or
//     tf32(x) copies (kind::tf32): one product per K slice either way, never the three-term split;
//   * accumulators: 2 halves x 2 buffers x 128 TMEM columns; eight epilogue warps (lane quarter x half) read them
//     with tcgen05.ld while the next unit's MMAs run;
//   * work units (query tile, centroid tile) are dealt to the 148 persistent CTAs as equal contiguous ranges (a
//     range may straddle two query tiles: the resident tile is swapped once), so there is no wave quantisation;
//   * pass A writes ONE number per (query, unit): the maximum proxy score of the unit's 128 centroids;
//     coarse_tau_kernel takes the k'-th largest of a query's unit maxima (each of the k' best units holds a row at
//     least that good) minus twice the rounding bound of the one-TF32 product as that query's threshold;
//   * pass B recomputes the same scores and appends the POSITIONS of everything above the threshold to the query's
//     candidate list (a 32-bit hit mask per 32 columns; lists private to one thread, plain stores; ~170 survivors
//     per query on C5's trained centroids);
//   * coarse_rank_kernel scores the survivors exactly, in the reference's order, sorts and writes the nprobe best.
//     A query whose list overflowed (thousands of centroids inside one rounding band) is ranked exhaustively.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "common.cuh"
#include "exact_arith.cuh"
#include "kernels.h"

namespace pyrope {
namespace {

constexpr int CQ = 256;       // queries per resident tile (two MMA halves of 128)
constexpr int CN = 128;       // centroids per work unit (UMMA N)
constexpr int CBK = 32;       // floats per K chunk = one 128-byte swizzle atom
constexpr int CKC = 4;        // K chunks held for the resident tile: dim <= 128
constexpr int XSTAGES = 5;
constexpr int C_THREADS = 384;  // TMA, MMA, TMEM-alloc, spare + 8 epilogue warps
constexpr int QCHUNK_BYTES = CQ * CBK * 4;   // 32 KiB
constexpr int XSTAGE_BYTES = CN * CBK * 4;   // 16 KiB
constexpr int C_TMEM_COLS = 512;
constexpr int TERM_FLOATS = 8 * 2 * 2 * CN;  // per epilogue warp: [2 buffers][bias CN | scale CN]
constexpr size_t C_SMEM = (size_t)CKC * QCHUNK_BYTES + (size_t)XSTAGES * XSTAGE_BYTES + TERM_FLOATS * sizeof(float) + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// One lane of a CONVERGED warp.  The single-thread roles (TMA producer, MMA issuer) must be entered through this and not
// through `lane == 0`: ptxas knows elect.sync yields exactly one lane and keeps descriptors in uniform registers, whereas
// under a plain divergent branch it wraps every UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop over the
// "possibly several" active threads — measured 166 cycles of issue per tcgen05.mma against a 64-cycle tensor-pipe floor.
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// the same without the wait: the caller overlaps the load of the next chunk with the arithmetic on this one
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B, 8-row atoms of 1024 bytes
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;  // stride between 8-row atoms
    d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t kCoarseIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
// kind::f16 with fp16 operands (format 0), fp32 accumulate: the same 10-bit mantissa as tf32 at twice the rate and half the
// operand bytes - a 128-byte swizzle atom holds 64 elements, one MMA consumes 16
constexpr uint32_t kCoarseIdescF16 = (1u << 4) | ((uint32_t)(CN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct CoarseParams {
    int64_t nq, n_scan;
    int dim;
    int64_t qtiles, ntiles, nq_pad;
    const float* scale;   // [n] per-centroid proxy terms: s = scale * (q.x) + bias
    const float* bias;
    float* umax;          // pass A out: [groups][nq_pad] maximum proxy score of a group: a whole unit (128 centroids) or,
    int fine;             //   fine = 1 (small tables: too few units to bound the k'-th best), each 32-column chunk of it
    const float* tau;     // pass B in: [nq_pad] accept s > tau
    // pass B out: survivors' centroid positions.  A query's units are spread over a few CTAs ("parts": the CTAs whose unit
    // ranges touch its query tile); each (query, part) list is private to ONE thread — plain stores, no atomics.
    uint32_t* qpos;       // [nq_pad][parts][cap]
    int32_t* qcnt;        // [nq_pad][parts] (zero-initialised; a count above cap marks an overflow)
    int cap, parts;
};

template <bool PASS_B, bool F16>
__global__ void __launch_bounds__(C_THREADS, 1)
coarse_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, CoarseParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* qs = smem;                                     // [CKC][CQ rows][128 B], swizzled
    uint8_t* xs = smem + CKC * QCHUNK_BYTES;                // [XSTAGES][CN rows][128 B]
    float* sterms = reinterpret_cast<float*>(xs + XSTAGES * XSTAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sterms + TERM_FLOATS);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * XSTAGES + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t xfull0 = smem_u32(&bars[0]), xempty0 = smem_u32(&bars[XSTAGES]);
    const uint32_t tfull0 = smem_u32(&bars[2 * XSTAGES]), tempty0 = smem_u32(&bars[2 * XSTAGES + 2]);
    const uint32_t qfull = smem_u32(&bars[2 * XSTAGES + 4]), qempty = smem_u32(&bars[2 * XSTAGES + 5]);

    // this CTA's contiguous range of work units; unit u = (query tile u / ntiles, centroid tile u % ntiles)
    const int64_t total = p.qtiles * p.ntiles;
    const int64_t u0 = total * (int64_t)blockIdx.x / gridDim.x, u1 = total * ((int64_t)blockIdx.x + 1) / gridDim.x;
    const int nunit = (int)(u1 - u0);
    constexpr int CBE = F16 ? 2 * CBK : CBK;  // elements per 128-byte K chunk
    constexpr int KI = F16 ? 16 : 8;          // elements one MMA consumes (32 bytes either way)
    const int KC = (p.dim + CBE - 1) / CBE;

    if (tid == 0) {
        for (int i = 0; i < XSTAGES; ++i) { mbar_init(xfull0 + 8 * i, 1); mbar_init(xempty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 8); }
        mbar_init(qfull, 1);
        mbar_init(qempty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(C_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    // Both single-thread roles walk the units with running counters (query tile, centroid tile, stage, phase): the MMA
    // issuer is ONE thread whose every instruction costs a full dependent latency, so its loop must hold nothing but the
    // barrier waits and the UTCHMMAs (a 64-bit division per unit and a modulo + three descriptor builds per K chunk made
    // it as slow as the tensor pipe: 3,900 cycles per unit against 2,048 of MMAs).
    const int64_t qt0 = u0 / p.ntiles;
    const int nt0 = (int)(u0 - qt0 * p.ntiles), ntiles_i = (int)p.ntiles;
    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            uint32_t s = 0, ph = 0, qloads = 0;
            int64_t qt = qt0;
            int nt = nt0;
            bool newq = true;
            for (int ui = 0; ui < nunit; ++ui) {
                if (newq) {  // swap the resident query tile: every MMA that read the old one has retired
                    if (qloads > 0) mbar_wait(qempty, (qloads - 1) & 1);
                    mbar_expect_tx(qfull, (uint32_t)(KC * QCHUNK_BYTES));
                    for (int kc = 0; kc < KC; ++kc)
                        tma_load_2d(smem_u32(qs + kc * QCHUNK_BYTES), &map_q, qfull, kc * CBE, (int)(qt * CQ));
                    ++qloads;
                    newq = false;
                }
                const int n0 = nt * CN;
                for (int kc = 0; kc < KC; ++kc) {
                    mbar_wait(xempty0 + 8 * s, ph ^ 1);
                    mbar_expect_tx(xfull0 + 8 * s, XSTAGE_BYTES);
                    tma_load_2d(smem_u32(xs) + s * XSTAGE_BYTES, &map_x, xfull0 + 8 * s, kc * CBE, n0);
                    if (++s == XSTAGES) { s = 0; ph ^= 1; }
                }
                if (++nt == ntiles_i) { nt = 0; ++qt; newq = true; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            uint32_t s = 0, ph = 0, qloads = 0;
            int nt = nt0;
            bool newq = true;
            // descriptors are base + offset: the stage / chunk offsets are added to the start-address field (16-byte units)
            const uint64_t xd_base = make_sw128_desc(smem_u32(xs));
            const uint64_t qd_base = make_sw128_desc(smem_u32(qs));
            constexpr uint64_t kXStep = XSTAGE_BYTES >> 4, kQStep = QCHUNK_BYTES >> 4, kHalf = (128 * 128) >> 4;
            for (int ui = 0; ui < nunit; ++ui) {
                if (newq) {
                    mbar_wait(qfull, qloads & 1);
                    ++qloads;
                    newq = false;
                }
                const uint32_t buf = ui & 1, aph = (ui >> 1) & 1;
                mbar_wait(tempty0 + 8 * buf, aph ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + buf * (2 * CN);
#pragma unroll
                for (int kc = 0; kc < CKC; ++kc) {
                    if (kc < KC) {
                        mbar_wait(xfull0 + 8 * s, ph);
                        tc_fence_after();
                        const uint64_t xd = xd_base + (uint64_t)s * kXStep;
                        const uint64_t qd0 = qd_base + (uint64_t)kc * kQStep, qd1 = qd0 + kHalf;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            if (kc * CBE + k4 * KI < p.dim) {
                                const uint64_t adv = (uint64_t)(k4 * 2);  // one MMA's K slice = 32 bytes = 2 x 16-byte units
                                if (F16) {
                                    tc_mma_f16(d0, qd0 + adv, xd + adv, kCoarseIdescF16, (kc | k4) != 0);
                                    tc_mma_f16(d0 + CN, qd1 + adv, xd + adv, kCoarseIdescF16, (kc | k4) != 0);
                                } else {
                                    tc_mma_tf32(d0, qd0 + adv, xd + adv, kCoarseIdesc, (kc | k4) != 0);
                                    tc_mma_tf32(d0 + CN, qd1 + adv, xd + adv, kCoarseIdesc, (kc | k4) != 0);
                                }
                            }
                        }
                        tc_commit(xempty0 + 8 * s);  // frees the centroid stage when these MMAs retire
                        if (++s == XSTAGES) { s = 0; ph ^= 1; }
                    }
                }
                tc_commit(tfull0 + 8 * buf);     // both halves of the unit complete
                if (++nt == ntiles_i) {          // last unit of this query tile: its retirement frees the resident tile
                    nt = 0;
                    newq = true;
                    tc_commit(qempty);
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: thread <-> query, warp <-> (TMEM lane quarter, query half) =================
        const int ew = (warp - 4) & 3, hf = (warp - 4) >> 2;
        float* sbias = sterms + (warp - 4) * (4 * CN);  // [2][CN] bias then [2][CN] scale, private to the warp
        float* sscale = sbias + 2 * CN;
        float nb[CN / 32], ns[CN / 32];
        auto load_terms = [&](int nt) {
#pragma unroll
            for (int j = 0; j < CN / 32; ++j) {
                const int64_t pos = nt * CN + lane + j * 32;
                nb[j] = -INFINITY; ns[j] = 0.f;
                if (pos < p.n_scan) { nb[j] = __ldg(p.bias + pos); ns[j] = __ldg(p.scale + pos); }
            }
        };
        auto store_terms = [&](uint32_t buf) {
#pragma unroll
            for (int j = 0; j < CN / 32; ++j) {
                sbias[buf * CN + lane + j * 32] = nb[j];
                sscale[buf * CN + lane + j * 32] = ns[j];
            }
            __syncwarp();
        };
        if (nunit > 0) { load_terms(nt0); store_terms(0); }
        int64_t qt = qt0;
        int nt = nt0;
        int64_t cur_qt = -1, gq = 0;
        bool qvalid = false;
        float tau = INFINITY;
        int cnt = 0;
        uint32_t* myq = nullptr;
        int32_t* mycnt = nullptr;
        for (int ui = 0; ui < nunit; ++ui) {
            if (qt != cur_qt) {
                if (PASS_B && mycnt && qvalid) *mycnt = cnt;  // close the list of the query tile this CTA leaves
                cur_qt = qt;
                gq = qt * CQ + hf * 128 + ew * 32 + lane;
                qvalid = gq < p.nq;
                if (PASS_B) {
                    tau = qvalid ? __ldg(p.tau + gq) : INFINITY;
                    // part = this CTA's rank among the CTAs whose ranges touch query tile qt (the first one holds unit qt * ntiles)
                    const int64_t first_cta = ((qt * p.ntiles + 1) * (int64_t)gridDim.x + total - 1) / total - 1;
                    const int part = (int)((int64_t)blockIdx.x - first_cta);
                    if (part >= 0 && part < p.parts) {
                        myq = p.qpos + ((size_t)gq * p.parts + part) * p.cap;
                        mycnt = p.qcnt + (size_t)gq * p.parts + part;
                        cnt = 0;
                    } else {  // cannot happen with the launcher's bound on parts; if it did, the query is ranked exhaustively
                        myq = p.qpos + (size_t)gq * p.parts * p.cap;
                        mycnt = p.qcnt + (size_t)gq * p.parts;
                        cnt = p.cap + 1;
                    }
                }
            }
            const uint32_t buf = ui & 1, aph = (ui >> 1) & 1;
            const int nt_next = nt + 1 == ntiles_i ? 0 : nt + 1;
            if (ui + 1 < nunit) load_terms(nt_next);  // consumed after this unit
            mbar_wait(tfull0 + 8 * buf, aph);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * (2 * CN) + hf * CN;
            float umax = -INFINITY;
            uint32_t rbuf[2][32];  // chunk c+1 streams out of TMEM while chunk c is filtered
            tc_ld32_issue(taddr0, rbuf[0]);
            tc_ld_wait();
#pragma unroll
            for (int c = 0; c < CN / 32; ++c) {
                if (c + 1 < CN / 32) tc_ld32_issue(taddr0 + (c + 1) * 32, rbuf[(c + 1) & 1]);
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rbuf[c & 1][i]);
                const float4* b4 = reinterpret_cast<const float4*>(sbias + buf * CN + c * 32);
                const float4* s4 = reinterpret_cast<const float4*>(sscale + buf * CN + c * 32);
                uint32_t mask = 0u;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = b4[j4], ss = s4[j4];
                    const float s0 = fmaf(v[4 * j4 + 0], ss.x, bb.x), s1 = fmaf(v[4 * j4 + 1], ss.y, bb.y);
                    const float s2 = fmaf(v[4 * j4 + 2], ss.z, bb.z), s3 = fmaf(v[4 * j4 + 3], ss.w, bb.w);
                    if (PASS_B) {
                        mask |= (s0 > tau ? 1u : 0u) << (4 * j4 + 0);
                        mask |= (s1 > tau ? 1u : 0u) << (4 * j4 + 1);
                        mask |= (s2 > tau ? 1u : 0u) << (4 * j4 + 2);
                        mask |= (s3 > tau ? 1u : 0u) << (4 * j4 + 3);
                    } else {
                        umax = fmaxf(fmaxf(umax, fmaxf(s0, s1)), fmaxf(s2, s3));
                    }
                }
                if (!PASS_B && p.fine) {
                    if (qvalid) p.umax[(nt * (CN / 32) + c) * p.nq_pad + gq] = umax;
                    umax = -INFINITY;
                }
                if (PASS_B) {
                    while (mask) {  // rare: ~1 survivor per 600 scores
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1u;
                        if (cnt < p.cap) myq[cnt] = (uint32_t)(nt * CN + c * 32 + j);
                        ++cnt;
                    }
                }
                if (c + 1 < CN / 32) tc_ld_wait();
            }
            if (!PASS_B && !p.fine && qvalid) p.umax[nt * p.nq_pad + gq] = umax;  // coalesced: consecutive lanes, consecutive queries
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            if (ui + 1 < nunit) store_terms(buf ^ 1);
            nt = nt_next;
            if (nt == 0) ++qt;
        }
        if (PASS_B && mycnt && qvalid) *mycnt = cnt;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C_TMEM_COLS) : "memory");
    }
}

// ---- per-query threshold: the k'-th largest unit maximum, lowered by the rounding band --------------------
// Scores are one-TF32 products: each operand rounded to 11 significant bits, so |score - exact proxy| <= E with
// E = 2^-10 (1 + 2^-7) |q| max_r(|scale_r| |x_r|)  (Cauchy-Schwarz).  There are at least k' rows scoring >= G (the k'-th
// unit maximum) in this arithmetic, hence exactly >= G - E; a row of the exact top k' therefore scores >= G - 2E here.
// 32 queries per block: the unit maxima arrive [unit][query] (coalesced), are transposed through shared memory and
// each of the 16 warps runs a register-resident bisection for 2 queries (the kernel's time is that serial chain: with 4
// warps x 8 queries it cost 0.047 ms whatever the batch size).
constexpr int TAU_MAXU = 1024;  // units per query handled in registers (32 per lane): 131,072 centroids
template <int R>  // registers per lane: groups <= 32 R
__global__ void __launch_bounds__(512) coarse_tau_kernel(const float* __restrict__ umax, int64_t nq_pad, int ntiles, int64_t nq,
                                                        int kprime, const float* __restrict__ Q, int dim,
                                                        const float* __restrict__ amax, float* tau_out,
                                                        float eband, float under,
                                                        float smax, const uint8_t* __restrict__ qbad) {
    extern __shared__ float s_u[];  // [32][ntiles + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * 32;
    const int ld = ntiles + 1;
    for (int i = tid; i < 32 * ntiles; i += blockDim.x) {
        const int t = i >> 5, ql = i & 31;
        s_u[ql * ld + t] = (q0 + ql < nq) ? __ldg(umax + (int64_t)t * nq_pad + q0 + ql) : -INFINITY;
    }
    __syncthreads();
    for (int ql = warp; ql < 32; ql += (int)(blockDim.x >> 5)) {
        const int64_t q = q0 + ql;
        if (q >= nq) break;
        float tau = -INFINITY;
        if (ntiles > kprime) {
            const int nr = (ntiles + 31) >> 5;  // occupied registers per lane (warp-uniform)
            uint32_t o[R];
            uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = r * 32 + lane;
                o[r] = (r < nr && i < ntiles) ? score_to_ord(s_u[ql * ld + i]) : 0u;
                if (r < nr && i < ntiles) { lo = min(lo, o[r]); hi = max(hi, o[r]); }
            }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            while (lo < hi) {  // largest T with count(ord >= T) >= k'
                const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
                int n = 0;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (r < nr) n += o[r] >= mid;
                n = __reduce_add_sync(0xffffffffu, n);
                if (n == kprime) {  // exactly k' values at or above mid: the k'-th largest is the smallest of them
                    uint32_t mn = 0xffffffffu;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r < nr && o[r] >= mid) mn = min(mn, o[r]);
                    lo = hi = __reduce_min_sync(0xffffffffu, mn);
                    break;
                }
                if (n > kprime) lo = mid; else hi = mid - 1u;
            }
            const float G = ord_to_score(lo);
            float qq = 0.f;
            for (int d = lane; d < dim; d += 32) { const float v = __ldg(Q + q * dim + d); qq = fmaf(v, v, qq); }
            qq = warp_sum(qq);
            // 2^-10 (1 + 2^-7) |q| A; fp16 operands add the absolute error of values below 2^-14 (<= 2^-25 each)
            const float E = eband * sqrtf(qq) * __ldg(amax) + under * (smax * sqrtf(qq) + __ldg(amax));
            // accept s > tau: everything >= G - 2E, with room for the fp32 rounding of the proxy itself and for the
            // difference between exact arithmetic and the reference's evaluation order (both ~1e-6 relative)
            tau = G - 2.f * E - 4e-6f * fabsf(G) - 1e-30f;
            tau = fminf(tau, G);
            tau = tau > -INFINITY ? nextafterf(tau, -INFINITY) : tau;
            if (qbad && qbad[q]) tau = -INFINITY;  // a component beyond the fp16 range: no bound holds, rank everything exactly
        }
        if (lane == 0) tau_out[q] = tau;
    }
}

// ---- exact ranking of the survivors, in the reference's arithmetic ---------------------------------------------
// VectorMath.L2Squared / DotProduct (VectorMath.cs:8-70): one 8-lane accumulator stepping 8 elements, pairwise
// horizontal sum ((v0+v1)+(v2+v3))+((v4+v5)+(v6+v7)), scalar tail; separate multiply and add.  FOUR consecutive lanes share
// a candidate: lane t owns accumulator lanes 2t and 2t+1 (one 64-bit load per step, all 16 loads of a d = 128 row in
// flight at once), so the first level of the horizontal tree is a local add and two shuffles finish it.
template <int OP>
__device__ __forceinline__ float a2_eval_quad(const float* q, const float* __restrict__ x, int n, int t) {
    int i = 0;
    float sum = 0.f;
    if (n >= 8) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 16
        for (; i <= n - 8; i += 8) {
            const float2 xv = __ldg(reinterpret_cast<const float2*>(x + i + 2 * t));
            const float2 qv = *reinterpret_cast<const float2*>(q + i + 2 * t);
            a0 = __fadd_rn(a0, exact::term<OP>(qv.x, xv.x));
            a1 = __fadd_rn(a1, exact::term<OP>(qv.y, xv.y));
        }
        float acc = __fadd_rn(a0, a1);
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
        sum = __fadd_rn(sum, acc);
    }
    for (; i < n; ++i) sum = __fadd_rn(sum, exact::term<OP>(q[i], __ldg(x + i)));
    return sum;
}

constexpr int RANK_KEYS = 1024;  // sort window: survivors (<= 512) or, exhaustively, 512 kept + 512 new (nprobe <= 256)
constexpr int RANK_SPOS = 512;   // survivor positions staged in shared memory
constexpr int RANK_MAXPARTS = 160;
template <int METRIC>
__global__ void __launch_bounds__(128) coarse_rank_kernel(const float* __restrict__ Q, int dim, const float* __restrict__ C,
                                                         const float* __restrict__ cnorms, int64_t nc,
                                                         const uint32_t* __restrict__ qpos, const int32_t* __restrict__ qcnt, int cap,
                                                         int parts, int64_t* pout, int P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);     // [RANK_KEYS]
    uint32_t* spos = reinterpret_cast<uint32_t*>(keys + RANK_KEYS);  // [RANK_SPOS]
    float* qv = reinterpret_cast<float*>(spos + RANK_SPOS);     // [dim]
    __shared__ float s_qn;
    __shared__ int s_off[RANK_MAXPARTS + 1];
    __shared__ int s_over;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, t = tid & 3, grp = tid >> 2, ngrp = blockDim.x >> 2;
    for (int d = tid; d < dim; d += blockDim.x) qv[d] = Q[q * dim + d];
    // offsets of the (few) per-CTA lists of this query: counts loaded in parallel, summed by one thread from shared memory
    if (tid == 0) s_over = 0;
    __syncthreads();
    for (int pt = tid; pt < parts; pt += blockDim.x) {
        const int c = qcnt[q * parts + pt];
        if (c > cap) s_over = 1;
        s_off[pt + 1] = min(c, cap);
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        s_off[0] = 0;
        for (int pt = 0; pt < parts; ++pt) { o += s_off[pt + 1]; s_off[pt + 1] = o; }
        if (o > RANK_SPOS) s_over = 1;
    }
    __syncthreads();
    if (METRIC == 2 && tid == 0) s_qn = exact::norm_eval(qv, dim);
    const int have = s_off[parts];
    const bool exhaustive = s_over != 0;
    if (!exhaustive)
        for (int pt = 0; pt < parts; ++pt)
            for (int i = s_off[pt] + tid; i < s_off[pt + 1]; i += blockDim.x) spos[i] = qpos[((size_t)q * parts + pt) * cap + (i - s_off[pt])];
    __syncthreads();
    auto score = [&](int64_t l) -> float {  // the four lanes of a group call this with the same l
        const float* cv = C + l * dim;
        if (METRIC == 0) return -a2_eval_quad<0>(qv, cv, dim, t);
        if (METRIC == 1) return a2_eval_quad<1>(qv, cv, dim, t);
        const float d = a2_eval_quad<1>(qv, cv, dim, t);
        const float cn = cnorms[l];
        return (s_qn < 1e-6f || cn < 1e-6f) ? 0.f : __fdiv_rn(d, __fmul_rn(s_qn, cn));
    };
    if (!exhaustive) {
        const int P2 = next_pow2(max(max(have, P), 2));  // the output loop reads P keys: pad with empties
        for (int i = have + tid; i < P2; i += blockDim.x) keys[i] = 0ull;
        for (int i0 = 0; i0 < have; i0 += ngrp) {  // uniform trip count: the shuffles inside score() need whole warps
            const int i = i0 + grp;
            const bool on = i < have;
            const uint32_t l = spos[on ? i : i0];
            const float sc = score(l);
            if (t == 0 && on) keys[i] = make_key(sc, l);
        }
        __syncthreads();
        bitonic_sort_desc<false>(keys, P2, tid, blockDim.x);
    } else {
        // a candidate list overflowed (a crowd of centroids inside one rounding band): rank every centroid
        for (int i = tid; i < RANK_KEYS; i += blockDim.x) keys[i] = 0ull;
        __syncthreads();
        for (int64_t c0 = 0; c0 < nc; c0 += RANK_KEYS / 2) {
            for (int i0 = 0; i0 < RANK_KEYS / 2; i0 += ngrp) {
                const int64_t l = c0 + i0 + grp;
                const bool on = l < nc;
                const float sc = score(on ? l : 0);
                if (t == 0) keys[RANK_KEYS / 2 + i0 + grp] = on ? make_key(sc, (uint32_t)l) : 0ull;
            }
            __syncthreads();
            bitonic_sort_desc<false>(keys, RANK_KEYS, tid, blockDim.x);  // best 1024 so far end up in the front half
        }
    }
    for (int i = tid; i < P; i += blockDim.x) {
        const uint64_t key = keys[i];
        pout[q * P + i] = key ? (int64_t)key_pos(key) : -1;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}
bool make_map(CUtensorMap* m, const void* base, int64_t rows, int dim, int box_rows, bool f16 = false) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)dim * (f16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)(f16 ? 2 * CBK : CBK), (cuuint32_t)box_rows};  // 128 bytes wide either way
    cuuint32_t estr[2] = {1, 1};
    return fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// groups whose maxima bound the threshold: whole units when there are at least 2 k' of them, else 32-column chunks
int coarse_groups(int64_t nc, int kprime, int* fine) {
    const int64_t ntiles = (nc + CN - 1) / CN;
    if (ntiles >= 2 * (int64_t)kprime && ntiles <= TAU_MAXU) { *fine = 0; return (int)ntiles; }
    const int64_t g = ntiles * (CN / 32);
    if (g >= 2 * (int64_t)kprime && g <= TAU_MAXU) { *fine = 1; return (int)g; }
    return 0;
}

struct CoarseLayout { size_t umax, tau, qcnt, qpos, total; int64_t nq_pad, qtiles, ntiles; int grid, parts, cap; };
CoarseLayout coarse_layout(int64_t nq, int64_t nc, int num_sms) {
    CoarseLayout L{};
    L.qtiles = (nq + CQ - 1) / CQ;
    L.nq_pad = L.qtiles * CQ;
    L.ntiles = (nc + CN - 1) / CN;
    L.grid = (int)std::min<int64_t>(num_sms, L.qtiles * L.ntiles);
    // CTAs whose contiguous unit ranges can touch one query tile
    L.parts = (int)std::min<int64_t>(L.grid, (L.grid + L.qtiles - 1) / L.qtiles + 2);
    // survivors per (query, part): ~110 per query in all at C5; room for skew, the whole budget when one CTA holds the tile
    L.cap = L.parts <= 2 ? kCoarseTcCap : kCoarseTcCap / 2;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t o = 0;
    L.qcnt = o; o += al(sizeof(int32_t) * (size_t)L.nq_pad * L.parts);
    L.tau = o; o += al(sizeof(float) * (size_t)L.nq_pad);
    L.umax = o; o += al(sizeof(float) * (size_t)std::min<int64_t>(L.ntiles * (CN / 32), std::max<int64_t>(L.ntiles, TAU_MAXU)) * (size_t)L.nq_pad);
    L.qpos = o; o += al(sizeof(uint32_t) * (size_t)L.nq_pad * L.parts * L.cap);
    L.total = o;
    return L;
}

// ---- fp16 operand copies ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t f2h_sat(float x) {
    uint16_t h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    return h;
}
__global__ void __launch_bounds__(256) tc_half_kernel(const float* __restrict__ X, int64_t n, uint16_t* __restrict__ out, float* absmax) {
    float mx = 0.f;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (int64_t)gridDim.x * blockDim.x * 4) {
        if (i + 3 < n) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(X + i));
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
            const uint32_t lo = (uint32_t)f2h_sat(v.x) | ((uint32_t)f2h_sat(v.y) << 16);
            const uint32_t hi = (uint32_t)f2h_sat(v.z) | ((uint32_t)f2h_sat(v.w) << 16);
            *reinterpret_cast<uint2*>(out + i) = make_uint2(lo, hi);
        } else {
            for (int64_t j = i; j < n; ++j) { mx = fmaxf(mx, fabsf(X[j])); out[j] = f2h_sat(X[j]); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // NaN compares false everywhere above: it never raises the maximum, and its fp16 copy is NaN (scores NaN, as the
    // reference's own arithmetic would give)
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(reinterpret_cast<unsigned int*>(absmax), __float_as_uint(mx));
}
__global__ void __launch_bounds__(256) tc_half_rows_kernel(const float* __restrict__ Q, int64_t nq, int dim, uint16_t* __restrict__ out,
                                                           uint8_t* __restrict__ bad) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= nq) return;
    float mx = 0.f;
    for (int d = lane; d < dim; d += 32) {
        const float v = __ldg(Q + r * dim + d);
        mx = fmaxf(mx, fabsf(v));
        out[r * dim + d] = f2h_sat(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) bad[r] = !(mx <= kTcHalfMaxAbs);
}

}  // namespace

// The unit maxima only bound the k'-th best score when there are comfortably more units than k'; smaller tables and
// dim > 128 stay on the streaming kernel (flat_tc.cu).
cudaError_t launch_tc_half(const float* X, int64_t n_elems, void* out16, float* absmax, cudaStream_t st) {
    if (n_elems <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<int64_t>((n_elems / 4 + 255) / 256 + 1, 148 * 32);
    tc_half_kernel<<<grid, 256, 0, st>>>(X, n_elems, static_cast<uint16_t*>(out16), absmax);
    return cudaGetLastError();
}
cudaError_t launch_tc_half_rows(const float* Q, int64_t nq, int dim, void* out16, uint8_t* bad, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    tc_half_rows_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(Q, nq, dim, static_cast<uint16_t*>(out16), bad);
    return cudaGetLastError();
}

bool coarse_tc_supported(int dim, int64_t nc, int nprobe) {
    int fine = 0;
    return dim % 4 == 0 && dim >= 8 && dim <= CKC * CBK && nprobe >= 1 && nprobe <= kCoarseTcCap / 2 &&
           coarse_groups(nc, nprobe + kCoarseTcMargin, &fine) > 0 && nc < ((int64_t)1 << 31) && !getenv("PYROPE_COARSE_STREAMING");
}
size_t coarse_tc_scratch_bytes(int64_t nq, int64_t nc, int num_sms) { return coarse_layout(nq, nc, num_sms).total; }
int coarse_tc_launches() { return 5; }  // clear, pass A, threshold, pass B, exact ranking

cudaError_t launch_coarse_tc(const CoarseTcParams& a, cudaStream_t st) {
    if (a.nq <= 0) return cudaSuccess;
    const CoarseLayout L = coarse_layout(a.nq, a.nc, a.num_sms);
    if (L.parts > RANK_MAXPARTS) return cudaErrorInvalidValue;
    unsigned char* base = reinterpret_cast<unsigned char*>(a.scratch);
    CUtensorMap mq, mx;
    // (Feeding the fp32 table itself to kind::tf32 - truncation, band 1.5x - was measured: 45% more survivors on the
    // C5 centroids, and the exact ranking gathers them from HBM either way: 0.35 ms instead of 0.28.)
    const float eband = 9.85e-4f;  // 2^-10 (1 + 2^-7): both operands rounded to nearest tf32 - or to nearest fp16
    const bool f16 = a.Q16 && a.C16;
    // fp16: a value below 2^-14 is rounded with an ABSOLUTE error of at most 2^-25; summed over the row,
    // sum_i (|q_i| + |c_i|) 2^-25 <= 2^-25 sqrt(d) (|q| + |c|)   (twice that, for the products of two such errors and slack)
    const float under = f16 ? 5.97e-8f * sqrtf((float)a.dim) : 0.f;
    const float smax = a.metric == kL2 ? 2.f : 1.f;  // |scale| of the proxy score (launch_tc_prepare)
    if (f16 ? (!make_map(&mq, a.Q16, a.nq, a.dim, CQ, true) || !make_map(&mx, a.C16, a.nc, a.dim, CN, true))
            : (!make_map(&mq, a.Qhi, a.nq, a.dim, CQ) || !make_map(&mx, a.Chi, a.nc, a.dim, CN)))
        return cudaErrorInvalidValue;
    CoarseParams p{};
    p.nq = a.nq; p.n_scan = a.nc; p.dim = a.dim; p.qtiles = L.qtiles; p.ntiles = L.ntiles; p.nq_pad = L.nq_pad;
    p.scale = a.scale; p.bias = a.bias;
    p.umax = reinterpret_cast<float*>(base + L.umax);
    p.tau = reinterpret_cast<float*>(base + L.tau);
    p.qpos = reinterpret_cast<uint32_t*>(base + L.qpos);
    p.qcnt = reinterpret_cast<int32_t*>(base + L.qcnt);
    p.cap = L.cap; p.parts = L.parts;
    cudaError_t e = cudaMemsetAsync(p.qcnt, 0, sizeof(int32_t) * (size_t)L.nq_pad * L.parts, st);
    if (e != cudaSuccess) return e;
    auto* passA = f16 ? coarse_tc_kernel<false, true> : coarse_tc_kernel<false, false>;
    auto* passB = f16 ? coarse_tc_kernel<true, true> : coarse_tc_kernel<true, false>;
    e = cudaFuncSetAttribute(passA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(passB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM);
    if (e != cudaSuccess) return e;
    static const bool dbg = getenv("PYROPE_COARSE_DEBUG") != nullptr;
    cudaEvent_t dev[5];
    auto mark = [&](int i) { if (dbg) { cudaEventCreate(&dev[i]); cudaEventRecord(dev[i], st); } };
    const unsigned grid = (unsigned)L.grid;
    const int kprime = a.nprobe + kCoarseTcMargin;
    const int ngroups = coarse_groups(a.nc, kprime, &p.fine);
    if (ngroups <= 0) return cudaErrorInvalidValue;
    const size_t tsm = sizeof(float) * 32 * ((size_t)ngroups + 1);
    mark(0);
    passA<<<grid, C_THREADS, C_SMEM, st>>>(mq, mx, p);
    mark(1);
    if (ngroups <= 512) {
        e = cudaFuncSetAttribute(coarse_tau_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        if (e != cudaSuccess) return e;
        coarse_tau_kernel<16><<<(unsigned)((a.nq + 31) / 32), 512, tsm, st>>>(
            p.umax, L.nq_pad, ngroups, a.nq, kprime, a.Q, a.dim, a.amax, reinterpret_cast<float*>(base + L.tau),
            eband, under, smax, f16 ? a.qbad : nullptr);
    } else {
        e = cudaFuncSetAttribute(coarse_tau_kernel<TAU_MAXU / 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        if (e != cudaSuccess) return e;
        coarse_tau_kernel<TAU_MAXU / 32><<<(unsigned)((a.nq + 31) / 32), 512, tsm, st>>>(
            p.umax, L.nq_pad, ngroups, a.nq, kprime, a.Q, a.dim, a.amax, reinterpret_cast<float*>(base + L.tau),
            eband, under, smax, f16 ? a.qbad : nullptr);
    }
    mark(2);
    passB<<<grid, C_THREADS, C_SMEM, st>>>(mq, mx, p);
    mark(3);
    const size_t rsm = sizeof(uint64_t) * RANK_KEYS + sizeof(uint32_t) * RANK_SPOS + sizeof(float) * (size_t)a.dim;
    if (a.metric == kL2)
        coarse_rank_kernel<0><<<(unsigned)a.nq, 128, rsm, st>>>(a.Q, a.dim, a.C, a.cnorms, a.nc, p.qpos, p.qcnt, p.cap, p.parts, a.probes_out, a.nprobe);
    else if (a.metric == kIP)
        coarse_rank_kernel<1><<<(unsigned)a.nq, 128, rsm, st>>>(a.Q, a.dim, a.C, a.cnorms, a.nc, p.qpos, p.qcnt, p.cap, p.parts, a.probes_out, a.nprobe);
    else
        coarse_rank_kernel<2><<<(unsigned)a.nq, 128, rsm, st>>>(a.Q, a.dim, a.C, a.cnorms, a.nc, p.qpos, p.qcnt, p.cap, p.parts, a.probes_out, a.nprobe);
    mark(4);
    if (dbg) {  // per-kernel times and survivors per query (debug aid, synchronises)
        cudaEventSynchronize(dev[4]);
        float ms[4];
        for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&ms[i], dev[i], dev[i + 1]);
        fprintf(stderr, "[coarse] pass A %.3f ms, threshold %.3f, pass B %.3f, exact ranking %.3f\n", ms[0], ms[1], ms[2], ms[3]);
        for (int i = 0; i < 5; ++i) cudaEventDestroy(dev[i]);
        std::vector<int32_t> hc((size_t)a.nq * p.parts);
        cudaStreamSynchronize(st);
        cudaMemcpy(hc.data(), p.qcnt, sizeof(int32_t) * hc.size(), cudaMemcpyDeviceToHost);
        long long sum = 0, mx = 0, over = 0;
        for (int64_t q = 0; q < a.nq; ++q) {
            long long tq = 0;
            bool ov = false;
            for (int pt = 0; pt < p.parts; ++pt) { const int c = hc[(size_t)q * p.parts + pt]; tq += c; ov |= c > p.cap; }
            sum += tq; mx = std::max(mx, tq); over += ov || tq > 1024;
        }
        fprintf(stderr, "[coarse] survivors per query: mean %.1f max %lld, %lld of %lld queries ranked exhaustively (parts %d x cap %d), groups %d%s\n",
                (double)sum / (double)a.nq, mx, over, (long long)a.nq, p.parts, p.cap, ngroups, p.fine ? " (32-column)" : "");
    }
    return cudaGetLastError();
}

}  // namespace pyrope
