// ivf.cu — K4 IVF_FLAT inverted-list scan with fused top-k, and the MaxScans budget kernel.
//
// Replaces IvfFlatVectorIndex.Search:200-218 (per probed list, per item: ComputeScore + heap).
// HBM-bound: one CTA per (query, probe group); each warp streams rows with 128-bit loads (a row of
// d floats is read by 32 lanes x float4, fully coalesced), the query sits in shared memory, and the
// CTA keeps one threshold-filtered candidate queue (common.cuh CtaQueue) across all its probes.
#include "common.cuh"
#include "kernels.h"

namespace pyrope {
namespace {

constexpr int NT = 256;
constexpr int CH = 512;      // rows per capacity check
constexpr int QCAP = 2048;   // queue capacity (k <= 1024)

template <int METRIC>
__global__ void __launch_bounds__(NT) ivfflat_scan_kernel(IvfFlatScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);            // [QCAP]
    float* qv = reinterpret_cast<float*>(keys + QCAP);                  // [dim]
    __shared__ int s_cnt;
    __shared__ uint64_t s_thr;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const int y = blockIdx.y;
    const int dim = p.dim;
    const bool vec_ok = (dim % 4 == 0);

    for (int i = tid; i < dim; i += NT) qv[i] = p.Q[q * dim + i];
    CtaQueue Qu{keys, &s_cnt, &s_thr, QCAP, p.k};
    Qu.reset(tid);
    __syncthreads();
    const float qn = (METRIC == kCosine) ? p.qnorm[q] : 0.f;

    uint64_t thr_reg = 0;
    int cnt_known = 0;  // conservative upper bound of *cnt, uniform across the CTA

    for (int pr = y; pr < p.nprobe; pr += p.groups) {
        const int64_t l = p.probes[q * p.nprobe + pr];
        if (l < 0) continue;
        int64_t beg = p.list_off[l], end = p.list_off[l + 1];
        if (p.allow) {
            int64_t a = p.allow[q * p.nprobe + pr];
            if (beg + a < end) end = beg + a;
        }
        for (int64_t c0 = beg; c0 < end; c0 += CH) {
            const int64_t cend = min(end, c0 + CH);
            if (cnt_known + CH > QCAP || (thr_reg == 0 && cnt_known >= 2 * p.k && cnt_known >= 64)) {
                __syncthreads();
                int actual = s_cnt;
                if (actual + CH > QCAP || (s_thr == 0 && actual >= 2 * p.k && actual >= 64)) {
                    Qu.prune(tid, NT);
                    actual = s_cnt;
                }
                __syncthreads();
                cnt_known = actual;
                thr_reg = s_thr;
            }
            // each warp takes groups of 4 consecutive rows
            for (int64_t r0 = c0 + warp * 4; r0 < cend; r0 += (NT / 32) * 4) {
                float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t r = r0 + u;
                    if (r >= cend) break;
                    const float* x = p.vecs + r * dim;
                    float a = 0.f;
                    if (vec_ok) {
                        for (int c = lane * 4; c < dim; c += 128) {
                            float4 xv = __ldg(reinterpret_cast<const float4*>(x + c));
                            float4 qq = *reinterpret_cast<const float4*>(qv + c);
                            if (METRIC == kL2) {
                                float d0 = qq.x - xv.x, d1 = qq.y - xv.y, d2 = qq.z - xv.z, d3 = qq.w - xv.w;
                                a = fmaf(d0, d0, a); a = fmaf(d1, d1, a); a = fmaf(d2, d2, a); a = fmaf(d3, d3, a);
                            } else {
                                a = fmaf(qq.x, xv.x, a); a = fmaf(qq.y, xv.y, a);
                                a = fmaf(qq.z, xv.z, a); a = fmaf(qq.w, xv.w, a);
                            }
                        }
                    } else {
                        for (int c = lane; c < dim; c += 32) {
                            float xv = __ldg(x + c), qq = qv[c];
                            if (METRIC == kL2) { float d = qq - xv; a = fmaf(d, d, a); }
                            else a = fmaf(qq, xv, a);
                        }
                    }
                    s[u] = a;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) s[u] = warp_sum(s[u]);
                if (lane < 4) {
                    const int64_t r = r0 + lane;
                    float v = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
                    if (r < cend && !(p.dead && p.dead[r])) {
                        float score;
                        if (METRIC == kL2) score = -v;
                        else if (METRIC == kIP) score = v;
                        else {
                            float xn = p.norms[r];
                            score = (qn < 1e-6f || xn < 1e-6f) ? 0.f : v / (qn * xn);
                        }
                        uint64_t key = make_key(score, (uint32_t)r);
                        if (key > thr_reg) {
                            int pos = atomicAdd(&s_cnt, 1);
                            if (pos < QCAP) keys[pos] = key;
                        }
                    }
                }
            }
            cnt_known += (int)(cend - c0);
        }
    }
    Qu.prune(tid, NT);
    const int keep = s_cnt;
    const int64_t ob = (q * p.out.parts_total + p.out.part_base + y) * (int64_t)p.k;
    for (int i = tid; i < p.k; i += NT) {
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

// IvfFlatVectorIndex.cs:172,202,209: `scanned` counts buffer rows first, then list rows in probe
// order; a probe may consume only what is left of MaxScans.  budget = MaxScans - buffer rows scanned.
__global__ void probe_allow_kernel(const int64_t* probes, int64_t nq, int nprobe,
                                   const int64_t* list_off, int64_t budget, int32_t* allow) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int64_t left = budget < 0 ? 0 : budget;
    for (int pr = 0; pr < nprobe; ++pr) {
        int64_t l = probes[q * nprobe + pr];
        int64_t a = 0;
        if (l >= 0) {
            int64_t len = list_off[l + 1] - list_off[l];
            a = len < left ? len : left;
            left -= a;
        }
        allow[q * nprobe + pr] = (int32_t)a;
    }
}

}  // namespace

cudaError_t launch_ivfflat_scan(const IvfFlatScanParams& p, cudaStream_t st) {
    if (p.nq <= 0) return cudaSuccess;
    size_t smem = sizeof(uint64_t) * QCAP + sizeof(float) * (size_t)((p.dim + 3) / 4 * 4);
    dim3 grid((unsigned)p.nq, (unsigned)p.groups);
    cudaError_t e = cudaSuccess;
    switch (p.metric) {
        case kL2:
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(ivfflat_scan_kernel<kL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            ivfflat_scan_kernel<kL2><<<grid, NT, smem, st>>>(p);
            break;
        case kIP:
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(ivfflat_scan_kernel<kIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            ivfflat_scan_kernel<kIP><<<grid, NT, smem, st>>>(p);
            break;
        default:
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(ivfflat_scan_kernel<kCosine>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            ivfflat_scan_kernel<kCosine><<<grid, NT, smem, st>>>(p);
            break;
    }
    return cudaGetLastError();
}

cudaError_t launch_probe_allow(const int64_t* probes, int64_t nq, int nprobe, const int64_t* list_off,
                               int64_t budget, int32_t* allow, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    probe_allow_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(probes, nq, nprobe, list_off, budget, allow);
    return cudaGetLastError();
}

}  // namespace pyrope
