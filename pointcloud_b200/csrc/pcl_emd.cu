// pcl_emd.cu -- auction-based approximate Earth Mover's Distance for sm_100a.
//
// Replaces the reference's 7*iters+1 launches (pointcloud_vision/loss/emd/emd_cuda.cu:256-269:
// clear, calc_unass_cnt, calc_unass_cnt_sum, calc_unass_idx, Bid, GetMax, Assign, CalcDist) with ONE
// persistent kernel.  A cloud is owned by a thread-block cluster of CS CTAs (CS in {1,2,4,8,16}, chosen so
// that B*CS fills the 148 SMs).  Every CTA keeps the whole auction state of its cloud in shared memory
// (targets, prices, assignment, inverse assignment) as a REPLICA:
//   1. every CTA compacts the list of unassigned bidders from its replica (identical in all CTAs);
//   2. the bidders are split over the CTAs of the cluster; inside a CTA one thread scans one
//      (bidder, target-chunk) item -- lanes of a warp share the chunk, so the target tile is read with
//      broadcast LDS.128 -- and chunk partials are merged exactly like emd_cuda.cu:165-173;
//   3. finished bids (object, increment) are written into EVERY CTA's bid arrays through distributed
//      shared memory, followed by the one cluster barrier of the iteration;
//   4. every CTA resolves all bids redundantly (GetMax / Assign, emd_cuda.cu:181-215) on its replica with
//      shared-memory atomics, so prices and assignments never have to be exchanged.
// Bid arithmetic is bit-faithful to the reference's SASS:
//   s = fma(dz,dz,fma(dx,dx,dy*dy)), r = sqrt.rn(s), v = (float)((3.0 - (double)r) - (double)price).
// The reference's GetMax race (last writer wins inside a +-1e-6 window) is resolved as
// "largest bidder index wins" (atomicMax), identical to oracle/emd_oracle.c.
#include <cooperative_groups.h>

#include "pcl_common.cuh"

namespace cg = cooperative_groups;

namespace pcl {
namespace {

constexpr int EMD_THREADS = 512;
constexpr int EMD_MAX_N = 4096;
constexpr unsigned short NONE16 = 0xffffu;

struct EmdSmem {
    float4 *tgt;            // N   {x,y,z,0}
    double *pd;             // N   price, held as double (exact image of the fp32 price)
    int *asg;               // N4  assignment (pred j -> target), -1 = unassigned; padded with 0
    float *inc;             // 2N  bid increments, double-buffered across iterations
    float *maxinc;          // N   per-object running max increment (reference: max_increments)
    int *maxidx;            // N   per-object winning bidder (reference: max_idx), -1 = none
    unsigned short *bid;    // 2N  object each bidder bids on, double-buffered
    unsigned short *inv;    // N   assignment_inv (target -> pred), NONE16 = free
    unsigned short *unass;  // N   compacted list of unassigned bidders
    float *pbest, *pbetter; // T   chunk partials
    int *pbi;               // T
    int *wsum;              // 32
};

__host__ __device__ inline size_t emd_smem_bytes(int N) {
    const size_t n4 = (size_t)(N + 3) / 4 * 4;
    return n4 * (16 + 8 + 4 + 8 + 4 + 4) + n4 * 2 * 5 + (size_t)EMD_THREADS * 12 + 32 * 4 + 64;
}

__device__ inline EmdSmem carve(unsigned char *base, int N) {
    const size_t n4 = (size_t)(N + 3) / 4 * 4;
    EmdSmem s;
    unsigned char *p = base;
    s.tgt = (float4 *)p; p += n4 * 16;
    s.pd = (double *)p; p += n4 * 8;
    s.asg = (int *)p; p += n4 * 4;
    s.inc = (float *)p; p += n4 * 8;
    s.maxinc = (float *)p; p += n4 * 4;
    s.maxidx = (int *)p; p += n4 * 4;
    s.pbest = (float *)p; p += EMD_THREADS * 4;
    s.pbetter = (float *)p; p += EMD_THREADS * 4;
    s.pbi = (int *)p; p += EMD_THREADS * 4;
    s.wsum = (int *)p; p += 32 * 4;
    s.bid = (unsigned short *)p; p += n4 * 4;
    s.inv = (unsigned short *)p; p += n4 * 2;
    s.unass = (unsigned short *)p; p += n4 * 2;
    return s;
}

// float max that is correct for mixed signs (reference: CAS loop, emd_cuda.cu:10-20)
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0.f) atomicMax((int *)addr, __float_as_int(v));
    else atomicMin((unsigned *)addr, __float_as_uint(v));
}

// value of target k for a bidder at (ax,ay,az): emd_cuda.cu:142-146, arithmetic pinned to the reference SASS
__device__ __forceinline__ float bid_value(const float4 &t, double p, float ax, float ay, float az) {
    const float dx = __fsub_rn(t.x, ax), dy = __fsub_rn(t.y, ay), dz = __fsub_rn(t.z, az);
    const float s = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const float r = __fsqrt_rn(s);
    return __double2float_rn(__dsub_rn(__dsub_rn(3.0, (double)r), p));
}

__device__ __forceinline__ void top2_update(float v, int k, float &best, float &better, int &bi) {
    if (v > best) { better = best; best = v; bi = k; }  // emd_cuda.cu:147-151
    else if (v > better) better = v;                    // :152-154
}

__global__ void __launch_bounds__(EMD_THREADS, 1)
emd_auction_kernel(Pts xyz1, Pts xyz2, int N, float eps, int iters, float *__restrict__ dist,
                   int *__restrict__ assignment, int *__restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / cs;
    const int tid = threadIdx.x, T = EMD_THREADS;
    const EmdSmem S = carve(smem_raw, N);
    const int n4 = (N + 3) / 4 * 4;

    // ---- init (emd_module.py:45-56) --------------------------------------------------------------
    for (int j = tid; j < n4; j += T) {
        if (j < N) {
            const float3 p = ld_xyz(xyz2, cloud, j);
            S.tgt[j] = make_float4(p.x, p.y, p.z, 0.f);
        }
        S.pd[j] = 0.0;
        S.asg[j] = (j < N) ? -1 : 0;
        S.inv[j] = NONE16;
        S.maxinc[j] = 0.f;
        S.maxidx[j] = -1;
    }
    cluster.sync();  // every CTA's arrays exist before anyone writes remote bids

    long long sum_u = 0;
    int iters_run = 0, extra_qualifiers = 0, cur = 0;
    const int E = (n4 / 4 + T - 1) / T * 4;  // contiguous elements per thread in the compaction (multiple of 4)

    for (int t = 0; t < iters; t++) {
        const bool last = (t == iters - 1);
        // ---- 1. list of unassigned bidders (emd_cuda.cu:23-93), ascending, identical in every CTA -------
        int cnt = 0;
        unsigned flags = 0;
        {
            const int base = tid * E;
            for (int e = 0; e < E; e += 4) {
                if (base + e < n4) {
                    const int4 a = *reinterpret_cast<const int4 *>(S.asg + base + e);
                    flags |= (unsigned)(a.x == -1) << e | (unsigned)(a.y == -1) << (e + 1) |
                             (unsigned)(a.z == -1) << (e + 2) | (unsigned)(a.w == -1) << (e + 3);
                }
            }
            cnt = __popc(flags);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += v;
        }
        if ((tid & 31) == 31) S.wsum[tid >> 5] = incl;
        __syncthreads();
        int wbase = 0, U = 0;
#pragma unroll
        for (int w = 0; w < T / 32; w++) {
            const int v = S.wsum[w];
            if (w < (tid >> 5)) wbase += v;
            U += v;
        }
        if (U == 0) break;  // uniform across the cluster: replicas are identical
        {
            int pos = wbase + incl - cnt;
            const int base = tid * E;
            while (flags) {
                const int e = __ffs(flags) - 1;
                flags &= flags - 1;
                S.unass[pos++] = (unsigned short)(base + e);
            }
        }
        __syncthreads();
        sum_u += U;
        iters_run = t + 1;

        // ---- 2. Bid (emd_cuda.cu:95-179) for this CTA's share of the bidders ----------------------------
        const int per = (U + cs - 1) / cs;
        const int lo = rank * per;
        const int Uc = max(0, min(U, lo + per) - lo);
        int KC = 1;
        if (Uc > 0 && Uc < T) {
            while (KC * 2 * Uc <= T && KC * 2 * 32 <= N) KC *= 2;
        }
        const int L = (N + KC - 1) / KC;
        unsigned short *bid_cur = S.bid + cur * n4;
        float *inc_cur = S.inc + cur * n4;

        auto publish = [&](int j, float best, float better, int bi) {
            const float inc = __fadd_rn(__fsub_rn(best, better), eps);  // emd_cuda.cu:175
            for (int r = 0; r < cs; r++) {
                unsigned short *rb = cluster.map_shared_rank(bid_cur, r);
                float *ri = cluster.map_shared_rank(inc_cur, r);
                rb[j] = (unsigned short)bi;
                ri[j] = inc;
            }
        };

        if (KC == 1) {
            for (int b = tid; b < Uc; b += T) {
                const int j = S.unass[lo + b];
                const float3 a = ld_xyz(xyz1, cloud, j);
                float best = -1e9f, better = -1e9f;
                int bi = -1;
#pragma unroll 4
                for (int k = 0; k < N; k++) top2_update(bid_value(S.tgt[k], S.pd[k], a.x, a.y, a.z), k, best, better, bi);
                publish(j, best, better, bi);
            }
        } else {
            const int items = Uc * KC;  // <= T
            if (tid < items) {
                const int c = tid / Uc, b = tid - c * Uc;
                const int j = S.unass[lo + b];
                const float3 a = ld_xyz(xyz1, cloud, j);
                float best = -1e9f, better = -1e9f;
                int bi = -1;
                const int k1 = min(N, (c + 1) * L);
#pragma unroll 4
                for (int k = c * L; k < k1; k++) top2_update(bid_value(S.tgt[k], S.pd[k], a.x, a.y, a.z), k, best, better, bi);
                S.pbest[tid] = best; S.pbetter[tid] = better; S.pbi[tid] = bi;
            }
            __syncthreads();
            if (tid < Uc) {  // merge chunk partials in ascending k (emd_cuda.cu:165-173)
                float best = S.pbest[tid], better = S.pbetter[tid];
                int bi = S.pbi[tid];
                for (int c = 1; c < KC; c++) {
                    const float pb = S.pbest[c * Uc + tid], pt = S.pbetter[c * Uc + tid];
                    if (pb > best) { better = fmaxf(best, pt); best = pb; bi = S.pbi[c * Uc + tid]; }
                    else better = fmaxf(better, pb);
                }
                publish(S.unass[lo + tid], best, better, bi);
            }
        }
        cluster.sync();  // all bids of this iteration are visible in every CTA

        // ---- 3. GetMax + Assign (emd_cuda.cu:181-215), replicated in every CTA -------------------------
        for (int q = tid; q < U; q += T) {
            const int j = S.unass[q];
            atomic_max_float(&S.maxinc[bid_cur[j]], inc_cur[j]);  // emd_cuda.cu:176
        }
        __syncthreads();
        for (int q = tid; q < U; q += T) {
            const int j = S.unass[q], o = bid_cur[j];
            const double bi = (double)inc_cur[j], mi = (double)S.maxinc[o];
            if (bi - 1e-6 <= mi && mi <= bi + 1e-6) atomicMax(&S.maxidx[o], j);  // :188-191, largest j wins
        }
        __syncthreads();
        // decisions are all taken before any state is modified (U <= 4096 => at most 8 passes per thread)
        unsigned winmask = 0;
        for (int q = tid, p = 0; q < U; q += T, p++) {
            const int j = S.unass[q], o = bid_cur[j];
            const bool winner = (S.maxidx[o] == j);
            if (last || winner) winmask |= 1u << p;  // emd_cuda.cu:201
            if (!winner && rank == 0) {               // statistics only: bidders inside the window that lost the race
                const double bi = (double)inc_cur[j], mi = (double)S.maxinc[o];
                if (bi - 1e-6 <= mi && mi <= bi + 1e-6) extra_qualifiers++;
            }
        }
        __syncthreads();
        for (int q = tid, p = 0; q < U; q += T, p++) {
            if (!((winmask >> p) & 1u)) continue;
            const int j = S.unass[q], o = bid_cur[j];  // emd_cuda.cu:203-211
            const int prev = S.inv[o];
            if (!last && prev != NONE16) S.asg[prev] = -1;
            S.inv[o] = (unsigned short)j;
            S.asg[j] = o;
            S.pd[o] = (double)__fadd_rn((float)S.pd[o], inc_cur[j]);
            S.maxinc[o] = -1e9f;
            S.maxidx[o] = -1;
        }
        __syncthreads();
        cur ^= 1;
    }

    // ---- CalcDist (emd_cuda.cu:217-226) + outputs; the cloud's points are split over the cluster ---------
    __syncthreads();
    for (int j = rank * T + tid; j < N; j += cs * T) {
        const int k = S.asg[j];
        float d = 0.f;
        if (k >= 0) {
            const float3 a = ld_xyz(xyz1, cloud, j);
            const float4 tp = S.tgt[k];
            const float dx = __fsub_rn(a.x, tp.x), dy = __fsub_rn(a.y, tp.y), dz = __fsub_rn(a.z, tp.z);
            d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        }
        dist[(size_t)cloud * N + j] = d;
        assignment[(size_t)cloud * N + j] = k;
    }
    if (stats && rank == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) extra_qualifiers += __shfl_xor_sync(0xffffffffu, extra_qualifiers, o);
        if ((tid & 31) == 0) S.wsum[tid >> 5] = extra_qualifiers;
        __syncthreads();
        if (tid == 0) {
            int e = 0;
            for (int w = 0; w < T / 32; w++) e += S.wsum[w];
            stats[cloud * 4 + 0] = (int)sum_u;
            stats[cloud * 4 + 1] = iters_run;
            stats[cloud * 4 + 2] = e;
            stats[cloud * 4 + 3] = cs;
        }
    }
}

// NmDistanceGradKernel (emd_cuda.cu:284-300) without the atomics (one writer per address) and without
// the zero-fill: grad = (2*graddist) * (xyz1 - xyz2[assignment]).
__global__ void __launch_bounds__(256)
emd_bwd_kernel(Pts xyz1, Pts xyz2, int N, const int *__restrict__ assignment, const float *__restrict__ graddist,
               float *__restrict__ grad) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const size_t o = (size_t)b * N + j;
    const int k = assignment[o];
    const float g = __fmul_rn(graddist[o], 2.f);
    const float3 a = ld_xyz(xyz1, b, j);
    float3 r = make_float3(0.f, 0.f, 0.f);
    if (k >= 0 && k < N) {
        const float3 tpt = ld_xyz(xyz2, b, k);
        r = make_float3(__fmul_rn(g, __fsub_rn(a.x, tpt.x)), __fmul_rn(g, __fsub_rn(a.y, tpt.y)),
                        __fmul_rn(g, __fsub_rn(a.z, tpt.z)));
    }
    grad[o * 3 + 0] = r.x; grad[o * 3 + 1] = r.y; grad[o * 3 + 2] = r.z;
}

// ---- loss epilogue (utils.py:257-304) ----------------------------------------------------------------
__global__ void __launch_bounds__(256)
emd_match_hist_kernel(Pts label, const int *__restrict__ assignment, int B, int N, int C,
                      unsigned long long *__restrict__ hist, int *__restrict__ matched) {
    extern __shared__ unsigned int sh[];
    for (int c = threadIdx.x; c < C; c += blockDim.x) sh[c] = 0;
    __syncthreads();
    const size_t total = (size_t)B * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(e / N);
        const int k = assignment[e];
        int lab = -1;
        if (k >= 0 && k < N) lab = (int)ld_any(label, (int64_t)b * label.bs + (int64_t)k * label.rs);  // .long(): truncation
        if (matched) matched[e] = lab;
        if (lab >= 0 && lab < C) atomicAdd(&sh[lab], 1u);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        if (sh[c]) atomicAdd(&hist[c], (unsigned long long)sh[c]);
}

// sums[0] = sum w*sqrt(d), sums[1] = sum w.  Two-stage, fixed order => deterministic.
constexpr int RED_BLOCKS = 64;
__global__ void __launch_bounds__(256)
emd_wreduce_stage1(const float *__restrict__ dist, const int *__restrict__ matched, const float *__restrict__ cw,
                   size_t total, int C, double *__restrict__ part) {
    __shared__ double s0[8], s1[8];
    double a0 = 0.0, a1 = 0.0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        float w = 1.f;
        if (cw) { const int l = matched[e]; w = (l >= 0 && l < C) ? cw[l] : 0.f; }
        a0 += (double)__fmul_rn(__fsqrt_rn(dist[e]), w);
        a1 += (double)w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a0; s1[threadIdx.x >> 5] = a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double b0 = 0.0, b1 = 0.0;
        for (int w = 0; w < 8; w++) { b0 += s0[w]; b1 += s1[w]; }
        part[blockIdx.x * 2] = b0; part[blockIdx.x * 2 + 1] = b1;
    }
}
__global__ void emd_wreduce_stage2(const double *__restrict__ part, int nb, float *__restrict__ sums) {
    if (threadIdx.x == 0) {
        double b0 = 0.0, b1 = 0.0;
        for (int i = 0; i < nb; i++) { b0 += part[i * 2]; b1 += part[i * 2 + 1]; }
        sums[0] = (float)b0; sums[1] = (float)b1;
    }
}

__global__ void __launch_bounds__(256)
emd_weighted_bwd_kernel(Pts xyz1, Pts xyz2, int N, const int *__restrict__ assignment, const float *__restrict__ dist,
                        const int *__restrict__ matched, const float *__restrict__ cw, int C,
                        const float *__restrict__ sums, const float *__restrict__ grad_out, float *__restrict__ grad) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const size_t o = (size_t)b * N + j;
    const int k = assignment[o];
    float w = 1.f;
    if (cw) { const int l = matched[o]; w = (l >= 0 && l < C) ? cw[l] : 0.f; }
    // d/d dist of  g * sum(w*sqrt(dist)) / sum(w)   (torch: sqrt backward = grad / (2*sqrt))
    const float up = __fdiv_rn(__fmul_rn(__ldg(grad_out), w), __ldg(sums + 1));
    const float gd = __fdiv_rn(up, __fmul_rn(2.f, __fsqrt_rn(dist[o])));
    const float g = __fmul_rn(gd, 2.f);
    const float3 a = ld_xyz(xyz1, b, j);
    float3 r = make_float3(0.f, 0.f, 0.f);
    if (k >= 0 && k < N) {
        const float3 tpt = ld_xyz(xyz2, b, k);
        r = make_float3(__fmul_rn(g, __fsub_rn(a.x, tpt.x)), __fmul_rn(g, __fsub_rn(a.y, tpt.y)),
                        __fmul_rn(g, __fsub_rn(a.z, tpt.z)));
    }
    grad[o * 3 + 0] = r.x; grad[o * 3 + 1] = r.y; grad[o * 3 + 2] = r.z;
}

int pick_cluster(int B, int N, int sm_count, size_t smem, cudaStream_t st) {
    int cs = 16;
    while (cs > 1 && (long)B * cs > sm_count) cs >>= 1;
    while (cs > 1) {  // is a cluster of this size schedulable with this much shared memory?
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(B * cs); cfg.blockDim = dim3(EMD_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, emd_auction_kernel, &cfg);
        if (e == cudaSuccess && ncl > 0) break;
        (void)cudaGetLastError();
        cs >>= 1;
    }
    (void)N;
    return cs;
}

}  // namespace
}  // namespace pcl

using namespace pcl;

extern "C" int pcl_emd_max_points(void) { return EMD_MAX_N; }

extern "C" size_t pcl_emd_workspace_bytes(int B, int N) {
    (void)B; (void)N;
    return align_up((size_t)RED_BLOCKS * 2 * sizeof(double), 256);
}

static int emd_check(const void *xyz1, int dtype1, const void *xyz2, int dtype2, int B, int N, const char *who) {
    if (B < 0 || N < 1) { set_error("%s: bad size B=%d N=%d", who, B, N); return PCL_E_SHAPE; }
    if (N > EMD_MAX_N) { set_error("%s: N=%d > %d points per cloud is not supported by the shared-memory auction", who, N, EMD_MAX_N); return PCL_E_UNSUPPORTED; }
    if (!dtype_ok(dtype1) || !dtype_ok(dtype2)) { set_error("%s: bad dtype", who); return PCL_E_ARG; }
    if (B > 0 && (!xyz1 || !xyz2)) { set_error("%s: null input", who); return PCL_E_ARG; }
    return PCL_OK;
}

extern "C" int pcl_emd_fwd(const void *xyz1, int dtype1, int64_t bs1, int64_t rs1, const void *xyz2, int dtype2,
                           int64_t bs2, int64_t rs2, int B, int N, float eps, int iters, float *dist,
                           int32_t *assignment, int32_t *stats, void *workspace, size_t workspace_bytes, void *stream) {
    (void)workspace; (void)workspace_bytes;
    int rc = emd_check(xyz1, dtype1, xyz2, dtype2, B, N, "emd_fwd");
    if (rc) return rc;
    if (iters < 0) { set_error("emd_fwd: iters=%d", iters); return PCL_E_ARG; }
    if (B > 0 && (!dist || !assignment)) { set_error("emd_fwd: null output"); return PCL_E_ARG; }
    if (B == 0) return PCL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceInfo di;
    if ((rc = device_info(&di))) return rc;
    const size_t smem = emd_smem_bytes(N);
    if (smem > (size_t)di.max_smem_optin) { set_error("emd_fwd: N=%d needs %zu B shared memory (> %d)", N, smem, di.max_smem_optin); return PCL_E_UNSUPPORTED; }
    static thread_local int attr_dev = -1;
    int dev = 0;
    PCL_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        PCL_CUDA(cudaFuncSetAttribute(emd_auction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        PCL_CUDA(cudaFuncSetAttribute(emd_auction_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_dev = dev;
    }
    const int cs = pick_cluster(B, N, di.sm_count, smem, st);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * cs); cfg.blockDim = dim3(EMD_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const Pts p1{xyz1, bs1, rs1, dtype1}, p2{xyz2, bs2, rs2, dtype2};
    PCL_CUDA(cudaLaunchKernelEx(&cfg, emd_auction_kernel, p1, p2, N, eps, iters, dist, (int *)assignment, (int *)stats));
    return PCL_OK;
}

extern "C" int pcl_emd_bwd(const void *xyz1, int dtype1, int64_t bs1, int64_t rs1, const void *xyz2, int dtype2,
                           int64_t bs2, int64_t rs2, int B, int N, const int32_t *assignment, const float *graddist,
                           float *grad_xyz1, void *stream) {
    if (B < 0 || N < 1) { set_error("emd_bwd: bad size B=%d N=%d", B, N); return PCL_E_SHAPE; }
    if (!dtype_ok(dtype1) || !dtype_ok(dtype2)) { set_error("emd_bwd: bad dtype"); return PCL_E_ARG; }
    if (B == 0) return PCL_OK;
    if (!xyz1 || !xyz2 || !assignment || !graddist || !grad_xyz1) { set_error("emd_bwd: null argument"); return PCL_E_ARG; }
    if (B > 65535) { set_error("emd_bwd: B=%d > 65535", B); return PCL_E_SHAPE; }
    const Pts p1{xyz1, bs1, rs1, dtype1}, p2{xyz2, bs2, rs2, dtype2};
    emd_bwd_kernel<<<dim3((N + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(p1, p2, N, assignment, graddist, grad_xyz1);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_emd_match_hist(const void *target_label, int dtype, int64_t bs, int64_t rs, const int32_t *assignment,
                                  int B, int N, int C, int64_t *hist, int32_t *matched_label, void *stream) {
    if (B < 0 || N < 1 || C < 1 || C > 4096) { set_error("emd_match_hist: bad size B=%d N=%d C=%d", B, N, C); return PCL_E_SHAPE; }
    if (!dtype_ok(dtype) || !hist) { set_error("emd_match_hist: bad argument"); return PCL_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    PCL_CUDA(cudaMemsetAsync(hist, 0, (size_t)C * sizeof(int64_t), st));
    if (B == 0) return PCL_OK;
    if (!target_label || !assignment) { set_error("emd_match_hist: null argument"); return PCL_E_ARG; }
    const Pts lab{target_label, bs, rs, dtype};
    const size_t total = (size_t)B * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 296) blocks = 296;
    emd_match_hist_kernel<<<blocks, 256, C * sizeof(unsigned), st>>>(lab, assignment, B, N, C, (unsigned long long *)hist, matched_label);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_emd_weighted_reduce(const float *dist, const int32_t *matched_label, const float *class_weights, int B,
                                       int N, int C, float *sums, void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || N < 1) { set_error("emd_weighted_reduce: bad size B=%d N=%d", B, N); return PCL_E_SHAPE; }
    if (!sums || (B > 0 && !dist) || (class_weights && !matched_label)) { set_error("emd_weighted_reduce: null argument"); return PCL_E_ARG; }
    if (!workspace || workspace_bytes < pcl_emd_workspace_bytes(B, N)) { set_error("emd_weighted_reduce: workspace too small"); return PCL_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    double *part = (double *)workspace;
    emd_wreduce_stage1<<<RED_BLOCKS, 256, 0, st>>>(dist, matched_label, class_weights, (size_t)B * N, C, part);
    PCL_CUDA(cudaGetLastError());
    emd_wreduce_stage2<<<1, 32, 0, st>>>(part, RED_BLOCKS, sums);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_emd_weighted_bwd(const void *xyz1, int dtype1, int64_t bs1, int64_t rs1, const void *xyz2, int dtype2,
                                    int64_t bs2, int64_t rs2, int B, int N, const int32_t *assignment, const float *dist,
                                    const int32_t *matched_label, const float *class_weights, int C, const float *sums,
                                    const float *grad_out, float *grad_xyz1, void *stream) {
    if (B < 0 || N < 1) { set_error("emd_weighted_bwd: bad size B=%d N=%d", B, N); return PCL_E_SHAPE; }
    if (!dtype_ok(dtype1) || !dtype_ok(dtype2)) { set_error("emd_weighted_bwd: bad dtype"); return PCL_E_ARG; }
    if (B == 0) return PCL_OK;
    if (!xyz1 || !xyz2 || !assignment || !dist || !sums || !grad_out || !grad_xyz1 || (class_weights && !matched_label)) {
        set_error("emd_weighted_bwd: null argument"); return PCL_E_ARG;
    }
    if (B > 65535) { set_error("emd_weighted_bwd: B=%d > 65535", B); return PCL_E_SHAPE; }
    const Pts p1{xyz1, bs1, rs1, dtype1}, p2{xyz2, bs2, rs2, dtype2};
    emd_weighted_bwd_kernel<<<dim3((N + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(
        p1, p2, N, assignment, dist, matched_label, class_weights, C, sums, grad_out, grad_xyz1);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}
