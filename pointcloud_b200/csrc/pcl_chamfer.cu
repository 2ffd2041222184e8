// pcl_chamfer.cu -- Chamfer distance forward/backward for sm_100a.
//
// Replaces the native layer under pytorch3d.loss.chamfer_distance as the reference calls it
// (pointcloud_vision/utils.py:211,228): knn_points(K=1) in both directions, squared L2, lowest
// index on exact ties, variable lengths, point-mean + batch-mean; backward scatters
// 2*g*(p - q[idx]) to both clouds.
//
// Forward kernel (D == 3): one CTA = 128 threads x QPT register-resident queries of one cloud and one
// direction; the target cloud streams through shared memory in float4 {x,y,z,0} tiles (one
// broadcast LDS.128 feeds QPT distance evaluations per thread); distance arithmetic is written with
// explicit __fmul_rn/__fadd_rn/__fmaf_rn so that the compiler cannot re-associate or contract it
// differently from the oracle (bit-exact distances => bit-exact argmin).
#include <stdlib.h>

#include "pcl_common.cuh"

namespace pcl {
namespace {

constexpr int CH_THREADS = 128;
constexpr int CH_TILE = 1024;  // targets per shared-memory tile (16 KB)
constexpr int CH_CHUNK = 32;   // argmin recovery granularity

template <bool FMA>
__device__ __forceinline__ float sqdist3(float qx, float qy, float qz, const float4 &t) {
    const float dx = __fsub_rn(qx, t.x), dy = __fsub_rn(qy, t.y), dz = __fsub_rn(qz, t.z);
    if constexpr (FMA) return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    else return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

template <int NWARPS>
__device__ __forceinline__ float block_sum(float v, float *red /* >= NWARPS floats */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NWARPS; i++) s += red[i];  // fixed order
    }
    return s;  // valid on thread 0
}

// grid: (ceil(maxP / (128*QPT)), B, 2 directions); block: 128 query lanes x KSP target parts.
// A CTA owns 128*QPT queries; its KSP groups of 4 warps scan interleaved 64-target chunks of every tile, so that
// even the small (B=32, N=2048) problem keeps ~28 warps per SM busy while one broadcast LDS.128 still feeds QPT
// evaluations.  Per evaluation only the running minimum is tracked (one FMNMX next to the 8 FMA-pipe operations);
// the argmin is recovered per chunk, and parts are merged with an explicit lowest-index rule.
template <bool FMA, int QPT, int KSP>
__global__ void __launch_bounds__(CH_THREADS * KSP)
chamfer_nn3_kernel(Pts x, const int64_t *__restrict__ x_len, Pts y, const int64_t *__restrict__ y_len, int P1, int P2,
                   float *__restrict__ dist_x, int *__restrict__ idx_x, float *__restrict__ dist_y,
                   int *__restrict__ idx_y, float *__restrict__ partial, int nblk) {
    constexpr int NT = CH_THREADS * KSP;
    __shared__ float4 tile[CH_TILE];
    __shared__ float red[NT / 32];
    static_assert(sizeof(float4) * CH_TILE >= (size_t)CH_THREADS * QPT * KSP * 8, "merge buffer aliases the tile");
    const int dir = blockIdx.z, n = blockIdx.y;
    const Pts q = dir ? y : x, t = dir ? x : y;
    const int PQ = dir ? P2 : P1;
    const int q0 = blockIdx.x * (CH_THREADS * QPT);
    if (q0 >= PQ) return;  // block-uniform
    const int lq = dir ? (y_len ? (int)y_len[n] : P2) : (x_len ? (int)x_len[n] : P1);
    const int lt = dir ? (x_len ? (int)x_len[n] : P1) : (y_len ? (int)y_len[n] : P2);
    float *dist = (dir ? dist_y : dist_x) + (size_t)n * PQ;
    int *idx = (dir ? idx_y : idx_x) + (size_t)n * PQ;
    const int ql = threadIdx.x & (CH_THREADS - 1), part = threadIdx.x / CH_THREADS;

    float qx[QPT], qy[QPT], qz[QPT], best[QPT];
    int bi[QPT], bchunk[QPT];
    bool improved[QPT];
#pragma unroll
    for (int r = 0; r < QPT; r++) {
        const int i = q0 + r * CH_THREADS + ql;
        float3 p = make_float3(0.f, 0.f, 0.f);
        if (i < lq) p = ld_xyz(q, n, i);
        qx[r] = p.x; qy[r] = p.y; qz[r] = p.z;
        best[r] = __int_as_float(0x7f800000);  // +inf
        bi[r] = 0x7fffffff; bchunk[r] = 0; improved[r] = false;
    }
    if (q0 < lq) {  // block-uniform: blocks made only of padded rows skip the scan
        for (int t0 = 0; t0 < lt; t0 += CH_TILE) {
            const int cnt = min(CH_TILE, lt - t0);
            for (int j = threadIdx.x; j < cnt; j += NT) {
                const float3 p = ld_xyz(t, n, t0 + j);
                tile[j] = make_float4(p.x, p.y, p.z, 0.f);
            }
            __syncthreads();
            // strict '<' between this part's chunks (ascending) keeps its earliest chunk; the re-scan below returns the
            // lowest index inside it -- together with the index-aware merge: "lowest index wins ties", as in knn's scan.
            for (int c0 = part * CH_CHUNK; c0 < cnt; c0 += CH_CHUNK * KSP) {
                const int c1 = min(c0 + CH_CHUNK, cnt);
                float mc[QPT];
#pragma unroll
                for (int r = 0; r < QPT; r++) mc[r] = __int_as_float(0x7f800000);
#pragma unroll 8
                for (int j = c0; j < c1; j++) {
                    const float4 tp = tile[j];
#pragma unroll
                    for (int r = 0; r < QPT; r++) mc[r] = fminf(mc[r], sqdist3<FMA>(qx[r], qy[r], qz[r], tp));
                }
#pragma unroll
                for (int r = 0; r < QPT; r++)
                    if (mc[r] < best[r]) { best[r] = mc[r]; bchunk[r] = c0; improved[r] = true; }
            }
#pragma unroll
            for (int r = 0; r < QPT; r++) {
                if (improved[r]) {  // the tile is still in shared memory: find the lowest index that attains the minimum
                    // Every lane re-scans a DIFFERENT chunk; chunks are 1 KB apart, i.e. the same banks.  Rotating the
                    // start by the lane id makes the 32 LDS.128 of a step hit 32 consecutive float4 (conflict-free).
                    const int c0 = bchunk[r], c1 = min(c0 + CH_CHUNK, cnt);
                    int found = 0x7fffffff;
#pragma unroll 4
                    for (int jj = 0; jj < CH_CHUNK; jj++) {
                        const int j = c0 + ((jj + (int)threadIdx.x) & (CH_CHUNK - 1));
                        if (j < c1 && sqdist3<FMA>(qx[r], qy[r], qz[r], tile[j]) == best[r]) found = min(found, j);
                    }
                    bi[r] = t0 + found;
                    improved[r] = false;
                }
            }
            __syncthreads();
        }
    }
    if constexpr (KSP > 1) {  // merge the parts: minimum distance, lowest index among equal minima
        float *mb = reinterpret_cast<float *>(tile);
        int *mi = reinterpret_cast<int *>(tile) + CH_THREADS * QPT * KSP;
#pragma unroll
        for (int r = 0; r < QPT; r++) {
            mb[(part * QPT + r) * CH_THREADS + ql] = best[r];
            mi[(part * QPT + r) * CH_THREADS + ql] = bi[r];
        }
        __syncthreads();
        if (part == 0) {
#pragma unroll
            for (int r = 0; r < QPT; r++) {
#pragma unroll
                for (int pp = 1; pp < KSP; pp++) {
                    const float ob = mb[(pp * QPT + r) * CH_THREADS + ql];
                    const int oi = mi[(pp * QPT + r) * CH_THREADS + ql];
                    if (ob < best[r] || (ob == best[r] && oi < bi[r])) { best[r] = ob; bi[r] = oi; }
                }
            }
        }
    }
    float s = 0.f;
    if (part == 0) {
#pragma unroll
        for (int r = 0; r < QPT; r++) {
            const int i = q0 + r * CH_THREADS + ql;
            if (i < PQ) {
                const bool valid = (i < lq) && (lt > 0);
                const float d = valid ? best[r] : 0.f;
                dist[i] = d; idx[i] = (valid && bi[r] != 0x7fffffff) ? bi[r] : 0;
                s += d;
            }
        }
    }
    s = block_sum<NT / 32>(s, red);
    if (threadIdx.x == 0) partial[((size_t)dir * gridDim.y + n) * nblk + blockIdx.x] = s;
}

// Generic feature width (ChamferDistance over all channels, utils.py:209-211).  One query per thread.
template <bool FMA, int D>
__global__ void __launch_bounds__(CH_THREADS)
chamfer_nnD_kernel(Pts x, const int64_t *__restrict__ x_len, Pts y, const int64_t *__restrict__ y_len, int P1, int P2,
                   float *__restrict__ dist_x, int *__restrict__ idx_x, float *__restrict__ dist_y,
                   int *__restrict__ idx_y, float *__restrict__ partial, int nblk) {
    constexpr int TILE = 512;
    __shared__ float tile[TILE * D];
    __shared__ float red[4];
    const int dir = blockIdx.z, n = blockIdx.y;
    const Pts q = dir ? y : x, t = dir ? x : y;
    const int PQ = dir ? P2 : P1;
    const int q0 = blockIdx.x * CH_THREADS;
    if (q0 >= PQ) return;
    const int lq = dir ? (y_len ? (int)y_len[n] : P2) : (x_len ? (int)x_len[n] : P1);
    const int lt = dir ? (x_len ? (int)x_len[n] : P1) : (y_len ? (int)y_len[n] : P2);
    float *dist = (dir ? dist_y : dist_x) + (size_t)n * PQ;
    int *idx = (dir ? idx_y : idx_x) + (size_t)n * PQ;
    const int i = q0 + threadIdx.x;
    float qv[D];
#pragma unroll
    for (int c = 0; c < D; c++) qv[c] = (i < lq) ? ld_any(q, (int64_t)n * q.bs + (int64_t)i * q.rs + c) : 0.f;
    float best = __int_as_float(0x7f800000);
    int bi = 0;
    if (q0 < lq) {
        for (int t0 = 0; t0 < lt; t0 += TILE) {
            const int cnt = min(TILE, lt - t0);
            for (int e = threadIdx.x; e < cnt * D; e += CH_THREADS) {
                const int j = e / D, c = e - j * D;
                tile[e] = ld_any(t, (int64_t)n * t.bs + (int64_t)(t0 + j) * t.rs + c);
            }
            __syncthreads();
#pragma unroll 2
            for (int j = 0; j < cnt; j++) {
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < D; c++) {
                    const float df = __fsub_rn(qv[c], tile[j * D + c]);
                    if constexpr (FMA) d = __fmaf_rn(df, df, d);
                    else d = __fadd_rn(d, __fmul_rn(df, df));
                }
                if (d < best) { best = d; bi = t0 + j; }
            }
            __syncthreads();
        }
    }
    float s = 0.f;
    if (i < PQ) {
        const bool valid = (i < lq) && (lt > 0);
        s = valid ? best : 0.f;
        dist[i] = s; idx[i] = valid ? bi : 0;
    }
    s = block_sum<4>(s, red);
    if (threadIdx.x == 0) partial[((size_t)dir * gridDim.y + n) * nblk + blockIdx.x] = s;
}

// loss_xy[dir] = sum_n (sum_i dist / clamp(len,1)) / max(B,1); fixed summation order => deterministic.
__global__ void chamfer_finish_kernel(const float *__restrict__ partial, int B, int nblk, int nbx, int nby, int P1, int P2,
                                      const int64_t *__restrict__ x_len, const int64_t *__restrict__ y_len,
                                      float *__restrict__ loss_xy) {
    __shared__ double acc[2][32];
    const int dir = threadIdx.x >> 5, lane = threadIdx.x & 31;  // 64 threads: one warp per direction
    double a = 0.0;
    for (int n = lane; n < B; n += 32) {
        const int used = dir ? nby : nbx;
        double s = 0.0;
        for (int k = 0; k < used; k++) s += (double)partial[((size_t)dir * B + n) * nblk + k];
        const int64_t len = dir ? (y_len ? y_len[n] : P2) : (x_len ? x_len[n] : P1);
        a += s / (double)(len > 1 ? len : 1);
    }
    acc[dir][lane] = a;
    __syncwarp();
    if (lane == 0) {
        double s = 0.0;
        for (int k = 0; k < 32; k++) s += acc[dir][k];
        loss_xy[dir] = (float)(s / (double)(B > 1 ? B : 1));
    }
}

// Backward: grid (ceil(maxP/256), B, 2).  All contributions go through fp32 red.global.add onto
// zero-filled outputs (same accumulation model as pytorch3d's CUDA knn backward).
__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(Pts x, const int64_t *__restrict__ x_len, Pts y, const int64_t *__restrict__ y_len, int B, int P1,
                   int P2, int D, const int *__restrict__ idx_x, const int *__restrict__ idx_y,
                   const float *__restrict__ grad_out, float *__restrict__ grad_x, float *__restrict__ grad_y) {
    const int dir = blockIdx.z, n = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const Pts q = dir ? y : x, t = dir ? x : y;
    const int PQ = dir ? P2 : P1, PT = dir ? P1 : P2;
    const int lq = dir ? (y_len ? (int)y_len[n] : P2) : (x_len ? (int)x_len[n] : P1);
    const int lt = dir ? (x_len ? (int)x_len[n] : P1) : (y_len ? (int)y_len[n] : P2);
    if (i >= lq || lt <= 0) return;
    const float g = __ldg(grad_out + dir);
    const float gd = g / (float)(B > 1 ? B : 1) / (float)(lq > 1 ? lq : 1);
    const int j = (dir ? idx_y : idx_x)[(size_t)n * PQ + i];
    float *gq = (dir ? grad_y : grad_x) + ((size_t)n * PQ + i) * D;
    float *gt = (dir ? grad_x : grad_y) + ((size_t)n * PT + j) * D;
    const int64_t qo = (int64_t)n * q.bs + (int64_t)i * q.rs, to = (int64_t)n * t.bs + (int64_t)j * t.rs;
    for (int c = 0; c < D; c++) {
        const float diff = __fmul_rn(__fmul_rn(2.0f, gd), __fsub_rn(ld_any(q, qo + c), ld_any(t, to + c)));
        atomicAdd(gq + c, diff);
        atomicAdd(gt + c, -diff);
    }
}

template <bool FMA>
int launch_fwd(const Pts &x, const int64_t *x_len, const Pts &y, const int64_t *y_len, int B, int P1, int P2, int D,
               float *dist_x, int *idx_x, float *dist_y, int *idx_y, float *partial, int nblk, int sm_count,
               cudaStream_t st) {
    const int maxP = P1 > P2 ? P1 : P2;
    if (D == 3) {
        // 128*QPT queries per CTA; KSP target parts per CTA keep the warp count per SM high on small problems
        const long queries = (long)B * ((long)P1 + P2);
        int qpt = 4;
        while (qpt > 1 && queries / (CH_THREADS * qpt) < 6L * sm_count) qpt >>= 1;
        int ksp = 1;  // target-part split: measured slower on B200 at every size tried (profiles/r1_chamfer_variants.txt); kept as an option
        if (const char *e = getenv("PCL_CHAMFER_QPT")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) qpt = v; }  // development aid
        if (const char *e = getenv("PCL_CHAMFER_KSP")) { const int v = atoi(e); if (v == 1 || v == 4) ksp = v; }
        dim3 grid((maxP + CH_THREADS * qpt - 1) / (CH_THREADS * qpt), B, 2);
#define PCL_LAUNCH_NN3(Q, K) \
    chamfer_nn3_kernel<FMA, Q, K><<<grid, CH_THREADS * K, 0, st>>>(x, x_len, y, y_len, P1, P2, dist_x, idx_x, dist_y, idx_y, partial, nblk)
        if (ksp == 4) {
            if (qpt == 4) PCL_LAUNCH_NN3(4, 4);
            else if (qpt == 2) PCL_LAUNCH_NN3(2, 4);
            else PCL_LAUNCH_NN3(1, 4);
        } else {
            if (qpt == 4) PCL_LAUNCH_NN3(4, 1);
            else if (qpt == 2) PCL_LAUNCH_NN3(2, 1);
            else PCL_LAUNCH_NN3(1, 1);
        }
#undef PCL_LAUNCH_NN3
    } else {
        dim3 grid((maxP + CH_THREADS - 1) / CH_THREADS, B, 2);
#define PCL_LAUNCH_NND(DD) \
    case DD: chamfer_nnD_kernel<FMA, DD><<<grid, CH_THREADS, 0, st>>>(x, x_len, y, y_len, P1, P2, dist_x, idx_x, dist_y, idx_y, partial, nblk); break
        switch (D) {
            PCL_LAUNCH_NND(1); PCL_LAUNCH_NND(2); PCL_LAUNCH_NND(4); PCL_LAUNCH_NND(5);
            PCL_LAUNCH_NND(6); PCL_LAUNCH_NND(7); PCL_LAUNCH_NND(8);
            default: set_error("chamfer: D=%d unsupported (1..8)", D); return PCL_E_UNSUPPORTED;
        }
#undef PCL_LAUNCH_NND
    }
    return PCL_OK;
}

int check_args(const void *x, int x_dtype, const void *y, int y_dtype, int B, int P1, int P2, int D) {
    if (B < 0 || P1 < 0 || P2 < 0) { set_error("chamfer: negative size B=%d P1=%d P2=%d", B, P1, P2); return PCL_E_SHAPE; }
    if (D < 1 || D > 8) { set_error("chamfer: D=%d unsupported (1..8)", D); return PCL_E_UNSUPPORTED; }
    if (B > 65535) { set_error("chamfer: B=%d > 65535", B); return PCL_E_SHAPE; }
    if (!dtype_ok(x_dtype) || !dtype_ok(y_dtype)) { set_error("chamfer: bad dtype"); return PCL_E_ARG; }
    if (B > 0 && ((P1 > 0 && !x) || (P2 > 0 && !y))) { set_error("chamfer: null input"); return PCL_E_ARG; }
    return PCL_OK;
}

inline int nblk_for(int P1, int P2) {
    const int maxP = P1 > P2 ? P1 : P2;
    return (maxP + CH_THREADS - 1) / CH_THREADS + 1;  // upper bound for every QPT
}

}  // namespace
}  // namespace pcl

using namespace pcl;

extern "C" size_t pcl_chamfer_workspace_bytes(int B, int P1, int P2) {
    if (B <= 0) return 256;
    return align_up((size_t)2 * B * nblk_for(P1, P2) * sizeof(float), 256);
}

extern "C" int pcl_chamfer_fwd(const void *x, int x_dtype, int64_t x_bs, int64_t x_rs, const int64_t *x_len,
                               const void *y, int y_dtype, int64_t y_bs, int64_t y_rs, const int64_t *y_len, int B,
                               int P1, int P2, int D, int mode, float *dist_x, int32_t *idx_x, float *dist_y,
                               int32_t *idx_y, float *loss_xy, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_args(x, x_dtype, y, y_dtype, B, P1, P2, D);
    if (rc) return rc;
    if (mode != PCL_CHAMFER_UNFUSED && mode != PCL_CHAMFER_FMA) { set_error("chamfer: bad mode %d", mode); return PCL_E_ARG; }
    if (!loss_xy || (B > 0 && ((P1 > 0 && (!dist_x || !idx_x)) || (P2 > 0 && (!dist_y || !idx_y))))) {
        set_error("chamfer: null output"); return PCL_E_ARG;
    }
    if (workspace_bytes < pcl_chamfer_workspace_bytes(B, P1, P2) || !workspace) {
        set_error("chamfer: workspace too small (%zu < %zu)", workspace_bytes, pcl_chamfer_workspace_bytes(B, P1, P2));
        return PCL_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DeviceInfo di;
    if ((rc = device_info(&di))) return rc;
    if (B == 0 || (P1 == 0 && P2 == 0)) {
        PCL_CUDA(cudaMemsetAsync(loss_xy, 0, 2 * sizeof(float), st));
        return PCL_OK;
    }
    const Pts xp{x, x_bs, x_rs, x_dtype}, yp{y, y_bs, y_rs, y_dtype};
    float *partial = (float *)workspace;
    const int nblk = nblk_for(P1, P2);
    // blocks that return early (beyond the shorter cloud) must still contribute zeros
    PCL_CUDA(cudaMemsetAsync(partial, 0, (size_t)2 * B * nblk * sizeof(float), st));
    rc = (mode == PCL_CHAMFER_FMA)
             ? launch_fwd<true>(xp, x_len, yp, y_len, B, P1, P2, D, dist_x, idx_x, dist_y, idx_y, partial, nblk, di.sm_count, st)
             : launch_fwd<false>(xp, x_len, yp, y_len, B, P1, P2, D, dist_x, idx_x, dist_y, idx_y, partial, nblk, di.sm_count, st);
    if (rc) return rc;
    PCL_CUDA(cudaGetLastError());
    chamfer_finish_kernel<<<1, 64, 0, st>>>(partial, B, nblk, nblk, nblk, P1, P2, x_len, y_len, loss_xy);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_chamfer_bwd(const void *x, int x_dtype, int64_t x_bs, int64_t x_rs, const int64_t *x_len,
                               const void *y, int y_dtype, int64_t y_bs, int64_t y_rs, const int64_t *y_len, int B,
                               int P1, int P2, int D, const int32_t *idx_x, const int32_t *idx_y,
                               const float *grad_out, float *grad_x, float *grad_y, void *stream) {
    int rc = check_args(x, x_dtype, y, y_dtype, B, P1, P2, D);
    if (rc) return rc;
    if (!grad_out || (B > 0 && ((P1 > 0 && (!grad_x || !idx_x)) || (P2 > 0 && (!grad_y || !idx_y))))) {
        set_error("chamfer_bwd: null argument"); return PCL_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) return PCL_OK;
    if (P1 > 0) PCL_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)B * P1 * D * sizeof(float), st));
    if (P2 > 0) PCL_CUDA(cudaMemsetAsync(grad_y, 0, (size_t)B * P2 * D * sizeof(float), st));
    if (P1 == 0 || P2 == 0) return PCL_OK;
    const Pts xp{x, x_bs, x_rs, x_dtype}, yp{y, y_bs, y_rs, y_dtype};
    const int maxP = P1 > P2 ? P1 : P2;
    dim3 grid((maxP + 255) / 256, B, 2);
    chamfer_bwd_kernel<<<grid, 256, 0, st>>>(xp, x_len, yp, y_len, B, P1, P2, D, idx_x, idx_y, grad_out, grad_x, grad_y);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}
