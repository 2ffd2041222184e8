// pcl_chamfer.cu -- Chamfer distance forward/backward for sm_100a.
//
// Replaces the native layer under pytorch3d.loss.chamfer_distance as the reference calls it
// (pointcloud_vision/utils.py:211,228): knn_points(K=1) in both directions, squared L2, lowest
// index on exact ties, variable lengths, point-mean + batch-mean; backward scatters
// 2*g*(p - q[idx]) to both clouds.
//
// Forward kernel (D == 3): one CTA = 128 threads x QPT register-resident queries of one cloud and one direction; the
// target cloud streams through shared memory in tiles stored per PAIR of targets ({x0,x1,y0,y1}, {z0,z1,|t0|^2,|t1|^2}:
// the operand layout of FFMA2).  The scan evaluates an APPROXIMATE expanded-form distance with packed FFMA2 and only
// remembers which 32-target chunks can hold the nearest neighbour; those chunks are re-scanned with the oracle's
// arithmetic (explicit __fsub_rn/__fmul_rn/__fadd_rn/__fmaf_rn), which alone decides distance, index and the
// lowest-index tie rule (derivation above chamfer_nn3_kernel and in DESIGN.md 3.2).
#include <stdlib.h>

#include "pcl_common.cuh"

namespace pcl {
namespace {

constexpr int CH_THREADS = 128;  // the generic-D kernel and the partial-sum layout: one slot per 128 queries
#ifndef PCL_C3_QPT
#define PCL_C3_QPT 2
#endif
#ifndef PCL_C3_WQ
#define PCL_C3_WQ 4
#endif
#ifndef PCL_C3_PARTS
#define PCL_C3_PARTS 1
#endif
constexpr int C3_QPT = PCL_C3_QPT;      // D == 3 kernel: register-resident queries per lane ...
constexpr int C3_WQ = PCL_C3_WQ;        // ... warps with different queries ...
constexpr int C3_PARTS = PCL_C3_PARTS;  // ... and warps that share the queries but scan different chunks of the targets (merged at the end)
constexpr int C3_QUERIES = 32 * C3_QPT * C3_WQ;  // queries per CTA
constexpr int C3_THREADS = 32 * C3_WQ * C3_PARTS;
constexpr int C3_TILE = 1024;    // targets per shared-memory tile (16 KB), cut into ...
constexpr int C3_CHUNK = 32;     // ... chunks of 32: the granularity of the approximate minimum and of the exact re-scan
#ifndef PCL_C3_GROUP
#define PCL_C3_GROUP 4
#endif
#ifndef PCL_C3_MINB
#define PCL_C3_MINB 4
#endif
constexpr int C3_GROUP = PCL_C3_GROUP;      // pairs of targets per register buffer of the scan (double-buffered: hides the LDS latency)

typedef unsigned long long u64;
// sm_100a packed FP32: FFMA2 = two IEEE-rounded FMAs per instruction, same FLOP rate as FFMA at half the issue slots
// (tools/microbench_packed.cu).  Only the approximate scan uses it: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 (even with -fmad=false), so the oracle's unfused sums cannot be written with packed instructions.
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }  // NaN operands are ignored

template <bool FMA>
__device__ __forceinline__ float sqdist3(float qx, float qy, float qz, const float4 &t) {
    const float dx = __fsub_rn(qx, t.x), dy = __fsub_rn(qy, t.y), dz = __fsub_rn(qz, t.z);
    if constexpr (FMA) return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    else return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// length of cloud n, clamped to [0, P]: pytorch3d raises on lengths > P; a kernel cannot, but it must not read past the cloud
__device__ __forceinline__ int len_of(const int64_t *len, int n, int P) {
    if (!len) return P;
    const int64_t v = len[n];
    return (int)(v < 0 ? 0 : (v > P ? P : v));
}

// Tail of both forward kernels: the CTA's partial sum goes to the workspace; the CTA that takes the LAST ticket of the grid
// (zeroed by the host call) then forms  loss_xy[dir] = sum_n (sum_i dist / clamp(len,1)) / max(B,1)  from all partials in a
// fixed order (deterministic scalar, fp64) -- no memset of the partials and no second launch.  Needs >= 64 threads per CTA.
__device__ __forceinline__ void finish_block(float s /* valid on thread 0 */, float *__restrict__ partial, int nblk, int dir, int n,
                                             unsigned *__restrict__ ticket, int P1, int P2, const int64_t *__restrict__ x_len,
                                             const int64_t *__restrict__ y_len, float *__restrict__ loss_xy) {
    __shared__ int last_flag;
    const int B = (int)gridDim.y;
    if (threadIdx.x == 0) {
        partial[((size_t)dir * B + n) * nblk + blockIdx.x] = s;
        __threadfence();
        last_flag = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y * gridDim.z - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!last_flag || threadIdx.x >= 64) return;
    __threadfence();
    __shared__ double acc[2][32];
    const int d = threadIdx.x >> 5, lane = threadIdx.x & 31;  // one warp per direction
    double a = 0.0;
    for (int c = lane; c < B; c += 32) {
        double t = 0.0;
        for (int k = 0; k < (int)gridDim.x; k++) t += (double)__ldcg(&partial[((size_t)d * B + c) * nblk + k]);
        const int len = d ? len_of(y_len, c, P2) : len_of(x_len, c, P1);
        a += t / (double)(len > 1 ? len : 1);
    }
    acc[d][lane] = a;
    __syncwarp();
    if (lane == 0) {
        double t = 0.0;
        for (int k = 0; k < 32; k++) t += acc[d][k];
        loss_xy[d] = (float)(t / (double)(B > 1 ? B : 1));
        loss_xy[2 + d] = (float)t;  // the un-normalised batch sum: what a batch-sharded caller all-reduces
    }
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

__device__ __forceinline__ float finite_or_inf(float v) { return (v <= 3.0e38f) ? v : __int_as_float(0x7f800000); }  // NaN -> +inf

// Exact re-scan of chunk c of the tile for one query.  Phase A evaluates the approximate distance once more (same
// bound, same threshold thr = running minimum + E2) and collects the pairs of targets that pass it -- the exact nearest
// neighbour always does; phase B gives those (one or two) the oracle's arithmetic and folds them into (best, bi) with the
// explicit rule "smaller distance, then lower index", the order-independent form of knn's ascending strict-'<' scan.
// Lanes re-scan DIFFERENT chunks; rotating the pair order by the lane id keeps the 8 lanes of an LDS.128 phase on 8
// different bank groups.
template <bool FMA>
__device__ __forceinline__ void rescan_chunk(const float4 *tA, const float4 *tB, bool act, int c, int cnt, int t0, float qx,
                                             float qy, float qz, u64 nqx, u64 nqy, u64 nqz, float thr, float &best, int &bi) {
    unsigned pm = 0;
    const int base = c * (C3_CHUNK / 2), rot = (int)threadIdx.x;
#pragma unroll
    for (int pp = 0; pp < C3_CHUNK / 2; pp++) {
        const int o = (pp + rot) & (C3_CHUNK / 2 - 1);
        const float4 a = tA[base + o], b = tB[base + o];
        const u64 v = ffma2(pk2(b.x, b.y), nqz, ffma2(pk2(a.z, a.w), nqy, ffma2(pk2(a.x, a.y), nqx, pk2(b.z, b.w))));
        pm |= (unsigned)(!(fminf(lo2(v), hi2(v)) > thr)) << o;
    }
    if (!act) pm = 0;
    while (__any_sync(0xffffffffu, pm != 0)) {
        if (pm) {
            const int o = __ffs(pm) - 1;
            pm &= pm - 1;
            const float4 a = tA[base + o], b = tB[base + o];
            const int j0 = 2 * (base + o), i0 = t0 + j0;
            const float d0 = sqdist3<FMA>(qx, qy, qz, make_float4(a.x, a.z, b.x, 0.f));
            const float d1 = sqdist3<FMA>(qx, qy, qz, make_float4(a.y, a.w, b.y, 0.f));
            if (j0 < cnt && (d0 < best || (d0 == best && i0 < bi))) { best = d0; bi = i0; }
            if (j0 + 1 < cnt && (d1 < best || (d1 == best && i0 + 1 < bi))) { best = d1; bi = i0 + 1; }
        }
    }
}

// Forward kernel for D == 3.  grid: (ceil(maxP / C3_QUERIES), B, 2 directions); block: WQ x PARTS warps, QPT queries per lane.
//
// The scan over the targets runs on an APPROXIMATE distance in expanded form,
//     a(q,t) = |t|^2 - 2 q.t  =  fma(tz, -2qz, fma(ty, -2qy, fma(tx, -2qx, |t|^2)))      (= |q-t|^2 - |q|^2),
// three FFMA2 per two targets instead of eight FP32 operations per target, and keeps only the minimum of a per
// (query, 32-target chunk).  Results never depend on it: with u = 2^-24, |a + |q|^2 - d| <= u (21 |t|^2 + 15 |q|^2) for
// the oracle's fp32 distance d (3 roundings in |t|^2, 3 in the FMAs on partial sums <= 2|t|^2 + |q|^2, <= 6 relative
// roundings in d <= 2|q|^2 + 2|t|^2), so the chunk that holds the exact nearest neighbour k* satisfies, when it is
// scanned and at any later time,
//     chunk minimum <= (running minimum of a) + E2,   E2 = 2^-18 (max|t|^2 seen so far + |q|^2)   (> twice that bound).
// Such a chunk is remembered per query when it is scanned (the most recent one in registers; when it is displaced and
// still qualifies it moves into a bit mask), and before the tile leaves shared memory the remembered chunks (one,
// rarely two) are re-scanned with the oracle's arithmetic (rescan_chunk), which alone decides
// distance and index.  Non-finite or huge coordinates make E2 infinite, i.e. everything is re-scanned: slower, still exact.
template <bool FMA>
__global__ void __launch_bounds__(C3_THREADS, PCL_C3_MINB)
chamfer_nn3_kernel(Pts x, const int64_t *__restrict__ x_len, Pts y, const int64_t *__restrict__ y_len, int P1, int P2,
                   float *__restrict__ dist_x, int *__restrict__ idx_x, float *__restrict__ dist_y,
                   int *__restrict__ idx_y, float *__restrict__ partial, int nblk, unsigned *__restrict__ ticket,
                   float *__restrict__ loss_xy) {
    __shared__ float4 tA[C3_TILE / 2 + C3_GROUP];  // per pair of targets {x0, x1, y0, y1} (+ slack for the read-ahead)
    __shared__ float4 tB[C3_TILE / 2 + C3_GROUP];  //                     {z0, z1, |t0|^2, |t1|^2}
    constexpr int C3_PPT = C3_TILE / C3_THREADS;   // staged points per thread and tile
    constexpr int CP = C3_CHUNK / 2;               // pairs per chunk
    __shared__ float red[C3_THREADS / 32];
    __shared__ int tmax_bits;
    const int dir = blockIdx.z, n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wq = (tid >> 5) % C3_WQ, part = (tid >> 5) / C3_WQ;
    const Pts q = dir ? y : x, t = dir ? x : y;
    const int PQ = dir ? P2 : P1;
    const int q0 = blockIdx.x * C3_QUERIES;
    if (q0 >= PQ) {  // block-uniform: beyond the shorter cloud -- contributes a zero partial (and may be the last CTA)
        finish_block(0.f, partial, nblk, dir, n, ticket, P1, P2, x_len, y_len, loss_xy);
        return;
    }
    const int lq = dir ? len_of(y_len, n, P2) : len_of(x_len, n, P1);
    const int lt = dir ? len_of(x_len, n, P1) : len_of(y_len, n, P2);
    float *dist = (dir ? dist_y : dist_x) + (size_t)n * PQ;
    int *idx = (dir ? idx_y : idx_x) + (size_t)n * PQ;
    const float INF = __int_as_float(0x7f800000);

    // query r of a lane: q0 + 32 (r WQ + wq) + lane; the PARTS warps with the same wq hold the same queries
    float qx[C3_QPT], qy[C3_QPT], qz[C3_QPT], q2[C3_QPT], best[C3_QPT], mrun[C3_QPT];
    u64 nqx[C3_QPT], nqy[C3_QPT], nqz[C3_QPT];
    int bi[C3_QPT];
#pragma unroll
    for (int r = 0; r < C3_QPT; r++) {
        const int i = q0 + (r * C3_WQ + wq) * 32 + lane;
        float3 p = make_float3(0.f, 0.f, 0.f);
        if (i < lq) p = ld_xyz(q, n, i);
        qx[r] = p.x; qy[r] = p.y; qz[r] = p.z;
        q2[r] = finite_or_inf(__fmaf_rn(p.z, p.z, __fmaf_rn(p.y, p.y, __fmul_rn(p.x, p.x))));
        nqx[r] = pk2(-2.f * p.x, -2.f * p.x); nqy[r] = pk2(-2.f * p.y, -2.f * p.y); nqz[r] = pk2(-2.f * p.z, -2.f * p.z);
        best[r] = INF; bi[r] = 0x7fffffff; mrun[r] = INF;
    }
    if (q0 < lq && lt > 0) {  // block-uniform: blocks made only of padded rows skip the scan
        if (tid == 0) tmax_bits = 0;
        float3 pre[C3_PPT];  // points of the next tile: loaded before the exact re-scans of the current one, stored after them
        auto prefetch = [&](int t0) {
#pragma unroll
            for (int k = 0; k < C3_PPT; k++) {
                const int j = t0 + k * C3_THREADS + tid;
                pre[k] = (j < lt) ? ld_xyz(t, n, j) : make_float3(0.f, 0.f, 0.f);
            }
        };
        prefetch(0);
        __syncthreads();
        for (int t0 = 0; t0 < lt; t0 += C3_TILE) {
            const int cnt = min(C3_TILE, lt - t0);
            const int nch = (cnt + C3_CHUNK - 1) / C3_CHUNK;
            {   // registers -> shared memory (pair layout), |t|^2, running maximum of |t|^2; padding: a = +inf, never a minimum
                float wmax = 0.f;
                float *fA = reinterpret_cast<float *>(tA), *fB = reinterpret_cast<float *>(tB);
#pragma unroll
                for (int k = 0; k < C3_PPT; k++) {
                    const int j = k * C3_THREADS + tid;
                    const float3 p = pre[k];
                    float w = __fmaf_rn(p.z, p.z, __fmaf_rn(p.y, p.y, __fmul_rn(p.x, p.x)));
                    if (j < cnt) wmax = fmaxf(wmax, finite_or_inf(w)); else w = INF;
                    const int o = (j >> 1) * 4 + (j & 1);
                    fA[o] = p.x; fA[o + 2] = p.y; fB[o] = p.z; fB[o + 2] = w;
                }
                const int wm = __reduce_max_sync(0xffffffffu, __float_as_int(wmax));  // wmax >= 0: integer order == float order
                if (lane == 0) atomicMax(&tmax_bits, wm);
            }
            __syncthreads();
            const float tmax = __int_as_float(tmax_bits);
            float e2[C3_QPT], thr[C3_QPT], lm[C3_QPT];
            int lcid[C3_QPT];       // chunk of the most recent candidate (lm = its approximate minimum)
            unsigned dmask[C3_QPT]; // displaced candidates that still qualified when they were displaced
#pragma unroll
            for (int r = 0; r < C3_QPT; r++) {
                e2[r] = fmaxf(__fmul_ru(__fadd_ru(tmax, q2[r]), 3.8147e-6f /* > 2^-18 */), 1e-36f);
                thr[r] = __fadd_ru(mrun[r], e2[r]);
                lm[r] = INF; lcid[r] = 0; dmask[r] = 0;
            }
            float4 ga[C3_GROUP], gb[C3_GROUP];  // current group of pairs; the next one is loaded while this one is used
#pragma unroll
            for (int i = 0; i < C3_GROUP; i++) { ga[i] = tA[part * CP + i]; gb[i] = tB[part * CP + i]; }
            for (int c = part; c < nch; c += C3_PARTS) {
                float mc[C3_QPT];
#pragma unroll
                for (int r = 0; r < C3_QPT; r++) mc[r] = INF;
#pragma unroll
                for (int g = 0; g < CP / C3_GROUP; g++) {
                    float4 na[C3_GROUP], nb[C3_GROUP];
                    // next group of this chunk, or the first group of this warp's next chunk (clamped into the slack at the end)
                    const int nxt = min((g + 1 < CP / C3_GROUP) ? c * CP + (g + 1) * C3_GROUP : (c + C3_PARTS) * CP, C3_TILE / 2);
#pragma unroll
                    for (int i = 0; i < C3_GROUP; i++) { na[i] = tA[nxt + i]; nb[i] = tB[nxt + i]; }
#pragma unroll
                    for (int i = 0; i < C3_GROUP; i++) {
                        const u64 tx = pk2(ga[i].x, ga[i].y), ty = pk2(ga[i].z, ga[i].w), tz = pk2(gb[i].x, gb[i].y), tw = pk2(gb[i].z, gb[i].w);
#pragma unroll
                        for (int r = 0; r < C3_QPT; r++) {
                            const u64 v = ffma2(tz, nqz[r], ffma2(ty, nqy[r], ffma2(tx, nqx[r], tw)));
                            mc[r] = fmin3(mc[r], lo2(v), hi2(v));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < C3_GROUP; i++) { ga[i] = na[i]; gb[i] = nb[i]; }
                }
#pragma unroll
                for (int r = 0; r < C3_QPT; r++) {  // branch-free: a chunk that may hold the exact nearest neighbour becomes the candidate
                    const bool cand = !(mc[r] > thr[r]);
                    const float nm = fminf(mrun[r], mc[r]);     // == mrun unless cand
                    const float nthr = __fadd_ru(nm, e2[r]);
                    const bool keep = cand && !(lm[r] > nthr);  // the displaced candidate still qualifies (lm = +inf: none yet)
                    dmask[r] |= keep ? (1u << lcid[r]) : 0u;
                    lm[r] = cand ? mc[r] : lm[r];
                    lcid[r] = cand ? c : lcid[r];
                    mrun[r] = nm; thr[r] = nthr;
                }
            }
            if (t0 + C3_TILE < lt) prefetch(t0 + C3_TILE);
            // exact re-scan of the remembered chunks while the tile is resident
#pragma unroll
            for (int r = 0; r < C3_QPT; r++) {
                unsigned m = dmask[r];
                if (!(lm[r] > thr[r])) m |= 1u << lcid[r];
                if (q0 + (r * C3_WQ + wq) * 32 + lane >= lq) m = 0;
                while (__any_sync(0xffffffffu, m != 0)) {
                    const bool act = m != 0;
                    const int c = act ? (__ffs(m) - 1) : 0;
                    m &= m - 1;
                    rescan_chunk<FMA>(tA, tB, act, c, cnt, t0, qx[r], qy[r], qz[r], nqx[r], nqy[r], nqz[r], thr[r], best[r], bi[r]);
                }
            }
            __syncthreads();
        }
    }
    // merge the parts (smaller distance, then lower index) -- the tile buffers are free now
    if constexpr (C3_PARTS > 1) {
        static_assert((C3_PARTS - 1) * C3_QUERIES * 4 <= (int)sizeof(float4) * (C3_TILE / 2), "merge buffer fits into a tile array");
        float *mb = reinterpret_cast<float *>(tA);
        int *mi = reinterpret_cast<int *>(tB);
        if (part > 0) {
#pragma unroll
            for (int r = 0; r < C3_QPT; r++) {
                mb[(part - 1) * C3_QUERIES + (r * C3_WQ + wq) * 32 + lane] = best[r];
                mi[(part - 1) * C3_QUERIES + (r * C3_WQ + wq) * 32 + lane] = bi[r];
            }
        }
        __syncthreads();
        if (part == 0) {
#pragma unroll
            for (int r = 0; r < C3_QPT; r++) {
#pragma unroll
                for (int pp = 1; pp < C3_PARTS; pp++) {
                    const float ob = mb[(pp - 1) * C3_QUERIES + (r * C3_WQ + wq) * 32 + lane];
                    const int oi = mi[(pp - 1) * C3_QUERIES + (r * C3_WQ + wq) * 32 + lane];
                    if (ob < best[r] || (ob == best[r] && oi < bi[r])) { best[r] = ob; bi[r] = oi; }
                }
            }
        }
    }
    float s = 0.f;
    if (part == 0) {
#pragma unroll
        for (int r = 0; r < C3_QPT; r++) {
            const int i = q0 + (r * C3_WQ + wq) * 32 + lane;
            if (i < PQ) {
                const bool valid = (i < lq) && (lt > 0);
                const float d = valid ? best[r] : 0.f;
                dist[i] = d; idx[i] = (valid && bi[r] != 0x7fffffff && best[r] < INF) ? bi[r] : 0;  // no finite distance: knn's initial index 0
                s += d;
            }
        }
    }
    s = block_sum<C3_THREADS / 32>(s, red);
    finish_block(s, partial, nblk, dir, n, ticket, P1, P2, x_len, y_len, loss_xy);
}

// Generic feature width (ChamferDistance over all channels, utils.py:209-211).  One query per thread.
template <bool FMA, int D>
__global__ void __launch_bounds__(CH_THREADS)
chamfer_nnD_kernel(Pts x, const int64_t *__restrict__ x_len, Pts y, const int64_t *__restrict__ y_len, int P1, int P2,
                   float *__restrict__ dist_x, int *__restrict__ idx_x, float *__restrict__ dist_y,
                   int *__restrict__ idx_y, float *__restrict__ partial, int nblk, unsigned *__restrict__ ticket,
                   float *__restrict__ loss_xy) {
    constexpr int TILE = 512;
    __shared__ float tile[TILE * D];
    __shared__ float red[4];
    const int dir = blockIdx.z, n = blockIdx.y;
    const Pts q = dir ? y : x, t = dir ? x : y;
    const int PQ = dir ? P2 : P1;
    const int q0 = blockIdx.x * CH_THREADS;
    if (q0 >= PQ) {
        finish_block(0.f, partial, nblk, dir, n, ticket, P1, P2, x_len, y_len, loss_xy);
        return;
    }
    const int lq = dir ? len_of(y_len, n, P2) : len_of(x_len, n, P1);
    const int lt = dir ? len_of(x_len, n, P1) : len_of(y_len, n, P2);
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
    finish_block(s, partial, nblk, dir, n, ticket, P1, P2, x_len, y_len, loss_xy);
}

// Backward: grid (ceil(maxP/256), B, 2).  All contributions go through fp32 red.global.add onto
// zero-filled outputs (same accumulation model as pytorch3d's CUDA knn backward).
__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(Pts x, const int64_t *__restrict__ x_len, Pts y, const int64_t *__restrict__ y_len, int B, int P1,
                   int P2, int D, const int *__restrict__ idx_x, const int *__restrict__ idx_y,
                   const float *__restrict__ grad_out, float g_imm_x, float g_imm_y, float *__restrict__ grad_x,
                   float *__restrict__ grad_y) {
    const int dir = blockIdx.z, n = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const Pts q = dir ? y : x, t = dir ? x : y;
    const int PQ = dir ? P2 : P1, PT = dir ? P1 : P2;
    const int lq = dir ? len_of(y_len, n, P2) : len_of(x_len, n, P1);
    const int lt = dir ? len_of(x_len, n, P1) : len_of(y_len, n, P2);
    if (i >= lq || lt <= 0) return;
    const float g = grad_out ? __ldg(grad_out + dir) : (dir ? g_imm_y : g_imm_x);  // upstream gradient: device scalar or immediate
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
               float *dist_x, int *idx_x, float *dist_y, int *idx_y, float *partial, int nblk, unsigned *ticket, float *loss_xy,
               cudaStream_t st) {
    const int maxP = P1 > P2 ? P1 : P2;
    if (D == 3) {
        dim3 grid((maxP + C3_QUERIES - 1) / C3_QUERIES, B, 2);
        chamfer_nn3_kernel<FMA><<<grid, C3_THREADS, 0, st>>>(x, x_len, y, y_len, P1, P2, dist_x, idx_x, dist_y, idx_y, partial, nblk, ticket, loss_xy);
    } else {
        dim3 grid((maxP + CH_THREADS - 1) / CH_THREADS, B, 2);
#define PCL_LAUNCH_NND(DD) \
    case DD: chamfer_nnD_kernel<FMA, DD><<<grid, CH_THREADS, 0, st>>>(x, x_len, y, y_len, P1, P2, dist_x, idx_x, dist_y, idx_y, partial, nblk, ticket, loss_xy); break
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
    constexpr int G = C3_QUERIES < CH_THREADS ? C3_QUERIES : CH_THREADS;  // queries per CTA of the finer-grained kernel
    return (maxP + G - 1) / G + 1;
}

}  // namespace
}  // namespace pcl

using namespace pcl;

extern "C" size_t pcl_chamfer_workspace_bytes(int B, int P1, int P2) {
    if (B <= 0) return 256;
    return 256 + align_up((size_t)2 * B * nblk_for(P1, P2) * sizeof(float), 256);  // [ticket][per-CTA partial sums]
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
        PCL_CUDA(cudaMemsetAsync(loss_xy, 0, 4 * sizeof(float), st));
        return PCL_OK;
    }
    const Pts xp{x, x_bs, x_rs, x_dtype}, yp{y, y_bs, y_rs, y_dtype};
    unsigned *ticket = (unsigned *)workspace;
    float *partial = (float *)((unsigned char *)workspace + 256);
    const int nblk = nblk_for(P1, P2);
    PCL_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));  // every CTA writes its partial (zeros beyond the shorter cloud); the last one sums
    rc = (mode == PCL_CHAMFER_FMA)
             ? launch_fwd<true>(xp, x_len, yp, y_len, B, P1, P2, D, dist_x, idx_x, dist_y, idx_y, partial, nblk, ticket, loss_xy, st)
             : launch_fwd<false>(xp, x_len, yp, y_len, B, P1, P2, D, dist_x, idx_x, dist_y, idx_y, partial, nblk, ticket, loss_xy, st);
    if (rc) return rc;
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

// shared by pcl_chamfer_bwd (upstream gradients on the device) and the composite step (immediate upstream gradients)
int pcl::chamfer_bwd_impl(const void *x, int x_dtype, int64_t x_bs, int64_t x_rs, const int64_t *x_len, const void *y, int y_dtype,
                          int64_t y_bs, int64_t y_rs, const int64_t *y_len, int B, int P1, int P2, int D, const int32_t *idx_x,
                          const int32_t *idx_y, const float *grad_out, float g_imm_x, float g_imm_y, float *grad_x, float *grad_y,
                          cudaStream_t st) {
    int rc = check_args(x, x_dtype, y, y_dtype, B, P1, P2, D);
    if (rc) return rc;
    if (B > 0 && ((P1 > 0 && (!grad_x || !idx_x)) || (P2 > 0 && (!grad_y || !idx_y)))) {
        set_error("chamfer_bwd: null argument"); return PCL_E_ARG;
    }
    if (B == 0) return PCL_OK;
    if (P1 > 0) PCL_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)B * P1 * D * sizeof(float), st));
    if (P2 > 0) PCL_CUDA(cudaMemsetAsync(grad_y, 0, (size_t)B * P2 * D * sizeof(float), st));
    if (P1 == 0 || P2 == 0) return PCL_OK;
    const Pts xp{x, x_bs, x_rs, x_dtype}, yp{y, y_bs, y_rs, y_dtype};
    const int maxP = P1 > P2 ? P1 : P2;
    dim3 grid((maxP + 255) / 256, B, 2);
    chamfer_bwd_kernel<<<grid, 256, 0, st>>>(xp, x_len, yp, y_len, B, P1, P2, D, idx_x, idx_y, grad_out, g_imm_x, g_imm_y, grad_x, grad_y);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_chamfer_bwd(const void *x, int x_dtype, int64_t x_bs, int64_t x_rs, const int64_t *x_len,
                               const void *y, int y_dtype, int64_t y_bs, int64_t y_rs, const int64_t *y_len, int B,
                               int P1, int P2, int D, const int32_t *idx_x, const int32_t *idx_y,
                               const float *grad_out, float *grad_x, float *grad_y, void *stream) {
    if (!grad_out) { set_error("chamfer_bwd: null argument"); return PCL_E_ARG; }
    return chamfer_bwd_impl(x, x_dtype, x_bs, x_rs, x_len, y, y_dtype, y_bs, y_rs, y_len, B, P1, P2, D, idx_x, idx_y, grad_out, 0.f, 0.f,
                            grad_x, grad_y, (cudaStream_t)stream);
}
