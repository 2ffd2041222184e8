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
//
// Bid arithmetic is bit-faithful to the reference's SASS:
//   s = fma(dz,dz,fma(dx,dx,dy*dy)), r = sqrt.rn(s), v = (float)((3.0 - (double)r) - (double)price).
// Only the two largest values of a scan matter, so the expensive part (IEEE sqrt, two F2F conversions, two
// DADDs -- the XU pipe runs at 16 lanes/clk/SM on B200 and bounds the reference's Bid) is evaluated only
// for candidates that pass an exact-safe FP32 filter in the squared domain:
//   skip k  <=>  s_k > (c_k - (T - margin))^2,  c_k = RU(3 - price_k) kept in the target tile's .w,
// where T is a proven lower bound of the bidder's final second-best value (its running second best, seeded
// with the exact current values of the two objects it preferred at its previous bid).  Skipped candidates
// are strictly below the final second best, so best / second-best / first-argmax are exactly the
// reference's (derivation in DESIGN.md "EMD filter").
// The reference's GetMax race (last writer wins inside a +-1e-6 window) is resolved as
// "largest bidder index wins" (atomicMax), identical to oracle/emd_oracle.c.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "pcl_common.cuh"

namespace cg = cooperative_groups;

namespace pcl {
namespace {

constexpr int EMD_THREADS = 512;
constexpr int EMD_MAX_N = 4096;
constexpr unsigned short NONE16 = 0xffffu;
constexpr unsigned NOLAST = 0xffffffffu;
constexpr float FILTER_MARGIN = 2e-6f;  // > 4.2e-7 worst-case rounding slack of the filter

struct EmdSmem {
    float4 *tgt;            // N   {x, y, z, c = RU(3 - price)}
    float *pf;              // N   price (fp32, as in the reference)
    uint2 *pub;             // 2N  published bids {object | second<<16, increment bits}, double-buffered
    float *maxinc;          // N   per-object running max increment (reference: max_increments)
    int *maxidx;            // N   per-object winning bidder (reference: max_idx), -1 = none
    unsigned *last;         // N   previous bid of every bidder (object | second<<16), NOLAST = never bid
    unsigned short *asg;    // N8  assignment (pred j -> target), NONE16 = unassigned; padded with 0
    unsigned short *inv;    // N   assignment_inv (target -> pred), NONE16 = free
    unsigned short *unass;  // N   compacted list of unassigned bidders
    float *pbest, *pbetter; // T   chunk partials
    unsigned *pbi;          // T
    int *wsum;              // 32
    float4 *x1;             // N   predictions {x,y,z,0} when they fit (else nullptr: read from global/L2)
};

__host__ __device__ inline size_t emd_smem_bytes(int N, bool with_x1) {
    const size_t n8 = (size_t)(N + 7) / 8 * 8;
    return n8 * (16 + 4 + 16 + 4 + 4 + 4) + n8 * 2 * 3 + (size_t)EMD_THREADS * 12 + 32 * 4 + 64 + (with_x1 ? n8 * 16 : 0);
}

__device__ inline EmdSmem carve(unsigned char *base, int N, bool with_x1) {
    const size_t n8 = (size_t)(N + 7) / 8 * 8;
    EmdSmem s;
    unsigned char *p = base;
    s.tgt = (float4 *)p; p += n8 * 16;
    s.pub = (uint2 *)p; p += n8 * 16;
    s.asg = (unsigned short *)p; p += n8 * 2;   // 16-byte aligned for the uint4 reads of the compaction
    s.inv = (unsigned short *)p; p += n8 * 2;
    s.unass = (unsigned short *)p; p += n8 * 2;
    s.pf = (float *)p; p += n8 * 4;
    s.maxinc = (float *)p; p += n8 * 4;
    s.maxidx = (int *)p; p += n8 * 4;
    s.last = (unsigned *)p; p += n8 * 4;
    s.pbest = (float *)p; p += EMD_THREADS * 4;
    s.pbetter = (float *)p; p += EMD_THREADS * 4;
    s.pbi = (unsigned *)p; p += EMD_THREADS * 4;
    s.wsum = (int *)p; p += 32 * 4;
    p = (unsigned char *)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    s.x1 = with_x1 ? (float4 *)p : nullptr;
    return s;
}

// float max that is correct for mixed signs (reference: CAS loop, emd_cuda.cu:10-20)
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0.f) atomicMax((int *)addr, __float_as_int(v));
    else atomicMin((unsigned *)addr, __float_as_uint(v));
}

__device__ __forceinline__ float sq3_ref(float dx, float dy, float dz) {  // the reference's contracted x*x+y*y+z*z
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// exact value of a target for a bidder (emd_cuda.cu:146), given the exact squared distance
__device__ __forceinline__ float bid_value_exact(float s, float price) {
    return __double2float_rn(__dsub_rn(__dsub_rn(3.0, (double)__fsqrt_rn(s)), (double)price));
}

struct Top2 {
    float best, better;  // emd_cuda.cu:112
    int bi, bi2;         // first argmax (the bid) and the index of the runner-up (seed for the next bid)
    float tm;            // filter threshold: (lower bound of the final second best) - margin
    int n_exact, n_slowgrp;  // statistics (profiling build only; dead code otherwise)
};

__device__ __forceinline__ void top2_exact(Top2 &r, float s, float price, int k) {
    r.n_exact++;
    const float v = bid_value_exact(s, price);
    if (v > r.best) { r.better = r.best; r.bi2 = r.bi; r.best = v; r.bi = k; }  // emd_cuda.cu:147-151
    else if (v > r.better) { r.better = v; r.bi2 = k; }                           // :152-154
    r.tm = fmaxf(r.tm, __fsub_rn(r.better, FILTER_MARGIN));
}

// Scan targets [k0,k1) for the bidder at (ax,ay,az).  e = u*u - s with u = c_k - tm: e < 0 proves
// value_k < (final second best), so the candidate cannot change best / second best / first argmax.
__device__ __forceinline__ void scan_targets(const EmdSmem &S, int k0, int k1, float ax, float ay, float az, Top2 &r) {
    int k = k0;
#define PCL_FILTER(T_, S_, E_)                                                                   \
    const float4 T_ = S.tgt[k_];                                                                 \
    const float S_ = sq3_ref(__fsub_rn(T_.x, ax), __fsub_rn(T_.y, ay), __fsub_rn(T_.z, az));     \
    const float u_##E_ = __fsub_rn(T_.w, r.tm);                                                   \
    const float E_ = __fmaf_rn(u_##E_, u_##E_, -S_);
    for (; k + 4 <= k1; k += 4) {
        float e0, e1, e2, e3, s0, s1, s2, s3;
        { const int k_ = k;     PCL_FILTER(t, s, e) e0 = e; s0 = s; }
        { const int k_ = k + 1; PCL_FILTER(t, s, e) e1 = e; s1 = s; }
        { const int k_ = k + 2; PCL_FILTER(t, s, e) e2 = e; s2 = s; }
        { const int k_ = k + 3; PCL_FILTER(t, s, e) e3 = e; s3 = s; }
        if (!(fmaxf(fmaxf(e0, e1), fmaxf(e2, e3)) < 0.f)) {  // rare: some candidate may be in the top 2
            r.n_slowgrp++;
            if (!(e0 < 0.f)) top2_exact(r, s0, S.pf[k], k);
            if (!(e1 < 0.f)) top2_exact(r, s1, S.pf[k + 1], k + 1);
            if (!(e2 < 0.f)) top2_exact(r, s2, S.pf[k + 2], k + 2);
            if (!(e3 < 0.f)) top2_exact(r, s3, S.pf[k + 3], k + 3);
        }
    }
    for (; k < k1; k++) {
        const int k_ = k;
        PCL_FILTER(t, s, e)
        if (!(e < 0.f)) top2_exact(r, s, S.pf[k], k);
    }
#undef PCL_FILTER
}

// Start of a scan: seed the threshold with the exact current values of the two objects this bidder
// preferred last time (their prices may have risen since; any two distinct objects give a valid bound).
__device__ __forceinline__ Top2 top2_init(const EmdSmem &S, unsigned lastpack, int N, float ax, float ay, float az) {
    Top2 r;
    r.best = -1e9f; r.better = -1e9f; r.bi = -1; r.bi2 = -1; r.tm = -1e9f; r.n_exact = 0; r.n_slowgrp = 0;
    const int k1 = (int)(lastpack & 0xffffu), k2 = (int)(lastpack >> 16);
    if (lastpack != NOLAST && k1 < N && k2 < N && k1 != k2) {
        const float4 t1 = S.tgt[k1], t2 = S.tgt[k2];
        const float v1 = bid_value_exact(sq3_ref(__fsub_rn(t1.x, ax), __fsub_rn(t1.y, ay), __fsub_rn(t1.z, az)), S.pf[k1]);
        const float v2 = bid_value_exact(sq3_ref(__fsub_rn(t2.x, ax), __fsub_rn(t2.y, ay), __fsub_rn(t2.z, az)), S.pf[k2]);
        r.tm = __fsub_rn(fminf(v1, v2), FILTER_MARGIN);
    }
    return r;
}

// PROF: per-phase clock64() totals of thread 0 of every CTA -> prof[blockIdx.x*8 + phase] (development aid,
// reached through the PCL_EMD_PROFILE environment variable; the product path instantiates PROF=false).
template <bool PROF>
__global__ void __launch_bounds__(EMD_THREADS, 1)
emd_auction_kernel(Pts xyz1, Pts xyz2, int N, float eps, int iters, int x1_smem, float *__restrict__ dist,
                   int *__restrict__ assignment, int *__restrict__ stats, long long *__restrict__ prof) {
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc = 0;
#define PCL_TICK(i)                                              \
    if constexpr (PROF) {                                        \
        const long long now_ = clock64();                        \
        pt[i] += now_ - pc;                                      \
        pc = now_;                                               \
    }
    if constexpr (PROF) pc = clock64();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / cs;
    const int tid = threadIdx.x, T = EMD_THREADS;
    const EmdSmem S = carve(smem_raw, N, x1_smem != 0);
    const int n8 = (N + 7) / 8 * 8;

    // ---- init (emd_module.py:45-56) --------------------------------------------------------------
    for (int j = tid; j < n8; j += T) {
        if (j < N) {
            const float3 p = ld_xyz(xyz2, cloud, j);
            S.tgt[j] = make_float4(p.x, p.y, p.z, 3.0f);
            if (x1_smem) {
                const float3 q = ld_xyz(xyz1, cloud, j);
                S.x1[j] = make_float4(q.x, q.y, q.z, 0.f);
            }
        }
        S.pf[j] = 0.f;
        S.asg[j] = (j < N) ? NONE16 : (unsigned short)0;
        S.inv[j] = NONE16;
        S.maxinc[j] = 0.f;
        S.maxidx[j] = -1;
        S.last[j] = NOLAST;
    }
    cluster.sync();  // every CTA's arrays exist before anyone writes remote bids
    PCL_TICK(0)

    auto pred_xyz = [&](int j) -> float3 {
        if (x1_smem) { const float4 q = S.x1[j]; return make_float3(q.x, q.y, q.z); }
        return ld_xyz(xyz1, cloud, j);
    };
    long long sum_u = 0;
    int iters_run = 0, extra_qualifiers = 0, cur = 0;
    const int E = (n8 / 8 + T - 1) / T * 8;  // contiguous elements per thread in the compaction (multiple of 8, <= 32)

    for (int t = 0; t < iters; t++) {
        const bool last = (t == iters - 1);
        // ---- 1. list of unassigned bidders (emd_cuda.cu:23-93), ascending, identical in every CTA -------
        unsigned flags = 0;
        {
            const int base = tid * E;
            for (int e = 0; e < E; e += 8) {
                if (base + e < n8) {
                    const uint4 a = *reinterpret_cast<const uint4 *>(S.asg + base + e);
                    const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        flags |= (unsigned)((w[h] & 0xffffu) == NONE16) << (e + 2 * h);
                        flags |= (unsigned)((w[h] >> 16) == NONE16) << (e + 2 * h + 1);
                    }
                }
            }
        }
        const int cnt = __popc(flags);
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
        PCL_TICK(1)

        // ---- 2. Bid (emd_cuda.cu:95-179) for this CTA's share of the bidders ----------------------------
        const int per = (U + cs - 1) / cs;
        const int lo = rank * per;
        const int Uc = max(0, min(U, lo + per) - lo);
        int KC = 1;
        if (Uc > 0 && Uc < T) KC = max(1, min(T / Uc, N / 32));
        uint2 *pub_cur = S.pub + cur * n8;

        auto publish = [&](int j, const Top2 &r) {
            const float inc = __fadd_rn(__fsub_rn(r.best, r.better), eps);  // emd_cuda.cu:175
            const uint2 v = make_uint2((unsigned)(r.bi & 0xffff) | ((unsigned)(r.bi2 & 0xffff) << 16), __float_as_uint(inc));
            for (int c = 0; c < cs; c++) cluster.map_shared_rank(pub_cur, c)[j] = v;
        };

        if (KC == 1) {
            for (int b0 = 0; b0 < Uc; b0 += T) {
                const int b = b0 + tid;
                if (b < Uc) {
                    const int j = S.unass[lo + b];
                    const float3 a = pred_xyz(j);
                    Top2 r = top2_init(S, S.last[j], N, a.x, a.y, a.z);
                    scan_targets(S, 0, N, a.x, a.y, a.z, r);
                    publish(j, r);
                }
            }
        } else {
            const int items = Uc * KC;  // <= T
            long long ts[5] = {0, 0, 0, 0, 0};
            if (tid < items) {
                const int c = tid / Uc, b = tid - c * Uc;
                const int j = S.unass[lo + b];
                const float3 a = pred_xyz(j);
                if constexpr (PROF) { ts[0] = clock64() + (long long)(a.x * 0.f); }
                Top2 r = top2_init(S, S.last[j], N, a.x, a.y, a.z);
                if constexpr (PROF) { ts[1] = clock64() + (long long)(r.tm * 0.f); }
                const int k0 = (int)(((long long)c * N) / KC), k1 = (int)(((long long)(c + 1) * N) / KC);
                scan_targets(S, k0, k1, a.x, a.y, a.z, r);
                if constexpr (PROF) { ts[2] = clock64() + (long long)(r.best * 0.f); }
                S.pbest[tid] = r.best; S.pbetter[tid] = r.better;
                S.pbi[tid] = (unsigned)(r.bi & 0xffff) | ((unsigned)(r.bi2 & 0xffff) << 16);
            }
            __syncthreads();
            if constexpr (PROF) {
                ts[3] = clock64();
                if (blockIdx.x == 0 && tid == 0 && prof && t < 50)
                    for (int i = 0; i < 4; i++) prof[(size_t)gridDim.x * 8 + 128 + t * 4 + i] = ts[i] - pc;
            }
            // Tree merge of the KC chunk partials of every bidder (log2 KC steps).  Equivalent to the reference's
            // sequential merge in ascending k (emd_cuda.cu:165-173): best = max, its index = the LOWEST k among
            // equal maxima (explicit index compare makes the merge order-independent), better = second largest
            // counting duplicates.
            {
                const int c = tid / Uc, b = tid - c * Uc;
                int span = 1;
                while (span < KC) span <<= 1;
                for (int st = span >> 1; st >= 1; st >>= 1) {
                    if (tid < items && c < st && c + st < KC) {
                        const int me = c * Uc + b, ot = (c + st) * Uc + b;
                        float best = S.pbest[me], better = S.pbetter[me];
                        unsigned pk = S.pbi[me];
                        const float ob = S.pbest[ot], obt = S.pbetter[ot];
                        const unsigned opk = S.pbi[ot];
                        const bool other_wins = (ob > best) || (ob == best && (opk & 0xffffu) < (pk & 0xffffu));
                        if (other_wins) {
                            // new second best = max(old best, other's second best)
                            const unsigned second = (best >= obt) ? (pk & 0xffffu) : (opk >> 16);
                            better = fmaxf(best, obt);
                            best = ob;
                            pk = (opk & 0xffffu) | (second << 16);
                        } else if (ob > better) {
                            better = ob;
                            pk = (pk & 0xffffu) | ((opk & 0xffffu) << 16);
                        }
                        S.pbest[me] = best; S.pbetter[me] = better; S.pbi[me] = pk;
                    }
                    __syncthreads();
                }
                if (tid < Uc) {
                    Top2 r;
                    r.best = S.pbest[tid]; r.better = S.pbetter[tid];
                    r.bi = (int)(S.pbi[tid] & 0xffffu); r.bi2 = (int)(S.pbi[tid] >> 16); r.tm = 0.f;
                    publish(S.unass[lo + tid], r);
                }
            }
        }
        if constexpr (PROF) {
            if (blockIdx.x == 0 && tid == 0 && prof && t < 64) {
                prof[(size_t)gridDim.x * 8 + t * 2] = U;
                prof[(size_t)gridDim.x * 8 + t * 2 + 1] = clock64() - pc;
            }
        }
        PCL_TICK(2)
        cluster.sync();  // all bids of this iteration are visible in every CTA
        PCL_TICK(3)

        // ---- 3. GetMax + Assign (emd_cuda.cu:181-215), replicated in every CTA -------------------------
        for (int q = tid; q < U; q += T) {
            const int j = S.unass[q];
            const uint2 pb = pub_cur[j];
            S.last[j] = pb.x;
            atomic_max_float(&S.maxinc[pb.x & 0xffffu], __uint_as_float(pb.y));  // emd_cuda.cu:176
        }
        __syncthreads();
        for (int q = tid; q < U; q += T) {
            const int j = S.unass[q];
            const uint2 pb = pub_cur[j];
            const int o = (int)(pb.x & 0xffffu);
            const double bi = (double)__uint_as_float(pb.y), mi = (double)S.maxinc[o];
            if (bi - 1e-6 <= mi && mi <= bi + 1e-6) atomicMax(&S.maxidx[o], j);  // :188-191, largest j wins
        }
        __syncthreads();
        // decisions are all taken before any state is modified (U <= 4096 => at most 8 passes per thread)
        unsigned winmask = 0;
        for (int q = tid, p = 0; q < U; q += T, p++) {
            const int j = S.unass[q];
            const uint2 pb = pub_cur[j];
            const int o = (int)(pb.x & 0xffffu);
            const bool winner = (S.maxidx[o] == j);
            if (last || winner) winmask |= 1u << p;  // emd_cuda.cu:201
            if (!winner && rank == 0) {               // statistics only: bidders inside the window that lost the race
                const double bi = (double)__uint_as_float(pb.y), mi = (double)S.maxinc[o];
                if (bi - 1e-6 <= mi && mi <= bi + 1e-6) extra_qualifiers++;
            }
        }
        __syncthreads();
        for (int q = tid, p = 0; q < U; q += T, p++) {
            if (!((winmask >> p) & 1u)) continue;
            const int j = S.unass[q];
            const uint2 pb = pub_cur[j];
            const int o = (int)(pb.x & 0xffffu);  // emd_cuda.cu:203-211
            const unsigned prev = S.inv[o];
            if (!last && prev != NONE16) S.asg[prev] = NONE16;
            S.inv[o] = (unsigned short)j;
            S.asg[j] = (unsigned short)o;
            const float pnew = __fadd_rn(S.pf[o], __uint_as_float(pb.y));
            S.pf[o] = pnew;
            S.tgt[o].w = __fsub_ru(3.0f, pnew);  // c = RU(3 - price): upper bound used by the filter
            S.maxinc[o] = -1e9f;
            S.maxidx[o] = -1;
        }
        __syncthreads();
        cur ^= 1;
        PCL_TICK(4)
    }

    // ---- CalcDist (emd_cuda.cu:217-226) + outputs; the cloud's points are split over the cluster ---------
    __syncthreads();
    for (int j = rank * T + tid; j < N; j += cs * T) {
        const unsigned k = S.asg[j];
        float d = 0.f;
        if (k != NONE16) {
            const float3 a = pred_xyz(j);
            const float4 tp = S.tgt[k];
            d = sq3_ref(__fsub_rn(a.x, tp.x), __fsub_rn(a.y, tp.y), __fsub_rn(a.z, tp.z));
        }
        dist[(size_t)cloud * N + j] = d;
        assignment[(size_t)cloud * N + j] = (k != NONE16) ? (int)k : -1;
    }
    PCL_TICK(5)
    if constexpr (PROF) {
        if (tid == 0 && prof)
            for (int i = 0; i < 8; i++) prof[(size_t)blockIdx.x * 8 + i] = pt[i];
    }
#undef PCL_TICK
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
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, emd_auction_kernel<false>, &cfg);
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
    (void)N;
    const size_t red = (size_t)RED_BLOCKS * 2 * sizeof(double), prof = ((size_t)(B > 0 ? B : 0) * 16 * 8 + 512) * sizeof(long long);
    return align_up(red > prof ? red : prof, 256);
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
    int rc = emd_check(xyz1, dtype1, xyz2, dtype2, B, N, "emd_fwd");
    if (rc) return rc;
    if (iters < 0) { set_error("emd_fwd: iters=%d", iters); return PCL_E_ARG; }
    if (B > 0 && (!dist || !assignment)) { set_error("emd_fwd: null output"); return PCL_E_ARG; }
    if (B == 0) return PCL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceInfo di;
    if ((rc = device_info(&di))) return rc;
    const bool with_x1 = emd_smem_bytes(N, true) <= (size_t)di.max_smem_optin;
    const size_t smem = emd_smem_bytes(N, with_x1);
    if (smem > (size_t)di.max_smem_optin) { set_error("emd_fwd: N=%d needs %zu B shared memory (> %d)", N, smem, di.max_smem_optin); return PCL_E_UNSUPPORTED; }
    static thread_local int attr_dev = -1;
    int dev = 0;
    PCL_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        PCL_CUDA(cudaFuncSetAttribute(emd_auction_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        PCL_CUDA(cudaFuncSetAttribute(emd_auction_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        PCL_CUDA(cudaFuncSetAttribute(emd_auction_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        PCL_CUDA(cudaFuncSetAttribute(emd_auction_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
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
    // development aid: PCL_EMD_PROFILE=1 makes the workspace receive per-phase clock totals (B*cs*8 int64)
    static const bool profile = getenv("PCL_EMD_PROFILE") != nullptr;
    if (profile && workspace && workspace_bytes >= ((size_t)B * cs * 8 + 512) * sizeof(long long)) {
        PCL_CUDA(cudaLaunchKernelEx(&cfg, emd_auction_kernel<true>, p1, p2, N, eps, iters, (int)with_x1, dist, (int *)assignment, (int *)stats, (long long *)workspace));
    } else {
        PCL_CUDA(cudaLaunchKernelEx(&cfg, emd_auction_kernel<false>, p1, p2, N, eps, iters, (int)with_x1, dist, (int *)assignment, (int *)stats, (long long *)nullptr));
    }
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
