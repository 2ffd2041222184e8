// pcl_emd_team.cu -- auction EMD as ONE persistent launch of owner and worker CTAs (sm_100a).
//
// Same algorithm, arithmetic and tie rules as pcl_emd.cu (reference: pointcloud_vision/loss/emd/emd_cuda.cu:23-282), different
// distribution of the work.  The cluster kernel gives every cloud a fixed cluster of CTAs for the whole auction: the launch ends
// with its slowest cloud (Table-shaped clouds: 3.6 M cycles against a mean of 2.7 M), 20 of the 148 SMs stay empty at B=32 and a
// small batch cannot use more than 16 SMs per cloud.  Here the grid is one CTA per SM:
//   * CTA c < B is the OWNER of cloud c.  It keeps the complete auction state in shared memory exactly like a cluster of one:
//     set-up sort, list of unassigned bidders, seed thresholds, GetMax / Assign (emd_cuda.cu:181-215), CalcDist and the fused epilogue.
//   * The Bid phase (emd_cuda.cu:95-179) of an iteration is cut into TASKS -- a block of bidders x all target tiles -- that ANY CTA
//     can execute: the owner itself, or one of the gridDim.x - B WORKER CTAs, which pull tasks of whichever cloud is furthest behind.
//     What a task needs is small and lives in an L2-resident mirror the owner keeps up to date: targets with c = RU(3 - price),
//     prices, tile boxes (46 KB at N=2048, loaded into the worker's shared memory per (cloud, iteration)) and one 16-byte record per
//     bidder {x, y, z, seed threshold}.  Bids come back as 16-byte records indexed by list position.
//   * Synchronisation is per cloud and one-directional: the owner publishes an iteration with a release store of
//     (iteration, ticket limit); tickets are claimed with a CAS; a finished task is a fence + atomicAdd on `done`; the owner serves
//     its own tickets while it waits, so it never depends on a CTA that is not running (no co-residency assumption, no deadlock
//     with fewer SMs than CTAs).  With at most `local_max` (32) bidders left the owner finishes the auction alone.
// Results are independent of who executes a task and in which order (order-independent top-2 update, tie rules on original
// indices): bit-exact against oracle/emd_oracle.c like the cluster kernel.
#include "pcl_emd_core.cuh"

namespace pcl {
namespace {


struct __align__(128) TeamCtl {  // one per cloud; zeroed by the host call before every launch
    unsigned long long avail;    // ((iteration + 1) << 32) | ticket limit, release-stored by the owner
    unsigned next;               // next ticket (claimed by CAS while next < limit)
    unsigned done;               // finished tickets
    int t, U, TB, KS, mode, base, pad0, pad2;  // header of the iteration the tickets [base, limit) belong to
    unsigned long long evals;    // evaluations executed by workers for this cloud (statistics)
    unsigned pad1[18];
};
static_assert(sizeof(TeamCtl) == 128, "TeamCtl is one 128-byte line");

struct TeamWs {
    TeamCtl *ctl;          // B control lines
    unsigned *finished;    // number of owners that are done
    unsigned char *clouds; // per-cloud mirror regions
    size_t stride, o_tgt, o_pf, o_tperm, o_box, o_brec, o_pub, o_jp;
};

__host__ __device__ inline size_t team_cloud_bytes(int N, size_t *o_tgt, size_t *o_pf, size_t *o_tperm, size_t *o_box, size_t *o_brec,
                                                   size_t *o_pub, size_t *o_jp) {
    const size_t n8 = (size_t)(N + 7) / 8 * 8, n32 = (size_t)(N + 31) / 32 * 32, nt = n32 / 32;
    size_t o = 0;
    auto put = [&](size_t *where, size_t bytes) { if (where) *where = o; o += (bytes + 255) / 256 * 256; };
    put(o_tgt, n32 * 16); put(o_pf, n8 * 4); put(o_tperm, n8 * 2); put(o_box, nt * 32); put(o_brec, n8 * 16); put(o_pub, n8 * 16);
    put(o_jp, n8 * 2);
    return o;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 16-byte copies global (L2, never the possibly stale L1) -> shared
__device__ __forceinline__ void copy16_in(void *dst, const void *src, int n16) {
    for (int i = threadIdx.x; i < n16; i += EMD_THREADS) reinterpret_cast<uint4 *>(dst)[i] = __ldcg(reinterpret_cast<const uint4 *>(src) + i);
}
__device__ __forceinline__ void copy16_out(void *dst, const void *src, int n16) {
    for (int i = threadIdx.x; i < n16; i += EMD_THREADS) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
}

// One bidder scanned by one warp, one lane per target of a tile (the warp-per-bidder mode of pcl_emd.cu): 32 boxes per ballot, up to
// PCL_WPB_CHUNK candidate tiles per step, rare filter survivors folded into a warp-uniform top two.  Returns the bid of the bidder.
struct WarpBid { float best, better; int bi, bi2, k3, k4; };
__device__ __forceinline__ WarpBid warp_scan_bidder(const EmdSmem &S, int NT, float ax, float ay, float az, float tm, unsigned long long &my_evals) {
    const int lane = threadIdx.x & 31;
    float best = -1e9f, better = -1e9f;
    int bi = -1, bi2 = -1, bio = 0x7fffffff, k3 = -1, k4 = -1;
    for (int tb = 0; tb < NT; tb += 32) {
        const int tl = tb + lane;
        bool cand = false;
        if (tl < NT) cand = !tile_skippable(S.tlo[tl], S.thi[tl], ax, ay, az, tm);
        unsigned cm = __ballot_sync(0xffffffffu, cand);
        constexpr int WC = PCL_WPB_CHUNK;
        while (cm) {
            int tix[WC];
            bool have[WC];
#pragma unroll
            for (int i = 0; i < WC; i++) {
                have[i] = cm != 0;
                tix[i] = tb + (have[i] ? __ffs(cm) - 1 : 0);
                cm &= cm - 1;
            }
            float sq[WC], cw[WC];
            bool pass[WC];
            bool any = false;
#pragma unroll
            for (int i = 0; i < WC; i++) {
                const float4 tq = S.tgt[tix[i] * TILE + lane];
                sq[i] = sq3_ref(__fsub_rn(tq.x, ax), __fsub_rn(tq.y, ay), __fsub_rn(tq.z, az));
                cw[i] = tq.w;
                const float u = __fsub_rn(tq.w, tm);
                pass[i] = have[i] && !(__fmaf_rn(u, u, -sq[i]) < 0.f);
                any |= pass[i];
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < WC; i++) my_evals += have[i] ? TILE : 0;
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            do {
                int sel = -1;
#pragma unroll
                for (int i = WC - 1; i >= 0; i--) sel = pass[i] ? i : sel;
                float ssel = 0.f;
                int ksel = 0;
#pragma unroll
                for (int i = 0; i < WC; i++) {
                    if (sel == i) { ssel = sq[i]; ksel = tix[i] * TILE + lane; pass[i] = false; }
                }
                float v = 0.f;
                int ko = 0;
                if (sel >= 0) { v = bid_value_exact(ssel, S.pf[ksel]); ko = S.tperm ? (int)S.tperm[ksel] : ksel; }
                unsigned pm = __ballot_sync(0xffffffffu, sel >= 0);
                if (__popc(pm) > 3) {  // many survivors: only the two largest can change (best, better)
                    const float vc = __fadd_rn(v, 0.f);
                    int key = __float_as_int(vc);
                    key ^= (key >> 31) & 0x7fffffff;
                    if (!(sel >= 0 && vc == vc)) key = (int)0x80000000;
                    const int key1 = __reduce_max_sync(0xffffffffu, key);
                    if (key1 != (int)0x80000000) {
                        const int ko1 = __reduce_min_sync(0xffffffffu, (key == key1) ? ko : 0x7fffffff);
                        const int l1 = __ffs(__ballot_sync(0xffffffffu, key == key1 && ko == ko1)) - 1;
                        const int keyr = (lane == l1) ? (int)0x80000000 : key;
                        const int key2 = __reduce_max_sync(0xffffffffu, keyr);
                        const int l2 = (key2 != (int)0x80000000) ? __ffs(__ballot_sync(0xffffffffu, keyr == key2)) - 1 : l1;
                        const float va = __shfl_sync(0xffffffffu, v, l1), vb = __shfl_sync(0xffffffffu, v, l2);
                        const int ka = __shfl_sync(0xffffffffu, ksel, l1), kb = __shfl_sync(0xffffffffu, ksel, l2);
                        const int kob = __shfl_sync(0xffffffffu, ko, l2);
                        if (va > best || (va == best && ko1 < bio)) { k4 = k3; k3 = bi2; better = best; bi2 = bi; best = va; bi = ka; bio = ko1; }
                        else if (va > better) { k4 = k3; k3 = bi2; better = va; bi2 = ka; }
                        else { k4 = k3; k3 = ka; }
                        if (key2 != (int)0x80000000) {
                            if (vb > best || (vb == best && kob < bio)) { k4 = k3; k3 = bi2; better = best; bi2 = bi; best = vb; bi = kb; bio = kob; }
                            else if (vb > better) { k4 = k3; k3 = bi2; better = vb; bi2 = kb; }
                            else { k4 = k3; k3 = kb; }
                        }
                    }
                    pm = 0;
                }
                while (pm) {
                    const int l = __ffs(pm) - 1;
                    pm &= pm - 1;
                    const float vl = __shfl_sync(0xffffffffu, v, l);
                    const int kol = __shfl_sync(0xffffffffu, ko, l), kl = __shfl_sync(0xffffffffu, ksel, l);
                    if (vl > best || (vl == best && kol < bio)) { k4 = k3; k3 = bi2; better = best; bi2 = bi; best = vl; bi = kl; bio = kol; }
                    else if (vl > better) { k4 = k3; k3 = bi2; better = vl; bi2 = kl; }
                    else { k4 = k3; k3 = kl; }
                }
                tm = fmaxf(tm, __fsub_rn(better, FILTER_MARGIN));
                any = false;
#pragma unroll
                for (int i = 0; i < WC; i++) {
                    const float u = __fsub_rn(cw[i], tm);
                    pass[i] = pass[i] && !(__fmaf_rn(u, u, -sq[i]) < 0.f);
                    any |= pass[i];
                }
            } while (__any_sync(0xffffffffu, any));
        }
    }
    return WarpBid{best, better, bi, bi2, k3, k4};
}

struct TaskHdr { int U, TB, KS, mode; };  // bidders of the iteration, bidders per task, tile slices per group, scan mode

// One task of a cloud's Bid phase, executed by a whole CTA whose shared memory holds the cloud's targets / prices / boxes (the
// owner's replica or a worker's copy of the mirror): TB consecutive bidders of the list.  mode 0: lane-per-bidder -- groups of 32
// neighbouring bidders, every group scanned in KS tile slices (TB/32*KS work items for the 16 warps, dynamic queue, slice partials
// merged by the tree of pcl_emd.cu == emd_cuda.cu:165-173).  mode 1: warp-per-bidder -- one bidder per warp at a time.
// Bids go to pub[list position] = {object | second << 16, increment bits, third | fourth << 16, 0}.
__device__ __forceinline__ void team_run_task(const EmdSmem &S, int NT, float eps, const TaskHdr &h, int task,
                                              const float4 *__restrict__ brec, const unsigned short *__restrict__ bjp,
                                              uint4 *__restrict__ pub, unsigned long long &my_evals) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int *const work_ctr = S.wsum + 48;
    const int b0 = task * h.TB;                            // first list position of the task
    const int nb = min(h.TB, h.U - b0);                    // bidders of the task
    if (tid == 0) *work_ctr = EMD_WARPS;
    __syncthreads();
    if (h.mode == 1) {
        for (int i = wid;;) {
            if (i >= nb) break;
            const float4 rec = __ldcg(&brec[b0 + i]);
            const WarpBid w = warp_scan_bidder(S, NT, rec.x, rec.y, rec.z, rec.w, my_evals);
            if (lane == 0) {
                const float inc = __fadd_rn(__fsub_rn(w.best, w.better), eps);  // emd_cuda.cu:175
                pub[b0 + i] = make_uint4((unsigned)(w.bi & 0xffff) | ((unsigned)(w.bi2 & 0xffff) << 16), __float_as_uint(inc),
                                         (unsigned)(w.k3 & 0xffff) | ((unsigned)(w.k4 & 0xffff) << 16), 0u);
                i = atomicAdd(work_ctr, 1);
            }
            i = __shfl_sync(0xffffffffu, i, 0);
        }
        return;
    }
    const int ng = (nb + 31) >> 5, KS = h.KS, GS = ng * 32;
    for (int it = wid;;) {
        if (it >= ng * KS) break;
        const int g = it % ng, sl = it / ng;
        const int bl = min(g * 32 + lane, nb - 1);        // surplus lanes shadow the last bidder (results discarded)
        const bool active = (g * 32 + lane) < nb;
        const float4 rec = __ldcg(&brec[b0 + bl]);
        const float ax = rec.x, ay = rec.y, az = rec.z;
        Top2 r = top2_init(rec.w);
        const int ntl = (NT - sl + KS - 1) / KS;           // tiles of this slice: sl, sl+KS, ...
        const int jp0 = (int)__ldcg(&bjp[b0 + g * 32]);
        const int home = min(max((jp0 / TILE - sl + KS / 2) / KS, 0), ntl - 1);
        for (int m = 0; m < ntl; m++) {                    // zig-zag outwards from the tile next to the bidders
            int q = home + ((m & 1) ? ((m + 1) >> 1) : -(m >> 1));
            q += (q < 0) ? ntl : 0;
            q -= (q >= ntl) ? ntl : 0;
            const int tl = sl + q * KS;
            if (__all_sync(0xffffffffu, tile_skippable(S.tlo[tl], S.thi[tl], ax, ay, az, r.tm))) continue;
            scan_tile(S, tl * TILE, ax, ay, az, r);
            my_evals += active ? TILE : 0;
        }
        const unsigned pack = (unsigned)(r.bi & 0xffff) | ((unsigned)(r.bi2 & 0xffff) << 16);
        const unsigned pack34 = (unsigned)(r.k3 & 0xffff) | ((unsigned)(r.k4 & 0xffff) << 16);
        if (KS == 1) {
            if (active) pub[b0 + bl] = make_uint4(pack, __float_as_uint(__fadd_rn(__fsub_rn(r.best, r.better), eps)), pack34, 0u);
        } else if (active) {
            S.pbest[sl * GS + bl] = r.best; S.pbetter[sl * GS + bl] = r.better; S.pbi[sl * GS + bl] = pack; S.pbi34[sl * GS + bl] = pack34;
        }
        if (lane == 0) it = atomicAdd(work_ctr, 1);
        it = __shfl_sync(0xffffffffu, it, 0);
    }
    if (KS > 1) {
        __syncthreads();
        int span = 1;
        while (span < KS) span <<= 1;
        for (int st = span >> 1; st >= 1; st >>= 1) {
            const int rows = min(st, KS - st);  // slices c in [0, rows) absorb slice c + st
            for (int idx = tid; idx < rows * nb; idx += EMD_THREADS) {
                const int c = idx / nb, b = idx - c * nb;
                const int me = c * GS + b, ot = (c + st) * GS + b;
                float best = S.pbest[me], better = S.pbetter[me];
                unsigned pk = S.pbi[me], pk34 = S.pbi34[me];
                const float ob = S.pbest[ot], obt = S.pbetter[ot];
                const unsigned opk = S.pbi[ot];
                bool other_wins = ob > best;
                if (ob == best && (opk & 0xffffu) != 0xffffu) {
                    const unsigned mine = pk & 0xffffu;
                    if (mine == 0xffffu) other_wins = true;
                    else {
                        const unsigned mo = S.tperm ? S.tperm[mine] : mine, oo = S.tperm ? S.tperm[opk & 0xffffu] : (opk & 0xffffu);
                        other_wins = oo < mo;
                    }
                }
                if (other_wins) {
                    const unsigned second = (best >= obt) ? (pk & 0xffffu) : (opk >> 16);
                    better = fmaxf(best, obt);
                    best = ob;
                    pk = (opk & 0xffffu) | (second << 16);
                    pk34 = S.pbi34[ot];
                } else if (ob > better) {
                    better = ob;
                    pk = (pk & 0xffffu) | ((opk & 0xffffu) << 16);
                }
                S.pbest[me] = best; S.pbetter[me] = better; S.pbi[me] = pk; S.pbi34[me] = pk34;
            }
            __syncthreads();
        }
        for (int b = tid; b < nb; b += EMD_THREADS)
            pub[b0 + b] = make_uint4(S.pbi[b], __float_as_uint(__fadd_rn(__fsub_rn(S.pbest[b], S.pbetter[b]), eps)), S.pbi34[b], 0u);
    }
}

// spin guard: a protocol bug must end in a trap (an error the host sees), never in a hung GPU
#define PCL_SPIN_LIMIT (1u << 25)

__device__ void team_worker(const EmdSmem &S, const TeamWs &W, int B, int N, float eps, long long *prof) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n8 = (N + 7) / 8 * 8, n32 = (N + 31) / 32 * 32, NT = n32 / TILE;
    int cached_c = -1, cached_t = -1;
    const int home = ((int)blockIdx.x - B) % B;
    unsigned long long my_evals = 0ull;
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc = prof ? clock64() : 0;  // development aid: idle, load, run, finish cycles; tasks, reloads
#define PCL_WTICK(i) if (prof) { const long long now_ = clock64(); pt[i] += now_ - pc; pc = now_; }
    for (;;) {
        if (wid == 0) {
            // Which task next?  First the worker's HOME cloud (workers are dealt to the clouds round-robin: one control line, no
            // contention with the workers of other clouds); else steal from the cloud that is furthest behind, ties broken by the
            // distance from home so that idle workers spread over the clouds instead of all racing for the same ticket.
            int c_sel = -1;
            unsigned ticket = 0;
            for (unsigned spin = 0;; spin++) {
                unsigned bestkey = 0xffffffffu, bestnx = 0;
                int bestc = -1;
                {
                    const unsigned long long av = ld_relaxed_u64(&W.ctl[home].avail);
                    const unsigned nx = ld_relaxed_u32(&W.ctl[home].next);
                    if (nx < (unsigned)av) { bestc = home; bestnx = nx; bestkey = 0; }
                }
                if (bestc < 0) {
                    for (int c0 = 0; c0 < B; c0 += 32) {
                        const int r = c0 + lane;                       // rotated position: cloud (home + 1 + r) mod B
                        int c = home + 1 + r;
                        c -= (c >= B) ? B : 0;
                        unsigned key = 0xffffffffu, nx = 0;
                        if (r < B - 1) {
                            const unsigned long long av = ld_relaxed_u64(&W.ctl[c].avail);
                            nx = ld_relaxed_u32(&W.ctl[c].next);
                            if (nx < (unsigned)av) key = ((unsigned)(av >> 32) << 16) | (unsigned)(r & 0xffff);
                        }
                        const unsigned k = __reduce_min_sync(0xffffffffu, key);
                        if (k < bestkey) {
                            bestkey = k;
                            const int src = __ffs(__ballot_sync(0xffffffffu, key == k)) - 1;
                            bestc = __shfl_sync(0xffffffffu, c, src);
                            bestnx = __shfl_sync(0xffffffffu, nx, src);
                        }
                    }
                }
                if (bestc >= 0) {
                    unsigned got = 0;
                    if (lane == 0) got = (atomicCAS(&W.ctl[bestc].next, bestnx, bestnx + 1) == bestnx) ? 1u : 0u;
                    got = __shfl_sync(0xffffffffu, got, 0);
                    if (got) { c_sel = bestc; ticket = bestnx; break; }
                    continue;  // lost the race: look again at once
                }
                if (ld_relaxed_u32(W.finished) >= (unsigned)B) break;  // every auction is over
                if (spin > PCL_SPIN_LIMIT) __trap();
                __nanosleep(100);
            }
            __threadfence();  // acquire side of the owner's release store: the ticket's iteration header, records and mirror are visible
            if (lane == 0) { S.wsum[56] = c_sel; S.wsum[57] = (int)ticket; }
        }
        __syncthreads();
        const int c = S.wsum[56];
        PCL_WTICK(0)
        if (c < 0) {
            if (prof && tid == 0) for (int i = 0; i < 8; i++) prof[(size_t)blockIdx.x * 16 + i] = pt[i];
            return;
        }
        const unsigned ticket = (unsigned)S.wsum[57];
        const TeamCtl *ctl = &W.ctl[c];
        TaskHdr h;
        h.U = __ldcg(&ctl->U); h.TB = __ldcg(&ctl->TB); h.KS = __ldcg(&ctl->KS); h.mode = __ldcg(&ctl->mode);
        const int t = __ldcg(&ctl->t), base = __ldcg(&ctl->base);
        const unsigned char *cl = W.clouds + (size_t)c * W.stride;
        if (c != cached_c || t != cached_t) {  // this CTA's copy of the cloud's hot state is for another cloud / iteration
            // targets + prices + boxes (+ original indices for a new cloud): every thread has all its 16-byte loads in flight at once
            const int c_tgt = n32, c_pf = n8 / 4, c_box = 2 * NT, c_tp = (c != cached_c && S.tperm) ? n8 / 8 : 0;
            const int total = c_tgt + c_pf + c_box + c_tp;
            constexpr int PER = 8;  // 16 B x 8 x 512 threads = 64 KB per round
            for (int base = 0; base < total; base += PER * EMD_THREADS) {
                uint4 v[PER];
#pragma unroll
                for (int i = 0; i < PER; i++) {
                    const int e = base + i * EMD_THREADS + tid;
                    const unsigned char *src = nullptr;
                    if (e < c_tgt) src = cl + W.o_tgt + (size_t)e * 16;
                    else if (e < c_tgt + c_pf) src = cl + W.o_pf + (size_t)(e - c_tgt) * 16;
                    else if (e < c_tgt + c_pf + c_box) src = cl + W.o_box + (size_t)(e - c_tgt - c_pf) * 16;
                    else if (e < total) src = cl + W.o_tperm + (size_t)(e - c_tgt - c_pf - c_box) * 16;
                    if (src) v[i] = __ldcg(reinterpret_cast<const uint4 *>(src));
                }
#pragma unroll
                for (int i = 0; i < PER; i++) {
                    const int e = base + i * EMD_THREADS + tid;
                    unsigned char *dst = nullptr;
                    if (e < c_tgt) dst = reinterpret_cast<unsigned char *>(S.tgt) + (size_t)e * 16;
                    else if (e < c_tgt + c_pf) dst = reinterpret_cast<unsigned char *>(S.pf) + (size_t)(e - c_tgt) * 16;
                    else if (e < c_tgt + c_pf + c_box) dst = reinterpret_cast<unsigned char *>(S.tlo) + (size_t)(e - c_tgt - c_pf) * 16;  // tlo, thi adjacent
                    else if (e < total) dst = reinterpret_cast<unsigned char *>(S.tperm) + (size_t)(e - c_tgt - c_pf - c_box) * 16;
                    if (dst) *reinterpret_cast<uint4 *>(dst) = v[i];
                }
            }
            cached_c = c; cached_t = t;
            pt[5]++;
            __syncthreads();
        }
        pt[4]++;
        PCL_WTICK(1)
        // (team_run_task starts with a block barrier: the copies are visible to every warp before the first scan)
        team_run_task(S, NT, eps, h, (int)ticket - base, reinterpret_cast<const float4 *>(cl + W.o_brec),
                      reinterpret_cast<const unsigned short *>(cl + W.o_jp), reinterpret_cast<uint4 *>(const_cast<unsigned char *>(cl) + W.o_pub), my_evals);
        PCL_WTICK(2)
        // statistics: evaluations executed for cloud c (before the task counts as done: the owner reads the total at the end)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_xor_sync(0xffffffffu, my_evals, o);
        if (lane == 0 && my_evals) atomicAdd(&W.ctl[c].evals, my_evals);
        my_evals = 0ull;
        __threadfence();   // this thread's bids are visible device-wide ...
        __syncthreads();   // ... for every thread of the CTA, before the task counts as done
        if (tid == 0) { __threadfence(); atomicAdd(&W.ctl[c].done, 1u); }
        PCL_WTICK(3)
    }
#undef PCL_WTICK
}

__global__ void __launch_bounds__(EMD_THREADS, 1)
emd_team_kernel(Pts xyz1, Pts xyz2, int B, int N, float eps, int iters, int flags, int pcap, int wpb_max, int tasks_target, int local_max,
                float *__restrict__ dist, int *__restrict__ assignment, int *__restrict__ stats, TeamWs W, float grad_scale,
                float *__restrict__ grad_xyz1, double *__restrict__ part, unsigned *__restrict__ ticket, float *__restrict__ sums_out,
                long long *__restrict__ prof) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, T = EMD_THREADS, lane = tid & 31, wid = tid >> 5;
    const EmdSmem S = carve(smem_raw, nullptr, N, flags, pcap);
    if ((int)blockIdx.x >= B) { team_worker(S, W, B, N, eps, prof); return; }
    // development aid (PCL_EMD_PROFILE): clock totals of thread 0 per phase -> prof[blockIdx.x * 16 + phase]
    long long pt[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pc = prof ? clock64() : 0;
#define PCL_TICK(i) if (prof) { const long long now_ = clock64(); pt[i] += now_ - pc; pc = now_; }

    const int cloud = blockIdx.x;
    int *const work_ctr = S.wsum + 48;
    const int n8 = (N + 7) / 8 * 8, n32 = (N + 31) / 32 * 32, NT = n32 / TILE;
    TeamCtl *const ctl = &W.ctl[cloud];
    unsigned char *const cl = W.clouds + (size_t)cloud * W.stride;
    float4 *const g_tgt = reinterpret_cast<float4 *>(cl + W.o_tgt);
    float *const g_pf = reinterpret_cast<float *>(cl + W.o_pf);
    float4 *const g_brec = reinterpret_cast<float4 *>(cl + W.o_brec);
    unsigned short *const g_jp = reinterpret_cast<unsigned short *>(cl + W.o_jp);
    uint4 *const g_pub = reinterpret_cast<uint4 *>(cl + W.o_pub);

    emd_setup(S, xyz1, xyz2, cloud, N, flags);
    __syncthreads();
    // the mirror of the hot state: targets (c = 3: price 0), prices, original target indices, tile boxes
    copy16_out(g_tgt, S.tgt, n32);
    copy16_out(g_pf, S.pf, n8 / 4);
    if (S.tperm) copy16_out(cl + W.o_tperm, S.tperm, n8 / 8);
    PCL_TICK(0)

    auto pred_xyz = [&](int jp) -> float3 {
        if (S.x1) { const float4 q = S.x1[jp]; return make_float3(q.x, q.y, q.z); }
        return ld_xyz(xyz1, cloud, S.pperm ? (int)S.pperm[jp] : jp);
    };
    long long sum_u = 0;
    unsigned long long my_evals = 0ull;
    int iters_run = 0, extra_qualifiers = 0;
    unsigned limit = 0;      // tickets handed out so far (== finished tickets between iterations)
    bool have_list = false;
    const int E = (n8 / 8 + T - 1) / T * 8;
    uint2 *const pub_cur = S.pub;  // bids of the current iteration by bidder (no double buffering: nobody else writes it)

    for (int t = 0; t < iters; t++) {
        const bool last = (t == iters - 1);
        // ---- 1. list of unassigned bidders (emd_cuda.cu:23-93) and the per-tile upper bound of c ----------------
        if (t > 0 && (!have_list || (t & 3) == 0)) {
            for (int t0 = wid; t0 < NT; t0 += 4 * EMD_WARPS) {
                int b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int tl = min(t0 + i * EMD_WARPS, NT - 1);
                    b[i] = __float_as_int(S.tgt[tl * TILE + lane].w);
                    b[i] ^= (b[i] >> 31) & 0x7fffffff;
                }
#pragma unroll
                for (int i = 0; i < 4; i++) b[i] = __reduce_max_sync(0xffffffffu, b[i]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int tl = t0 + i * EMD_WARPS;
                    b[i] ^= (b[i] >> 31) & 0x7fffffff;
                    if (lane == 0 && tl < NT) S.tlo[tl].w = __int_as_float(b[i]);
                }
            }
        }
        int U = 0;
        if (have_list) {
            U = S.wsum[40];
            if (U == 0) break;
            if (tid == 0) *work_ctr = EMD_WARPS;
        } else {
            unsigned fl = 0;
            {
                const int base = tid * E;
                for (int e = 0; e < E; e += 8) {
                    if (base + e < n8) {
                        const uint4 a = *reinterpret_cast<const uint4 *>(S.asg + base + e);
                        const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int h = 0; h < 4; h++) {
                            fl |= (unsigned)((w[h] & 0xffffu) == NONE16) << (e + 2 * h);
                            fl |= (unsigned)((w[h] >> 16) == NONE16) << (e + 2 * h + 1);
                        }
                    }
                }
            }
            const int cnt = __popc(fl);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) S.wsum[wid] = incl;
            __syncthreads();
            int wbase = 0;
#pragma unroll
            for (int w = 0; w < EMD_WARPS; w++) {
                const int v = S.wsum[w];
                if (w < wid) wbase += v;
                U += v;
            }
            if (U == 0) break;
            if (tid == 0) *work_ctr = EMD_WARPS;
            {
                int pos = wbase + incl - cnt;
                const int base = tid * E;
                while (fl) {
                    const int e = __ffs(fl) - 1;
                    fl &= fl - 1;
                    S.unass[pos++] = (unsigned short)(base + e);
                }
            }
        }
        __syncthreads();
        sum_u += U;
        iters_run = t + 1;
        PCL_TICK(1)

        // ---- 2. Bid (emd_cuda.cu:95-179) ------------------------------------------------------------------------
        const bool team = U > local_max;
        if (!team) {
            // few bidders: the owner scans them itself, one warp per bidder (a trip through L2 and a worker would cost more)
            for (int b = wid;;) {
                if (b >= U) break;
                const int jp = S.unass[b];
                const float3 a = pred_xyz(jp);
                unsigned lp = S.last[jp], lp34 = S.last34[jp];
                if (flags & EMD_F_SORT) first_seeds(jp, N, lp, lp34);
                float tm = -1e9f;
                {
                    const int k1 = (int)(lp & 0xffffu), k2 = (int)(lp >> 16);
                    if (lp != NOLAST && k1 < N && k2 < N && k1 != k2) {  // seeds: lanes 0..3 evaluate up to four of them in parallel
                        const int k3 = (int)(lp34 & 0xffffu), k4 = (int)(lp34 >> 16);
                        const bool ok3 = k3 < N && k3 != k1 && k3 != k2, ok4 = k4 < N && k4 != k1 && k4 != k2 && k4 != k3;
                        float v = -3e38f;
                        const int ks = lane == 0 ? k1 : lane == 1 ? k2 : lane == 2 ? k3 : k4;
                        if (lane < 2 || (lane == 2 && ok3) || (lane == 3 && ok4)) v = seed_value(S, ks, a.x, a.y, a.z);
                        const float v0 = __shfl_sync(0xffffffffu, v, 0), v1 = __shfl_sync(0xffffffffu, v, 1),
                                    v2 = __shfl_sync(0xffffffffu, v, 2), v3 = __shfl_sync(0xffffffffu, v, 3);
                        const float hi01 = fmaxf(v0, v1), lo01 = fminf(v0, v1), hi23 = fmaxf(v2, v3), lo23 = fminf(v2, v3);
                        tm = __fsub_rn(fmaxf(fminf(hi01, hi23), fmaxf(lo01, lo23)), FILTER_MARGIN);
                    }
                }
                const WarpBid w = warp_scan_bidder(S, NT, a.x, a.y, a.z, tm, my_evals);
                if (lane == 0) {
                    pub_cur[jp] = make_uint2((unsigned)(w.bi & 0xffff) | ((unsigned)(w.bi2 & 0xffff) << 16),
                                             __float_as_uint(__fadd_rn(__fsub_rn(w.best, w.better), eps)));
                    S.last34[jp] = (unsigned)(w.k3 & 0xffff) | ((unsigned)(w.k4 & 0xffff) << 16);
                    b = atomicAdd(work_ctr, 1);
                }
                b = __shfl_sync(0xffffffffu, b, 0);
            }
            __syncthreads();
            PCL_TICK(7)
        } else {
            // bidder records {x, y, z, seed threshold} by list position + the internal index (scan start hint) ...
            for (int b = tid; b < U; b += T) {
                const int jp = S.unass[b];
                const float3 a = pred_xyz(jp);
                unsigned lp = S.last[jp], lp34 = S.last34[jp];
                if (flags & EMD_F_SORT) first_seeds(jp, N, lp, lp34);
                g_brec[b] = make_float4(a.x, a.y, a.z, seed_threshold(S, lp, lp34, N, a.x, a.y, a.z));
                g_jp[b] = (unsigned short)jp;
            }
            // ... and the tile boxes with this iteration's upper bounds of c
            copy16_out(cl + W.o_box, S.tlo, 2 * NT);
            TaskHdr h;
            h.U = U;
            h.mode = (U <= wpb_max) ? 1 : 0;
            if (h.mode == 1) {  // warp-per-bidder: tasks of 16..64 bidders (one to four rounds of the 16 warps)
                h.TB = 16 * max(1, min(4, (U + 16 * tasks_target - 1) / (16 * tasks_target)));
                h.KS = 1;
            } else {            // lane-per-bidder: tasks of 1..8 groups of 32 bidders, ~32 work items (group x tile slice) per task
                const int Gn = (U + 31) >> 5, tg = max(1, min(8, (Gn + tasks_target - 1) / tasks_target));
                h.TB = tg * 32;
                h.KS = max(1, min(min(32 / tg, NT), pcap / h.TB));
            }
            const int ntasks = (U + h.TB - 1) / h.TB;
            __threadfence();  // mirror updates of the previous commit, records and boxes: visible before the iteration is published
            __syncthreads();
            if (tid == 0) {
                ctl->t = t; ctl->U = h.U; ctl->TB = h.TB; ctl->KS = h.KS; ctl->mode = h.mode; ctl->base = (int)limit;
                __threadfence();
                st_release_u64(&ctl->avail, ((unsigned long long)(unsigned)(t + 1) << 32) | (unsigned long long)(limit + (unsigned)ntasks));
            }
            const unsigned base = limit;
            limit += (unsigned)ntasks;
            PCL_TICK(2)
            // serve own tickets while any is left, then wait for the tasks the workers took
            for (;;) {
                if (tid == 0) {
                    int got = -1;
                    for (;;) {
                        const unsigned nx = ld_relaxed_u32(&ctl->next);
                        if (nx >= limit) break;
                        if (atomicCAS(&ctl->next, nx, nx + 1) == nx) { got = (int)(nx - base); break; }
                    }
                    S.wsum[56] = got;
                }
                __syncthreads();
                const int task = S.wsum[56];
                if (task < 0) break;
                team_run_task(S, NT, eps, h, task, g_brec, g_jp, g_pub, my_evals);
                __threadfence();
                __syncthreads();
                if (tid == 0) { __threadfence(); atomicAdd(&ctl->done, 1u); }
                pt[9]++;
            }
            pt[10] += ntasks;
            PCL_TICK(3)
            if (tid == 0) {
                for (unsigned spin = 0; ld_acquire_u32(&ctl->done) < limit; spin++) {
                    if (spin > PCL_SPIN_LIMIT) __trap();
                    __nanosleep(100);
                }
            }
            __syncthreads();
            PCL_TICK(4)
            // all bids of the iteration are in the mirror: bring them home, indexed by bidder as GetMax / Assign expect them
            for (int q = tid; q < U; q += T) {
                const uint4 rec = __ldcg(&g_pub[q]);
                const int jp = S.unass[q];
                pub_cur[jp] = make_uint2(rec.x, rec.y);
                S.last34[jp] = rec.z;
            }
            __syncthreads();
            PCL_TICK(5)
        }

        // ---- 3. GetMax + Assign (emd_cuda.cu:181-215) -------------------------------------------------------------
        for (int q = tid; q < U; q += T) {
            const int jp = S.unass[q];
            const uint2 pb = pub_cur[jp];
            S.last[jp] = pb.x;
            atomic_max_float(&S.maxinc[pb.x & 0xffffu], __uint_as_float(pb.y));  // emd_cuda.cu:176
        }
        __syncthreads();
        for (int q = tid; q < U; q += T) {
            const int jp = S.unass[q];
            const uint2 pb = pub_cur[jp];
            const int o = (int)(pb.x & 0xffffu);
            const double bi = (double)__uint_as_float(pb.y), mi = (double)S.maxinc[o];
            if (bi - 1e-6 <= mi && mi <= bi + 1e-6) {  // :188-191; the largest ORIGINAL bidder index wins
                atomicMax(&S.maxidx[o], S.pperm ? (int)S.pperm[jp] : jp);
                extra_qualifiers += 1;
            }
        }
        __syncthreads();
        unsigned keep = NONE16;
        bool loser = false;
        for (int q = tid; q < U; q += T) {
            const int jp = S.unass[q];
            const uint2 pb = pub_cur[jp];
            const int o = (int)(pb.x & 0xffffu);  // emd_cuda.cu:203-211
            const bool winner = (S.maxidx[o] == (S.pperm ? (int)S.pperm[jp] : jp));
            extra_qualifiers -= winner ? 1 : 0;
            if (!(last || winner)) { keep = (unsigned)jp; loser = true; continue; }
            const unsigned prev = S.inv[o];
            if (!last && prev != NONE16) { S.asg[prev] = NONE16; keep = prev; }
            S.inv[o] = (unsigned short)jp;
            S.asg[jp] = (unsigned short)o;
            const float pnew = __fadd_rn(S.pf[o], __uint_as_float(pb.y));
            const float cnew = __fsub_ru(3.0f, pnew);  // c = RU(3 - price): upper bound used by the filter
            S.pf[o] = pnew;
            S.tgt[o].w = cnew;
            if (team) { g_pf[o] = pnew; g_tgt[o].w = cnew; }  // U never grows: once the owner works alone the mirror is not read again
            S.maxinc[o] = -1e9f;
            S.maxidx[o] = -1;
        }
        have_list = (U <= 32);
        if (have_list && wid == 0) {
            const unsigned lose = __ballot_sync(0xffffffffu, loser);
            const unsigned evic = __ballot_sync(0xffffffffu, keep != NONE16) & ~lose;
            const unsigned below = (1u << lane) - 1u;
            const int p = ((lose >> lane) & 1u) ? __popc(lose & below) : __popc(lose) + __popc(evic & below);
            __syncwarp();
            if (keep != NONE16) S.unass[p] = (unsigned short)keep;
            if (lane == 0) S.wsum[40] = __popc(lose) + __popc(evic);
        }
        __syncthreads();
        PCL_TICK(6)
    }

    // ---- CalcDist (emd_cuda.cu:217-226) + outputs in ORIGINAL index order + fused loss epilogue (see pcl_emd.cu) ----
    __syncthreads();
    double sq_sum = 0.0;
    for (int jp = tid; jp < N; jp += T) {
        const unsigned k = S.asg[jp];
        float d = 0.f;
        float3 gr = make_float3(0.f, 0.f, 0.f);
        if (k != NONE16) {
            const float3 a = pred_xyz(jp);
            const float4 tp = S.tgt[k];
            const float dx = __fsub_rn(a.x, tp.x), dy = __fsub_rn(a.y, tp.y), dz = __fsub_rn(a.z, tp.z);
            d = sq3_ref(dx, dy, dz);
            if (grad_xyz1) {
                const float g2 = __fmul_rn(__fdiv_rn(grad_scale, __fmul_rn(2.f, __fsqrt_rn(d))), 2.f);
                gr = make_float3(__fmul_rn(g2, dx), __fmul_rn(g2, dy), __fmul_rn(g2, dz));
            }
        }
        const int jo = S.pperm ? (int)S.pperm[jp] : jp;
        const size_t o = (size_t)cloud * N + jo;
        dist[o] = d;
        assignment[o] = (k != NONE16) ? (S.tperm ? (int)S.tperm[k] : (int)k) : -1;
        if (grad_xyz1) { grad_xyz1[o * 3 + 0] = gr.x; grad_xyz1[o * 3 + 1] = gr.y; grad_xyz1[o * 3 + 2] = gr.z; }
        sq_sum += (double)__fsqrt_rn(d);
    }
    if (part) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq_sum += __shfl_xor_sync(0xffffffffu, sq_sum, o);
        double *red = reinterpret_cast<double *>(S.pbest);
        if (lane == 0) red[wid] = sq_sum;
        __syncthreads();
        if (tid == 0) {
            double b = 0.0;
            for (int w = 0; w < EMD_WARPS; w++) b += red[w];
            part[cloud] = b;
            __threadfence();
            S.wsum[41] = (atomicAdd(ticket, 1u) == (unsigned)B - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (S.wsum[41] && wid == 0) {
            __threadfence();
            double b = 0.0;
            for (int i = lane; i < B; i += 32) b += __ldcg(&part[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
            if (lane == 0) {
                const float total = (float)b, count = (float)((double)B * (double)N);
                sums_out[0] = total; sums_out[1] = count; sums_out[2] = total / count;
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            extra_qualifiers += __shfl_xor_sync(0xffffffffu, extra_qualifiers, o);
            my_evals += __shfl_xor_sync(0xffffffffu, my_evals, o);
        }
        __syncthreads();
        if (lane == 0) { S.wsum[wid] = extra_qualifiers; if (my_evals) atomicAdd(&ctl->evals, my_evals); }
        __syncthreads();
        if (tid == 0) {
            int e = 0;
            for (int w = 0; w < EMD_WARPS; w++) e += S.wsum[w];
            __threadfence();
            int *st = stats + (size_t)cloud * 8;
            st[0] = (int)sum_u; st[1] = iters_run; st[2] = e; st[3] = 0;  // cluster size 0 = team kernel
            const unsigned long long ce = atomicAdd(&ctl->evals, 0ull);
            st[4] = (int)(ce & 0xffffffffull); st[5] = (int)(ce >> 32);
            st[6] = flags; st[7] = NT;
        }
    }
    __syncthreads();
    PCL_TICK(8)
    if (prof && tid == 0) for (int i = 0; i < 12; i++) prof[(size_t)blockIdx.x * 16 + i] = pt[i];
#undef PCL_TICK
    if (tid == 0) { __threadfence(); atomicAdd(W.finished, 1u); }  // the workers leave when every owner is here
}

}  // namespace

size_t emd_team_workspace_bytes(int B, int N) {
    if (B <= 0 || N > EMD_SMEM_ONLY_N) return 0;
    return align_up((size_t)(B + 1) * sizeof(TeamCtl), 256) + (size_t)B * team_cloud_bytes(N, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

// Launches the team kernel on `st`.  team_ws: emd_team_workspace_bytes(B, N) bytes.  Everything else as emd_auction_kernel.
int emd_team_launch(const Pts &p1, const Pts &p2, int B, int N, float eps, int iters, int flags, int pcap, int wpb_max, int tasks_target,
                    int local_max, int grid, size_t smem, float *dist, int *assignment, int *stats, void *team_ws, float grad_scale,
                    float *grad_xyz1, double *part, unsigned *ticket, float *sums, long long *prof, cudaStream_t st) {
    TeamWs W;
    const size_t ctl_bytes = align_up((size_t)(B + 1) * sizeof(TeamCtl), 256);
    W.ctl = reinterpret_cast<TeamCtl *>(team_ws);
    W.finished = reinterpret_cast<unsigned *>(W.ctl + B);
    W.clouds = reinterpret_cast<unsigned char *>(team_ws) + ctl_bytes;
    W.stride = team_cloud_bytes(N, &W.o_tgt, &W.o_pf, &W.o_tperm, &W.o_box, &W.o_brec, &W.o_pub, &W.o_jp);
    PCL_CUDA(cudaMemsetAsync(team_ws, 0, ctl_bytes, st));
    static thread_local int attr_dev = -1;
    int dev = 0;
    PCL_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        DeviceInfo di;
        int rc = device_info(&di);
        if (rc) return rc;
        PCL_CUDA(cudaFuncSetAttribute(emd_team_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        attr_dev = dev;
    }
    emd_team_kernel<<<grid, EMD_THREADS, smem, st>>>(p1, p2, B, N, eps, iters, flags, pcap, wpb_max, tasks_target, local_max, dist, assignment,
                                                      stats, W, grad_scale, grad_xyz1, part, ticket, sums, prof);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

}  // namespace pcl
