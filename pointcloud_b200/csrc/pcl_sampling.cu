// pcl_sampling.cu -- farthest point sampling and ball query for sm_100a (the callers on
// the producer side of the loss path -- PointNet2's sample_and_group, models/pointnet2_utils.py:89-144 -- and the
// dataset / sensor sampler, utils.py:86-94).
//
// FPS replaces pointnet2_ops._ext.furthest_point_sampling / pytorch3d.ops.sample_farthest_points (third party, absent):
// start at index 0, keep the running minimum squared distance of every point to the selected set (initial 1e10),
// select the point with the largest running minimum, lowest index on ties -- the torch algorithm the reference keeps
// as a comment (pointnet2_utils.py:64-86).  One CTA per cloud; the cloud and its running minima live in REGISTERS
// (PPT points per thread), one __syncthreads per selected point: per-warp (max, arg, xyz) partials are reduced with
// REDUX and re-reduced redundantly by every warp from double-buffered shared memory.
//
// Ball query replaces query_ball_point (pointnet2_utils.py:93-113, a full sort over N per centroid): one warp per
// centroid scans the cloud in index order with ballot compaction and stops after nsample hits.
#include "pcl_common.cuh"

namespace pcl {
namespace {

constexpr int FPS_THREADS = 512;
constexpr int FPS_WARPS = FPS_THREADS / 32;

template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS)
fps_kernel(Pts xyz, int N, int npoint, const int *__restrict__ start, int skip_origin, int *__restrict__ out) {
    __shared__ int p_bits[2][FPS_WARPS], p_idx[2][FPS_WARPS];
    __shared__ float p_xyz[2][FPS_WARPS][3];
    const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float px[PPT], py[PPT], pz[PPT], mind[PPT];
    bool ok[PPT];
#pragma unroll
    for (int i = 0; i < PPT; i++) {
        const int k = tid + i * FPS_THREADS;
        ok[i] = k < N;
        float3 p = make_float3(0.f, 0.f, 0.f);
        if (ok[i]) p = ld_xyz(xyz, cloud, k);
        px[i] = p.x; py[i] = p.y; pz[i] = p.z;
        mind[i] = 1e10f;
        if (skip_origin && __fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(p.z, p.z)) <= 1e-3f) ok[i] = false;
    }
    int last = start ? start[cloud] : 0;
    last = min(max(last, 0), N - 1);
    float3 lp = ld_xyz(xyz, cloud, last);
    int buf = 0;
    for (int j = 0; j < npoint; j++) {
        if (tid == 0) out[(size_t)cloud * npoint + j] = last;
        if (j == npoint - 1) break;
        float best = -1.f, bx = 0.f, by = 0.f, bz = 0.f;
        int besti = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < PPT; i++) {
            if (ok[i]) {
                const float dx = __fsub_rn(px[i], lp.x), dy = __fsub_rn(py[i], lp.y), dz = __fsub_rn(pz[i], lp.z);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                const float m = d < mind[i] ? d : mind[i];
                mind[i] = m;
                if (m > best) { best = m; besti = tid + i * FPS_THREADS; bx = px[i]; by = py[i]; bz = pz[i]; }  // ascending k
            }
        }
        // warp: maximum value (non-negative floats and -1 order like their int bits), then the lowest index attaining it
        const int bits = __float_as_int(best);
        const int wmax = __reduce_max_sync(0xffffffffu, bits);
        const int wmin = __reduce_min_sync(0xffffffffu, bits == wmax ? besti : 0x7fffffff);
        if (bits == wmax && besti == wmin) {  // exactly one lane (indices are unique) -- or none when nothing is selectable
            p_bits[buf][wid] = wmax; p_idx[buf][wid] = wmin;
            p_xyz[buf][wid][0] = bx; p_xyz[buf][wid][1] = by; p_xyz[buf][wid][2] = bz;
        } else if (wmin == 0x7fffffff && lane == 0) {
            p_bits[buf][wid] = wmax; p_idx[buf][wid] = 0x7fffffff;
        }
        __syncthreads();
        // every warp reduces the FPS_WARPS partials redundantly: no second barrier, buffers alternate
        const int b2 = lane < FPS_WARPS ? p_bits[buf][lane] : (int)0x80000000;
        const int i2 = lane < FPS_WARPS ? p_idx[buf][lane] : 0x7fffffff;
        const int gmax = __reduce_max_sync(0xffffffffu, b2);
        const int gmin = __reduce_min_sync(0xffffffffu, b2 == gmax ? i2 : 0x7fffffff);
        if (gmin == 0x7fffffff) {  // no selectable point at all (skip_origin removed everything): index 0, like the oracle
            last = 0;
            lp = ld_xyz(xyz, cloud, 0);
        } else {
            const unsigned who = __ballot_sync(0xffffffffu, b2 == gmax && i2 == gmin);
            const int src = __ffs(who) - 1;
            last = gmin;
            lp.x = p_xyz[buf][src][0]; lp.y = p_xyz[buf][src][1]; lp.z = p_xyz[buf][src][2];
        }
        buf ^= 1;
    }
}

// one warp per centroid; group_idx (B, S, nsample) int32
__global__ void __launch_bounds__(256)
ball_query_kernel(Pts xyz, Pts centers, int N, int S, float r2, int nsample, int *__restrict__ group_idx) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (warp >= S) return;
    const float3 c = ld_xyz(centers, b, warp);
    int *out = group_idx + ((size_t)b * S + warp) * nsample;
    int cnt = 0, first = N;  // N = "nothing within the radius" (pointnet2_utils.py:107-112 leaves N in that case)
    for (int k0 = 0; k0 < N && cnt < nsample; k0 += 32) {
        const int k = k0 + lane;
        bool in = false;
        if (k < N) {
            const float3 p = ld_xyz(xyz, b, k);
            const float dx = __fsub_rn(c.x, p.x), dy = __fsub_rn(c.y, p.y), dz = __fsub_rn(c.z, p.z);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            in = !(d > r2);  // group_idx[sqrdists > radius ** 2] = N
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (m) {
            if (first == N) first = k0 + __ffs(m) - 1;
            const int pos = cnt + __popc(m & ((1u << lane) - 1u));
            if (in && pos < nsample) out[pos] = k;
            cnt += __popc(m);
        }
    }
    cnt = min(cnt, nsample);
    for (int p = cnt + lane; p < nsample; p += 32) out[p] = first;  // pad with the first hit (pointnet2_utils.py:110-112)
}

}  // namespace
}  // namespace pcl

using namespace pcl;

extern "C" int pcl_fps_max_points(void) { return FPS_THREADS * 32; }

extern "C" int pcl_fps(const void *xyz, int dtype, int64_t bs, int64_t rs, int B, int N, int npoint, const int32_t *start_idx,
                       int skip_origin, int32_t *idx_out, void *stream) {
    if (B < 0 || N < 1 || npoint < 0) { set_error("fps: bad size B=%d N=%d npoint=%d", B, N, npoint); return PCL_E_SHAPE; }
    if (N > FPS_THREADS * 32) { set_error("fps: N=%d > %d points per cloud is not supported", N, FPS_THREADS * 32); return PCL_E_UNSUPPORTED; }
    if (!dtype_ok(dtype)) { set_error("fps: bad dtype"); return PCL_E_ARG; }
    if (B == 0 || npoint == 0) return PCL_OK;
    if (!xyz || !idx_out) { set_error("fps: null argument"); return PCL_E_ARG; }
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const Pts p{xyz, bs, rs, dtype};
    cudaStream_t st = (cudaStream_t)stream;
    const int ppt = (N + FPS_THREADS - 1) / FPS_THREADS;
#define PCL_FPS(P) fps_kernel<P><<<B, FPS_THREADS, 0, st>>>(p, N, npoint, start_idx, skip_origin, idx_out)
    if (ppt <= 1) PCL_FPS(1);
    else if (ppt <= 2) PCL_FPS(2);
    else if (ppt <= 4) PCL_FPS(4);
    else if (ppt <= 8) PCL_FPS(8);
    else if (ppt <= 16) PCL_FPS(16);
    else PCL_FPS(32);
#undef PCL_FPS
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_ball_query(const void *xyz, int dtype, int64_t bs, int64_t rs, const void *new_xyz, int ndtype, int64_t nbs,
                              int64_t nrs, int B, int N, int S, float radius2, int nsample, int32_t *group_idx, void *stream) {
    if (B < 0 || N < 1 || S < 0 || nsample < 1) { set_error("ball_query: bad size B=%d N=%d S=%d nsample=%d", B, N, S, nsample); return PCL_E_SHAPE; }
    if (!dtype_ok(dtype) || !dtype_ok(ndtype)) { set_error("ball_query: bad dtype"); return PCL_E_ARG; }
    if (B == 0 || S == 0) return PCL_OK;
    if (B > 65535) { set_error("ball_query: B=%d > 65535", B); return PCL_E_SHAPE; }
    if (!xyz || !new_xyz || !group_idx) { set_error("ball_query: null argument"); return PCL_E_ARG; }
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const Pts p{xyz, bs, rs, dtype}, c{new_xyz, nbs, nrs, ndtype};
    ball_query_kernel<<<dim3((S * 32 + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(p, c, N, S, radius2, nsample, group_idx);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}
