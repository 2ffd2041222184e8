// pcl_epilogue.cu -- feature term of the EMD loss, fused.
//
// The reference forms it with ~20 small torch kernels per step around the auction (pointcloud_vision/utils.py:257-301):
//   Segmenter  : class weights w = cw[label of the matched target]; F.cross_entropy(logits, label, weight=cw)
//                = sum_i w_i nll_i / sum_i w_i (utils.py:295); histogram of argmax(logits) for the logged KL term (:278-279)
//   Autoencoder: F.mse_loss(pred[..., 3:], target.take_along_dim(assignment)[..., 3:]) (utils.py:257-258,301)
// Here each is one forward kernel producing (numerator, denominator) -- kept apart so that a batch-sharded caller can
// all-reduce them -- and one backward kernel.  Sums are accumulated in fp64 in a fixed order (deterministic).
#include "pcl_common.cuh"

namespace pcl {
namespace {

constexpr int EP_BLOCKS = 64, EP_THREADS = 256, EP_MAX_C = 64;

__device__ __forceinline__ void block_sum2(double a0, double a1, double *part) {
    __shared__ double s0[EP_THREADS / 32], s1[EP_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a0; s1[threadIdx.x >> 5] = a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double b0 = 0.0, b1 = 0.0;
        for (int w = 0; w < EP_THREADS / 32; w++) { b0 += s0[w]; b1 += s1[w]; }
        part[blockIdx.x * 2] = b0; part[blockIdx.x * 2 + 1] = b1;
    }
}

__global__ void ep_stage2(const double *__restrict__ part, int nb, float den_override, float *__restrict__ sums) {
    if (threadIdx.x == 0) {
        double b0 = 0.0, b1 = 0.0;
        for (int i = 0; i < nb; i++) { b0 += part[i * 2]; b1 += part[i * 2 + 1]; }
        sums[0] = (float)b0;
        sums[1] = (den_override >= 0.f) ? den_override : (float)b1;
    }
}

// one thread per point: log-sum-exp over the C logits, nll of the matched label, class weight; argmax histogram
__global__ void __launch_bounds__(EP_THREADS)
seg_ce_fwd_kernel(Pts logits, const int *__restrict__ matched, const float *__restrict__ cw, int B, int N, int C,
                  double *__restrict__ part, unsigned long long *__restrict__ pred_hist) {
    __shared__ unsigned int sh[EP_MAX_C];
    for (int c = threadIdx.x; c < C; c += blockDim.x) sh[c] = 0;
    __syncthreads();
    double a0 = 0.0, a1 = 0.0;
    const size_t total = (size_t)B * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int64_t o = (int64_t)(e / N) * logits.bs + (int64_t)(e % N) * logits.rs;
        float m = ld_any(logits, o);
        int am = 0;
        for (int c = 1; c < C; c++) { const float v = ld_any(logits, o + c); if (v > m) { m = v; am = c; } }  // first maximum
        float se = 0.f;
        for (int c = 0; c < C; c++) se += expf(ld_any(logits, o + c) - m);
        const int l = matched[e];
        if (l >= 0 && l < C) {
            const float w = cw[l];
            const float nll = (m + logf(se)) - ld_any(logits, o + l);
            a0 += (double)__fmul_rn(w, nll);
            a1 += (double)w;
        }
        if (pred_hist) atomicAdd(&sh[am], 1u);
    }
    __syncthreads();
    if (pred_hist)
        for (int c = threadIdx.x; c < C; c += blockDim.x)
            if (sh[c]) atomicAdd(&pred_hist[c], (unsigned long long)sh[c]);
    block_sum2(a0, a1, part);
}

// d/d logits of g * sum_i w_i nll_i = g * w_i * (softmax_i - onehot(label_i))
__global__ void __launch_bounds__(EP_THREADS)
seg_ce_bwd_kernel(Pts logits, const int *__restrict__ matched, const float *__restrict__ cw, int B, int N, int C,
                  const float *__restrict__ grad_sums, float *__restrict__ grad) {
    const size_t total = (size_t)B * N;
    const float g = __ldg(grad_sums);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int64_t o = (int64_t)(e / N) * logits.bs + (int64_t)(e % N) * logits.rs;
        float m = ld_any(logits, o);
        for (int c = 1; c < C; c++) m = fmaxf(m, ld_any(logits, o + c));
        float se = 0.f;
        for (int c = 0; c < C; c++) se += expf(ld_any(logits, o + c) - m);
        const int l = matched[e];
        const bool ok = l >= 0 && l < C;
        const float gw = ok ? __fmul_rn(g, cw[l]) : 0.f;
        const float inv = 1.f / se;
        for (int c = 0; c < C; c++) {
            const float p = expf(ld_any(logits, o + c) - m) * inv;
            grad[e * C + c] = gw * (p - ((c == l) ? 1.f : 0.f));
        }
    }
}

// sum over (b, i, f) of (feat[b,i,f] - tfeat[b, assignment[b,i], f])^2
__global__ void __launch_bounds__(EP_THREADS)
feat_mse_fwd_kernel(Pts feat, Pts tfeat, const int *__restrict__ assignment, int B, int N, int F, double *__restrict__ part) {
    double a0 = 0.0;
    const size_t total = (size_t)B * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int64_t b = (int64_t)(e / N);
        int k = assignment[e];
        k = (k >= 0 && k < N) ? k : 0;  // unmatched points do not occur with iters >= 1 (forced assignment, emd_cuda.cu:201)
        const int64_t o = b * feat.bs + (int64_t)(e % N) * feat.rs, ot = b * tfeat.bs + (int64_t)k * tfeat.rs;
        for (int f = 0; f < F; f++) {
            const float d = __fsub_rn(ld_any(feat, o + f), ld_any(tfeat, ot + f));
            a0 += (double)__fmul_rn(d, d);
        }
    }
    block_sum2(a0, 0.0, part);
}

__global__ void __launch_bounds__(EP_THREADS)
feat_mse_bwd_kernel(Pts feat, Pts tfeat, const int *__restrict__ assignment, int B, int N, int F,
                    const float *__restrict__ grad_sums, float *__restrict__ grad) {
    const size_t total = (size_t)B * N;
    const float g2 = __fmul_rn(2.f, __ldg(grad_sums));
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int64_t b = (int64_t)(e / N);
        int k = assignment[e];
        k = (k >= 0 && k < N) ? k : 0;
        const int64_t o = b * feat.bs + (int64_t)(e % N) * feat.rs, ot = b * tfeat.bs + (int64_t)k * tfeat.rs;
        for (int f = 0; f < F; f++) grad[e * F + f] = __fmul_rn(g2, __fsub_rn(ld_any(feat, o + f), ld_any(tfeat, ot + f)));
    }
}

// Per-cloud class filter of the target + left-packed copy of the kept xyz (stable order), zero padding and the number of
// kept points -- FilteringChamferDistance's per-cloud Python loop (utils.py:110-124,222-226) as one launch without a host
// synchronisation.  One CTA per cloud; a running offset carries the block-wide prefix from chunk to chunk.
constexpr int CF_THREADS = 256, CF_MAX_LABELS = 16;
struct LabelSet { int n; long long v[CF_MAX_LABELS]; };

__global__ void __launch_bounds__(CF_THREADS)
class_filter_kernel(Pts target, int N, int label_channel, LabelSet labels, float *__restrict__ out_xyz, long long *__restrict__ lengths) {
    __shared__ int wsum[CF_THREADS / 32];
    __shared__ int base_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    float *out = out_xyz + (size_t)b * N * 3;
    for (int i0 = 0; i0 < N; i0 += CF_THREADS) {
        const int i = i0 + tid;
        bool keep = false;
        float3 p = make_float3(0.f, 0.f, 0.f);
        if (i < N) {
            const int64_t o = (int64_t)b * target.bs + (int64_t)i * target.rs;
            const long long lab = (long long)ld_any(target, o + label_channel);  // .long(): truncation (utils.py:119)
            for (int k = 0; k < labels.n; k++) keep = keep || (lab == labels.v[k]);
            if (keep) p = make_float3(ld_any(target, o), ld_any(target, o + 1), ld_any(target, o + 2));
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[wid] = __popc(m);
        __syncthreads();
        int off = base_s, tot = 0;
        for (int w = 0; w < CF_THREADS / 32; w++) { const int v = wsum[w]; off += (w < wid) ? v : 0; tot += v; }
        if (keep) {
            const int pos = off + __popc(m & ((1u << lane) - 1u));
            out[pos * 3 + 0] = p.x; out[pos * 3 + 1] = p.y; out[pos * 3 + 2] = p.z;
        }
        __syncthreads();
        if (tid == 0) base_s += tot;
    }
    __syncthreads();
    const int len = base_s;
    for (int e = len * 3 + tid; e < N * 3; e += CF_THREADS) out[e] = 0.f;  // F.pad zeros (utils.py:226)
    if (tid == 0) lengths[b] = len;
}

int ep_blocks(size_t total) {  // grid-stride kernels: at most four CTAs per SM of the current device
    DeviceInfo di;
    const size_t cap = 4 * (size_t)(device_info(&di) == PCL_OK ? di.sm_count : 148);
    const size_t b = (total + EP_THREADS - 1) / EP_THREADS;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace pcl

using namespace pcl;

extern "C" size_t pcl_emd_feature_workspace_bytes(void) { return align_up((size_t)EP_BLOCKS * 2 * sizeof(double), 256); }

extern "C" int pcl_emd_seg_ce_fwd(const void *logits, int dtype, int64_t bs, int64_t rs, const int32_t *matched_label,
                                  const float *class_weights, int B, int N, int C, float *sums, int64_t *pred_hist,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || N < 1 || C < 1 || C > EP_MAX_C) { set_error("emd_seg_ce_fwd: bad size B=%d N=%d C=%d (C <= %d)", B, N, C, EP_MAX_C); return PCL_E_SHAPE; }
    if (!dtype_ok(dtype) || !sums || !class_weights) { set_error("emd_seg_ce_fwd: bad argument"); return PCL_E_ARG; }
    if (!workspace || workspace_bytes < pcl_emd_feature_workspace_bytes()) { set_error("emd_seg_ce_fwd: workspace too small"); return PCL_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    if (pred_hist) PCL_CUDA(cudaMemsetAsync(pred_hist, 0, (size_t)C * sizeof(int64_t), st));
    if (B > 0 && (!logits || !matched_label)) { set_error("emd_seg_ce_fwd: null input"); return PCL_E_ARG; }
    const Pts lg{logits, bs, rs, dtype};
    double *part = (double *)workspace;
    seg_ce_fwd_kernel<<<EP_BLOCKS, EP_THREADS, 0, st>>>(lg, matched_label, class_weights, B, N, C, part, (unsigned long long *)pred_hist);
    PCL_CUDA(cudaGetLastError());
    ep_stage2<<<1, 32, 0, st>>>(part, EP_BLOCKS, -1.f, sums);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_emd_seg_ce_bwd(const void *logits, int dtype, int64_t bs, int64_t rs, const int32_t *matched_label,
                                  const float *class_weights, int B, int N, int C, const float *grad_sums, float *grad_logits,
                                  void *stream) {
    if (B < 0 || N < 1 || C < 1 || C > EP_MAX_C) { set_error("emd_seg_ce_bwd: bad size B=%d N=%d C=%d", B, N, C); return PCL_E_SHAPE; }
    if (!dtype_ok(dtype)) { set_error("emd_seg_ce_bwd: bad dtype"); return PCL_E_ARG; }
    if (B == 0) return PCL_OK;
    if (!logits || !matched_label || !class_weights || !grad_sums || !grad_logits) { set_error("emd_seg_ce_bwd: null argument"); return PCL_E_ARG; }
    const Pts lg{logits, bs, rs, dtype};
    seg_ce_bwd_kernel<<<ep_blocks((size_t)B * N), EP_THREADS, 0, (cudaStream_t)stream>>>(lg, matched_label, class_weights, B, N, C, grad_sums, grad_logits);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_emd_feat_mse_fwd(const void *feat, int dtype1, int64_t bs1, int64_t rs1, const void *tfeat, int dtype2,
                                    int64_t bs2, int64_t rs2, const int32_t *assignment, int B, int N, int F, float *sums,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || N < 1 || F < 0) { set_error("emd_feat_mse_fwd: bad size B=%d N=%d F=%d", B, N, F); return PCL_E_SHAPE; }
    if (!dtype_ok(dtype1) || !dtype_ok(dtype2) || !sums) { set_error("emd_feat_mse_fwd: bad argument"); return PCL_E_ARG; }
    if (!workspace || workspace_bytes < pcl_emd_feature_workspace_bytes()) { set_error("emd_feat_mse_fwd: workspace too small"); return PCL_E_WORKSPACE; }
    if (B > 0 && F > 0 && (!feat || !tfeat || !assignment)) { set_error("emd_feat_mse_fwd: null input"); return PCL_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const Pts a{feat, bs1, rs1, dtype1}, t{tfeat, bs2, rs2, dtype2};
    double *part = (double *)workspace;
    feat_mse_fwd_kernel<<<EP_BLOCKS, EP_THREADS, 0, st>>>(a, t, assignment, B, N, F, part);
    PCL_CUDA(cudaGetLastError());
    ep_stage2<<<1, 32, 0, st>>>(part, EP_BLOCKS, (float)((double)B * N * F), sums);  // denominator: the element count (mse 'mean')
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_emd_feat_mse_bwd(const void *feat, int dtype1, int64_t bs1, int64_t rs1, const void *tfeat, int dtype2,
                                    int64_t bs2, int64_t rs2, const int32_t *assignment, int B, int N, int F,
                                    const float *grad_sums, float *grad_feat, void *stream) {
    if (B < 0 || N < 1 || F < 0) { set_error("emd_feat_mse_bwd: bad size B=%d N=%d F=%d", B, N, F); return PCL_E_SHAPE; }
    if (!dtype_ok(dtype1) || !dtype_ok(dtype2)) { set_error("emd_feat_mse_bwd: bad dtype"); return PCL_E_ARG; }
    if (B == 0 || F == 0) return PCL_OK;
    if (!feat || !tfeat || !assignment || !grad_sums || !grad_feat) { set_error("emd_feat_mse_bwd: null argument"); return PCL_E_ARG; }
    const Pts a{feat, bs1, rs1, dtype1}, t{tfeat, bs2, rs2, dtype2};
    feat_mse_bwd_kernel<<<ep_blocks((size_t)B * N), EP_THREADS, 0, (cudaStream_t)stream>>>(a, t, assignment, B, N, F, grad_sums, grad_feat);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}

extern "C" int pcl_class_filter(const void *target, int dtype, int64_t bs, int64_t rs, int B, int N, int label_channel,
                                const int64_t *labels /* host */, int n_labels, float *out_xyz, int64_t *lengths, void *stream) {
    if (B < 0 || N < 1 || label_channel < 0) { set_error("class_filter: bad size B=%d N=%d channel=%d", B, N, label_channel); return PCL_E_SHAPE; }
    if (n_labels < 0 || n_labels > CF_MAX_LABELS || (n_labels > 0 && !labels)) { set_error("class_filter: 0..%d labels", CF_MAX_LABELS); return PCL_E_ARG; }
    if (!dtype_ok(dtype)) { set_error("class_filter: bad dtype"); return PCL_E_ARG; }
    if (B == 0) return PCL_OK;
    if (!target || !out_xyz || !lengths) { set_error("class_filter: null argument"); return PCL_E_ARG; }
    LabelSet ls;
    ls.n = n_labels;
    for (int k = 0; k < CF_MAX_LABELS; k++) ls.v[k] = (k < n_labels) ? (long long)labels[k] : 0;
    const Pts t{target, bs, rs, dtype};
    class_filter_kernel<<<B, CF_THREADS, 0, (cudaStream_t)stream>>>(t, N, label_channel, ls, out_xyz, (long long *)lengths);
    PCL_CUDA(cudaGetLastError());
    return PCL_OK;
}
