// pcl_api.cu -- error plumbing, device query and the host-buffer composite entry point of libpcl_b200.
#include <stdarg.h>
#include <string.h>

#include "pcl_common.cuh"

namespace pcl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    (void)cudaGetLastError();
    return PCL_E_CUDA;
}

int device_info(DeviceInfo *out) {
    static thread_local int cached_dev = -1;
    static thread_local DeviceInfo cached;
    int dev = 0;
    PCL_CUDA(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        DeviceInfo d;
        PCL_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
        PCL_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        PCL_CUDA(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        PCL_CUDA(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if (d.cc_major != 10) {
            set_error("libpcl_b200 is built for sm_100a only; device %d is sm_%d%d", dev, d.cc_major, d.cc_minor);
            return PCL_E_UNSUPPORTED;
        }
        cached = d;
        cached_dev = dev;
    }
    *out = cached;
    return PCL_OK;
}

namespace {
__global__ void fill2_kernel(float *p, float v) { p[0] = v; p[1] = v; }
__global__ void emd_mean_kernel(const float *sums, float *out) { *out = sums[0] / sums[1]; }
}  // namespace

}  // namespace pcl

using namespace pcl;

extern "C" int pcl_version(void) { return PCL_VERSION; }
extern "C" const char *pcl_last_error(void) { return g_err; }

extern "C" int pcl_device_info(int *sm_count, int *cc_major, int *cc_minor, int *max_smem_optin) {
    DeviceInfo d;
    int rc = device_info(&d);
    if (rc) return rc;
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    if (max_smem_optin) *max_smem_optin = d.max_smem_optin;
    return PCL_OK;
}

// ---- host-buffer composite step ---------------------------------------------------------------------------
namespace {
struct HostStepLayout {
    size_t pred, target, dist_x, idx_x, dist_y, idx_y, grad_x, grad_y, emd_dist, emd_asg, grad_emd, scalars, ch_ws, emd_ws, total;
};
HostStepLayout host_layout(int B, int N) {
    HostStepLayout L;
    const size_t pts = align_up((size_t)B * N * 3 * sizeof(float), 256), per = align_up((size_t)B * N * 4, 256);
    size_t o = 0;
    L.pred = o; o += pts;  L.target = o; o += pts;
    L.dist_x = o; o += per; L.idx_x = o; o += per; L.dist_y = o; o += per; L.idx_y = o; o += per;
    L.grad_x = o; o += pts; L.grad_y = o; o += pts;
    L.emd_dist = o; o += per; L.emd_asg = o; o += per; L.grad_emd = o; o += pts;
    L.scalars = o; o += 256;  // [0..1] loss_xy, [2..3] emd sums, [4..5] ones, [6] emd mean
    L.ch_ws = o; o += pcl_chamfer_workspace_bytes(B, N, N);
    L.emd_ws = o; o += pcl_emd_workspace_bytes(B, N);
    L.total = o;
    return L;
}
}  // namespace

extern "C" size_t pcl_loss_host_scratch_bytes(int B, int N) {
    if (B < 0 || N < 0) return 0;
    return host_layout(B, N).total;
}

extern "C" int pcl_chamfer_emd_step_host(const float *pred_host, const float *target_host, int B, int N, float eps,
                                         int iters, int chamfer_mode, float *loss_host, float *grad_pred_chamfer_host,
                                         float *grad_pred_emd_host, void *dev_scratch, size_t dev_scratch_bytes,
                                         void *stream) {
    if (B < 1 || N < 1) { set_error("step_host: bad size B=%d N=%d", B, N); return PCL_E_SHAPE; }
    if (!pred_host || !target_host || !loss_host || !dev_scratch) { set_error("step_host: null argument"); return PCL_E_ARG; }
    const HostStepLayout L = host_layout(B, N);
    if (dev_scratch_bytes < L.total) { set_error("step_host: scratch too small (%zu < %zu)", dev_scratch_bytes, L.total); return PCL_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *d = (unsigned char *)dev_scratch;
    float *pred = (float *)(d + L.pred), *target = (float *)(d + L.target);
    float *sc = (float *)(d + L.scalars);
    const size_t bytes = (size_t)B * N * 3 * sizeof(float);
    PCL_CUDA(cudaMemcpyAsync(pred, pred_host, bytes, cudaMemcpyHostToDevice, st));
    PCL_CUDA(cudaMemcpyAsync(target, target_host, bytes, cudaMemcpyHostToDevice, st));
    fill2_kernel<<<1, 1, 0, st>>>(sc + 4, 1.0f);
    const int64_t bs = (int64_t)N * 3, rs = 3;
    int rc;
    // Chamfer forward + backward (upstream gradient 1)
    rc = pcl_chamfer_fwd(pred, PCL_F32, bs, rs, nullptr, target, PCL_F32, bs, rs, nullptr, B, N, N, 3, chamfer_mode,
                         (float *)(d + L.dist_x), (int32_t *)(d + L.idx_x), (float *)(d + L.dist_y),
                         (int32_t *)(d + L.idx_y), sc + 0, d + L.ch_ws, pcl_chamfer_workspace_bytes(B, N, N), stream);
    if (rc) return rc;
    rc = pcl_chamfer_bwd(pred, PCL_F32, bs, rs, nullptr, target, PCL_F32, bs, rs, nullptr, B, N, N, 3,
                         (int32_t *)(d + L.idx_x), (int32_t *)(d + L.idx_y), sc + 4, (float *)(d + L.grad_x),
                         (float *)(d + L.grad_y), stream);
    if (rc) return rc;
    // EMD forward, mean sqrt(dist) (utils.py:304 with weights == 1), backward
    rc = pcl_emd_fwd(pred, PCL_F32, bs, rs, target, PCL_F32, bs, rs, B, N, eps, iters, (float *)(d + L.emd_dist),
                     (int32_t *)(d + L.emd_asg), nullptr, d + L.emd_ws, pcl_emd_workspace_bytes(B, N), stream);
    if (rc) return rc;
    rc = pcl_emd_weighted_reduce((float *)(d + L.emd_dist), nullptr, nullptr, B, N, 0, sc + 2, d + L.emd_ws,
                                 pcl_emd_workspace_bytes(B, N), stream);
    if (rc) return rc;
    rc = pcl_emd_weighted_bwd(pred, PCL_F32, bs, rs, target, PCL_F32, bs, rs, B, N, (int32_t *)(d + L.emd_asg),
                              (float *)(d + L.emd_dist), nullptr, nullptr, 0, sc + 2, sc + 4,
                              (float *)(d + L.grad_emd), stream);
    if (rc) return rc;
    emd_mean_kernel<<<1, 1, 0, st>>>(sc + 2, sc + 6);
    PCL_CUDA(cudaGetLastError());
    // results back to the host: {chamfer_x, chamfer_y} and the EMD mean
    PCL_CUDA(cudaMemcpyAsync(loss_host, sc + 0, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    PCL_CUDA(cudaMemcpyAsync(loss_host + 2, sc + 6, sizeof(float), cudaMemcpyDeviceToHost, st));
    if (grad_pred_chamfer_host) PCL_CUDA(cudaMemcpyAsync(grad_pred_chamfer_host, d + L.grad_x, bytes, cudaMemcpyDeviceToHost, st));
    if (grad_pred_emd_host) PCL_CUDA(cudaMemcpyAsync(grad_pred_emd_host, d + L.grad_emd, bytes, cudaMemcpyDeviceToHost, st));
    return PCL_OK;
}
