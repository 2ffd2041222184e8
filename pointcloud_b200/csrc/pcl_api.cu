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

// ---- composite step: Chamfer fwd+bwd and EMD fwd+bwd of one batch, device or host buffers ------------------------
namespace {
struct StepLayout {
    size_t dist_x, idx_x, dist_y, idx_y, grad_y, emd_dist, emd_asg, scalars, ch_ws, emd_ws, total;
};
StepLayout step_layout(int B, int N) {
    StepLayout L;
    const size_t pts = align_up((size_t)B * N * 3 * sizeof(float), 256), per = align_up((size_t)B * N * 4, 256);
    size_t o = 0;
    L.dist_x = o; o += per; L.idx_x = o; o += per; L.dist_y = o; o += per; L.idx_y = o; o += per;
    L.grad_y = o; o += pts;
    L.emd_dist = o; o += per; L.emd_asg = o; o += per;
    L.scalars = o; o += 256;  // reserved
    L.ch_ws = o; o += pcl_chamfer_workspace_bytes(B, N, N);
    L.emd_ws = o; o += pcl_emd_workspace_bytes(B, N);
    L.total = o;
    return L;
}
struct HostStepLayout {
    size_t pred, target, grad_x, grad_emd, losses, step, total;
};
HostStepLayout host_layout(int B, int N) {
    HostStepLayout L;
    const size_t pts = align_up((size_t)B * N * 3 * sizeof(float), 256);
    size_t o = 0;
    L.pred = o; o += pts; L.target = o; o += pts; L.grad_x = o; o += pts; L.grad_emd = o; o += pts;
    L.losses = o; o += 256;
    L.step = o; o += step_layout(B, N).total;
    L.total = o;
    return L;
}
// Library-owned side stream: Chamfer runs next to the auction (which occupies 128 of the 148 SMs at ~45 % issue
// utilisation and leaves shared memory / registers for a few more CTAs per SM), joined before the call returns control
// of `stream` to the caller's next operation.
struct SideStream {
    int dev = -1;
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
int side_stream(SideStream **out) {
    static thread_local SideStream ss;
    int dev = 0;
    PCL_CUDA(cudaGetDevice(&dev));
    if (ss.dev != dev) {
        if (ss.side) { cudaStreamDestroy(ss.side); cudaEventDestroy(ss.fork); cudaEventDestroy(ss.join); ss = SideStream(); }
        PCL_CUDA(cudaStreamCreateWithFlags(&ss.side, cudaStreamNonBlocking));
        PCL_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
        PCL_CUDA(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
        ss.dev = dev;
    }
    *out = &ss;
    return PCL_OK;
}
}  // namespace

extern "C" size_t pcl_chamfer_emd_step_scratch_bytes(int B, int N) {
    if (B < 0 || N < 0) return 0;
    return step_layout(B, N).total;
}

extern "C" int pcl_chamfer_emd_step(const void *pred, int dtype1, int64_t bs1, int64_t rs1, const void *target, int dtype2,
                                    int64_t bs2, int64_t rs2, int B, int N, float eps, int iters, int chamfer_mode,
                                    float *losses, float *grad_pred_chamfer, float *grad_pred_emd, void *scratch,
                                    size_t scratch_bytes, void *stream) {
    if (B < 1 || N < 1) { set_error("step: bad size B=%d N=%d", B, N); return PCL_E_SHAPE; }
    if (!pred || !target || !losses || !grad_pred_chamfer || !grad_pred_emd || !scratch) { set_error("step: null argument"); return PCL_E_ARG; }
    const StepLayout L = step_layout(B, N);
    if (scratch_bytes < L.total) { set_error("step: scratch too small (%zu < %zu)", scratch_bytes, L.total); return PCL_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    SideStream *ss = nullptr;
    int rc = side_stream(&ss);
    if (rc) return rc;
    unsigned char *d = (unsigned char *)scratch;
    PCL_CUDA(cudaEventRecord(ss->fork, st));
    PCL_CUDA(cudaStreamWaitEvent(ss->side, ss->fork, 0));
    // caller's stream: ONE kernel for the whole EMD side -- auction, CalcDist, sqrt-mean (utils.py:304 with weights == 1) and the
    // gradient of that mean (emd_module.py:63-72), losses[4..6] = {sum sqrt(dist), B*N, mean}.  It takes its 128 SMs right away ...
    rc = pcl_emd_fwd_fused(pred, dtype1, bs1, rs1, target, dtype2, bs2, rs2, B, N, eps, iters, (float *)(d + L.emd_dist),
                           (int32_t *)(d + L.emd_asg), nullptr, 1.0f / ((float)B * (float)N), grad_pred_emd, losses + 4, d + L.emd_ws,
                           pcl_emd_workspace_bytes(B, N), stream);
    if (rc) return rc;
    // ... then the side stream: Chamfer forward + backward (upstream gradient 1) fill the remaining SMs.  The zero fills of the scatter
    // targets do not depend on the forward: issued first, they are off the critical path of the late-training steps (where Chamfer on the
    // 20 free SMs, not the auction, ends the step)
    PCL_CUDA(cudaMemsetAsync(grad_pred_chamfer, 0, (size_t)B * N * 3 * sizeof(float), ss->side));
    PCL_CUDA(cudaMemsetAsync(d + L.grad_y, 0, (size_t)B * N * 3 * sizeof(float), ss->side));
    rc = pcl_chamfer_fwd(pred, dtype1, bs1, rs1, nullptr, target, dtype2, bs2, rs2, nullptr, B, N, N, 3, chamfer_mode,
                         (float *)(d + L.dist_x), (int32_t *)(d + L.idx_x), (float *)(d + L.dist_y), (int32_t *)(d + L.idx_y),
                         losses, d + L.ch_ws, pcl_chamfer_workspace_bytes(B, N, N), ss->side);
    if (rc) return rc;
    rc = chamfer_bwd_impl(pred, dtype1, bs1, rs1, nullptr, target, dtype2, bs2, rs2, nullptr, B, N, N, 3, (int32_t *)(d + L.idx_x),
                          (int32_t *)(d + L.idx_y), nullptr, 1.f, 1.f, grad_pred_chamfer, (float *)(d + L.grad_y), ss->side, true);
    if (rc) return rc;
    PCL_CUDA(cudaEventRecord(ss->join, ss->side));
    PCL_CUDA(cudaStreamWaitEvent(st, ss->join, 0));
    return PCL_OK;
}

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
    float *pred = (float *)(d + L.pred), *target = (float *)(d + L.target), *losses = (float *)(d + L.losses);
    const size_t bytes = (size_t)B * N * 3 * sizeof(float);
    PCL_CUDA(cudaMemcpyAsync(pred, pred_host, bytes, cudaMemcpyHostToDevice, st));
    PCL_CUDA(cudaMemcpyAsync(target, target_host, bytes, cudaMemcpyHostToDevice, st));
    const int64_t bs = (int64_t)N * 3, rs = 3;
    int rc = pcl_chamfer_emd_step(pred, PCL_F32, bs, rs, target, PCL_F32, bs, rs, B, N, eps, iters, chamfer_mode, losses,
                                  (float *)(d + L.grad_x), (float *)(d + L.grad_emd), d + L.step, step_layout(B, N).total, stream);
    if (rc) return rc;
    // results back to the host: the 8-float `losses` vector of pcl_chamfer_emd_step
    PCL_CUDA(cudaMemcpyAsync(loss_host, losses, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (grad_pred_chamfer_host) PCL_CUDA(cudaMemcpyAsync(grad_pred_chamfer_host, d + L.grad_x, bytes, cudaMemcpyDeviceToHost, st));
    if (grad_pred_emd_host) PCL_CUDA(cudaMemcpyAsync(grad_pred_emd_host, d + L.grad_emd, bytes, cudaMemcpyDeviceToHost, st));
    return PCL_OK;
}
