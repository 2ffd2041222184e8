// tools/microbench.cu -- per-SM instruction throughput probes on B200 used to choose the EMD bid arithmetic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Each kernel runs ITER iterations of U independent dependency chains of one operation per thread;
// reported: warp-instructions per clock per SM (relative to the FFMA line measured in the same run).
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITER = 4096;
constexpr int U = 8;

enum Op { FFMA, FADD, FMNMX, IADD, DADD, F2D, D2F, RSQ, SQRT_RN, BID_REF, BID_MANUAL, BID_F32ONLY, LDS_B128, FMUL, DIST_UNFUSED, DIST_FMA, DIST_UNFUSED_LDS };

__device__ __forceinline__ float manual_d2f(double t) {  // RNE double->float for normal-range positive results
    const unsigned lo = (unsigned)__double2loint(t), hi = (unsigned)__double2hiint(t);
    const unsigned trunc = __funnelshift_l(lo, hi, 3);
    const unsigned rem = lo << 3;
    const unsigned long long v = ((unsigned long long)trunc << 32 | rem) + 0x7fffffffull + (trunc & 1u);
    return __uint_as_float((unsigned)(v >> 32) + 0x40000000u);
}
__device__ __forceinline__ double manual_f2d(float r) {  // exact for normal positive floats
    const unsigned b = __float_as_uint(r);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}

template <int OP>
__global__ void __launch_bounds__(256) probe(float *out, float seed, double dseed) {
    float f[U];
    double d[U];
    int n[U];
#pragma unroll
    for (int u = 0; u < U; u++) { f[u] = seed + threadIdx.x * 1e-3f + u; d[u] = dseed + threadIdx.x * 1e-3 + u; n[u] = threadIdx.x + u; }
    __shared__ float4 tile[256];
    tile[threadIdx.x] = make_float4(seed, seed * 2, seed * 3, seed * 4);
    __syncthreads();
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (OP == FFMA) f[u] = __fmaf_rn(f[u], 1.0001f, 0.5f);
            else if (OP == FADD) f[u] = __fadd_rn(f[u], seed);
            else if (OP == FMUL) f[u] = __fmul_rn(f[u], 1.0000001f);
            else if (OP == DIST_UNFUSED || OP == DIST_FMA || OP == DIST_UNFUSED_LDS) {
                float4 t = make_float4(seed + i, seed * 2, seed * 3 + u, 0.f);
                if (OP == DIST_UNFUSED_LDS) t = tile[(i + (u >> 2)) & 255];
                const float qx = (float)n[u], qy = (float)d[u], qz = seed * u;
                const float dx = __fsub_rn(qx, t.x), dy = __fsub_rn(qy, t.y), dz = __fsub_rn(qz, t.z);
                float dd;
                if (OP == DIST_FMA) dd = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                else dd = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                f[u] = fminf(f[u], dd);
            }
            else if (OP == FMNMX) f[u] = fmaxf(f[u], seed + (float)i);
            else if (OP == IADD) n[u] = (n[u] ^ i) + u;
            else if (OP == DADD) d[u] = __dadd_rn(d[u], dseed);
            else if (OP == F2D) { d[u] = (double)f[u]; f[u] = __int_as_float(__double2loint(d[u]) ^ __double2hiint(d[u])); }
            else if (OP == D2F) { f[u] = __double2float_rn(d[u]); d[u] = __hiloint2double(__float_as_int(f[u]), i); }
            else if (OP == RSQ) f[u] = rsqrtf(f[u]) + 2.f;
            else if (OP == SQRT_RN) f[u] = __fsqrt_rn(f[u]) + 2.f;
            else if (OP == LDS_B128) { const float4 t = tile[(i + u) & 255]; f[u] += t.x; }
            else {
                // one bid evaluation; f[u] plays best, n[u] best index
                const float4 t = tile[(i * U + u) & 255];
                const float dx = __fsub_rn(t.x, seed), dy = __fsub_rn(t.y, seed), dz = __fsub_rn(t.z, seed);
                const float s = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                float v;
                if (OP == BID_REF) {
                    const float r = __fsqrt_rn(s);
                    v = __double2float_rn(__dsub_rn(__dsub_rn(3.0, (double)r), dseed));
                } else if (OP == BID_MANUAL) {
                    const float sc = fmaxf(s, 1e-30f);
                    float y;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(sc));
                    const float g = __fmul_rn(sc, y), h = __fmul_rn(y, 0.5f);
                    const float r = __fmaf_rn(__fmaf_rn(-g, g, sc), h, g);
                    v = manual_d2f(__dsub_rn(__dsub_rn(3.0, manual_f2d(r)), dseed));
                } else {
                    v = __fsub_rn(__fsub_rn(3.0f, __fsqrt_rn(s)), t.w);
                }
                const bool p = v > f[u];
                d[u] = __hiloint2double(0, __float_as_int(fmaxf(__int_as_float(__double2loint(d[u])), fminf(v, f[u]))));
                f[u] = fmaxf(f[u], v);
                n[u] = p ? i : n[u];
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < U; u++) acc += f[u] + (float)d[u] + (float)n[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
double run(const char *name, float *out, int sms, double ffma_rate) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sms * 8;  // 2048 threads per SM
    probe<OP><<<blocks, 256>>>(out, 1.5f, 0.25);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) probe<OP><<<blocks, 256>>>(out, 1.5f, 0.25);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = 5.0 * blocks * 256.0 * ITER * U;  // lane-ops
    const double rate = ops / (ms * 1e-3) / sms;          // lane-ops per second per SM
    printf("%-12s %8.3f ms  %8.2f Glane-op/s/SM  rel_to_FFMA=%.3f  (cudaErr=%d)\n", name, ms / 5, rate * 1e-9,
           ffma_rate > 0 ? rate / ffma_rate : 1.0, (int)cudaGetLastError());
    return rate;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    float *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * sizeof(float));
    const int sms = p.multiProcessorCount;
    const double r0 = run<FFMA>("FFMA", out, sms, 0);
    printf("  => FFMA lane-ops/clk/SM at %d kHz nominal: %.1f\n", p.clockRate, r0 / (p.clockRate * 1e3));
    run<FADD>("FADD", out, sms, r0);
    run<FMNMX>("FMNMX", out, sms, r0);
    run<IADD>("LOP+IADD(2)", out, sms, r0);
    run<DADD>("DADD", out, sms, r0);
    run<F2D>("F2F.64.32+2", out, sms, r0);
    run<D2F>("F2F.32.64+1", out, sms, r0);
    run<RSQ>("MUFU.RSQ+1", out, sms, r0);
    run<SQRT_RN>("sqrt.rn+1", out, sms, r0);
    run<LDS_B128>("LDS.128+1", out, sms, r0);
    run<FMUL>("FMUL", out, sms, r0);
    run<DIST_UNFUSED>("dist_unfused(9)", out, sms, r0);
    run<DIST_FMA>("dist_fma(7)", out, sms, r0);
    run<DIST_UNFUSED_LDS>("dist_unf+LDS/4", out, sms, r0);
    run<BID_REF>("bid_ref", out, sms, r0);
    run<BID_MANUAL>("bid_manual", out, sms, r0);
    run<BID_F32ONLY>("bid_f32only", out, sms, r0);
    return 0;
}
