// tools/microbench_packed.cu -- issue-rate probes of the sm_100a packed FP32 instructions (FFMA2 / FADD2 / FMUL2)
// and the three-input FMNMX3, alone and in the instruction mixes of the Chamfer / EMD inner loops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_packed tools/microbench_packed.cu
// Reported: lane-FLOP-instructions (one FFMA2 = 2) and warp instructions per clock per SM at the nominal clock.
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITER = 2048;
constexpr int U = 8;
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmin2(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

enum Op { P_FFMA, P_FFMA2, P_FADD2, P_FMUL2, P_FMNMX, P_FMNMX3, MIX_APPROX_SCALAR, MIX_APPROX_PACKED, MIX_APPROX_PACKED_LDS,
          MIX_EXACT_SCALAR, MIX_EXACT_PACKED, MIX_EXACT_PACKED_LDS, MIX_EMD_SCALAR, MIX_EMD_PACKED, MIX_EMD_EXPAND_PACKED };

// per probe: how many "lane-ops" (scalar-equivalent FP instructions) and warp instructions one inner step of one chain issues
template <int OP>
__global__ void __launch_bounds__(256) probe(float *out, float seed) {
    float f[U];
    u64 d[U];
#pragma unroll
    for (int u = 0; u < U; u++) { f[u] = seed + threadIdx.x * 1e-3f + u; d[u] = pk(f[u], f[u] + 1.f); }
    __shared__ float4 tile[512];
    tile[threadIdx.x] = make_float4(seed, seed * 2, seed * 3, seed * 4);
    tile[threadIdx.x + 256] = make_float4(seed * 5, seed * 6, seed * 7, seed * 8);
    __syncthreads();
    const u64 c1 = pk(1.0001f, 0.9999f), c2 = pk(0.5f, 0.25f);
    const u64 qx = pk(seed * 0.3f, seed * 0.3f), qy = pk(seed * 0.7f, seed * 0.7f), qz = pk(seed * 0.9f, seed * 0.9f);
    const float sx = seed * 0.3f, sy = seed * 0.7f, sz = seed * 0.9f;
    for (int i = 0; i < ITER; i++) {
        if (OP == P_FFMA) {
#pragma unroll
            for (int u = 0; u < U; u++) f[u] = ffma(f[u], 1.0001f, 0.5f);
        } else if (OP == P_FFMA2) {
#pragma unroll
            for (int u = 0; u < U; u++) d[u] = ffma2(d[u], c1, c2);
        } else if (OP == P_FADD2) {
#pragma unroll
            for (int u = 0; u < U; u++) d[u] = fadd2(d[u], c2);
        } else if (OP == P_FMUL2) {
#pragma unroll
            for (int u = 0; u < U; u++) d[u] = fmul2(d[u], c1);
        } else if (OP == P_FMNMX) {
#pragma unroll
            for (int u = 0; u < U; u++) f[u] = fmin2(f[u], seed + (float)i);
        } else if (OP == P_FMNMX3) {
#pragma unroll
            for (int u = 0; u < U; u++) f[u] = fmin3(f[u], seed + (float)i, seed);
        } else if (OP == MIX_APPROX_SCALAR) {  // per eval: 3 FFMA + 1 FMNMX; U evals per step (targets from registers)
            const float tx = seed + i, ty = seed * 2 + i, tz = seed * 3, tw = seed * 4;
#pragma unroll
            for (int u = 0; u < U; u++) f[u] = fmin2(f[u], ffma(tz, sz + u, ffma(ty, sy + u, ffma(tx, sx + u, tw))));
        } else if (OP == MIX_APPROX_PACKED || OP == MIX_APPROX_PACKED_LDS) {  // per 2 evals: 3 FFMA2 + 1 FMNMX3; U queries x 2 targets
            u64 tx, ty, tz, tw;
            if (OP == MIX_APPROX_PACKED_LDS) {
                const float4 a = tile[(2 * i) & 511], b = tile[(2 * i + 1) & 511];  // {x0,x1,y0,y1} {z0,z1,w0,w1}
                tx = pk(a.x, a.y); ty = pk(a.z, a.w); tz = pk(b.x, b.y); tw = pk(b.z, b.w);
            } else { tx = pk(seed + i, seed - i); ty = pk(seed * 2 + i, seed * 2); tz = pk(seed * 3, seed * 3 + i); tw = pk(seed * 4, seed * 5); }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const u64 a = ffma2(tz, qz + u, ffma2(ty, qy + u, ffma2(tx, qx + u, tw)));
                f[u] = fmin3(f[u], lo(a), hi(a));
            }
        } else if (OP == MIX_EXACT_SCALAR) {  // 3 FADD + 3 FMUL + 2 FADD + FMNMX per eval
            const float tx = seed + i, ty = seed * 2 + i, tz = seed * 3;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const float dx = fadd(tx, sx + u), dy = fadd(ty, sy + u), dz = fadd(tz, sz + u);
                f[u] = fmin2(f[u], fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
            }
        } else if (OP == MIX_EXACT_PACKED || OP == MIX_EXACT_PACKED_LDS) {  // 3 FADD2 + 3 FMUL2 + 2 FADD2 + FMNMX3 per 2 evals
            u64 tx, ty, tz;
            if (OP == MIX_EXACT_PACKED_LDS) {
                const float4 a = tile[(2 * i) & 511], b = tile[(2 * i + 1) & 511];
                tx = pk(a.x, a.y); ty = pk(a.z, a.w); tz = pk(b.x, b.y);
            } else { tx = pk(seed + i, seed - i); ty = pk(seed * 2 + i, seed * 2); tz = pk(seed * 3, seed * 3 + i); }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const u64 dx = fadd2(tx, qx + u), dy = fadd2(ty, qy + u), dz = fadd2(tz, qz + u);
                const u64 s = fadd2(fadd2(fmul2(dx, dx), fmul2(dy, dy)), fmul2(dz, dz));
                f[u] = fmin3(f[u], lo(s), hi(s));
            }
        } else if (OP == MIX_EMD_SCALAR) {  // EMD filter per eval: 3 FADD, FMUL, 2 FFMA, FADD, FFMA (+ max per 4: here 1 FMNMX per eval)
            const float tx = seed + i, ty = seed * 2 + i, tz = seed * 3, tw = seed * 4;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const float dx = fadd(tx, sx + u), dy = fadd(ty, sy + u), dz = fadd(tz, sz + u);
                const float s = ffma(dz, dz, ffma(dx, dx, fmul(dy, dy)));
                const float w = fadd(tw, sx);
                f[u] = fmin2(f[u], ffma(w, w, -s));
            }
        } else if (OP == MIX_EMD_PACKED) {  // same, 2 targets per instruction
            const u64 tx = pk(seed + i, seed - i), ty = pk(seed * 2 + i, seed * 2), tz = pk(seed * 3, seed * 3 + i), tw = pk(seed * 4, seed * 5);
            const u64 m1 = pk(-1.f, -1.f);
#pragma unroll
            for (int u = 0; u < U; u++) {
                const u64 dx = fadd2(tx, qx + u), dy = fadd2(ty, qy + u), dz = fadd2(tz, qz + u);
                const u64 s = ffma2(dz, dz, ffma2(dx, dx, fmul2(dy, dy)));
                const u64 w = fadd2(tw, qx);
                const u64 e = ffma2(w, w, fmul2(s, m1));
                f[u] = fmin3(f[u], lo(e), hi(e));
            }
        } else if (OP == MIX_EMD_EXPAND_PACKED) {  // expansion form: 4 FFMA2 + FMNMX3 per 2 targets
            const u64 tx = pk(seed + i, seed - i), ty = pk(seed * 2 + i, seed * 2), tz = pk(seed * 3, seed * 3 + i), tw = pk(seed * 4, seed * 5);
#pragma unroll
            for (int u = 0; u < U; u++) {
                const u64 a = ffma2(tw, c1 + u, ffma2(tz, qz + u, ffma2(ty, qy + u, ffma2(tx, qx + u, tw))));
                f[u] = fmin3(f[u], lo(a), hi(a));
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < U; u++) acc += f[u] + lo(d[u]) + hi(d[u]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char *name, float *out, int sms, double khz, double evals_per_step, double winstr_per_step, double fmapipe_lane_ops_per_step,
         int threads_per_sm = 2048) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sms * (threads_per_sm / 256);
    probe<OP><<<blocks, 256>>>(out, 1.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) probe<OP><<<blocks, 256>>>(out, 1.5f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double steps = 5.0 * blocks * 256.0 * ITER;  // per-thread inner steps
    const double clk = (ms * 1e-3) * khz * 1e3;         // clocks at nominal
    printf("%-24s thr/SM=%4d %8.3f ms | evals/clk/SM %7.2f | warp-instr/clk/SM %5.2f | fma-pipe lane-ops/clk/SM %6.1f (err=%d)\n", name,
           threads_per_sm, ms / 5, steps * evals_per_step / clk / sms, steps * winstr_per_step / 32 / clk / sms,
           steps * fmapipe_lane_ops_per_step / clk / sms, (int)cudaGetLastError());
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    float *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * sizeof(float));
    const int sms = p.multiProcessorCount;
    const double khz = p.clockRate;
    for (int tps : {2048, 512}) {
        run<P_FFMA>("FFMA", out, sms, khz, U, U, U, tps);
        run<P_FFMA2>("FFMA2", out, sms, khz, 2 * U, U, 2 * U, tps);
        run<P_FADD2>("FADD2", out, sms, khz, 2 * U, U, 2 * U, tps);
        run<P_FMUL2>("FMUL2", out, sms, khz, 2 * U, U, 2 * U, tps);
        run<P_FMNMX>("FMNMX", out, sms, khz, U, U, 0, tps);
        run<P_FMNMX3>("FMNMX3", out, sms, khz, U, U, 0, tps);
        run<MIX_APPROX_SCALAR>("approx scalar 3F+1M", out, sms, khz, U, 4 * U + 3, 3 * U, tps);
        run<MIX_APPROX_PACKED>("approx packed 3F2+1M3 /2", out, sms, khz, 2 * U, 4 * U + 6, 6 * U, tps);
        run<MIX_APPROX_PACKED_LDS>("approx packed + 2 LDS", out, sms, khz, 2 * U, 4 * U + 2, 6 * U, tps);
        run<MIX_EXACT_SCALAR>("exact scalar 8F+1M", out, sms, khz, U, 9 * U + 3, 8 * U, tps);
        run<MIX_EXACT_PACKED>("exact packed 8F2+1M3 /2", out, sms, khz, 2 * U, 9 * U + 6, 16 * U, tps);
        run<MIX_EXACT_PACKED_LDS>("exact packed + 2 LDS", out, sms, khz, 2 * U, 9 * U + 2, 16 * U, tps);
        run<MIX_EMD_SCALAR>("emd filter scalar 8F+1M", out, sms, khz, U, 9 * U + 3, 8 * U, tps);
        run<MIX_EMD_PACKED>("emd filter packed /2", out, sms, khz, 2 * U, 10 * U + 6, 18 * U, tps);
        run<MIX_EMD_EXPAND_PACKED>("emd expand 4F2+1M3 /2", out, sms, khz, 2 * U, 5 * U + 6, 8 * U, tps);
    }
    return 0;
}
