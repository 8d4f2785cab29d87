// Per-op issue rates of the special-function unit and friends on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rates mufu_rates.cu && ./mufu_rates
#include <cstdio>
#include <cuda_runtime.h>

#define OP_RSQ 0
#define OP_EX2 1
#define OP_RCP 2
#define OP_SQRT 3
#define OP_LG2 4
#define OP_TANH 5
#define OP_SIN 6
#define OP_FFMA 7
#define OP_FMNMX 8
#define OP_F2I 9
#define OP_MIX3 10      // rsq, ex2, rcp round robin
#define OP_MIX_FMA 11   // 3 mufu + 19 ffma per group
#define OP_EX2_NOFTZ 12

template <int OP>
__device__ __forceinline__ float op(float v) {
    float y;
    if (OP == OP_RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_EX2) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_EX2_NOFTZ) asm volatile("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_RCP) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_SQRT) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_LG2) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_TANH) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_SIN) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(v));
    else if (OP == OP_FMNMX) asm volatile("max.f32 %0, %1, 0f3F800000;" : "=f"(y) : "f"(v));
    else if (OP == OP_F2I) { int i; asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(i) : "f"(v)); y = __int_as_float(i); }
    else y = v;
    return y;
}

template <int OP, int NCH>
__global__ void __launch_bounds__(512) rate_kernel(float *out, int iters) {
    float v[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) v[i] = 1.0f + 0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            if (OP == OP_MIX3) {
#pragma unroll
                for (int i = 0; i + 2 < NCH; i += 3) {
                    v[i] = op<OP_RSQ>(v[i]); v[i + 1] = op<OP_EX2>(v[i + 1]); v[i + 2] = op<OP_RCP>(v[i + 2]);
                }
            } else if (OP == OP_MIX_FMA) {
#pragma unroll
                for (int i = 0; i + 2 < NCH; i += 3) {
                    v[i] = op<OP_RSQ>(v[i]); v[i + 1] = op<OP_EX2>(v[i + 1]); v[i + 2] = op<OP_RCP>(v[i + 2]);
#pragma unroll
                    for (int f = 0; f < 19; ++f) v[(i + f) % NCH] = fmaf(v[(i + f) % NCH], 0.999f, 0.001f);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NCH; ++i) v[i] = op<OP>(v[i]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

template <int OP, int NCH>
void run(const char *name, int sms, double clk_ghz, int block, int blocks_per_sm, double ops_per_rep) {
    float *out; cudaMalloc(&out, 16);
    const int iters = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    rate_kernel<OP, NCH><<<sms * blocks_per_sm, block>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        rate_kernel<OP, NCH><<<sms * blocks_per_sm, block>>>(out, iters);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    const double lane_ops = (double)sms * blocks_per_sm * block * iters * 4.0 * ops_per_rep;
    const double per_clk_sm = lane_ops / (best * 1e-3) / (clk_ghz * 1e9) / sms;
    printf("%-28s warps/SM=%2d  %8.3f ms  %7.2f lane-ops/clk/SM\n", name, block * blocks_per_sm / 32, best, per_clk_sm);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount; const double g = clk / 1e6;
    printf("%s, %d SMs, %.3f GHz nominal (rates assume nominal clock)\n", p.name, sms, g);
#define R(OP, name) run<OP, 12>(name, sms, g, 512, 4, 12.0); run<OP, 12>(name, sms, g, 512, 1, 12.0);
    R(OP_RSQ, "MUFU.RSQ") R(OP_EX2, "MUFU.EX2 (ftz)") R(OP_EX2_NOFTZ, "ex2.approx (no ftz)") R(OP_RCP, "MUFU.RCP")
    R(OP_SQRT, "MUFU.SQRT") R(OP_LG2, "MUFU.LG2") R(OP_TANH, "MUFU.TANH") R(OP_SIN, "MUFU.SIN")
    R(OP_FFMA, "FFMA") R(OP_FMNMX, "FMNMX") R(OP_F2I, "F2I") R(OP_MIX3, "mix rsq+ex2+rcp")
    run<OP_MIX_FMA, 12>("mix 3 mufu + 19 ffma", sms, g, 512, 4, 4 * 22.0);
    run<OP_MIX_FMA, 12>("mix 3 mufu + 19 ffma", sms, g, 512, 1, 4 * 22.0);
    return 0;
}
