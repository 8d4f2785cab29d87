// How do MUFU, FFMA / FFMA2 and LDS.128 overlap on one sm_100a SMSP?  Each group = 3 independent
// MUFU + NF FMA-pipe ops (+ NL LDS.128); 16 warps per SM (4 per SMSP), many independent chains.
// Prints SMSP cycles per group (MUFU floor = 24).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float rsq(float v) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float ex2(float v) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float lg2(float v) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }

template <int NF, int PACKED, int NL, int NM>
__global__ void __launch_bounds__(512, 1) mix(float *out, int iters) {
    __shared__ float4 sbuf[1024];
    for (int i = threadIdx.x; i < 1024; i += 512) sbuf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    __syncthreads();
    float m[12];
    float2 f[16];
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = 1.0f + 0.001f * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = make_float2(1.0f + i, 0.5f * i);
    const float2 a2 = make_float2(0.999f, 0.999f), b2 = make_float2(0.001f, 0.001f);
    int idx = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (NM >= 1) m[3 * g] = rsq(m[3 * g]);
            if (NM >= 2) m[3 * g + 1] = ex2(m[3 * g + 1]);
            if (NM >= 3) m[3 * g + 2] = lg2(m[3 * g + 2]);
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                const int j = (g * NF + k) % 16;
                if (PACKED) f[j] = __ffma2_rn(f[j], a2, b2);
                else f[j].x = fmaf(f[j].x, 0.999f, 0.001f);
            }
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                const float4 v = sbuf[(idx + 32 * (g * NL + k)) & 1023];
                acc.x += v.x;  // 1 FADD per LDS keeps it alive
            }
        }
        idx = (idx + 1) & 1023;
    }
    float s = acc.x;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += m[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i].x + f[i].y;
    if (s == 12345.678f) out[0] = s;
}

template <int NF, int PACKED, int NL, int NM>
void run(float *out, int sms, double clk) {
    const int iters = 4000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    mix<NF, PACKED, NL, NM><<<sms, 512>>>(out, iters); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a); mix<NF, PACKED, NL, NM><<<sms, 512>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    const double groups_per_smsp = 4.0 /*warps*/ * iters * 4.0;
    printf("mufu=%d  %s=%2d  lds128=%d : %6.2f SMSP-cycles per group\n", NM, PACKED ? "ffma2" : "ffma ", NF, NL,
           best * 1e-3 * clk * 1e9 / groups_per_smsp);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, 16);
    const int sms = p.multiProcessorCount; const double g = clk / 1e6;
    run<0, 0, 0, 3>(out, sms, g); run<4, 0, 0, 3>(out, sms, g); run<8, 0, 0, 3>(out, sms, g); run<12, 0, 0, 3>(out, sms, g);
    run<16, 0, 0, 3>(out, sms, g); run<20, 0, 0, 3>(out, sms, g); run<24, 0, 0, 3>(out, sms, g); run<28, 0, 0, 3>(out, sms, g);
    run<4, 1, 0, 3>(out, sms, g); run<6, 1, 0, 3>(out, sms, g); run<8, 1, 0, 3>(out, sms, g); run<10, 1, 0, 3>(out, sms, g);
    run<12, 1, 0, 3>(out, sms, g); run<14, 1, 0, 3>(out, sms, g);
    run<10, 1, 1, 3>(out, sms, g); run<10, 1, 2, 3>(out, sms, g); run<0, 0, 1, 3>(out, sms, g); run<0, 0, 2, 3>(out, sms, g);
    run<16, 0, 0, 0>(out, sms, g); run<10, 1, 0, 0>(out, sms, g); run<10, 1, 0, 2>(out, sms, g); run<10, 1, 0, 1>(out, sms, g);
    run<0, 0, 1, 0>(out, sms, g); run<0, 0, 4, 0>(out, sms, g);
    return 0;
}
