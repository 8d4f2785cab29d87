// Cost of FFMA2 / FFMA / FADD2 as a function of how many distinct register operands they read.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *out, int iters) {
    float2 f[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) f[i] = make_float2(1.0f + 0.001f * (threadIdx.x + i), 0.5f + 0.001f * i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            const int a = (j + 7) % 24, b = (j + 13) % 24;
            if (MODE == 0) f[j] = __ffma2_rn(f[j], make_float2(0.999f, 0.999f), make_float2(0.001f, 0.001f));
            if (MODE == 1) f[j] = __ffma2_rn(f[a], f[b], f[j]);                     // 3 distinct pairs
            if (MODE == 2) f[j] = __ffma2_rn(f[a], f[a], f[j]);                     // 2 distinct pairs
            if (MODE == 3) f[j] = __ffma2_rn(f[a], make_float2(f[b].x, f[b].x), f[j]);  // pair, scalar bcast, pair
            if (MODE == 4) f[j].x = fmaf(f[a].x, f[b].y, f[j].x);                   // scalar FFMA, 3 regs
            if (MODE == 5) f[j] = __fadd2_rn(f[a], f[j]);                           // FADD2 2 pairs
            if (MODE == 6) f[j] = __fmul2_rn(f[a], f[b]);                           // FMUL2 2 distinct -> new dst
            if (MODE == 7) { f[j] = __ffma2_rn(f[a], f[b], f[j]); ++j; if (j < 24) f[j] = __ffma2_rn(f[a], f[b], f[j]); }  // shared a,b back to back
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 24; ++i) s += f[i].x + f[i].y;
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char *name, float *out, int sms, double clk) {
    const int iters = 4000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<sms, 512>>>(out, iters); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a); k<MODE><<<sms, 512>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("%-44s %5.2f SMSP-cycles per instruction\n", name, best * 1e-3 * clk * 1e9 / (4.0 * iters * 24.0));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, 16);
    const int sms = p.multiProcessorCount; const double g = clk / 1e6;
    run<0>("FFMA2 d=d*imm+imm", out, sms, g);
    run<1>("FFMA2 3 distinct register pairs", out, sms, g);
    run<2>("FFMA2 a*a+d (2 distinct pairs)", out, sms, g);
    run<3>("FFMA2 pair*scalar-broadcast+pair", out, sms, g);
    run<4>("FFMA  3 distinct registers", out, sms, g);
    run<5>("FADD2 2 distinct pairs", out, sms, g);
    run<6>("FMUL2 2 distinct pairs", out, sms, g);
    run<7>("FFMA2 pairs sharing a,b back to back", out, sms, g);
    return 0;
}
