// The instruction mix of the chromatin pair block WITHOUT its dependencies: per group (= 2 packs = 4 bead
// pairs) NM MUFU, N2 packed FP32 ops with two distinct register operands, N3 FFMA2 with three distinct register
// pairs, NS scalar FFMA with three distinct registers, NL LDS.128 -- all on independent chains whose
// results are consumed one loop iteration later.  16 warps per SM.  Prints SMSP-cycles per bead pair: what a
// perfectly scheduled pair block with this mix could reach (the real block: 12 / 24 / 6 / 12 / 3.25).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix2.bin mix2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float rsq(float v) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float ex2(float v) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float rcp(float v) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }  // (lg2: ptxas cancels rcp(rcp(x)); same pipe, same rate)

template <int NM, int N2, int N3, int NS, int NL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) mix2(float *out, int iters) {
    __shared__ float4 sbuf[1024];
    for (int i = threadIdx.x; i < 1024; i += WARPS * 32) sbuf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    __syncthreads();
    float m[12], s[6];
    float2 f[8], a[4], b[4];
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = 1.0f + 0.001f * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i] = 0.25f * i;
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = make_float2(1.0f + i, 0.5f * i + threadIdx.x);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = make_float2(0.999f + 1e-6f * threadIdx.x, 0.998f), b[i] = make_float2(1e-3f * i, 2e-3f * threadIdx.x);
    int idx = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        // interleaved by construction: the j-th op of every class at position j / N of the group
        constexpr int TOT = 24;
#pragma unroll
        for (int p = 0; p < TOT; ++p) {
#pragma unroll
            for (int k = 0; k < NM; ++k)
                if (k * TOT / NM == p) {
                    const int j = k % 12;
                    m[j] = (j % 3 == 0) ? rsq(m[j]) : (j % 3 == 1) ? ex2(m[j]) : rcp(m[j]);
                }
#pragma unroll
            for (int k = 0; k < N2; ++k)
                if (k * TOT / N2 == p) f[k % 8] = __ffma2_rn(f[k % 8], a[k % 4], f[k % 8]);
#pragma unroll
            for (int k = 0; k < N3; ++k)
                if (k * TOT / N3 == p) f[(k + 3) % 8] = __ffma2_rn(a[k % 4], b[(k + 1) % 4], f[(k + 3) % 8]);
#pragma unroll
            for (int k = 0; k < NS; ++k)
                if (k * TOT / NS == p) s[k % 6] = fmaf(a[k % 4].x, b[(k + 2) % 4].y, s[k % 6]);
#pragma unroll
            for (int k = 0; k < NL; ++k)
                if (k * TOT / NL == p) {
                    const float4 v = sbuf[(idx + 32 * k) & 1023];
                    acc.x += v.x;  // 1 FADD per LDS keeps it alive
                }
        }
        idx = (idx + 1) & 1023;
    }
    float r = acc.x;
#pragma unroll
    for (int i = 0; i < 12; ++i) r += m[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) r += s[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r += f[i].x + f[i].y;
    if (r == 12345.678f) out[0] = r;
}

template <int NM, int N2, int N3, int NS, int NL, int WARPS = 16>
void run(float *out, int sms, double clk) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix2<NM, N2, N3, NS, NL, WARPS><<<sms, WARPS * 32>>>(out, iters); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); mix2<NM, N2, N3, NS, NL, WARPS><<<sms, WARPS * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double pairs_per_smsp = (WARPS / 4.0) * iters * 4.0;
    printf("warps=%2d mufu=%2d packed2=%2d packed3=%d ffma=%2d lds128=%d : %6.2f SMSP-cycles per pair\n", WARPS, NM, N2, N3, NS, NL,
           best * 1e-3 * clk * 1e9 / pairs_per_smsp);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, 16);
    const int sms = p.multiProcessorCount; const double g = clk / 1e6;
    run<12, 24, 6, 12, 3>(out, sms, g);   // the pair block
    run<12, 24, 6, 12, 4>(out, sms, g);
    run<12, 24, 6, 12, 0>(out, sms, g);   // no shared-memory traffic
    run<0, 24, 6, 12, 3>(out, sms, g);    // no MUFU
    run<0, 24, 6, 12, 0>(out, sms, g);    // FMA pipe alone
    run<12, 0, 0, 0, 0>(out, sms, g);     // MUFU alone
    run<12, 0, 0, 0, 3>(out, sms, g);     // MUFU + LDS
    run<8, 24, 6, 12, 3>(out, sms, g);    // 2 MUFU per pair
    run<10, 24, 6, 12, 3>(out, sms, g);   // 2.5 MUFU per pair
    run<12, 24, 0, 24, 3>(out, sms, g);   // column sums scalar as well
    run<12, 18, 6, 12, 3>(out, sms, g);   // 3 packed ops per pack fewer
    run<12, 24, 6, 12, 3, 12>(out, sms, g);
    run<12, 24, 6, 12, 3, 8>(out, sms, g);
    return 0;
}
