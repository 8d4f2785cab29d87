// Does FFMA2 with a uniform-register broadcast operand (data row fetched with LDCU from constant memory,
// every lane of the warp on the same row) avoid the 3-cycle cost of "pair * scalar-broadcast + pair" with
// the scalar in a vector register (3.04 cycles, ffma2_operands.cu)?  Polynomial row body, K = 4, two chains
// packed per thread, one thread per chain pair.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o poly_ur.bin poly_ur.cu
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float4 crow[4096];

template <int JP>
__global__ void __launch_bounds__(896, 1) k_ur(float2 *o, int n, int evals) {
    float2 c[JP][4], a[JP][4];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int j = 0; j < JP; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) c[j][k] = o[(t * JP + j) * 4 + k];
    for (int e = 0; e < evals; ++e) {
#pragma unroll
        for (int j = 0; j < JP; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) a[j][k] = make_float2(0.f, 0.f);
#pragma unroll 8
        for (int i = 0; i < n; ++i) {
            const float4 r = crow[i];
            const float2 x = make_float2(r.x, r.x);
#pragma unroll
            for (int j = 0; j < JP; ++j) {
                float2 tt = __ffma2_rn(c[j][3], x, c[j][2]);
                tt = __ffma2_rn(tt, x, c[j][1]);
                tt = __ffma2_rn(tt, x, c[j][0]);
                const float2 res = __fadd2_rn(tt, make_float2(-r.w, -r.w));
                a[j][0] = __fadd2_rn(a[j][0], res);
                a[j][1] = __ffma2_rn(res, x, a[j][1]);
                a[j][2] = __ffma2_rn(res, make_float2(r.y, r.y), a[j][2]);
                a[j][3] = __ffma2_rn(res, make_float2(r.z, r.z), a[j][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < JP; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) c[j][k] = __ffma2_rn(a[j][k], make_float2(-1e-6f, -1e-6f), c[j][k]);
    }
#pragma unroll
    for (int j = 0; j < JP; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) o[(t * JP + j) * 4 + k] = c[j][k];
}

template <int JP>
void run(int warps_per_cta, int ctas, const char *name) {
    const int n = 1000, evals = 21;
    float2 *o;
    const size_t cnt = (size_t)ctas * warps_per_cta * 32 * JP * 4;
    cudaMalloc(&o, cnt * sizeof(float2));
    cudaMemset(o, 0, cnt * sizeof(float2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_ur<JP><<<ctas, warps_per_cta * 32>>>(o, n, evals);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k_ur<JP><<<ctas, warps_per_cta * 32>>>(o, n, evals);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double chain_rows = (double)ctas * warps_per_cta * 32 * JP * 2 * n * evals;
    const double flops = chain_rows * 14.0;
    // SMSP cycles per warp chain-row on the busiest SMSP
    const double warps_smsp = (warps_per_cta + 3) / 4;
    const double cyc = best * 1e-3 * clk * 1e3 / (warps_smsp * JP * 2.0 * n * evals);
    printf("%-40s ctas=%d warps/cta=%d  %.4f ms  %.2f TFLOP/s  %.2f SMSP-cycles per warp chain-row  (%s)\n", name, ctas,
           warps_per_cta, best, flops / best / 1e9, cyc, cudaGetErrorString(cudaGetLastError()));
    cudaFree(o);
}

int main() {
    static float4 h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = make_float4(-2.f + 0.004f * (i % 1000), 0.5f, 0.25f, 1.0f);
    cudaMemcpyToSymbol(crow, h, sizeof(h));
    run<1>(4, 148, "J=2, 1 warp per SMSP");
    run<1>(8, 148, "J=2, 2 warps per SMSP");
    run<1>(28, 148, "J=2, 7 warps per SMSP");
    run<2>(4, 148, "J=4, 1 warp per SMSP");
    run<2>(8, 148, "J=4, 2 warps per SMSP");
    run<2>(16, 148, "J=4, 4 warps per SMSP");
    return 0;
}
