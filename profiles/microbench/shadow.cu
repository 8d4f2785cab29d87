// Which instructions issue "in the shadow" of a packed FP32 op?  An FFMA2 occupies the FMA pipe for 2 cycles
// (3 with three distinct register pairs).  Per group: NF FFMA2 (two distinct register pairs each) plus NX
// extra instructions of one kind on independent chains; 16 warps per SM.  If the extra instruction can be
// issued while the FMA pipe is still busy with the second half of an FFMA2, the time per group stays at
// NF * 2.04 until the issue slots (one per cycle) run out; if not, it grows by the instruction's own cost.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -diag-suppress 39 -o shadow.bin shadow.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { X_NONE, X_LOP, X_IADD, X_FFMA, X_MUFU, X_LDS32, X_LDS128, X_FMNMX, X_FFMA_IMM, X_FFMA2_3 };
static const char *xname[] = {"none", "LOP3 (ALU pipe)", "IADD3 (ALU pipe)", "FFMA 3 regs", "MUFU.EX2", "LDS.32", "LDS.128", "FMNMX (ALU pipe)", "FFMA reg*imm+imm", "FFMA2 3 reg pairs"};

template <int NF, int KIND, int NX>
__global__ void __launch_bounds__(512, 1) shadow(float *out, int iters, float fa, float fb, int ia) {
    __shared__ float4 sbuf[1024];
    for (int i = threadIdx.x; i < 1024; i += 512) sbuf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    __syncthreads();
    float2 f[10], a2 = make_float2(fa, fa * 0.5f), b2 = make_float2(fb, fb + 1.f);
    float x[12];
    unsigned u[12];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 10; ++i) f[i] = make_float2(1.0f + i, 0.5f * i + threadIdx.x);
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = 1.0f + 0.01f * (i + threadIdx.x), u[i] = threadIdx.x * 7u + i;
    int idx = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        constexpr int TOT = 60;
#pragma unroll
        for (int p = 0; p < TOT; ++p) {
#pragma unroll
            for (int k = 0; k < NF; ++k)
                if (k * TOT / (NF > 0 ? NF : 1) == p) f[k % 10] = __ffma2_rn(f[k % 10], a2, f[k % 10]);
#pragma unroll
            for (int k = 0; k < NX; ++k)
                if (k * TOT / (NX > 0 ? NX : 1) == p) {
                    const int j = k % 12;
                    if (KIND == X_LOP) u[j] = (u[j] ^ (unsigned)ia) & (u[(j + 1) % 12] | 0x55u);
                    if (KIND == X_IADD) u[j] = u[j] + u[(j + 5) % 12] + (unsigned)ia;
                    if (KIND == X_FFMA) x[j] = fmaf(x[j], fa, x[(j + 1) % 12]);
                    if (KIND == X_FFMA_IMM) x[j] = fmaf(x[j], 0.999f, 0.001f);
                    if (KIND == X_MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
                    if (KIND == X_FMNMX) x[j] = fminf(x[j], x[(j + 1) % 12] + 0.f);
                    if (KIND == X_LDS32) acc.x += ((const float *)sbuf)[(idx + 32 * k) & 4095];
                    if (KIND == X_LDS128) acc.x += sbuf[(idx + 32 * k) & 1023].x;
                    if (KIND == X_FFMA2_3) f[j % 10] = __ffma2_rn(a2, b2, f[j % 10]);
                }
        }
        idx = (idx + 1) & 1023;
    }
    float r = acc.x;
#pragma unroll
    for (int i = 0; i < 10; ++i) r += f[i].x + f[i].y;
#pragma unroll
    for (int i = 0; i < 12; ++i) r += x[i] + (float)u[i];
    if (r == 12345.678f) out[0] = r;
}

template <int NF, int KIND, int NX>
void run(float *out, int sms, double clk) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    shadow<NF, KIND, NX><<<sms, 512>>>(out, iters, 0.999f, 0.001f, 3); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); shadow<NF, KIND, NX><<<sms, 512>>>(out, iters, 0.999f, 0.001f, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double groups_per_smsp = 4.0 * iters;
    printf("ffma2=%2d + %2d x %-20s : %6.2f SMSP-cycles per group\n", NF, NX, xname[KIND], best * 1e-3 * clk * 1e9 / groups_per_smsp);
}

template <int KIND>
void sweep(float *out, int sms, double g) {
    run<10, KIND, 2>(out, sms, g); run<10, KIND, 5>(out, sms, g); run<10, KIND, 10>(out, sms, g); run<10, KIND, 20>(out, sms, g);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, 16);
    const int sms = p.multiProcessorCount; const double g = clk / 1e6;
    run<10, X_NONE, 0>(out, sms, g);
    sweep<X_LOP>(out, sms, g); sweep<X_IADD>(out, sms, g); sweep<X_FMNMX>(out, sms, g); sweep<X_FFMA>(out, sms, g);
    sweep<X_FFMA_IMM>(out, sms, g); sweep<X_MUFU>(out, sms, g); sweep<X_LDS32>(out, sms, g); sweep<X_LDS128>(out, sms, g);
    sweep<X_FFMA2_3>(out, sms, g);
    run<0, X_LOP, 20>(out, sms, g); run<0, X_FFMA, 20>(out, sms, g); run<0, X_FMNMX, 20>(out, sms, g); run<0, X_LDS32, 20>(out, sms, g);
    return 0;
}
