// What would a dense-tile formulation of the chromatin pair sweep cost?  (A measurement for DESIGN.md 6b: the shipped
// kernel evaluates 4x4 blocks that are private to a lane and is bound by FP32 issue at ~29 SMSP-cycles per warp-pair.)
//
// Here a warp owns a 16 x 16 tile of bead pairs in the fragment layout of mma.sync.m16n8k16 (thread (g, t) holds rows
// {g, g + 8} x columns {2t, 2t + 1} of each 8-column block: 8 pairs per lane and step) and the tensor cores take over
// everything that is a small dense product, with operands split into bf16 pieces so that the sums keep fp32 accuracy:
//   r^2_ij = |x_i|^2 + |x_j|^2 - 2 x_i.x_j   2 MMAs per 8 columns (hi / mid / lo pieces laid out along K = 32)
//   G_i    = sum_j coef_ij [x_j, 1]          4 MMAs (coef in three bf16 pieces against the pieces of the positions)
//   F_j    = sum_i coef_ij [x_i, 1]          4 MMAs on the TRANSPOSED pieces (movmatrix.trans: an MMA contracts over
//                                            the index spread over lane % 4, i.e. over columns, whichever operand the
//                                            accumulator fragment is reused as)
// What stays on the FMA pipe per pair: d = r^2 * rsqrt, -(1 + C 2^d), m + y, m^2 + m, two multiplies, and the three-way
// split of coef (cvt.rn.bf16x2 + unpack + subtract, twice); 3 MUFU as before.  The harness has the real data flow
// (nothing can be hoisted) but does not form a valid force: it answers "how many SMSP-cycles per pair".
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mmabench.bin mmabench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float rsq(float v) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float ex2(float v) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float rcp(float v) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float hi, float lo) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t p) {  // (value in the upper half, value in the lower half)
    return make_float2(__uint_as_float(p & 0xffff0000u), __uint_as_float(p << 16));
}
__device__ __forceinline__ uint32_t movtrans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// SPLIT = number of bf16 pieces of coef (3: fp32 accuracy, 2: 16 bits); NMUFU = 3 or 2 (rcp knocked out)
template <int SPLIT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) mmabench(const uint32_t *init, float *out, int steps) {
    extern __shared__ __align__(16) unsigned char smem[];
    // per warp: column operands of 64 column tiles in fragment order (what a TMA-fed stream would hold)
    //   [tile][lane]: uint4 r2b (2 MMAs x 2 col blocks... as 2 x uint4), uint4 xb (pieces of the positions), float4 y x 2
    constexpr int TILES = 2, PER = TILES * 32 * 2;  // elements of each array per warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 *colr = reinterpret_cast<uint4 *>(smem) + (size_t)warp * PER;
    uint4 *colx = reinterpret_cast<uint4 *>(smem + (size_t)WARPS * PER * 16) + (size_t)warp * PER;
    float4 *ys = reinterpret_cast<float4 *>(smem + (size_t)WARPS * PER * 32) + (size_t)warp * PER;
    float2 *fs = reinterpret_cast<float2 *>(smem + (size_t)WARPS * PER * 48) + (size_t)warp * PER;
    for (int i = lane; i < TILES * 32 * 2; i += 32) {
        const uint32_t v = init[(i * 7 + warp) & 4095];
        colr[i] = make_uint4(v & 0x3f803f80u | 0x3c003c00u, v >> 3 & 0x3f803f80u | 0x3c003c00u, 0x3c803c80u, 0x3d003d00u);
        colx[i] = make_uint4(0x3f803e80u, 0x3e003f00u, 0x3c003c00u, 0x3b003b00u);
        ys[i] = make_float4(0.1f, 0.2f, 0.05f, 0.3f);
        fs[i] = make_float2(0.f, 0.f);
    }
    __syncwarp();
    // row operands: persistent for the whole sweep
    uint32_t ra[2][4], xr[3][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int i = 0; i < 4; ++i) ra[m][i] = 0x3e803e80u + ((init[(lane * 8 + m * 4 + i) & 4095] & 0x7fu) << 16);
#pragma unroll
    for (int s = 0; s < 3; ++s) xr[s][0] = 0x3f003e00u >> s, xr[s][1] = 0x3e803f80u >> s;
    float G[4] = {0.f, 0.f, 0.f, 0.f};
    const float C = -0.03f;
    int tile = 0;
    for (int st = 0; st < steps; ++st) {
        const uint4 cr0 = colr[(tile * 2 + 0) * 32 + lane], cr1 = colr[(tile * 2 + 1) * 32 + lane];
        const uint4 cx0 = colx[(tile * 2 + 0) * 32 + lane], cx1 = colx[(tile * 2 + 1) * 32 + lane];
        const float4 y0 = ys[(tile * 2 + 0) * 32 + lane], y1 = ys[(tile * 2 + 1) * 32 + lane];
        // ---- r^2 of the 16 x 16 tile: 2 MMAs (K = 32) per 8-column block ---------------------------------
        float r2[2][4];
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
            r2[cb][0] = r2[cb][1] = r2[cb][2] = r2[cb][3] = 1e-6f;
            const uint4 c = cb ? cr1 : cr0;
            mma_bf16(r2[cb], ra[0], c.x, c.y);
            mma_bf16(r2[cb], ra[1], c.z, c.w);
        }
        // ---- the pair function on 8 values, two at a time -------------------------------------------------
        uint32_t piece[3][4];  // [bf16 piece][A-fragment register]: {cb0 rows g, cb0 rows g+8, cb1 rows g, cb1 rows g+8}
#pragma unroll
        for (int cb = 0; cb < 2; ++cb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 rr = make_float2(r2[cb][2 * h], r2[cb][2 * h + 1]);
                const float4 yv = cb ? y1 : y0;
                const float2 y2 = h ? make_float2(yv.z, yv.w) : make_float2(yv.x, yv.y);
                const float2 inv = make_float2(rsq(rr.x), rsq(rr.y));
                const float2 d = __fmul2_rn(rr, inv);
                const float2 e = make_float2(ex2(d.x), ex2(d.y));
                const float2 sn = __ffma2_rn(e, make_float2(C, C), make_float2(-1.f, -1.f));
                const float2 mn = make_float2(rcp(sn.x), rcp(sn.y));
                const float2 rs = __fadd2_rn(mn, y2);
                const float2 wn = __ffma2_rn(mn, mn, mn);
                float2 coef = __fmul2_rn(__fmul2_rn(rs, wn), inv);
                // three-way bf16 split of the two coefficients (columns 2t, 2t+1 of one row = one fragment register)
                const int reg = cb * 2 + h;
                piece[0][reg] = pack_bf16(coef.y, coef.x);
                if (SPLIT >= 2) {
                    const float2 p = unpack_bf16(piece[0][reg]);
                    coef = __fadd2_rn(coef, make_float2(-p.y, -p.x));
                    piece[1][reg] = pack_bf16(coef.y, coef.x);
                }
                if (SPLIT >= 3) {
                    const float2 p = unpack_bf16(piece[1][reg]);
                    coef = __fadd2_rn(coef, make_float2(-p.y, -p.x));
                    piece[2][reg] = pack_bf16(coef.y, coef.x);
                }
            }
        // ---- row sums: coef pieces x [x_j, 1] pieces (K = the 16 columns) -----------------------------------
        {
            const uint32_t a0[4] = {piece[0][0], piece[0][1], piece[0][2], piece[0][3]};
            mma_bf16(G, a0, cx0.x, cx0.y);
            mma_bf16(G, a0, cx1.x, cx1.y);
            if (SPLIT >= 2) {
                const uint32_t a1[4] = {piece[1][0], piece[1][1], piece[1][2], piece[1][3]};
                mma_bf16(G, a1, cx0.x, cx0.y);
            }
            if (SPLIT >= 3) {
                const uint32_t a2[4] = {piece[2][0], piece[2][1], piece[2][2], piece[2][3]};
                mma_bf16(G, a2, cx0.z, cx0.w);
            }
        }
        // ---- column sums: the same pieces transposed (4 8x8 blocks each) x [x_i, 1] pieces (K = the 16 rows) ---
        {
            float F[4];
            const float2 f0 = fs[(tile * 2 + 0) * 32 + lane], f1 = fs[(tile * 2 + 1) * 32 + lane];
            F[0] = f0.x, F[1] = f0.y, F[2] = f1.x, F[3] = f1.y;
#pragma unroll
            for (int s = 0; s < SPLIT; ++s) {
                // A' = coef^T: a0' = T(rows 0-7, cols 0-7), a1' = T(rows 0-7, cols 8-15), a2' = T(rows 8-15, cols 0-7), ..
                const uint32_t at[4] = {movtrans(piece[s][0]), movtrans(piece[s][2]), movtrans(piece[s][1]),
                                        movtrans(piece[s][3])};
                mma_bf16(F, at, xr[0][0], xr[0][1]);
                if (s == 0) mma_bf16(F, at, xr[1][0], xr[1][1]);
            }
            fs[(tile * 2 + 0) * 32 + lane] = make_float2(F[0], F[1]);
            fs[(tile * 2 + 1) * 32 + lane] = make_float2(F[2], F[3]);
        }
        if (++tile >= TILES) tile = 0;
        __syncwarp();
    }
    out[blockIdx.x * WARPS * 32 + threadIdx.x] = G[0] + G[1] + G[2] + G[3] + fs[lane].x;
}

template <int SPLIT, int WARPS>
void run(const char *name, const uint32_t *init, float *out, int sms, double clk) {
    const int steps = 4000;
    const size_t smem = (size_t)WARPS * (2 * 32 * 2) * 56;
    cudaFuncSetAttribute(mmabench<SPLIT, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, mmabench<SPLIT, WARPS>);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mmabench<SPLIT, WARPS><<<sms, WARPS * 32, smem>>>(init, out, steps);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        mmabench<SPLIT, WARPS><<<sms, WARPS * 32, smem>>>(init, out, steps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double warp_pairs_per_smsp = (double)WARPS / 4.0 * steps * 8.0;   // 8 pairs per lane and step
    const double cyc = best * 1e-3 * clk * 1e9 / warp_pairs_per_smsp;
    printf("%-44s warps=%2d regs=%3d  %7.3f ms  %6.2f SMSP-cycles/warp-pair  (%4.1f%% of FP32 peak at 31 flop/pair) %s\n",
           name, WARPS, fa.numRegs, best, cyc, 100.0 * 31.0 / (2.0 * cyc), err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = (uint32_t)i * 2654435761u;
    uint32_t *init; float *out; cudaMalloc(&init, sizeof(h)); cudaMalloc(&out, 148 * 1024 * 4);
    cudaMemcpy(init, h, sizeof(h), cudaMemcpyHostToDevice);
    const int sms = p.multiProcessorCount; const double g = clk / 1e6;
    run<3, 16>("dense 16x16 tiles, coef in 3 bf16 pieces", init, out, sms, g);
    run<2, 16>("dense 16x16 tiles, coef in 2 bf16 pieces", init, out, sms, g);
    run<1, 16>("dense 16x16 tiles, coef in 1 bf16 piece", init, out, sms, g);
    run<3, 12>("dense 16x16 tiles, 3 pieces", init, out, sms, g);
    run<3, 8>("dense 16x16 tiles, 3 pieces", init, out, sms, g);
    return 0;
}
