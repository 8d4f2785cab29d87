// A tcgen05 formulation of the chromatin pair sweep, measured on real forces.
//
// The shipped kernel evaluates 4x4 blocks of bead pairs that are private to a lane: 18 FMA-pipe ops + 3 MUFU per pair,
// FP32-issue bound at ~29 SMSP-cycles per warp-pair (33 in the kernel).  Here everything that is a small dense product
// moves to the 5th-generation tensor cores, with operands split into TF32 pieces so that the sums keep fp32 accuracy:
//
//   r^2_ij = |p_i|^2 + |p_j|^2 - 2 p_i.p_j          D1[128 x 64]  = A1[128 x 24] . B1[24 x 64]        (3 k-steps)
//   G_i    = sum_j coef_ij [p_j, 1]                 D2[128 x 16] += Coef[128 x 64] . Uc[64 x 16]       (2 x 8 k-steps)
//   F_j    = sum_i coef_ij [p_i, 1]                 D3[ 64 x  8]  = Coef^T[64 x 128] . Ur[128 x 8]     (2 x 16 k-steps)
//
// p = (scaled position) - (centre of the row tile, a TF32 number): the Gram form of r^2 cancels, so the operands are
// kept small; p = ph + pm + pl (TF32 pieces, exact), all products that matter are formed (24 = 18 + 6 norm pieces).
// coef = ch + cl (two TF32 pieces).  The SAME shared-memory coefficient tile is the K-major A operand of the row sums
// and the MN-major A operand (M = 64) of the column sums -- the two canonical no-swizzle layouts coincide -- and the
// feature chunks of the column beads double as the MN-major B operand of the row sums.
//
// What stays on the SIMT pipes per pair: tcgen05.ld of r^2 (1/16), the contact y (1/8 LDS.128: registers are reused for
// the second chain), rsqrt, d = r^2 rsqrt, ex2, -(1 + C 2^d), rcp, m + y, m^2 + m, two multiplies, the two-piece split
// (LOP + FADD) and 1/2 STS.128: 8 FMA-pipe ops + 1 ALU op + 3 MUFU.  With all three MUFU that is bound by the
// special-function unit (24 SMSP-cycles per warp-pair); NPOLY of every 16 pairs take 2^d from an FMA-pipe polynomial.
//
// A CTA (one per SM) works on TWO chains at a time: step (t, A), step (t, B), step (t + 1, A) ... over the tiles
// t = (row tile I of 128 beads, column sub-tile J >= 2 I of 64 beads); diagonal sub-tiles mask j <= i and skip dead
// 16-column chunks.  While the 8 SIMT warps work on (t, A) the tensor core forms r^2 of (t + 1, .) and the sums of (t, B).
// Warp 8 issues the MMAs and the bulk copy of the next contact tile, warp 9 builds the operand features.
//
// The harness computes the likelihood force sums  F_b = sum_{j != b} coef_bj (u_j - u_b)  of every chain, checks two
// chains against a float64 host loop and reports SMSP-cycles per warp-pair (processed and useful pairs).
//
// STATUS (profiles/r2_microbench_tcpair.txt, r2_microbench_tcprobe.txt): the r^2 leg is verified on the B200 (max abs
// error 5.4e-4 = 1.1e-4 relative: the tensor core's fp32 accumulation of terms of size |p|^2 ~ 1000).  The row / column
// sum legs are issued with MN-major NO-SWIZZLE operands, which kind::tf32 accepts but answers with zeros (tcprobe.cu: a
// TF32 MN-major operand only works in the 128B-swizzle / 32B-base layout, and that layout type is rejected for K-major
// operands, so one image of the coefficient tile cannot serve both sums) -- the force check of this file therefore FAILS
// and only its timing means something, as a lower bound: 63.9 SMSP-cycles per tile slot against 33.1 of the shipped
// kernel; 44.5 with the sums and the coefficient stores knocked out (the 8-op + 3-MUFU SIMT residue with two warps per
// SMSP), +20 for the 48 small-N MMAs of a step (~28 cycles each: the 4 KB A operand crosses the 128 B/clk shared-memory
// port).  Shared-memory traffic alone (8 B stored + 8..16 B read per pair) bounds the design at 20..28 cycles.  Not pursued.
// -DKO_SUMS / -DKO_STS build the knock-outs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcpair.bin tcpair.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int TM = 128, TN = 64;
constexpr int NCH = 6;                    // 16-byte feature chunks per bead (K = 24)
constexpr int GRP = NCH * 128;            // bytes per group of 8 beads in a feature buffer
constexpr int COEF_PIECE = TM * TN * 4;   // 32 KiB
// shared memory map
constexpr int OFF_COEF = 0;                            // [chain 2][piece 2][32 KiB]
constexpr int OFF_Y = OFF_COEF + 4 * COEF_PIECE;       // 32 KiB, float4 index xor (row & 7)
constexpr int OFF_A1 = OFF_Y + TM * TN * 4;            // [chain 2][16 groups][768]
constexpr int OFF_B1 = OFF_A1 + 2 * (TM / 8) * GRP;    // [chain 2][parity 2][8 groups][768]
constexpr int OFF_UR = OFF_B1 + 4 * (TN / 8) * GRP;    // [chain 2][parity 2][16 groups][256]
constexpr int OFF_BAR = OFF_UR + 4 * (TM / 8) * 256;
constexpr int SMEM_BYTES = OFF_BAR + 128;
// tensor memory columns
constexpr int COL_D1 = 0, COL_G = 128, COL_F = 160, TMEM_COLS = 256;
#ifndef NSW
#define NSW 8                              // SIMT warps: 8 (32 columns of a sub-tile each) or 16 (16 columns each)
#endif
constexpr int CW = 64 / (NSW / 4), NCHUNK = CW / 16;
constexpr int NTHREADS = (NSW + 2) * 32;

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float rsq(float v) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float ex2(float v) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float rcp(float v) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
// 2^t for t <= 126 on the FMA pipe (the shipped kernel's poly_ex2: Cody-Waite split, degree-5 polynomial)
__device__ __forceinline__ float poly_ex2(float t) {
    t = fminf(t, 126.0f);
    const float magic = 12582912.0f;
    const float tm = t + magic;
    const float f = t - (tm - magic);
    float p = 1.3390863366e-3f;
    p = fmaf(p, f, 9.6760319183e-3f);
    p = fmaf(p, f, 5.5503571142e-2f);
    p = fmaf(p, f, 2.4022107485e-1f);
    p = fmaf(p, f, 6.9314718803e-1f);
    p = fmaf(p, f, 1.0000000755f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(tm) << 23));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    int spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1 << 22)) __trap();  // a bounded wait: a protocol error must not hang the box
    } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void step_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts4(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::tf32 (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Params {
    const float4 *pos;     // [chain][NP] scaled positions (x, y, z, 0)
    const float4 *ctr;     // [chain][NT] centre of each row tile (TF32 numbers)
    const float *ytiles;   // [tile][128][64], float4 index xor (row & 7)
    float *force;          // [chain][NP][4] (zeroed by the host; red.add)
    float *dbg;            // [2 tiles][128][64] r^2 of chain 0 (or null)
    int NP, NT, n_tiles, pairs_per_cta;
    float Bc;              // -2^B
};

__device__ __forceinline__ void tile_next(int &I, int &J, int NT) {
    if (++J == 2 * NT) { ++I; J = 2 * I; }
}

// One bead's feature chunks.  ROW: [-2ph,1] [-2pm,nh] [-2ph,1] [-2pl,nm] [-2ph,1] [-2pm,nl];  column: [ph,nh] [ph,1] [pm,nm] [ph,1] [pl,nl] [pm,1]
template <bool ROW>
__device__ __forceinline__ void bead_features(const float4 u, const float4 c, uint32_t dst, uint32_t ur_dst) {
    const float px = u.x - c.x, py = u.y - c.y, pz = u.z - c.z;
    const float hx = tf32_trunc(px), hy = tf32_trunc(py), hz = tf32_trunc(pz);
    const float rx = px - hx, ry = py - hy, rz = pz - hz;
    const float mx = tf32_trunc(rx), my = tf32_trunc(ry), mz = tf32_trunc(rz);
    const float lx = rx - mx, ly = ry - my, lz = rz - mz;
    const float n = fmaf(pz, pz, fmaf(py, py, px * px));
    const float nh = tf32_trunc(n), nr = n - nh, nm = tf32_trunc(nr), nl = nr - nm;
    if (ROW) {
        sts4(dst, -2.f * hx, -2.f * hy, -2.f * hz, 1.f);
        sts4(dst + 128, -2.f * mx, -2.f * my, -2.f * mz, nh);
        sts4(dst + 256, -2.f * hx, -2.f * hy, -2.f * hz, 1.f);
        sts4(dst + 384, -2.f * lx, -2.f * ly, -2.f * lz, nm);
        sts4(dst + 512, -2.f * hx, -2.f * hy, -2.f * hz, 1.f);
        sts4(dst + 640, -2.f * mx, -2.f * my, -2.f * mz, nl);
        sts4(ur_dst, hx, hy, hz, 1.f);
        sts4(ur_dst + 128, mx, my, mz, 0.f);
    } else {
        sts4(dst, hx, hy, hz, nh);
        sts4(dst + 128, hx, hy, hz, 1.f);
        sts4(dst + 256, mx, my, mz, nm);
        sts4(dst + 384, hx, hy, hz, 1.f);
        sts4(dst + 512, lx, ly, lz, nl);
        sts4(dst + 640, mx, my, mz, 1.f);
    }
}

// The SIMT part of a step: the warp's 32 rows x 32 columns of the sub-tile.
template <int NPOLY, bool LOADY, bool MASKED>
__device__ __forceinline__ void simt_step(uint32_t tmem, uint32_t sm, int chain, int q, int h, int lane, int I, int J,
                                          float Bc, float (&y)[CW], float *dbg) {
    const int il = 32 * q + lane;                      // row within the tile
    float r2[NCHUNK][16];
    bool dead[NCHUNK];
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        dead[c] = MASKED && (J * TN + CW * h + 16 * c + 15 <= I * TM + 32 * q);
        if (!dead[c]) tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + COL_D1 + chain * 64 + CW * h + 16 * c, r2[c]);
    }
    if (LOADY) {
        const uint32_t yrow = sm + OFF_Y + il * 256;
#pragma unroll
        for (int f = 0; f < CW / 4; ++f) {
            const float4 v = lds4(yrow + ((((CW / 4) * h + f) ^ (il & 7)) << 4));
            y[4 * f] = v.x, y[4 * f + 1] = v.y, y[4 * f + 2] = v.z, y[4 * f + 3] = v.w;
        }
    }
    tmem_wait_ld();
    if (dbg != nullptr && chain == 0 && I == 0 && J < 2) {
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (!dead[c]) dbg[(J * TM + il) * TN + CW * h + 16 * c + k] = r2[c][k];
    }
    const uint32_t crow = sm + OFF_COEF + chain * 2 * COEF_PIECE + (il & 7) * 16 + (il >> 3) * 2048;
    const int gi = I * TM + il;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        float ch[16], cl[16];
        if (dead[c]) {
#pragma unroll
            for (int k = 0; k < 16; ++k) ch[k] = 0.f, cl[k] = 0.f;
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float rr = r2[c][k];
                if (MASKED) rr = fmaxf(rr, 1e-6f);
                const float inv = rsq(rr);
                const float d = rr * inv;
                const float e = (k < NPOLY) ? poly_ex2(d) : ex2(d);
                const float sn = fmaf(e, Bc, -1.0f);
                const float mn = rcp(sn);
                const float rs = mn + y[16 * c + k];
                const float wn = fmaf(mn, mn, mn);
                float cf = rs * wn * inv;
                if (MASKED) cf = (J * TN + CW * h + 16 * c + k > gi) ? cf : 0.f;
                ch[k] = tf32_trunc(cf);
                cl[k] = cf - ch[k];
#ifdef DEBUG_PRINT
                if (blockIdx.x == 0 && I == 0 && J == 0 && chain == 0 && il == 1 && k < 4 && dbg != nullptr)
                    printf("coef i %d j %d: r2 %g inv %g d %g e %g sn %g mn %g y %g rs %g wn %g cf %g ch %g cl %g\n", gi, J * TN + CW * h + 16 * c + k, rr, inv, d, e, sn, mn, y[16 * c + k], rs, wn, cf, ch[k], cl[k]);
#endif
            }
        }
#ifdef KO_STS
        float acc = 0.f;   // keep the arithmetic alive without the stores
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += ch[k] + cl[k];
        if (acc == 12345.678f) sts4(crow, acc, acc, acc, acc);
#else
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
            const uint32_t a = crow + ((CW / 4) * h + 4 * c + t4) * 128;
            sts4(a, ch[4 * t4], ch[4 * t4 + 1], ch[4 * t4 + 2], ch[4 * t4 + 3]);
            sts4(a + COEF_PIECE, cl[4 * t4], cl[4 * t4 + 1], cl[4 * t4 + 2], cl[4 * t4 + 3]);
        }
#endif
    }
}

// column sums of the step that has just completed on the tensor core: D3 (M = 64: row j sits in lane (j % 16) + 32 (j / 16))
__device__ __forceinline__ void flush_F(const Params &P, uint32_t tmem, int chain_l, int chain_g, int q, int lane, int I, int J) {
    float v[8];
    tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + COL_F + 8 * chain_l, v);
    tmem_wait_ld();
#ifdef DEBUG_PRINT
    if (blockIdx.x == 0 && chain_g == 0 && I == 0 && J < 2 && lane < 18 && (lane & 7) == 1)
        printf("F tile (%d,%d) q %d lane %d: %g %g %g %g | %g %g %g %g\n", I, J, q, lane, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
#endif
    if (lane < 16) {
        const int j = J * TN + 16 * q + lane;
        const float4 u = P.pos[(size_t)chain_g * P.NP + j], c = P.ctr[(size_t)chain_g * P.NT + I];
        const float S = v[3];
        float *f = P.force + ((size_t)chain_g * P.NP + j) * 4;
        atomicAdd(f + 0, (v[0] + v[4]) - (u.x - c.x) * S);
        atomicAdd(f + 1, (v[1] + v[5]) - (u.y - c.y) * S);
        atomicAdd(f + 2, (v[2] + v[6]) - (u.z - c.z) * S);
    }
}
// row sums of a finished row tile: D2 columns [Th, S | Tm, . | Th, S | Tl, .]
__device__ __forceinline__ void flush_G(const Params &P, uint32_t tmem, int chain_l, int chain_g, int q, int lane, int I) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + COL_G + 16 * chain_l, v);
    tmem_wait_ld();
#ifdef DEBUG_PRINT
    if (blockIdx.x == 0 && chain_g == 0 && I == 0 && (lane & 15) == 1)
        printf("G row tile %d q %d lane %d: %g %g %g %g | %g %g %g %g | %g %g %g %g | %g %g %g %g\n", I, q, lane, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
#endif
    const int i = I * TM + 32 * q + lane;
    const float4 u = P.pos[(size_t)chain_g * P.NP + i], c = P.ctr[(size_t)chain_g * P.NT + I];
    const float S = v[3];
    float *f = P.force + ((size_t)chain_g * P.NP + i) * 4;
    atomicAdd(f + 0, (v[0] + v[4] + v[12]) - (u.x - c.x) * S);
    atomicAdd(f + 1, (v[1] + v[5] + v[13]) - (u.y - c.y) * S);
    atomicAdd(f + 2, (v[2] + v[6] + v[14]) - (u.z - c.z) * S);
}

template <int NPOLY>
__global__ void __launch_bounds__(NTHREADS, 1) tcpair(const Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint32_t tmem_slot;
    const uint32_t sm = s_u32(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_y = sm + OFF_BAR, bar_r2 = sm + OFF_BAR + 8, bar_cd = sm + OFF_BAR + 24;  // r2[2], coef_done[2]
    if (threadIdx.x == 0) {
        mbar_init(bar_y, 1);
        mbar_init(bar_r2, 1), mbar_init(bar_r2 + 8, 1);
        mbar_init(bar_cd, 1), mbar_init(bar_cd + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NSW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int NT = P.NT, n_tiles = P.n_tiles;
    constexpr uint32_t ID_R2 = make_idesc(128, 64, 0, 0), ID_G = make_idesc(128, 16, 0, 1), ID_F = make_idesc(64, 8, 1, 1);

    for (int pr = 0; pr < P.pairs_per_cta; ++pr) {
        const int chain0 = (blockIdx.x * P.pairs_per_cta + pr) * 2;   // chains chain0, chain0 + 1
        const uint32_t base_par = (uint32_t)(pr * n_tiles);           // completed phases of every per-tile barrier before this pair
        if (warp < NSW) {
            // ------------------------------------------------------------------ SIMT warps
            const int q = warp & 3, h = warp >> 2;
            float y[CW];
            int I = 0, J = 0, pI = 0, pJ = 0;
            step_barrier();   // the features of tile 0 are in place
            for (int t = 0; t < n_tiles; ++t) {
                const bool diag = J < 2 * I + 2;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c == 0) mbar_wait(bar_y, (base_par + t) & 1);
                    mbar_wait(bar_r2 + 8 * c, (base_par + t) & 1);
                    if (t > 0) {
                        mbar_wait(bar_cd + 8 * c, (base_par + t - 1) & 1);
                        tc_fence_after();
                        if (h == 0) flush_F(P, tmem, c, chain0 + c, q, lane, pI, pJ);
                        else if (h == 1 && pI != I) flush_G(P, tmem, c, chain0 + c, q, lane, pI);
                    } else {
                        tc_fence_after();
                    }
                    float *dbg = (pr == 0 && blockIdx.x == 0) ? P.dbg : nullptr;
                    if (c == 0) {
                        if (diag) simt_step<NPOLY, true, true>(tmem, sm, 0, q, h, lane, I, J, P.Bc, y, dbg);
                        else simt_step<NPOLY, true, false>(tmem, sm, 0, q, h, lane, I, J, P.Bc, y, dbg);
                    } else {
                        if (diag) simt_step<NPOLY, false, true>(tmem, sm, 1, q, h, lane, I, J, P.Bc, y, nullptr);
                        else simt_step<NPOLY, false, false>(tmem, sm, 1, q, h, lane, I, J, P.Bc, y, nullptr);
                    }
                    proxy_fence();
                    tc_fence_before();
                    step_barrier();
                }
                pI = I, pJ = J;
                tile_next(I, J, NT);
            }
            // the sums of the last tile
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                mbar_wait(bar_cd + 8 * c, (base_par + n_tiles - 1) & 1);
                tc_fence_after();
                if (h == 0) flush_F(P, tmem, c, chain0 + c, q, lane, pI, pJ);
                else if (h == 1) flush_G(P, tmem, c, chain0 + c, q, lane, pI);
            }
            tc_fence_before();
            step_barrier();   // end of the chain pair: the tensor memory accumulators are free again
        } else if (warp == NSW) {
            // ------------------------------------------------------------------ MMA issue + contact stream
            const size_t ytile_bytes = (size_t)TM * TN * 4;
            int I = 0, J = 0;
            // prologue: features of tile 0 are ready after the first barrier
            step_barrier();
            tc_fence_after();
            if (lane == 0) {
                mbar_expect_tx(bar_y, (uint32_t)ytile_bytes);
                bulk_g2s(sm + OFF_Y, P.ytiles, (uint32_t)ytile_bytes, bar_y);
                for (int c = 0; c < 2; ++c) {
                    const uint32_t a1 = sm + OFF_A1 + c * (TM / 8) * GRP, b1 = sm + OFF_B1 + (c * 2 + 0) * (TN / 8) * GRP;
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk)
                        umma_tf32(tmem + COL_D1 + 64 * c, make_desc(a1 + kk * 256, 128, GRP), make_desc(b1 + kk * 256, 128, GRP), ID_R2, kk > 0);
                    umma_commit(bar_r2 + 8 * c);
                }
            }
            __syncwarp();
            for (int t = 0; t < n_tiles; ++t) {
                int nI = I, nJ = J;
                tile_next(nI, nJ, NT);
                for (int c = 0; c < 2; ++c) {
                    step_barrier();
                    tc_fence_after();
                    if (lane == 0) {
                        if (c == 0 && t + 1 < n_tiles) {   // every warp has its contacts of tile t in registers
                            mbar_expect_tx(bar_y, (uint32_t)ytile_bytes);
                            bulk_g2s(sm + OFF_Y, reinterpret_cast<const unsigned char *>(P.ytiles) + (size_t)(t + 1) * ytile_bytes, (uint32_t)ytile_bytes, bar_y);
                        }
                        const uint32_t coef = sm + OFF_COEF + c * 2 * COEF_PIECE;
                        const uint32_t b1 = sm + OFF_B1 + (c * 2 + (t & 1)) * (TN / 8) * GRP;
                        const uint32_t ur = sm + OFF_UR + (c * 2 + (I & 1)) * (TM / 8) * 256;
#ifndef KO_SUMS
                        // row sums: A = coef (K-major, K = j), B = column features chunks 1..4 (MN-major)
#pragma unroll
                        for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                umma_tf32(tmem + COL_G + 16 * c, make_desc(coef + pc * COEF_PIECE + kk * 256, 128, 2048),
                                          make_desc(b1 + 128 + kk * GRP, GRP, 128), ID_G, !(J == 2 * I && pc == 0 && kk == 0));
                        // column sums: A = coef^T (MN-major, M = j, K = i), B = row features (MN-major)
#pragma unroll
                        for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                            for (int kk = 0; kk < 16; ++kk)
                                umma_tf32(tmem + COL_F + 8 * c, make_desc(coef + pc * COEF_PIECE + kk * 2048, 2048, 128),
                                          make_desc(ur + kk * 256, 256, 128), ID_F, !(pc == 0 && kk == 0));
#endif
                        umma_commit(bar_cd + 8 * c);
                        if (t + 1 < n_tiles) {
                            const uint32_t a1 = sm + OFF_A1 + c * (TM / 8) * GRP, nb1 = sm + OFF_B1 + (c * 2 + ((t + 1) & 1)) * (TN / 8) * GRP;
#pragma unroll
                            for (int kk = 0; kk < 3; ++kk)
                                umma_tf32(tmem + COL_D1 + 64 * c, make_desc(a1 + kk * 256, 128, GRP), make_desc(nb1 + kk * 256, 128, GRP), ID_R2, kk > 0);
                            umma_commit(bar_r2 + 8 * c);
                        }
                    }
                    __syncwarp();
                }
                I = nI, J = nJ;
            }
            step_barrier();
        } else {
            // ------------------------------------------------------------------ feature producer
            int I = 0, J = 0;
            for (int c = 0; c < 2; ++c) {
                const float4 *pos = P.pos + (size_t)(chain0 + c) * P.NP;
                const float4 ctr = P.ctr[(size_t)(chain0 + c) * NT];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int b = lane + 32 * r;
                    bead_features<true>(pos[b], ctr, sm + OFF_A1 + c * (TM / 8) * GRP + (b & 7) * 16 + (b >> 3) * GRP,
                                        sm + OFF_UR + (c * 2 + 0) * (TM / 8) * 256 + (b & 7) * 16 + (b >> 3) * 256);
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int b = lane + 32 * r;
                    bead_features<false>(pos[b], ctr, sm + OFF_B1 + (c * 2 + 0) * (TN / 8) * GRP + (b & 7) * 16 + (b >> 3) * GRP, 0);
                }
            }
            proxy_fence();
            step_barrier();
            for (int t = 0; t < n_tiles; ++t) {
                int nI = I, nJ = J;
                tile_next(nI, nJ, NT);
                for (int c = 0; c < 2; ++c) {
                    if (t + 1 < n_tiles) {
                        const float4 *pos = P.pos + (size_t)(chain0 + c) * P.NP;
                        const float4 ctr = P.ctr[(size_t)(chain0 + c) * NT + nI];
                        // the column buffer of tile t + 1 was the one of tile t - 1: its row sums must have completed
                        if (t > 0) mbar_wait(bar_cd + 8 * c, (base_par + t - 1) & 1);
                        if (nI != I) {
                            mbar_wait(bar_r2 + 8 * c, (base_par + t) & 1);   // r^2 of the last tile of row I read the old row features
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const int b = lane + 32 * r;
                                bead_features<true>(pos[nI * TM + b], ctr, sm + OFF_A1 + c * (TM / 8) * GRP + (b & 7) * 16 + (b >> 3) * GRP,
                                                    sm + OFF_UR + (c * 2 + (nI & 1)) * (TM / 8) * 256 + (b & 7) * 16 + (b >> 3) * 256);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const int b = lane + 32 * r;
                            bead_features<false>(pos[nJ * TN + b], ctr,
                                                 sm + OFF_B1 + (c * 2 + ((t + 1) & 1)) * (TN / 8) * GRP + (b & 7) * 16 + (b >> 3) * GRP, 0);
                        }
                        proxy_fence();
                    }
                    step_barrier();
                }
                I = nI, J = nJ;
            }
            step_barrier();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NSW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------- host
static float h_tf32_round(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    memcpy(&v, &u, 4);
    return v;
}

template <int NPOLY>
static void run(const Params &P, int grid, int n, int chains, const std::vector<float4> &pos, const std::vector<float> &ydense,
                float Bc, double clk_ghz, float *d_force, float *d_dbg, bool check) {
    CK(cudaFuncSetAttribute(tcpair<NPOLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int NP = P.NP;
    CK(cudaMemset(d_force, 0, (size_t)chains * NP * 4 * sizeof(float)));
    tcpair<NPOLY><<<grid, NTHREADS, SMEM_BYTES>>>(P);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (check) {
        std::vector<float> force((size_t)chains * NP * 4);
        CK(cudaMemcpy(force.data(), d_force, force.size() * 4, cudaMemcpyDeviceToHost));
        if (d_dbg) {
            std::vector<float> dbg(2 * TM * TN);
            CK(cudaMemcpy(dbg.data(), d_dbg, dbg.size() * 4, cudaMemcpyDeviceToHost));
            double worst = 0, worst_rel = 0;
            for (int J = 0; J < 2; ++J)
                for (int i = 0; i < TM; ++i)
                    for (int jj = 0; jj < TN; ++jj) {
                        const int j = J * TN + jj;
                        if (j <= i || J * TN + (jj / 16) * 16 + 15 <= (i / 32) * 32) continue;
                        const float4 a = pos[i], b = pos[j];
                        const double r2 = (double)(a.x - b.x) * (a.x - b.x) + (double)(a.y - b.y) * (a.y - b.y) + (double)(a.z - b.z) * (a.z - b.z);
                        const double e = fabs(dbg[(J * TM + i) * TN + jj] - r2);
                        if (e > worst) worst = e;
                        if (e / r2 > worst_rel) worst_rel = e / r2;
                    }
            printf("  r^2 of tiles (0,0), (0,1) of chain 0 against float64: max abs err %.3e, max rel err %.3e\n", worst, worst_rel);
        }
#ifdef DEBUG_PRINT
        {
            const float4 *x = pos.data();
            auto coef = [&](int i, int j) {
                const double dx = (double)x[j].x - x[i].x, dy = (double)x[j].y - x[i].y, dz = (double)x[j].z - x[i].z;
                const double d = sqrt(dx * dx + dy * dy + dz * dz);
                const double m = 1.0 / (1.0 - (double)Bc * exp2(d));
                return (m - (double)ydense[(size_t)i * n + j]) * m * (1.0 - m) / d;
            };
            for (int j : {1, 9, 17, 25, 33, 49}) {
                double S = 0, T = 0;
                for (int i = 0; i < j && i < 128; ++i) S += coef(i, j), T += coef(i, j) * (x[i].x);
                printf("expect F tile (0,0) j %d: S %g  T_x(uncentred) %g\n", j, S, T);
            }
            for (int i : {1, 17, 33}) {
                double S = 0;
                for (int j = i + 1; j < n; ++j) S += coef(i, j);
                printf("expect G row tile 0 i %d: S %g\n", i, S);
            }
        }
#endif
        const int which[2] = {0, chains - 1};
        for (int w = 0; w < 2; ++w) {
            const int ch = which[w];
            const float4 *x = pos.data() + (size_t)ch * NP;
            std::vector<double> ref((size_t)n * 3, 0.0);
            for (int i = 0; i < n; ++i)
                for (int j = i + 1; j < n; ++j) {
                    const double dx = (double)x[j].x - x[i].x, dy = (double)x[j].y - x[i].y, dz = (double)x[j].z - x[i].z;
                    const double d = sqrt(dx * dx + dy * dy + dz * dz);
                    const double m = 1.0 / (1.0 - (double)Bc * exp2(d));
                    const double cf = (m - (double)ydense[(size_t)i * n + j]) * m * (1.0 - m) / d;
                    ref[3 * i] += cf * dx, ref[3 * i + 1] += cf * dy, ref[3 * i + 2] += cf * dz;
                    ref[3 * j] -= cf * dx, ref[3 * j + 1] -= cf * dy, ref[3 * j + 2] -= cf * dz;
                }
            double max_ref = 0, max_err = 0, s2 = 0, e2 = 0;
            for (int i = 0; i < n; ++i)
                for (int k = 0; k < 3; ++k) {
                    const double r = ref[3 * i + k], g = force[((size_t)ch * NP + i) * 4 + k];
                    max_ref = fmax(max_ref, fabs(r)), max_err = fmax(max_err, fabs(g - r));
                    s2 += r * r, e2 += (g - r) * (g - r);
                }
            printf("  chain %d force sums against float64: max |F| %.4f, max abs err %.3e, rms rel err %.3e\n", ch, max_ref, max_err, sqrt(e2 / s2));
        }
    }
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a));
        tcpair<NPOLY><<<grid, NTHREADS, SMEM_BYTES>>>(P);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * clk_ghz * 1e9;
    const double evals_per_sm = 2.0 * P.pairs_per_cta;
    const double useful = (double)n * (n - 1) / 2, processed = (double)P.n_tiles * TM * TN;
    const double cpp_useful = cyc / (evals_per_sm * useful / 128.0), cpp_proc = cyc / (evals_per_sm * processed / 128.0);
    printf("NPOLY %2d: %.3f ms per launch (%d chains) = %.2f us per chain-evaluation and SM; %.2f SMSP-cycles per warp-pair (useful pairs), %.2f (tile slots); "
           "31 flop per pair: %.1f %% of the FP32 peak\n",
           NPOLY, best, chains, best * 1e3 / evals_per_sm, cpp_useful, cpp_proc, 100.0 * 31.0 / (2.0 * cpp_useful));
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1000;
    const int pairs_per_cta = argc > 2 ? atoi(argv[2]) : 2;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const int grid = prop.multiProcessorCount;
    const int NT = (n + TM - 1) / TM, NP = NT * TM, n_tiles = NT * (NT + 1);   // sum over I of (2 NT - 2 I)
    const int chains = grid * pairs_per_cta * 2;
    const float alpha = 2.0f, d_c = 2.5f, S = alpha * 1.44269504f;
    const float Bc = -exp2f(-alpha * d_c * 1.44269504f);
    printf("tcpair: n = %d beads (%d row tiles, %d tile visits of 128 x 64), %d chains on %d SMs, %d KiB shared memory\n", n, NT, n_tiles, chains, grid, SMEM_BYTES / 1024);

    // ground truth: a 3-D random walk with unit steps; chains = truth + noise; everything scaled by S
    srand(1234);
    auto gauss = []() {
        double u = (rand() + 1.0) / (RAND_MAX + 2.0), v = (rand() + 1.0) / (RAND_MAX + 2.0);
        return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
    };
    std::vector<double> truth((size_t)n * 3);
    double p[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) truth[3 * i + k] = (p[k] += gauss() * 0.577);
    std::vector<float> ydense((size_t)n * n, 0.f);
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            double d2 = 0;
            for (int k = 0; k < 3; ++k) d2 += (truth[3 * i + k] - truth[3 * j + k]) * (truth[3 * i + k] - truth[3 * j + k]);
            ydense[(size_t)i * n + j] = (float)(1.0 / (1.0 + exp(alpha * (sqrt(d2) - d_c))) + 0.05 * gauss());
        }
    std::vector<float4> pos((size_t)chains * NP), ctr((size_t)chains * NT);
    for (int c = 0; c < chains; ++c) {
        for (int i = 0; i < NP; ++i) {
            float4 v;
            if (i < n) v = make_float4(S * (float)(truth[3 * i] + 0.1 * gauss()), S * (float)(truth[3 * i + 1] + 0.1 * gauss()), S * (float)(truth[3 * i + 2] + 0.1 * gauss()), 0.f);
            else v = make_float4(3000.f + 300.f * (i - n), 3000.f, 3000.f, 0.f);   // padding beads: far from everything (coef = 0)
            pos[(size_t)c * NP + i] = v;
        }
        for (int I = 0; I < NT; ++I) {
            double s[3] = {0, 0, 0};
            int cnt = 0;
            for (int i = I * TM; i < (I + 1) * TM && i < n; ++i, ++cnt) s[0] += pos[(size_t)c * NP + i].x, s[1] += pos[(size_t)c * NP + i].y, s[2] += pos[(size_t)c * NP + i].z;
            ctr[(size_t)c * NT + I] = make_float4(h_tf32_round((float)(s[0] / cnt)), h_tf32_round((float)(s[1] / cnt)), h_tf32_round((float)(s[2] / cnt)), 0.f);
        }
    }
    // contact tiles in processing order
    std::vector<float> ytiles((size_t)n_tiles * TM * TN, 0.f);
    {
        int t = 0;
        for (int I = 0; I < NT; ++I)
            for (int J = 2 * I; J < 2 * NT; ++J, ++t)
                for (int r = 0; r < TM; ++r)
                    for (int cc = 0; cc < TN; ++cc) {
                        const int i = I * TM + r, j = J * TN + cc;
                        const float v = (i < n && j < n && j > i) ? ydense[(size_t)i * n + j] : 0.f;
                        const int f4 = (cc >> 2) ^ (r & 7);
                        ytiles[((size_t)t * TM + r) * TN + f4 * 4 + (cc & 3)] = v;
                    }
    }
    float4 *d_pos, *d_ctr;
    float *d_y, *d_force, *d_dbg;
    CK(cudaMalloc(&d_pos, pos.size() * sizeof(float4)));
    CK(cudaMalloc(&d_ctr, ctr.size() * sizeof(float4)));
    CK(cudaMalloc(&d_y, ytiles.size() * 4));
    CK(cudaMalloc(&d_force, (size_t)chains * NP * 16));
    CK(cudaMalloc(&d_dbg, 2 * TM * TN * 4));
    CK(cudaMemcpy(d_pos, pos.data(), pos.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ctr, ctr.data(), ctr.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_y, ytiles.data(), ytiles.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_dbg, 0, 2 * TM * TN * 4));
    Params P{d_pos, d_ctr, d_y, d_force, d_dbg, NP, NT, n_tiles, pairs_per_cta, Bc};
    const double ghz = clk / 1e6;
    run<0>(P, grid, n, chains, pos, ydense, Bc, ghz, d_force, d_dbg, true);
    P.dbg = nullptr;
    run<6>(P, grid, n, chains, pos, ydense, Bc, ghz, d_force, nullptr, true);
    run<8>(P, grid, n, chains, pos, ydense, Bc, ghz, d_force, nullptr, false);
    run<10>(P, grid, n, chains, pos, ydense, Bc, ghz, d_force, nullptr, false);
    run<12>(P, grid, n, chains, pos, ydense, Bc, ghz, d_force, nullptr, false);
    return 0;
}
