// The 4x4 bead-pair block of the chromatin kernel, in several code shapes.
//
// All variants compute, for rows r (beads of the lane's own quad) and columns c (beads of the
// partner quad):
//     d = sqrt(|x_r - x_c|^2 + soft), e = 2^(A d + B), m = 1/(1+e), res = m - y_rc,
//     coef = res * m(1-m) / d
//     G_r += coef (x_c - x_r)      (= minus the force sum on the row bead; folded at flush time)
//     F_c += coef (x_c - x_r)      (force sum on the column bead)
//     chi += res^2                  (ENERGY only)
// Sign convention: rows are kept negated (nx = -x_r) so that the pair difference is one packed
// add; the row accumulator G therefore has the same sign as the column accumulator F and the
// caller subtracts it when flushing.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace binfb {

constexpr float PAIR_SOFT = 1e-12f;

struct f2 {
    float2 v;
};
__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// 2^t for t in [-126, 126] on the FMA pipe: Cody-Waite split t = n + f, f in [-1/2, 1/2],
// degree-5 near-minimax polynomial for 2^f (max rel. error 2.4e-7 incl. fp32 rounding), exponent patched in with one integer op.
__device__ __forceinline__ float poly_ex2(float t) {
    t = fminf(fmaxf(t, -126.0f), 126.0f);
    const float magic = 12582912.0f;  // 1.5 * 2^23
    const float tm = t + magic;
    const float f = t - (tm - magic);
    float p = 1.3390863366e-3f;  // Chebyshev-node interpolant of 2^f on [-1/2, 1/2]
    p = fmaf(p, f, 9.6760319183e-3f);
    p = fmaf(p, f, 5.5503571142e-2f);
    p = fmaf(p, f, 2.4022107485e-1f);
    p = fmaf(p, f, 6.9314718803e-1f);
    p = fmaf(p, f, 1.0000000755f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(tm) << 23));
}

// packed version of poly_ex2 for t <= 126 (callers guarantee t > -126): 3 + 5 packed FMA-pipe
// ops and 4 scalar ALU-pipe ops for two values
__device__ __forceinline__ float2 poly_ex2_2(float2 t) {
    t = mk2(fminf(t.x, 126.0f), fminf(t.y, 126.0f));
    const float2 magic = mk2(12582912.0f, 12582912.0f);
    const float2 tm = add2(t, magic);
    const float2 nn = fma2(tm, mk2(-1.f, -1.f), magic);  // -(round(t))
    const float2 f = add2(t, nn);
    float2 p = fma2(mk2(1.3390863366e-3f, 1.3390863366e-3f), f, mk2(9.6760319183e-3f, 9.6760319183e-3f));
    p = fma2(p, f, mk2(5.5503571142e-2f, 5.5503571142e-2f));
    p = fma2(p, f, mk2(2.4022107485e-1f, 2.4022107485e-1f));
    p = fma2(p, f, mk2(6.9314718803e-1f, 6.9314718803e-1f));
    p = fma2(p, f, mk2(1.0000000755f, 1.0000000755f));
    return mk2(__int_as_float(__float_as_int(p.x) + (__float_as_int(tm.x) << 23)),
               __int_as_float(__float_as_int(p.y) + (__float_as_int(tm.y) << 23)));
}

// ---------------------------------------------------------------------------------------------
// scalar reference shape (one pair at a time; the compiler interleaves)
// ---------------------------------------------------------------------------------------------
// Excluded-volume prior (EV): E_ev = k_ev sum_{i<j} max(0, d_ev - d_ij)^4 (a quartic repulsion between
// beads closer than d_ev).  Its force -4 k_ev s^3 (x_i - x_j)/d has the same geometry as the
// likelihood force, so it rides on the same accumulators: coef += cev * s^3 / d with
// cev = 4 k_ev / (alpha beta tau) cancelling the factor -alpha beta tau the caller applies to the sums.
// SCALED: see pair_packed_gs (positions pre-multiplied by the slope of the exponent; A carries the scaled
// softening, B carries -2^B, dev / cev the scaled excluded-volume constants).
template <bool ENERGY, bool EV = false, bool SCALED = false, bool ALG = false>
__device__ __forceinline__ void pair_scalar(float nxi, float nyi, float nzi, float xj, float yj,
                                            float zj, float y, float A, float B, float &gx,
                                            float &gy, float &gz, float &fx, float &fy, float &fz,
                                            float &chi, float dev = 0.f, float cev = 0.f, float *ev = nullptr) {
    const float dx = xj + nxi, dy = yj + nyi, dz = zj + nzi;
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, SCALED ? A : PAIR_SOFT)));
    const float inv = mufu_rsqrt(r2);
    const float d = r2 * inv;
    float rs, wn;
    if (ALG) {  // algebraic contact function (see pair_packed_gs); SCALED only
        const float zn = d + B;                         // -z
        const float ri = mufu_rsqrt(fmaf(zn, zn, 1.0f));
        rs = fmaf(zn * ri, 0.5f, y) - 0.5f;             // y - m
        wn = ri * ri * ri;                              // 2 dm/dz
    } else {
        const float sn = SCALED ? fmaf(mufu_ex2(d), B, -1.0f) : fmaf(mufu_ex2(fmaf(d, A, B)), -1.0f, -1.0f);
        const float mn = mufu_rcp(sn);  // -m
        rs = mn + y;                    // -(m - y)
        wn = fmaf(mn, mn, mn);          // -(m - m^2)
    }
    float coef = rs * wn * inv;
    if (EV) {
        const float s = fmaxf(dev - d, 0.f), s2 = s * s;
        coef = fmaf(s2 * s * cev, inv, coef);
        if (ENERGY) *ev = fmaf(s2, s2, *ev);
    }
    gx = fmaf(coef, dx, gx), gy = fmaf(coef, dy, gy), gz = fmaf(coef, dz, gz);
    fx = fmaf(coef, dx, fx), fy = fmaf(coef, dy, fy), fz = fmaf(coef, dz, fz);
    if (ENERGY) chi = fmaf(rs, rs, chi);
}

// ---------------------------------------------------------------------------------------------
// packed shape: two columns (c, c+1) of one row per instruction (FFMA2 / FADD2 / FMUL2).
// nx2 = (-x_r, -x_r) etc. are broadcast pairs; xj2, F2 are natural register pairs.
// POLY: evaluate 2^t on the FMA pipe instead of MUFU.EX2 for this pack.
// ---------------------------------------------------------------------------------------------
template <bool ENERGY, bool POLY>
__device__ __forceinline__ void pair_packed(float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                            float2 zj2, float2 y2, float2 A2, float2 B2, float2 &gx2,
                                            float2 &gy2, float2 &gz2, float2 &fx2, float2 &fy2,
                                            float2 &fz2, float2 &chi2) {
    const float2 dx = add2(xj2, nx2), dy = add2(yj2, ny2), dz = add2(zj2, nz2);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, mk2(PAIR_SOFT, PAIR_SOFT))));
    const float2 inv = mk2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 d = mul2(r2, inv);
    const float2 t = fma2(d, A2, B2);
    float2 e;
    if (POLY) e = poly_ex2_2(t);
    else e = mk2(mufu_ex2(t.x), mufu_ex2(t.y));
    const float2 sn = fma2(e, mk2(-1.f, -1.f), mk2(-1.f, -1.f));
    const float2 mn = mk2(mufu_rcp(sn.x), mufu_rcp(sn.y));
    const float2 rs = add2(mn, y2);
    const float2 wn = fma2(mn, mn, mn);
    const float2 coef = mul2(mul2(rs, wn), inv);
    gx2 = fma2(coef, dx, gx2), gy2 = fma2(coef, dy, gy2), gz2 = fma2(coef, dz, gz2);
    fx2 = fma2(coef, dx, fx2), fy2 = fma2(coef, dy, fy2), fz2 = fma2(coef, dz, fz2);
    if (ENERGY) chi2 = fma2(rs, rs, chi2);
}

// ---------------------------------------------------------------------------------------------
// packed shape with SCALAR row accumulators: the three row accumulations (coef*d + G, three distinct
// register pairs = 3 dispatch cycles each as FFMA2) become six scalar FFMA into one scalar G per
// component (1.15 cycles each), which also halves the registers the row accumulators occupy
// (12 instead of 24 per lane).
// ---------------------------------------------------------------------------------------------
// SCALED: the positions arrive multiplied by A (the slope of the exponent, A > 0), so that the distance
// itself is the exponent: d' = A d, e = 2^B 2^d'.  The FMA that formed A d + B disappears (2^B rides on
// the FMA that forms -(1 + e)) and nothing else changes: coef' = coef / A and dx' = A dx, so the force sums
// coef' dx' come out in the original units.  In this mode A2 carries the scaled softening (A^2 soft) and B2
// carries -2^B; dev / cev are the scaled excluded-volume constants (A d_ev, cev / A^3; the energy sum
// comes out multiplied by A^4).
//
// ALG (SCALED only): the ALGEBRAIC contact function of SURVEY.md A.2, mock = 1/2 (1 + z / sqrt(1 + z^2)) with
// z = alpha (d_c - d): two rsqrt and no exponential.  The positions arrive multiplied by alpha, so that
// -z = d' - alpha d_c is one add (B2 carries -alpha d_c, A2 the scaled softening).  With ri = 1/sqrt(1 + z^2):
//   m = 1/2 - 1/2 (-z) ri,   dm/dz = 1/2 ri^3,   coefficient = (y - m) * ri^3 / d'
// which is -2 x the coefficient convention of the logistic form ((m - y) dm/dz / d); the caller folds the -1/2 into
// the force scale (ChromDev::alpha = -alpha / 2), so the block pays no instruction for it.
template <bool ENERGY, bool EV = false, bool SCALED = false, bool ALG = false>
__device__ __forceinline__ void pair_packed_gs(float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                               float2 zj2, float2 y2, float2 A2, float2 B2, float &gx, float &gy,
                                               float &gz, float2 &fx2, float2 &fy2, float2 &fz2, float2 &chi2,
                                               float dev = 0.f, float cev = 0.f, float2 *ev2 = nullptr) {
    const float2 dx = add2(xj2, nx2), dy = add2(yj2, ny2), dz = add2(zj2, nz2);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, SCALED ? A2 : mk2(PAIR_SOFT, PAIR_SOFT))));
    const float2 inv = mk2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 d = mul2(r2, inv);
    float2 rs, wn;
    if (ALG) {
        const float2 zn = add2(d, B2);                                     // -z
        const float2 s1 = fma2(zn, zn, mk2(1.f, 1.f));
        const float2 ri = mk2(mufu_rsqrt(s1.x), mufu_rsqrt(s1.y));
        rs = add2(fma2(mul2(zn, ri), mk2(0.5f, 0.5f), y2), mk2(-0.5f, -0.5f));  // y - m
        wn = mul2(mul2(ri, ri), ri);                                       // 2 dm/dz
    } else {
        float2 sn;
        if (SCALED) {
            const float2 e = mk2(mufu_ex2(d.x), mufu_ex2(d.y));
            sn = fma2(e, B2, mk2(-1.f, -1.f));
        } else {
            const float2 t = fma2(d, A2, B2);
            const float2 e = mk2(mufu_ex2(t.x), mufu_ex2(t.y));
            sn = fma2(e, mk2(-1.f, -1.f), mk2(-1.f, -1.f));
        }
        const float2 mn = mk2(mufu_rcp(sn.x), mufu_rcp(sn.y));
        rs = add2(mn, y2);
        wn = fma2(mn, mn, mn);
    }
    float2 coef = mul2(mul2(rs, wn), inv);
    if (EV) {
        float2 s = fma2(d, mk2(-1.f, -1.f), mk2(dev, dev));   // d_ev - d
        s = mk2(fmaxf(s.x, 0.f), fmaxf(s.y, 0.f));
        const float2 s2 = mul2(s, s);
        coef = fma2(mul2(mul2(s2, s), mk2(cev, cev)), inv, coef);
        if (ENERGY) *ev2 = fma2(s2, s2, *ev2);
    }
    gx = fmaf(coef.y, dx.y, fmaf(coef.x, dx.x, gx));
    gy = fmaf(coef.y, dy.y, fmaf(coef.x, dy.x, gy));
    gz = fmaf(coef.y, dz.y, fmaf(coef.x, dz.x, gz));
    fx2 = fma2(coef, dx, fx2), fy2 = fma2(coef, dy, fy2), fz2 = fma2(coef, dz, fz2);
    if (ENERGY) chi2 = fma2(rs, rs, chi2);
}

// ---------------------------------------------------------------------------------------------
// knock-out study (profiles/microbench/pairbench.cu): pair_packed_gs<false, false, true> with single pieces
// replaced by one ALU-pipe integer op (so that the FMA-pipe load stays what it was) or removed.  Not used by the
// kernels.  KO bits: 1 rsqrt, 2 ex2, 4 rcp, 8 row accumulators G, 16 column accumulators F.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float alu_fake(float x) { return __int_as_float(__float_as_int(x) ^ 0x00400000); }
template <int KO>
__device__ __forceinline__ void pair_packed_ko(float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                               float2 zj2, float2 y2, float2 A2, float2 B2, float &gx, float &gy,
                                               float &gz, float2 &fx2, float2 &fy2, float2 &fz2) {
    const float2 dx = add2(xj2, nx2), dy = add2(yj2, ny2), dz = add2(zj2, nz2);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, A2)));
    const float2 inv = (KO & 1) ? mk2(alu_fake(r2.x), alu_fake(r2.y)) : mk2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 d = mul2(r2, inv);
    const float2 e = (KO & 2) ? mk2(alu_fake(d.x), alu_fake(d.y)) : mk2(mufu_ex2(d.x), mufu_ex2(d.y));
    const float2 sn = fma2(e, B2, mk2(-1.f, -1.f));
    const float2 mn = (KO & 4) ? mk2(alu_fake(sn.x), alu_fake(sn.y)) : mk2(mufu_rcp(sn.x), mufu_rcp(sn.y));
    const float2 rs = add2(mn, y2);
    const float2 wn = fma2(mn, mn, mn);
    const float2 coef = mul2(mul2(rs, wn), inv);
    if (!(KO & 8)) {
        gx = fmaf(coef.y, dx.y, fmaf(coef.x, dx.x, gx));
        gy = fmaf(coef.y, dy.y, fmaf(coef.x, dy.x, gy));
        gz = fmaf(coef.y, dz.y, fmaf(coef.x, dz.x, gz));
    } else {
        gx += coef.x * 1e-30f;  // keep coef alive
    }
    if (!(KO & 16)) fx2 = fma2(coef, dx, fx2), fy2 = fma2(coef, dy, fy2), fz2 = fma2(coef, dz, fz2);
    else fx2 = add2(fx2, coef);
}

// 1 / x for x <= -1 (the negated logistic denominator -(1 + e)) WITHOUT the special-function unit: magic-constant
// seed (one integer subtraction on the ALU pipe, |error| < 12.5 %) and two cubically convergent steps
// x <- x + x (r + r^2), r = 1 - s x on the FMA pipe: 0.125 -> 2e-3 -> 8e-9, i.e. as accurate as MUFU.RCP.
// 6 packed FMA-pipe ops for two values.  Used for a FRACTION of the packs to balance the two pipes
// (the XU needs 8 SMSP-cycles per warp-wide MUFU, this costs ~6 FMA-pipe cycles per pair).
__device__ __forceinline__ float2 rcp_neg_fma2(float2 sn) {
    float2 x = mk2(__int_as_float(0xFEF311C7u - __float_as_uint(sn.x)), __int_as_float(0xFEF311C7u - __float_as_uint(sn.y)));
    x = mk2(-x.x, -x.y);  // (folded into the operand modifiers of the first FMA)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const float2 r = fma2(sn, mk2(-x.x, -x.y), mk2(1.f, 1.f));  // 1 - sn x
        const float2 t = fma2(r, r, r);
        x = fma2(x, t, x);
    }
    return x;
}

template <bool ENERGY, bool FMARCP>
__device__ __forceinline__ void pair_packed_gs_fr(float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                                  float2 zj2, float2 y2, float2 A2, float2 B2, float &gx, float &gy,
                                                  float &gz, float2 &fx2, float2 &fy2, float2 &fz2, float2 &chi2) {
    const float2 dx = add2(xj2, nx2), dy = add2(yj2, ny2), dz = add2(zj2, nz2);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, A2)));
    const float2 inv = mk2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 d = mul2(r2, inv);
    float2 e = mk2(mufu_ex2(d.x), mufu_ex2(d.y));
    if (FMARCP) e = mk2(fminf(e.x, 1e30f), fminf(e.y, 1e30f));  // keep the Newton steps finite (m < 1e-30 there)
    const float2 sn = fma2(e, B2, mk2(-1.f, -1.f));
    const float2 mn = FMARCP ? rcp_neg_fma2(sn) : mk2(mufu_rcp(sn.x), mufu_rcp(sn.y));
    const float2 rs = add2(mn, y2);
    const float2 wn = fma2(mn, mn, mn);
    const float2 coef = mul2(mul2(rs, wn), inv);
    gx = fmaf(coef.y, dx.y, fmaf(coef.x, dx.x, gx));
    gy = fmaf(coef.y, dy.y, fmaf(coef.x, dy.x, gy));
    gz = fmaf(coef.y, dz.y, fmaf(coef.x, dz.x, gz));
    fx2 = fma2(coef, dx, fx2), fy2 = fma2(coef, dy, fy2), fz2 = fma2(coef, dz, fz2);
    if (ENERGY) chi2 = fma2(rs, rs, chi2);
}

// scaled positions + scalar row accumulators + ONE reciprocal per pack (1/a = b/(ab), 1/b = a/(ab)): 2.5 MUFU
// per pair.  The exponent is clamped to 30 so that (1 + C 2^d)^2 cannot overflow (m < 2^-30/C there).
template <bool ENERGY>
__device__ __forceinline__ void pair_packed_gs_sr(float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                                  float2 zj2, float2 y2, float2 A2, float2 B2, float &gx, float &gy,
                                                  float &gz, float2 &fx2, float2 &fy2, float2 &fz2, float2 &chi2) {
    const float2 dx = add2(xj2, nx2), dy = add2(yj2, ny2), dz = add2(zj2, nz2);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, A2)));
    const float2 inv = mk2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 d = mul2(r2, inv);
    const float2 e = mk2(mufu_ex2(fminf(d.x, 40.0f)), mufu_ex2(fminf(d.y, 40.0f)));
    const float2 sn = fma2(e, B2, mk2(-1.f, -1.f));  // -(1 + e)
    const float ip = mufu_rcp(sn.x * sn.y);          // 1 / ((1+e_x)(1+e_y)) > 0
    const float2 mn = mul2(mk2(sn.y, sn.x), mk2(ip, ip));
    const float2 rs = add2(mn, y2);
    const float2 wn = fma2(mn, mn, mn);
    const float2 coef = mul2(mul2(rs, wn), inv);
    gx = fmaf(coef.y, dx.y, fmaf(coef.x, dx.x, gx));
    gy = fmaf(coef.y, dy.y, fmaf(coef.x, dy.x, gy));
    gz = fmaf(coef.y, dz.y, fmaf(coef.x, dz.x, gz));
    fx2 = fma2(coef, dx, fx2), fy2 = fma2(coef, dy, fy2), fz2 = fma2(coef, dz, fz2);
    if (ENERGY) chi2 = fma2(rs, rs, chi2);
}

// ---------------------------------------------------------------------------------------------
// hybrid shape.  Measured on B200 (profiles/microbench/ffma2_operands.cu): FFMA2 with three distinct
// register operands occupies the FMA pipe 3.04 cycles (register-bank reads), two-operand packed ops
// 2.04, a scalar FFMA with three distinct registers 1.15.  So two-operand work stays packed (half
// the issue slots) and every three-operand FMA -- the six force accumulations and A*d+B -- is
// scalar: 21.3 instead of 23.5 FMA-pipe cycles per pair.
// ---------------------------------------------------------------------------------------------
template <bool ENERGY>
__device__ __forceinline__ void pair_hybrid(float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                            float2 zj2, float2 y2, float A, float B, float2 &gx2,
                                            float2 &gy2, float2 &gz2, float2 &fx2, float2 &fy2,
                                            float2 &fz2, float2 &chi2) {
    const float2 dx = add2(xj2, nx2), dy = add2(yj2, ny2), dz = add2(zj2, nz2);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, mk2(PAIR_SOFT, PAIR_SOFT))));
    const float2 inv = mk2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 d = mul2(r2, inv);
    const float2 e = mk2(mufu_ex2(fmaf(d.x, A, B)), mufu_ex2(fmaf(d.y, A, B)));
    const float2 sn = fma2(e, mk2(-1.f, -1.f), mk2(-1.f, -1.f));
    const float2 mn = mk2(mufu_rcp(sn.x), mufu_rcp(sn.y));
    const float2 rs = add2(mn, y2);
    const float2 wn = fma2(mn, mn, mn);
    const float2 coef = mul2(mul2(rs, wn), inv);
    gx2.x = fmaf(coef.x, dx.x, gx2.x), gx2.y = fmaf(coef.y, dx.y, gx2.y);
    gy2.x = fmaf(coef.x, dy.x, gy2.x), gy2.y = fmaf(coef.y, dy.y, gy2.y);
    gz2.x = fmaf(coef.x, dz.x, gz2.x), gz2.y = fmaf(coef.y, dz.y, gz2.y);
    fx2.x = fmaf(coef.x, dx.x, fx2.x), fx2.y = fmaf(coef.y, dx.y, fx2.y);
    fy2.x = fmaf(coef.x, dy.x, fy2.x), fy2.y = fmaf(coef.y, dy.y, fy2.y);
    fz2.x = fmaf(coef.x, dz.x, fz2.x), fz2.y = fmaf(coef.y, dz.y, fz2.y);
    if (ENERGY) chi2 = fma2(rs, rs, chi2);
}

// ---------------------------------------------------------------------------------------------
// packed shape with a shared reciprocal: the two logistic denominators of a pack are inverted with
// ONE MUFU.RCP of their product (1/a = b / (ab), 1/b = a / (ab)).  t is clamped to <= 30 so that
// (1 + 2^t)^2 <= 2^61 cannot overflow; the clamp changes m by < 1e-9 (m < 2^-30 there).
// 2.5 MUFU + 20.5 FMA-pipe cycles per pair instead of 3 + 19.
// ---------------------------------------------------------------------------------------------
template <bool ENERGY>
__device__ __forceinline__ void pair_packed_sr(float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                               float2 zj2, float2 y2, float2 A2, float2 B2, float2 &gx2,
                                               float2 &gy2, float2 &gz2, float2 &fx2, float2 &fy2,
                                               float2 &fz2, float2 &chi2) {
    const float2 dx = add2(xj2, nx2), dy = add2(yj2, ny2), dz = add2(zj2, nz2);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, mk2(PAIR_SOFT, PAIR_SOFT))));
    const float2 inv = mk2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 d = mul2(r2, inv);
    const float2 t = fma2(d, A2, B2);
    const float2 e = mk2(mufu_ex2(fminf(t.x, 30.0f)), mufu_ex2(fminf(t.y, 30.0f)));
    const float2 sn = fma2(e, mk2(-1.f, -1.f), mk2(-1.f, -1.f));  // -(1 + e)
    const float ip = mufu_rcp(sn.x * sn.y);                        // 1 / ((1+e_x)(1+e_y)) > 0
    const float2 mn = mul2(mk2(sn.y, sn.x), mk2(ip, ip));          // (-m_x, -m_y)
    const float2 rs = add2(mn, y2);
    const float2 wn = fma2(mn, mn, mn);
    const float2 coef = mul2(mul2(rs, wn), inv);
    gx2 = fma2(coef, dx, gx2), gy2 = fma2(coef, dy, gy2), gz2 = fma2(coef, dz, gz2);
    fx2 = fma2(coef, dx, fx2), fy2 = fma2(coef, dy, fy2), fz2 = fma2(coef, dz, fz2);
    if (ENERGY) chi2 = fma2(rs, rs, chi2);
}

// ---------------------------------------------------------------------------------------------
// staged shape: NP packs advance through the MUFU stages together (all rsqrt, then all ex2, then
// all rcp), so that NP*2 special-function ops are in flight behind independent FMA work.
// pack p uses row nrow[p] (negated broadcast position), partner column pair xj[p], contacts y[p],
// row accumulators g[p] and column accumulators f[p] (callers alias them as needed).
// ---------------------------------------------------------------------------------------------
// volatile variants: ptxas keeps volatile asm statements in program order, which pins the stage order
__device__ __forceinline__ float vmufu_rsqrt(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float vmufu_ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float vmufu_rcp(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <bool ENERGY, int NP>
struct StagedPairs {
    float2 dx[NP], dy[NP], dz[NP], inv[NP], e[NP];
    __device__ __forceinline__ void stage1(int p, float2 nx, float2 ny, float2 nz, float2 xj, float2 yj, float2 zj) {
        dx[p] = add2(xj, nx), dy[p] = add2(yj, ny), dz[p] = add2(zj, nz);
        const float2 r2 = fma2(dz[p], dz[p], fma2(dy[p], dy[p], fma2(dx[p], dx[p], mk2(PAIR_SOFT, PAIR_SOFT))));
        inv[p] = mk2(vmufu_rsqrt(r2.x), vmufu_rsqrt(r2.y));
        e[p] = r2;
    }
    __device__ __forceinline__ void stage2(int p, float2 A2, float2 B2) {
        const float2 t = fma2(mul2(e[p], inv[p]), A2, B2);
        e[p] = mk2(vmufu_ex2(t.x), vmufu_ex2(t.y));
    }
    __device__ __forceinline__ void stage3(int p) {
        const float2 sn = fma2(e[p], mk2(-1.f, -1.f), mk2(-1.f, -1.f));
        e[p] = mk2(vmufu_rcp(sn.x), vmufu_rcp(sn.y));  // -m
    }
    __device__ __forceinline__ void stage4(int p, float2 y2, float2 &gx, float2 &gy, float2 &gz, float2 &fx,
                                           float2 &fy, float2 &fz, float2 &chi2) {
        const float2 mn = e[p];
        const float2 rs = add2(mn, y2);
        const float2 wn = fma2(mn, mn, mn);
        const float2 coef = mul2(mul2(rs, wn), inv[p]);
        gx = fma2(coef, dx[p], gx), gy = fma2(coef, dy[p], gy), gz = fma2(coef, dz[p], gz);
        fx = fma2(coef, dx[p], fx), fy = fma2(coef, dy[p], fy), fz = fma2(coef, dz[p], fz);
        if (ENERGY) chi2 = fma2(rs, rs, chi2);
    }
};

// packed two-operand work, ALL twelve force accumulations as scalar FFMA grouped by coefficient (three
// consecutive FFMA share one multiplicand: operand-reuse cache) -- harness only
__device__ __forceinline__ void pair_packed_gfs(float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                                float2 zj2, float2 y2, float2 A2, float2 B2, float &gx, float &gy,
                                                float &gz, float2 &fx2, float2 &fy2, float2 &fz2) {
    const float2 dx = add2(xj2, nx2), dy = add2(yj2, ny2), dz = add2(zj2, nz2);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, A2)));
    const float2 inv = mk2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 d = mul2(r2, inv);
    const float2 e = mk2(mufu_ex2(d.x), mufu_ex2(d.y));
    const float2 sn = fma2(e, B2, mk2(-1.f, -1.f));
    const float2 mn = mk2(mufu_rcp(sn.x), mufu_rcp(sn.y));
    const float2 rs = add2(mn, y2);
    const float2 wn = fma2(mn, mn, mn);
    const float2 coef = mul2(mul2(rs, wn), inv);
    gx = fmaf(coef.x, dx.x, gx), gy = fmaf(coef.x, dy.x, gy), gz = fmaf(coef.x, dz.x, gz);
    fx2.x = fmaf(coef.x, dx.x, fx2.x), fy2.x = fmaf(coef.x, dy.x, fy2.x), fz2.x = fmaf(coef.x, dz.x, fz2.x);
    gx = fmaf(coef.y, dx.y, gx), gy = fmaf(coef.y, dy.y, gy), gz = fmaf(coef.y, dz.y, gz);
    fx2.y = fmaf(coef.y, dx.y, fx2.y), fy2.y = fmaf(coef.y, dy.y, fy2.y), fz2.y = fmaf(coef.y, dz.y, fz2.y);
}

// ---------------------------------------------------------------------------------------------
// software-pipelined shape (scaled positions): a pack walks through four stages, one per "slot", so that
// every special-function result is consumed a whole slot after it was requested and the instruction stream
// of a warp carries the same mix of MUFU and FMA work at every point (the plain shapes above alternate
// between FMA-heavy and MUFU-heavy phases and drain at every step boundary):
//   S1  dx, r^2, request rsqrt      S2  d = r^2 * inv, request 2^d
//   S3  -(1 + C 2^d), request rcp   S4  coefficient, row sums G (scalar), column sums F (packed)
// A2 = scaled softening, B2 = -2^B (see pair_packed_gs, SCALED).  The MUFU requests are volatile asm: ptxas
// keeps them in program order, which pins the slot structure.
// ---------------------------------------------------------------------------------------------
struct PipePack {
    float2 dx, dy, dz, inv, v;  // v: r^2 after S1, 2^d after S2, -m after S3
};
__device__ __forceinline__ void pipe_s1(PipePack &p, float2 nx2, float2 ny2, float2 nz2, float2 xj2, float2 yj2,
                                        float2 zj2, float2 A2) {
    p.dx = add2(xj2, nx2), p.dy = add2(yj2, ny2), p.dz = add2(zj2, nz2);
    p.v = fma2(p.dz, p.dz, fma2(p.dy, p.dy, fma2(p.dx, p.dx, A2)));
    p.inv = mk2(vmufu_rsqrt(p.v.x), vmufu_rsqrt(p.v.y));
}
__device__ __forceinline__ void pipe_s2(PipePack &p) {
    const float2 d = mul2(p.v, p.inv);
    p.v = mk2(vmufu_ex2(d.x), vmufu_ex2(d.y));
}
__device__ __forceinline__ void pipe_s3(PipePack &p, float2 B2) {
    const float2 sn = fma2(p.v, B2, mk2(-1.f, -1.f));
    p.v = mk2(vmufu_rcp(sn.x), vmufu_rcp(sn.y));
}
template <bool ENERGY>
__device__ __forceinline__ void pipe_s4(const PipePack &p, float2 y2, float &gx, float &gy, float &gz, float2 &fx2,
                                        float2 &fy2, float2 &fz2, float2 &chi2) {
    const float2 rs = add2(p.v, y2);
    const float2 wn = fma2(p.v, p.v, p.v);
    const float2 coef = mul2(mul2(rs, wn), p.inv);
    gx = fmaf(coef.y, p.dx.y, fmaf(coef.x, p.dx.x, gx));
    gy = fmaf(coef.y, p.dy.y, fmaf(coef.x, p.dy.x, gy));
    gz = fmaf(coef.y, p.dz.y, fmaf(coef.x, p.dz.x, gz));
    fx2 = fma2(coef, p.dx, fx2), fy2 = fma2(coef, p.dy, fy2), fz2 = fma2(coef, p.dz, fz2);
    if (ENERGY) chi2 = fma2(rs, rs, chi2);
}

}  // namespace binfb
