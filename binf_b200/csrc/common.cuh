// Shared device helpers: Philox4x32-10 streams, MUFU wrappers, warp reductions,
// mbarrier + bulk-copy (TMA, UBLKCP) PTX wrappers.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace binfb {

// ------------------------------------------------------------------------------------------
// Philox4x32-10.  key = 64-bit seed; counter = (chain id lo, chain id hi, draw index, stream
// kind << 24 | element block).  One call per (chain, draw, element): independent of grid
// shape and of how chains are sharded across GPUs.
// ------------------------------------------------------------------------------------------
enum RngKind : uint32_t { RNG_MOMENTUM = 1, RNG_ACCEPT = 2, RNG_GAMMA = 3, RNG_SWAP = 4 };

struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_round(u32x4 &c, uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
    u32x4 r;
    r.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    r.y = (uint32_t)p1;
    r.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    r.w = (uint32_t)p0;
    c = r;
}

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint64_t seed, uint64_t chain,
                                                        uint32_t draw_lo, uint32_t kind_elem) {
    u32x4 c = {(uint32_t)chain, (uint32_t)(chain >> 32), draw_lo, kind_elem};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// uniform in (0, 1]: never 0, so log() is safe
__host__ __device__ __forceinline__ float u32_to_unit_open0(uint32_t x) {
    return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
// uniform in [0, 1)
__host__ __device__ __forceinline__ float u32_to_unit(uint32_t x) {
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}
__host__ __device__ __forceinline__ double u64_to_unit_open0(uint32_t hi, uint32_t lo) {
    const uint64_t v = ((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11);  // 53 bits
    return ((double)v + 1.0) * (1.0 / 9007199254740992.0);
}

// standard normal for (chain, draw, element e) -- the momentum stream.  One Philox block serves the four
// elements 4b .. 4b+3 (two Box-Muller pairs, cosine and sine branch of each), so a caller that draws
// consecutive elements with compile-time indices (the K coefficients of a chain) pays for one block, one
// logarithm / square root and one sincos per two normals; a caller with a run-time index pays what a private
// block per element would cost.
__device__ __forceinline__ float rng_normal(uint64_t seed, uint64_t chain, uint64_t draw,
                                            uint32_t elem) {
    const u32x4 r = philox4x32_10(seed ^ (draw >> 32) * 0x9E3779B97F4A7C15ull, chain,
                                  (uint32_t)draw, ((uint32_t)RNG_MOMENTUM << 24) | (elem >> 2));
    const bool second = (elem & 2u) != 0u;
    const float u1 = u32_to_unit_open0(second ? r.z : r.x);
    const float u2 = u32_to_unit(second ? r.w : r.y);
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return sqrtf(-2.0f * logf(u1)) * ((elem & 1u) ? s : c);
}

__device__ __forceinline__ float rng_uniform(uint64_t seed, uint64_t chain, uint64_t draw,
                                             uint32_t kind) {
    const u32x4 r = philox4x32_10(seed ^ (draw >> 32) * 0x9E3779B97F4A7C15ull, chain,
                                  (uint32_t)draw, kind << 24);
    return u32_to_unit(r.x);
}

// Gamma(shape, 1), shape >= 1, Marsaglia-Tsang in float64 on a Philox stream.
__device__ inline double rng_gamma(uint64_t seed, uint64_t chain, uint64_t draw, double shape) {
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t it = 0; it < 64; ++it) {
        const u32x4 r = philox4x32_10(seed ^ (draw >> 32) * 0x9E3779B97F4A7C15ull, chain,
                                      (uint32_t)draw, ((uint32_t)RNG_GAMMA << 24) | it);
        const double u1 = u64_to_unit_open0(r.x, r.y);
        const u32x4 r2 = philox4x32_10(seed ^ (draw >> 32) * 0x9E3779B97F4A7C15ull, chain,
                                       (uint32_t)draw, ((uint32_t)RNG_GAMMA << 24) | (it + 64u));
        const double u2 = u64_to_unit_open0(r2.x, r2.y);
        const double u3 = u64_to_unit_open0(r2.z, r2.w);
        const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(u3) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
    }
    return d;  // unreachable in practice (acceptance > 95 % per iteration)
}

// ------------------------------------------------------------------------------------------
// MUFU (special function unit) approximations: 1 instruction each
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float mufu_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------------------------------
// warp / sub-warp butterfly all-reduce.  Every stage adds the same two numbers on both lanes,
// so all lanes end with bitwise identical results.
// ------------------------------------------------------------------------------------------
template <int WIDTH, typename T>
__device__ __forceinline__ T group_allreduce_sum(T v) {
#pragma unroll
    for (int m = WIDTH / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// ------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine; SASS UBLKCP) -- global -> shared
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                              uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace binfb
