// Fused HMC / log-prob+gradient for the chromatin bead-chain posterior -- sm_100a.
//
// What it replaces: for the contact-frequency model of SURVEY.md A.2 expressed behind the
// reference API, every pdf.gradient / pdf.log_prob call made by HMCSampler._leapfrog / .sample
// (binf/samplers/hmc.py:92-125,136-164), i.e. Posterior._evaluate_gradient
// (binf/pdf/posteriors.py:173-187) -> Likelihood._evaluate_gradient = J(theta).dot(dE/dmock)
// (binf/pdf/likelihoods.py:148-155) with the GaussianErrorModel (binf/example/likelihood.py:54-61),
// the structure prior, the Metropolis test and the step-size adaption (hmc.py:151-157), and the
// conjugate precision update (binf/example/samplers.py:27-51).  The 3n x M Jacobian (12 GB at
// n = 1000) is never formed: the pair loop applies it on the fly.
//
// Work decomposition
//   * R warps ("roles") per chain, W chains per CTA (W*R <= 16 consumer warps) sharing one stream
//     of contact data y; R = 1 for small n (16 chains fit in shared memory), R = 2 at n = 1000.
//   * beads are grouped in quads (4 beads).  Quad a is paired with quads (a+k) mod Q,
//     k = 1..Q/2 ("circulant half shell"): every unordered quad pair is visited exactly once and
//     every quad has the same number of partners, so there is no triangular waste.  Lane l of the
//     warp owns quad a = 32*rb + l of row block rb; its 4 bead positions and force accumulators
//     stay in registers for the whole row block.  The R roles of a chain split the partner
//     offsets k of a row block into R contiguous ranges.  Each step is a 4x4 block of bead pairs: 12
//     positions + 12 partner-force accumulators + 16 contacts in registers, 16 pair evaluations.
//   * within a step the 32 lanes address 32 distinct partner quads, so the read-modify-write of the
//     partner forces in shared memory is conflict- and race-free without atomics.
//   * contacts are pre-laid-out on the host in exactly the order the lanes consume them
//     ([row block][slot][role][row r][lane] float4) and streamed through a 4-stage shared-memory
//     ring with 1-D bulk async copies (TMA engine, UBLKCP) completing on mbarriers; the last warp to
//     release a stage issues its refill (see `Ring`); all warps read the same stage, so L2->SM
//     traffic is 1/W of naive.
//   * shared memory per chain: positions and force sums per quad as [x0..x3 | y0..y3 | z0..z3]
//     (48-byte lane stride: one address register, no bank conflicts).
//   * small batches switch to an alternative plan with twice the roles per chain, kept race-free by
//     a per-step chain barrier (LOCKSTEP); see chrom_plan / chrom_launch.
//   * scheduling: a work item is (trajectory, leapfrog pass, group of W chains).  CTAs are
//     persistent and claim items from an atomic counter; a per-group pass counter (release/acquire)
//     orders the passes of one group.  Between passes q and p round-trip through L2 (24 KB per
//     chain per pass, <0.1 % of the pass) -- this removes the 3.46-waves quantisation a
//     "whole trajectory per CTA" launch has at C = 4096 on 148 SMs.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "internal.h"
#include "pair_block.cuh"

namespace binfb {

// experiment switches (profiles/experiments/build_variants.py).  The pair block is pair_packed_gs of
// pair_block.cuh (packed columns, scalar row accumulators); the other shapes measured there and in
// profiles/README.md (plain packed, hybrid, shared reciprocal) were 0.5-2 % slower in this kernel.
#ifndef BINFB_PEEL
#define BINFB_PEEL 0
#endif
#ifndef BINFB_PREFETCH
#define BINFB_PREFETCH 0  // load the next step's partner positions one step ahead
#endif
#ifndef BINFB_WRAP2
#define BINFB_WRAP2 0  // partner address wraps by comparing with the end address (no step counter)
#endif
#ifndef BINFB_DEFER
#define BINFB_DEFER 0  // consume the ticket of a stage release one stage later (hides the shared-memory atomic's latency)
#endif
#ifndef BINFB_YJIT
#define BINFB_YJIT 0  // load each row's contacts right before its two packs instead of all 16 up front
#endif
#ifndef BINFB_ROTATE
#define BINFB_ROTATE 1  // start each chain group at a different row block
#endif
#ifndef BINFB_FMARCP
#define BINFB_FMARCP 0  // N of the 8 packs of a step take the logistic reciprocal on the FMA pipe (pair_packed_gs_fr)
#endif
#ifndef BINFB_LOCKFLAGS
#define BINFB_LOCKFLAGS 0  // LOCKSTEP kernels: one-directional progress flags between neighbouring roles instead of a
                           // chain barrier after every step (see chrom_sweep)
#endif
#ifndef BINFB_CTAS_PER_SM
#define BINFB_CTAS_PER_SM 1  // persistent CTAs per SM (experiment: two half-size CTAs, with "chrom.warps" = 4 and
                             // -DBINFB_CHROM_NS=2 so that two of them fit in shared memory)
#endif
#ifndef BINFB_SRCP
#define BINFB_SRCP 0    // N of the 8 packs of a step share one MUFU.RCP between their two pairs (pair_packed_gs_sr)
#endif

constexpr float CHROM_SOFT = PAIR_SOFT;
constexpr int STEP_FLOAT4 = 4 * 32;  // float4 per warp-step (16 contacts per lane)
constexpr int STEP_BYTES = STEP_FLOAT4 * 16;
constexpr int FLUSH_ROLE_BYTES = 12 * 32 * 4;  // one role's row sums of a row block: [r*3+comp][lane] float
constexpr int CHAIN_SCRATCH_BYTES = 144;  // per chain, in front of its positions: one double per role (<= 16)
                                          // and the mbarrier of the position copy
#ifndef BINFB_CHROM_NS
#define BINFB_CHROM_NS 4
#endif
#ifndef BINFB_CHROM_SS
#define BINFB_CHROM_SS 4
#endif
constexpr int CHROM_NS = BINFB_CHROM_NS;  // ring depth (stages)

struct ChromDev {
    int n, n_pad, Q, KS, NRB, q_even;
    int R, Lr, SS, S_pad;  // roles per chain, slots per row block, steps per stage, padded slots
    const float4 *ystream;
    float A, B;  // exp(alpha (d - d_c)) = 2^(A d + B)
    // The kernel keeps positions multiplied by S = A (> 0) in shared memory and in the working copy qw, so
    // that the scaled distance IS the exponent (pair_block.cuh, SCALED): e = 2^B 2^(S d).
    float S, invS;      // A and 1 / A
    float softS, nC;    // A^2 * soft and -2^B
    float alpha, k_bb, l0, inv_s2;
    float ev_k, ev_d;  // excluded volume: k_ev (0 = off) and d_ev
    double M;
    // workspace
    float *qw, *pw;
    double *h0, *chi2_0, *chi2_state;
    float *tau_w;
    int *counter, *pass_done;
};

enum { CHROM_MODE_GRAD = 0, CHROM_MODE_HMC = 1 };

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Host pipelining of binfb_hmc_run_host (capi.cu): behind the per-group pass counters the scheduler words hold
//   gate[0]      0 = the whole batch is resident (device entry point); else 1 + number of leading chains whose
//                positions have arrived -- written by the copy-in stream after every chunk of the H2D copy
//                (cuStreamWriteValue32), read here before the first pass of a group;
//   gate[1]      chain groups per output chunk (0 = none);
//   gate[2 + j]  number of groups of output chunk j that have finished their last pass -- the copy-out stream
//                waits on it (cuStreamWaitValue32) before it copies the chunk back.
// So the H2D copy of the batch hides under the first pass of the groups that are already there and the D2H copy
// under the last pass of those still running: one launch, no chunked kernels.
constexpr int CHROM_GATE_WORDS = 2 + 64;
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t a) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t a, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// barrier over the warps of one chain (ids 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void chain_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void unpack4(const float4 v, float (&a)[4]) {
    a[0] = v.x, a[1] = v.y, a[2] = v.z, a[3] = v.w;
}

// Shared memory of one chain.  Positions and force sums are stored per quad of beads as three
// consecutive float4 (x0..x3 | y0..y3 | z0..z3): one address register reaches all three with
// immediate offsets, and the 48-byte lane stride of an LDS.128 / STS.128 over 32 consecutive quads
// is bank-conflict free (48 = 3 * 16 with 3 odd: 8 consecutive lanes cover all 8 16-byte bank groups).
struct ChainSmem {
    float *pos, *frc;  // [Q][3][4]
    double *red;       // [16] cross-role reduction scratch (CHAIN_SCRATCH_BYTES)
    float *flush;      // [F][12][32] row sums of F roles at a time (nullptr: the roles fold them in one after the other)
    int flush_roles;   // F: a power of two <= R
    uint32_t pos_bar;  // shared address of the mbarrier the bulk copy of the positions completes on
};
__device__ __forceinline__ int qidx(int bead, int comp) { return (bead >> 2) * 12 + comp * 4 + (bead & 3); }

// mbarrier / bulk-copy / vector load-store wrappers on raw 32-bit shared addresses (no generic->shared
// conversion and no 64-bit pointer arithmetic in the stage loop)
__device__ __forceinline__ bool bar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// Orders generic-proxy accesses to global memory (ordinary ld/st) against async-proxy accesses (the bulk copies
// of the TMA engine) to the same locations.  The working positions qw are written with __stcg by one CTA and
// fetched with cp.async.bulk by another CTA of the same launch; release/acquire on pass_done orders the two
// CTAs, but the PTX memory model additionally requires a proxy fence on the generic -> async edge.
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
// keep a value in a register: the compiler cannot rematerialise it from its inputs (address arithmetic
// that it would otherwise redo at every use inside the stage loop)
__device__ __forceinline__ uint32_t pin_reg(uint32_t v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}
// true in exactly one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(p));
    return p != 0;
}
template <int OFF>
__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(addr), "n"(OFF), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// The contact ring: CHROM_NS stages of SS warp-steps in shared memory, filled by 1-D bulk async
// copies (TMA engine, SASS UBLKCP) that complete on the `full` mbarrier of the slot.  A slot is handed
// back through a plain shared-memory counter: the LAST of the n_warps consumers to release stage j
// refills the slot with stage j + CHROM_NS itself.  The refill is issued at the earliest possible
// moment, nobody ever waits for an "empty" slot, and a fast warp may run up to CHROM_NS - 1 stages
// ahead of the slowest one.  (Measured alternatives, profiles/README.md: a loader lane that waits on
// `empty` mbarriers, and a loader lane that polls them once per step -- 7 % and 6 % slower.)
struct Ring {
    uint32_t ystage;  // shared address of stage 0
    uint32_t full;    // shared address of full[0] (8 bytes each); the release counters follow at + 8 * CHROM_NS
    const unsigned char *src;  // global source: the contact stream of one pass
    int n_stage_pass;
    int n_warps;
    int rot;  // the pass starts at stream stage `rot` and wraps around (see chrom_kernel)
};

// ring slot of stage gi (the ring depth is a power of two except for the 3-stage ring of the 16-role kernel)
template <int NS>
__device__ __forceinline__ uint32_t ring_slot(uint32_t gi) {
    if constexpr ((NS & (NS - 1)) == 0) return gi & (NS - 1);
    else return gi % NS;
}

template <int STAGE_BYTES, int NS>
__device__ __forceinline__ void ring_issue(const Ring &ring, uint32_t gi, int s_local) {
    const uint32_t sl = ring_slot<NS>(gi);
    const uint32_t bar = ring.full + sl * 8u;
    bar_expect_tx(bar, STAGE_BYTES);
    int st = s_local + ring.rot;
    if (st >= ring.n_stage_pass) st -= ring.n_stage_pass;
    bulk_g2s(ring.ystage + sl * STAGE_BYTES, ring.src + (size_t)st * STAGE_BYTES, STAGE_BYTES, bar);
}

// release stage (gi, s_local): called by one elected lane of every consumer warp after the __syncwarp
// that follows the warp's last read of the slot
template <int STAGE_BYTES, int NS>
__device__ __forceinline__ void ring_release(const Ring &ring, uint32_t gi, int s_local) {
    const uint32_t c = ring.full + NS * 8u + ring_slot<NS>(gi) * 4u;
    uint32_t old;
    // the slot's loads have returned (their values were consumed before the __syncwarp in front of this
    // call), so a relaxed add is enough to order them before the refill the last arriver issues
    asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(c) : "memory");
    if (old == (uint32_t)ring.n_warps - 1u) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(c), "r"(0u) : "memory");
        if (s_local + NS < ring.n_stage_pass) ring_issue<STAGE_BYTES, NS>(ring, gi + NS, s_local + NS);
    }
}

// the two halves of ring_release, for callers that look at the ticket one stage later (BINFB_DEFER)
template <int NS>
__device__ __forceinline__ uint32_t ring_arrive(const Ring &ring, uint32_t gi) {
    const uint32_t c = ring.full + NS * 8u + ring_slot<NS>(gi) * 4u;
    uint32_t old;
    asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(c) : "memory");
    return old;
}
template <int STAGE_BYTES, int NS>
__device__ __forceinline__ void ring_finish(const Ring &ring, uint32_t gi, int s_local, uint32_t old) {
    if (old == (uint32_t)ring.n_warps - 1u) {
        const uint32_t c = ring.full + NS * 8u + ring_slot<NS>(gi) * 4u;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(c), "r"(0u) : "memory");
        if (s_local + NS < ring.n_stage_pass) ring_issue<STAGE_BYTES, NS>(ring, gi + NS, s_local + NS);
    }
}

// The pair sweep of one chain, shared by the R warps ("roles") of that chain.  Role r owns the
// partner steps k in [r*Lr, (r+1)*Lr) of every row block; SPR = SS/R of them per ring stage, and
// Lr is a multiple of SPR so that a stage never straddles two row blocks.  Fills frc with
// sum_j coef_ij (x_i - x_j) (the likelihood force up to the factor -alpha*beta*tau) and returns
// this lane's share of chi^2.
//
// The 4x4 block is evaluated with packed FP32 (FFMA2/FADD2/FMUL2: two partner columns per
// instruction, see pair_block.cuh): 9.5 FMA-pipe issue slots + 3 MUFU per pair.
//
// Race freedom of the partner-force read-modify-write in shared memory: within one warp-step the
// 32 lanes address 32 distinct quads ((a+k) mod Q is a bijection of a).  Two roles r < r' of
// one chain work on partner offsets k and k' >= k + Lr - drift, where drift < CHROM_NS*SPR slots is
// enforced by the shared stage ring; they can only meet on a quad if k' - k <= 31, which the
// host-side plan excludes (Lr - CHROM_NS*SPR >= 33 whenever R > 1).
struct SweepRegs {
    float2 nx2[4], ny2[4], nz2[4];  // own quad, negated, as broadcast pairs
    float g[4][3];                  // G = -(force sum) of the own quad
    int k;                          // partner offset of the next step
    uint32_t paddr;                 // shared address of the partner quad's positions (48 bytes per quad)
    uint32_t pwrap;                 // shared address of quad 0's positions
    int wrap;                       // steps until the partner index wraps from Q - 1 to 0
#if BINFB_PREFETCH
    float4 nxt[3];                  // positions of the NEXT step's partner quad (read-only data: safe to
                                    // load across the __syncwarp that orders the force updates)
#endif
    double chi2;
    double ev;                      // sum of max(0, d_ev - d)^4 over this lane's pairs (EV energy passes)
    float dev, cev;                 // excluded volume: d_ev and 4 k_ev / (alpha beta tau)
};

// one regular step: the 16 pairs (own quad) x (partner quad), every lane (inactive lanes compute on
// quad 0 and never store).  frc_off = byte distance from a quad's positions to its force sums.
template <bool ENERGY, bool EV, bool ALG>
__device__ __forceinline__ void step_fast(SweepRegs &s, uint32_t frc_off, uint32_t yaddr, float2 A2, float2 B2,
                                          bool active) {
    const uint32_t pa = s.paddr, fa = s.paddr + frc_off;
#if BINFB_PREFETCH
    const float4 xj = s.nxt[0], yj = s.nxt[1], zj = s.nxt[2];
#if BINFB_PREFETCH == 1
    {
        const uint32_t pn = s.wrap == 1 ? s.pwrap : pa + 48u;
        s.nxt[0] = lds4<0>(pn), s.nxt[1] = lds4<16>(pn), s.nxt[2] = lds4<32>(pn);
    }
#endif
#else
    const float4 xj = lds4<0>(pa), yj = lds4<16>(pa), zj = lds4<32>(pa);
#endif
    const float4 fx = lds4<0>(fa), fy = lds4<16>(fa), fz = lds4<32>(fa);
    const float2 xj2[2] = {mk2(xj.x, xj.y), mk2(xj.z, xj.w)}, yj2[2] = {mk2(yj.x, yj.y), mk2(yj.z, yj.w)},
                 zj2[2] = {mk2(zj.x, zj.y), mk2(zj.z, zj.w)};
    float2 fx2[2] = {mk2(fx.x, fx.y), mk2(fx.z, fx.w)}, fy2[2] = {mk2(fy.x, fy.y), mk2(fy.z, fy.w)},
           fz2[2] = {mk2(fz.x, fz.y), mk2(fz.z, fz.w)};
    float2 c2 = mk2(0.f, 0.f), ev2 = mk2(0.f, 0.f);
    float4 yv[4];
#if !BINFB_YJIT
    yv[0] = lds4<0>(yaddr), yv[1] = lds4<512>(yaddr), yv[2] = lds4<1024>(yaddr), yv[3] = lds4<1536>(yaddr);
#endif
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#if BINFB_YJIT
        yv[r] = r == 0 ? lds4<0>(yaddr) : r == 1 ? lds4<512>(yaddr) : r == 2 ? lds4<1024>(yaddr) : lds4<1536>(yaddr);
#endif
#if BINFB_FMARCP || BINFB_SRCP
        // experiment (profiles/experiments): a fraction of the packs with another reciprocal
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            constexpr int NALT = BINFB_FMARCP ? BINFB_FMARCP : BINFB_SRCP;
            constexpr int STRIDE = 8 / NALT;
            const bool alt = !EV && !ALG && ((r * 2 + h) % STRIDE) == STRIDE - 1;
            const float2 y2 = h ? mk2(yv[r].z, yv[r].w) : mk2(yv[r].x, yv[r].y);
            if (alt && BINFB_FMARCP)
                pair_packed_gs_fr<ENERGY, true>(s.nx2[r], s.ny2[r], s.nz2[r], xj2[h], yj2[h], zj2[h], y2, A2, B2,
                                                s.g[r][0], s.g[r][1], s.g[r][2], fx2[h], fy2[h], fz2[h], c2);
            else if (alt)
                pair_packed_gs_sr<ENERGY>(s.nx2[r], s.ny2[r], s.nz2[r], xj2[h], yj2[h], zj2[h], y2, A2, B2,
                                          s.g[r][0], s.g[r][1], s.g[r][2], fx2[h], fy2[h], fz2[h], c2);
            else
                pair_packed_gs<ENERGY, EV, true, ALG>(s.nx2[r], s.ny2[r], s.nz2[r], xj2[h], yj2[h], zj2[h], y2, A2, B2,
                                                      s.g[r][0], s.g[r][1], s.g[r][2], fx2[h], fy2[h], fz2[h], c2,
                                                      s.dev, s.cev, &ev2);
        }
#else
        pair_packed_gs<ENERGY, EV, true, ALG>(s.nx2[r], s.ny2[r], s.nz2[r], xj2[0], yj2[0], zj2[0], mk2(yv[r].x, yv[r].y),
                                   A2, B2, s.g[r][0], s.g[r][1], s.g[r][2], fx2[0], fy2[0], fz2[0], c2, s.dev,
                                   s.cev, &ev2);
        pair_packed_gs<ENERGY, EV, true, ALG>(s.nx2[r], s.ny2[r], s.nz2[r], xj2[1], yj2[1], zj2[1], mk2(yv[r].z, yv[r].w),
                                   A2, B2, s.g[r][0], s.g[r][1], s.g[r][2], fx2[1], fy2[1], fz2[1], c2, s.dev,
                                   s.cev, &ev2);
#endif
    }
#if BINFB_PREFETCH == 2
    {   // positions are read-only during the sweep: the next step's partner may be loaded across the
        // __syncwarp that orders the force updates; issued here, in the tail of the step, the 12 registers
        // are only live across the step boundary
        const uint32_t pn = s.wrap == 1 ? s.pwrap : pa + 48u;
        s.nxt[0] = lds4<0>(pn), s.nxt[1] = lds4<16>(pn), s.nxt[2] = lds4<32>(pn);
    }
#endif
    if (active) {
        sts4<0>(fa, fx2[0].x, fx2[0].y, fx2[1].x, fx2[1].y);
        sts4<16>(fa, fy2[0].x, fy2[0].y, fy2[1].x, fy2[1].y);
        sts4<32>(fa, fz2[0].x, fz2[0].y, fz2[1].x, fz2[1].y);
        if (ENERGY) s.chi2 += (double)(c2.x + c2.y);
        if (ENERGY && EV) s.ev += (double)(ev2.x + ev2.y);
    }
}

// the special steps: k == 0 (the 6 pairs inside the lane's own quad), k == KS with an even quad
// count (only the lower half of the quads owns the (q, q + Q/2) block), k > KS (padding: nothing)
template <bool ENERGY, bool EV, bool ALG>
__device__ __forceinline__ void step_special(SweepRegs &s, uint32_t frc_off, uint32_t yaddr, float A, float B,
                                             int KS, bool upper_half) {
    const int k = s.k;
    if (k > KS) return;
    float chi = 0.f, evs = 0.f;
    if (k == 0) {
        float yv[4][4];
        unpack4(lds4<0>(yaddr), yv[0]), unpack4(lds4<512>(yaddr), yv[1]), unpack4(lds4<1024>(yaddr), yv[2]);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = r + 1; c < 4; ++c) {
                float tx = 0.f, ty = 0.f, tz = 0.f;
                pair_scalar<ENERGY, EV, true, ALG>(s.nx2[r].x, s.ny2[r].x, s.nz2[r].x, -s.nx2[c].x, -s.ny2[c].x, -s.nz2[c].x,
                                        yv[r][c], A, B, s.g[r][0], s.g[r][1], s.g[r][2], tx, ty,
                                        tz, chi, s.dev, s.cev, &evs);
                s.g[c][0] -= tx, s.g[c][1] -= ty, s.g[c][2] -= tz;
            }
    } else if (!upper_half) {
        const uint32_t pa = s.paddr, fa = s.paddr + frc_off;
        float xj[4], yj[4], zj[4], fjx[4], fjy[4], fjz[4], yv[4][4];
        unpack4(lds4<0>(pa), xj), unpack4(lds4<16>(pa), yj), unpack4(lds4<32>(pa), zj);
        unpack4(lds4<0>(fa), fjx), unpack4(lds4<16>(fa), fjy), unpack4(lds4<32>(fa), fjz);
        unpack4(lds4<0>(yaddr), yv[0]), unpack4(lds4<512>(yaddr), yv[1]);
        unpack4(lds4<1024>(yaddr), yv[2]), unpack4(lds4<1536>(yaddr), yv[3]);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                pair_scalar<ENERGY, EV, true, ALG>(s.nx2[r].x, s.ny2[r].x, s.nz2[r].x, xj[c], yj[c], zj[c], yv[r][c], A, B,
                                        s.g[r][0], s.g[r][1], s.g[r][2], fjx[c], fjy[c], fjz[c],
                                        chi, s.dev, s.cev, &evs);
        sts4<0>(fa, fjx[0], fjx[1], fjx[2], fjx[3]);
        sts4<16>(fa, fjy[0], fjy[1], fjy[2], fjy[3]);
        sts4<32>(fa, fjz[0], fjz[1], fjz[2], fjz[3]);
    }
    if (ENERGY) s.chi2 += (double)chi;
    if (ENERGY && EV) s.ev += (double)evs;
}

template <bool ENERGY, int R, int SPR, bool LOCKSTEP, int NS, bool EV, bool ALG>
__device__ __forceinline__ double chrom_sweep(const ChromDev &cd, const ChainSmem &sm, const Ring &ring,
                                              uint32_t &stage_idx_io, bool chain_valid, int lane, int role,
                                              int bar_id, int rb0, float cev, double &ev_out) {
    constexpr int STAGE_BYTES = R * SPR * STEP_BYTES;
    float4 *pos4 = reinterpret_cast<float4 *>(sm.pos), *frc4 = reinterpret_cast<float4 *>(sm.frc);
    const uint32_t pos_base = smem_u32(sm.pos);
    const uint32_t frc_off = pin_reg((uint32_t)cd.n_pad * 12u);
    // scaled positions: the "A" slot of the pair block carries the scaled softening, the "B" slot -2^B
    const float A = cd.softS, B = cd.nC;
    const float2 A2 = mk2(A, A), B2 = mk2(B, B);
    const int Q = cd.Q, KS = cd.KS, Lr = cd.Lr, halfQ = cd.Q >> 1;
    const bool q_even = cd.q_even != 0;
    const int k_fast = q_even ? KS - 1 : KS;  // offsets 1..k_fast need no special handling
    const int n_sg = Lr / SPR;                 // stages per row block
    const int k0 = role * Lr;
#if BINFB_PEEL
    // stages [sg_lo, sg_hi) of this role hold regular steps only
    const int sg_lo = k0 == 0 ? 1 : 0;
    int sg_hi = (k_fast - k0 + 1) / SPR;
    sg_hi = sg_hi < sg_lo ? sg_lo : (sg_hi > n_sg ? n_sg : sg_hi);
#endif
    const uint32_t stage_base = stage_idx_io;
    uint32_t stage_idx = stage_base;
    const uint32_t ylane = pin_reg(ring.ystage + (uint32_t)role * STEP_BYTES + (uint32_t)lane * 16u);
    SweepRegs s;
    uint32_t ticket = 0xffffffffu;  // BINFB_DEFER: result of the previous stage's release (none yet)
    s.chi2 = 0.0, s.ev = 0.0;
    s.dev = cd.ev_d * cd.S, s.cev = cev * (cd.invS * cd.invS * cd.invS);  // scaled units (pair_block.cuh)

#if BINFB_LOCKFLAGS
    // LOCKSTEP without the per-step chain barrier.  Role r at step s works on the 32 partner quads starting at
    // 32 rb + r Lr + s; two roles can only meet on a quad if they are neighbours (Lr >= 32) and the LOWER role is
    // ahead of the upper one by at least Lr - 31 steps.  So it is enough that role r does not start a step before
    // role r + 1 has completed (steps completed by r) - (Lr - 32) steps: every role but the first publishes its
    // step count (release) after the step's stores, every role but the last checks its upper neighbour's (acquire)
    // before a step.  The upper roles never wait for the lower ones, so there is no cycle with the shared stage ring.
    // The counters alias the chain's reduction scratch, which is idle during the sweep.
    const uint32_t lk_flags = smem_u32(sm.red);
    uint32_t lk_done = 0;
    const int lk_margin = Lr - 32;
    if (LOCKSTEP) {
        if (lane == 0) st_release_cta_shared(lk_flags + (uint32_t)role * 4u, 0u);
        chain_bar(bar_id, R * 32);
    }
#endif
    for (int rbi = 0; rbi < cd.NRB; ++rbi) {
        int rb = rbi + rb0;
        if (rb >= cd.NRB) rb -= cd.NRB;
        const int a = rb * 32 + lane;
        const bool active = chain_valid && a < Q;
        const int aa = active ? a : 0;
        const bool upper_half = q_even && a >= halfQ;
        {
            float xi[4], yi[4], zi[4];
            unpack4(pos4[3 * aa], xi), unpack4(pos4[3 * aa + 1], yi), unpack4(pos4[3 * aa + 2], zi);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                s.nx2[r] = mk2(-xi[r], -xi[r]), s.ny2[r] = mk2(-yi[r], -yi[r]), s.nz2[r] = mk2(-zi[r], -zi[r]);
                s.g[r][0] = s.g[r][1] = s.g[r][2] = 0.f;
            }
        }
        s.k = k0;  // partner offset of slot 0
        {
            int b = aa + k0;
            if (b >= Q) b -= Q;
            s.paddr = pos_base + (uint32_t)b * 48u;
            s.pwrap = pos_base;
#if BINFB_WRAP2
            s.wrap = (int)(pos_base + (uint32_t)Q * 48u);  // end address
#else
            s.wrap = Q - b;
#endif
#if BINFB_PREFETCH
            s.nxt[0] = lds4<0>(s.paddr), s.nxt[1] = lds4<16>(s.paddr), s.nxt[2] = lds4<32>(s.paddr);
#endif
        }
        // stages [sg_begin, sg_end) of this row block; GENERIC: a step may be one of the special ones
        auto run_stages = [&](auto generic_tag, int sg_begin, int sg_end) {
            constexpr bool GENERIC = decltype(generic_tag)::value;
#pragma unroll 1
            for (int sg = sg_begin; sg < sg_end; ++sg) {
                const uint32_t slot = ring_slot<NS>(stage_idx);
                {  // (probing the barrier one step early does not pay: the result is consumed at once)
                    const uint32_t fb = ring.full + slot * 8u, fp = (stage_idx / NS) & 1u;
                    while (!bar_try_wait(fb, fp)) {
                    }
                }
                const uint32_t ybase = ylane + slot * STAGE_BYTES;
#pragma unroll
                for (int u = 0; u < SPR; ++u) {
#if BINFB_LOCKFLAGS
                    if (LOCKSTEP && role < R - 1) {
                        const int need = (int)lk_done - lk_margin;
                        if (need > 0)
                            while ((int)ld_acquire_cta_shared(lk_flags + (uint32_t)(role + 1) * 4u) < need) {
                            }
                    }
#endif
                    if (!GENERIC || (unsigned)(s.k - 1) < (unsigned)k_fast)
                        step_fast<ENERGY, EV, ALG>(s, frc_off, ybase + u * R * STEP_BYTES, A2, B2, active);
                    else {
                        if (active) step_special<ENERGY, EV, ALG>(s, frc_off, ybase + u * R * STEP_BYTES, A, B, KS, upper_half);
#if BINFB_PREFETCH
                        const uint32_t pn = s.wrap == 1 ? s.pwrap : s.paddr + 48u;
                        s.nxt[0] = lds4<0>(pn), s.nxt[1] = lds4<16>(pn), s.nxt[2] = lds4<32>(pn);
#endif
                    }
                    if (GENERIC) ++s.k;
                    s.paddr += 48u;
#if BINFB_WRAP2
                    if (s.paddr == (uint32_t)s.wrap) s.paddr = s.pwrap;
#else
                    if (--s.wrap == 0) s.paddr -= (uint32_t)Q * 48u, s.wrap = Q;
#endif
                    // LOCKSTEP: all roles of the chain finish step s before any starts s + 1, so their
                    // partner offsets always differ by exactly a multiple of Lr >= 32 (see chrom_plan)
#if BINFB_LOCKFLAGS
                    __syncwarp();
                    if (LOCKSTEP) {
                        ++lk_done;
                        if (role > 0 && lane == 0) st_release_cta_shared(lk_flags + (uint32_t)role * 4u, lk_done);
                    }
#else
                    if (LOCKSTEP) chain_bar(bar_id, R * 32);
                    else __syncwarp();
#endif
                }
#if BINFB_DEFER
                if (elect_one()) {
                    ring_finish<STAGE_BYTES, NS>(ring, stage_idx - 1u, (int)(stage_idx - stage_base) - 1, ticket);
                    ticket = ring_arrive<NS>(ring, stage_idx);
                }
#else
                if (elect_one()) ring_release<STAGE_BYTES, NS>(ring, stage_idx, (int)(stage_idx - stage_base));
#endif
                ++stage_idx;
            }
            if (!GENERIC) s.k += (sg_end - sg_begin) * SPR;
        };
#if BINFB_PEEL
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            // part 0: generic [0, sg_lo) then regular [sg_lo, sg_hi); part 1: generic [sg_hi, n_sg)
            run_stages(std::true_type{}, part == 0 ? 0 : sg_hi, part == 0 ? sg_lo : n_sg);
            if (part == 0) run_stages(std::false_type{}, sg_lo, sg_hi);
        }
#else
        run_stages(std::true_type{}, 0, n_sg);
#endif
        // ---- end of the row block: fold the register-resident accumulators of the own quad into
        //      shared memory (f -= G).  All R roles hold sums for the SAME 32 quads.  With scratch space
        //      (many roles per chain: n = 5000 runs 16) every role parks its 12 sums per lane, and the chain's
        //      threads then add up the R contributions of one (quad, component) each: 3 barriers per row
        //      block instead of R + 1.  Without it the roles take turns. --------------------------------
        if (R >= 4 && sm.flush != nullptr) {  // (R is a template parameter)
            const int F = sm.flush_roles;
            chain_bar(bar_id, R * 32);  // no partner update of this row block is in flight any more
#pragma unroll 1
            for (int round = 0; round * F < R; ++round) {
                if (round > 0) chain_bar(bar_id, R * 32);  // the previous round has been read
                if (role / F == round) {
                    float *mine = sm.flush + (size_t)(role % F) * (12 * 32) + lane;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) mine[(c * 4 + r) * 32] = active ? s.g[r][c] : 0.f;
                }
                chain_bar(bar_id, R * 32);
                for (int v = role * 32 + lane; v < 12 * 32; v += R * 32) {
                    const int ql = v & 31, idx = v >> 5;  // idx = comp * 4 + r: the offset inside a quad's 12 floats
                    float t = 0.f;
                    for (int rr = 0; rr < F; ++rr) t += sm.flush[(size_t)rr * (12 * 32) + idx * 32 + ql];
                    const int aq = rb * 32 + ql;
                    if (chain_valid && aq < Q) sm.frc[aq * 12 + idx] -= t;
                }
            }
        } else {
#pragma unroll 1
            for (int rr = 0; rr < R; ++rr) {
                if (R > 1) chain_bar(bar_id, R * 32);
                if (rr == role && active) {
                    float4 v = frc4[3 * a];
                    v.x -= s.g[0][0], v.y -= s.g[1][0];
                    v.z -= s.g[2][0], v.w -= s.g[3][0];
                    frc4[3 * a] = v;
                    v = frc4[3 * a + 1];
                    v.x -= s.g[0][1], v.y -= s.g[1][1];
                    v.z -= s.g[2][1], v.w -= s.g[3][1];
                    frc4[3 * a + 1] = v;
                    v = frc4[3 * a + 2];
                    v.x -= s.g[0][2], v.y -= s.g[1][2];
                    v.z -= s.g[2][2], v.w -= s.g[3][2];
                    frc4[3 * a + 2] = v;
                }
            }
        }
        if (R > 1) chain_bar(bar_id, R * 32);
        else __syncwarp();
    }
    stage_idx_io = stage_idx;
    ev_out = s.ev;
    return s.chi2;
}

__device__ __forceinline__ double warp_sum(double v) { return group_allreduce_sum<32>(v); }

// sum over all lanes of all R warps of a chain; every thread gets the same value
__device__ __forceinline__ double chain_sum(double v, const ChainSmem &sm, int lane, int role, int R,
                                            int bar_id) {
    v = warp_sum(v);
    if (R == 1) return v;
    if (lane == 0) sm.red[role] = v;
    chain_bar(bar_id, R * 32);
    double t = 0.0;
    for (int r = 0; r < R; ++r) t += sm.red[r];
    chain_bar(bar_id, R * 32);
    return t;
}

struct ChromCall {
    int mode;
    HmcArgs h;   // HMC mode
    GradArgs g;  // GRAD mode
    int W;       // chains per CTA (W * R warps)
    int n_groups, total_items;
    int set_groups;  // groups per set: a set of chain groups runs ALL its passes before the next set starts, so
                     // that the q / p hand-off between consecutive passes of a group stays in L2 (see chrom_launch)
};

// work item -> (chain group o, pass index seq).  Items are ordered set by set; inside a set pass-major, so
// consecutive passes of one group are set_groups items apart.  Only the last set may be smaller.
__device__ __forceinline__ void chrom_item(const ChromCall &call, int it, int n_seq, int &o, int &seq) {
    const int per_set = call.set_groups * n_seq;
    const int set = it / per_set, r = it - set * per_set;
    const int g0 = set * call.set_groups;
    const int gs = min(call.set_groups, call.n_groups - g0);
    seq = r / gs;
    o = g0 + (r - seq * gs);
}

template <int R, int SPR, bool LOCKSTEP, int NS, bool EV, bool ALG>
__global__ void __launch_bounds__(512, 1) chrom_kernel(ChromDev cd, ChromCall call) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int STAGE_BYTES = R * SPR * STEP_BYTES;
    const int W = call.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // (warp w runs on SMSP w % 4, so with R = 2 an SMSP hosts one role only; mixing the roles per SMSP
    // measured 2 % slower)
    const int chain_local = warp / R, role = warp % R;
    const int bar_id = 1 + chain_local;
    // ---- shared memory carve-up: [stages][full barriers, release counters, item][W x chain]
    Ring ring;
    ring.ystage = smem_u32(smem_raw);
    ring.full = ring.ystage + NS * STAGE_BYTES;
    ring.src = reinterpret_cast<const unsigned char *>(cd.ystream);
    ring.n_stage_pass = cd.S_pad / SPR;
    ring.n_warps = W * R;
    unsigned char *ctl = smem_raw + (size_t)NS * STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(ctl);  // full[NS], then the release counters [NS]
    uint32_t *cnts = reinterpret_cast<uint32_t *>(ctl + NS * 8);
    int *s_item = reinterpret_cast<int *>(ctl + NS * 12);
    unsigned char *chains = ctl + 128;
    const size_t per_chain = (size_t)6 * cd.n_pad * sizeof(float) + CHAIN_SCRATCH_BYTES;
    ChainSmem sm;
    {
        unsigned char *b0 = chains + per_chain * chain_local;
        sm.red = reinterpret_cast<double *>(b0);
        sm.pos_bar = smem_u32(b0 + 128);
        sm.pos = reinterpret_cast<float *>(b0 + CHAIN_SCRATCH_BYTES);
        sm.frc = sm.pos + 3 * cd.n_pad;
        sm.flush = nullptr, sm.flush_roles = 0;
        if constexpr (R >= 4) {
            // the launcher appends W * R * FLUSH_ROLE_BYTES behind the chains where that fits (chrom_launch);
            // the kernel sees it in the size of its dynamic shared memory (no extra kernel parameter: the
            // kernels with fewer roles stay exactly as they were)
            uint32_t dyn;
            asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
            const size_t need = (size_t)NS * STAGE_BYTES + 128 + per_chain * W;
            int F = (size_t)dyn > need ? (int)(((size_t)dyn - need) / ((size_t)W * FLUSH_ROLE_BYTES)) : 0;
            if (F > R) F = R;
            while (F & (F - 1)) F &= F - 1;  // round down to a power of two
            sm.flush_roles = F;
            if (F >= 2)
                sm.flush = reinterpret_cast<float *>(chains + per_chain * W + (size_t)chain_local * F * FLUSH_ROLE_BYTES);
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) mbar_init(&bars[i], 1), cnts[i] = 0u;
        for (int w = 0; w < W; ++w) mbar_init(reinterpret_cast<uint64_t *>(chains + per_chain * w + 128), 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int D = 3 * cd.n;
    const int passes = call.mode == CHROM_MODE_HMC ? call.h.L + 1 : 1;
    const int n_seq = passes * (call.mode == CHROM_MODE_HMC ? call.h.n_traj : 1);  // passes of a group in this launch
    const int cthreads = R * 32, ctid = role * 32 + lane;  // threads of this chain
    uint32_t stage_idx = 0;  // running stage counter, identical in every warp
    uint32_t pos_copies = 0; // bulk copies of positions this chain slot has received (mbarrier phase)

    for (;;) {
        // ---- claim a work item and wait for the previous pass of its chain group ----------
        if (threadIdx.x == 0) {
            const int it = atomicAdd(cd.counter, 1);
            if (it < call.total_items) {
                // every warp has released every stage of the previous item (block barrier below):
                // start streaming the contacts of this pass right away.  Chain group o walks the row
                // blocks starting at block o mod NRB: concurrently running CTAs then read different
                // parts of the contact stream instead of hammering the same L2 lines in lockstep (the
                // order is a function of the group, so results do not depend on the CTA schedule).
                int o, need;  // need = passes of this group that must be done
                chrom_item(call, it, n_seq, o, need);
                ring.rot = BINFB_ROTATE ? (o % cd.NRB) * (cd.Lr / SPR) : 0;
                for (int i = 0; i < NS && i < ring.n_stage_pass; ++i)
                    ring_issue<STAGE_BYTES, NS>(ring, stage_idx + (uint32_t)i, i);
                if (need > 0)
                    while (ld_acquire(cd.pass_done + o) < need) __nanosleep(100);
                else {
                    // first pass of the group: under a pipelined host call its positions may still be in flight
                    const int *gate = cd.pass_done + call.n_groups;
                    int g = ld_acquire_sys(gate);
                    if (g != 0) {
                        const int Cn = call.mode == CHROM_MODE_HMC ? call.h.C : call.g.C;
                        int last = (o + 1) * call.W;
                        if (last > Cn) last = Cn;
                        const unsigned long long t0 = global_timer_ns();
                        while (g - 1 < last && global_timer_ns() - t0 < 10000000000ull) {  // (10 s: never hang the GPU)
                            __nanosleep(200);
                            g = ld_acquire_sys(gate);
                        }
                    }
                }
            }
            *s_item = it;
        }
        __syncthreads();
        const int item = *s_item;
        if (item >= call.total_items) break;
        int o, seq;  // seq = tr * passes + k
        chrom_item(call, item, n_seq, o, seq);
        ring.rot = BINFB_ROTATE ? (o % cd.NRB) * (cd.Lr / SPR) : 0;
        const int tr = seq / passes, k = seq % passes;

        {
            const int c = o * W + chain_local;
            const int C = call.mode == CHROM_MODE_HMC ? call.h.C : call.g.C;
            const bool valid = c < C;
            const bool hmc = call.mode == CHROM_MODE_HMC;
            const HmcArgs &h = call.h;
            const bool energy = !hmc || k == 0 || k == h.L;
            float eps_c = 0.f, beta_c = 1.f;
            double kin0 = 0.0;
            const size_t off = (size_t)(valid ? c : 0) * D;
            if (valid) {
                // ---- phase A: positions (+ drift) into shared memory ---------------------
                const float *src = hmc ? (k == 0 ? h.q : cd.qw) : call.g.q;
                if (hmc) {
                    eps_c = __ldcg(h.eps + c);
                    beta_c = h.beta ? h.beta[c] : 1.f;
                } else {
                    beta_c = call.g.beta ? call.g.beta[c] : 1.f;
                }
                float kin = 0.f;
                if (hmc && k == 0) {
                    for (int e = ctid; e < D; e += cthreads) {
                        const float v = __ldcg(src + off + e);
                        const float pv = h.p0 ? h.p0[off + e]
                                              : rng_normal(h.seed, h.chain_base + c,
                                                           h.draw + (uint64_t)tr, (uint32_t)e);
                        __stcg(cd.pw + off + e, pv);
                        kin = fmaf(pv, pv, kin);
                        const int bead = e / 3, comp = e - 3 * bead;
                        sm.pos[qidx(bead, comp)] = v * cd.S;
                    }
                } else if (hmc) {
                    // passes k > 0: the previous pass left the drifted positions q + eps p (hmc.py:119,122)
                    // in the working copy, already in the shared-memory layout: one bulk async copy
                    // (TMA engine) brings them in while the forces are zeroed
                    if (ctid == 0) {
                        const uint32_t bytes = (uint32_t)cd.n_pad * 12u;
                        fence_proxy_async_global();  // reader side of the generic -> async edge on qw
                        bar_expect_tx(sm.pos_bar, bytes);
                        bulk_g2s(smem_u32(sm.pos), cd.qw + (size_t)c * 3 * cd.n_pad, bytes, sm.pos_bar);
                    }
                } else {
                    constexpr int U = 8;
                    for (int e0 = ctid; e0 < D; e0 += cthreads * U) {
                        float v[U];
#pragma unroll
                        for (int uu = 0; uu < U; ++uu) {
                            const int e = e0 + uu * cthreads;
                            v[uu] = e < D ? __ldcg(src + off + e) : 0.f;
                        }
#pragma unroll
                        for (int uu = 0; uu < U; ++uu) {
                            const int e = e0 + uu * cthreads;
                            if (e < D) {
                                const int bead = e / 3, comp = e - 3 * bead;
                                sm.pos[qidx(bead, comp)] = v[uu] * cd.S;
                            }
                        }
                    }
                }
                for (int i = ctid; i < 3 * cd.n_pad; i += cthreads) sm.frc[i] = 0.f;
                if (hmc && k > 0) {
                    while (!bar_try_wait(sm.pos_bar, pos_copies & 1u)) {
                    }
                    ++pos_copies;
                }
                for (int i = cd.n + ctid; i < cd.n_pad; i += cthreads) {
                    // padding beads: far away from everything (scaled distance = exponent > 128) => contact 0, force 0
                    sm.pos[qidx(i, 0)] = 1.0e4f * (float)(1 + i - cd.n);
                    sm.pos[qidx(i, 1)] = 3.0e4f, sm.pos[qidx(i, 2)] = -2.0e4f;
                }
                if (hmc && k == 0) kin0 = chain_sum((double)kin, sm, lane, role, R, bar_id);
            }
            if (R > 1) chain_bar(bar_id, cthreads);
            else __syncwarp();
            // ---- the precision of this pass.  With the excluded-volume term it must be known BEFORE the
            //      sweep (the EV force is folded into the pair coefficients relative to alpha beta tau),
            //      so a fused precision-first Gibbs update draws tau here from chi^2 of the current state,
            //      which the host launcher (first trajectory) or the previous trajectory left in
            //      chi2_state; without EV it is drawn after pass 0 from the chi^2 that pass computes.
            float tau_pre = 1.f, cev = 0.f;
            if (EV && valid) {
                if (!hmc) tau_pre = call.g.tau[c];
                else if (k > 0) tau_pre = __ldcg(cd.tau_w + c);
                else {
                    tau_pre = __ldcg(h.tau + c);
                    if (h.gibbs_mode == BINFB_GIBBS_TAU_FIRST) {
                        const double shape = 0.5 * (double)beta_c * cd.M + h.gamma_shape - 1.0;
                        const double rate = 0.5 * (double)beta_c * __ldcg(cd.chi2_state + c) + h.gamma_rate;
                        const double gd = h.gamma_draws ? h.gamma_draws[c]
                                                        : rng_gamma(h.seed, h.chain_base + c, h.draw + (uint64_t)tr, shape);
                        tau_pre = (float)(gd / rate);
                    }
                }
                cev = 4.0f * cd.ev_k / (cd.alpha * beta_c * tau_pre);
            }
            // ---- phase B: pair sweep -----------------------------------------------------
            const int rb0 = BINFB_ROTATE ? o % cd.NRB : 0;
            double ev_sum = 0.0;
            double chi2 = energy ? chrom_sweep<true, R, SPR, LOCKSTEP, NS, EV, ALG>(cd, sm, ring, stage_idx, valid, lane, role,
                                                                              bar_id, rb0, cev, ev_sum)
                                 : chrom_sweep<false, R, SPR, LOCKSTEP, NS, EV, ALG>(cd, sm, ring, stage_idx, valid, lane, role,
                                                                               bar_id, rb0, cev, ev_sum);
            if (R > 1) chain_bar(bar_id, cthreads);
            else __syncwarp();
            if (valid) {
                // ---- phase C: forces, kick, energies ----------------------------------------
                if (energy) chi2 = chain_sum(chi2, sm, lane, role, R, bar_id);
                float tau_c;
                if (EV) {
                    tau_c = tau_pre;
                    if (hmc && k == 0 && ctid == 0) {
                        __stcg(cd.tau_w + c, tau_c);
                        if (h.gibbs_mode == BINFB_GIBBS_TAU_FIRST) __stcg(h.tau + c, tau_c);
                    }
                } else if (hmc) {
                    if (k == 0) {
                        tau_c = __ldcg(h.tau + c);
                        if (h.gibbs_mode == BINFB_GIBBS_TAU_FIRST) {
                            const double shape = 0.5 * (double)beta_c * cd.M + h.gamma_shape - 1.0;
                            const double rate = 0.5 * (double)beta_c * chi2 + h.gamma_rate;
                            const double gd = h.gamma_draws
                                                  ? h.gamma_draws[c]
                                                  : rng_gamma(h.seed, h.chain_base + c,
                                                              h.draw + (uint64_t)tr, shape);
                            tau_c = (float)(gd / rate);
                            if (ctid == 0) __stcg(h.tau + c, tau_c);
                        }
                        if (ctid == 0) __stcg(cd.tau_w + c, tau_c);
                    } else {
                        tau_c = __ldcg(cd.tau_w + c);
                    }
                } else {
                    tau_c = call.g.tau[c];
                }
                const float scale = -cd.alpha * beta_c * tau_c;
                const float kick = hmc ? ((k == 0 || k == h.L) ? 0.5f * eps_c : eps_c) : 0.f;
                float e_prior = 0.f, kin = 0.f;
                constexpr int UC = 4;
                for (int i0 = ctid; i0 < cd.n; i0 += cthreads * UC) {
                    float pold[UC][3];
                    if (hmc) {
#pragma unroll
                        for (int uu = 0; uu < UC; ++uu) {
                            const int i = i0 + uu * cthreads;
                            const float *pp = cd.pw + off + 3 * (i < cd.n ? i : 0);
                            pold[uu][0] = __ldcg(pp), pold[uu][1] = __ldcg(pp + 1), pold[uu][2] = __ldcg(pp + 2);
                        }
                    }
#pragma unroll
                    for (int uu = 0; uu < UC; ++uu) {
                        const int i = i0 + uu * cthreads;
                        if (i >= cd.n) continue;
                        const int ix = qidx(i, 0);
                        // x, y, z and the bond vectors are in scaled units (S x); a bond force
                        // k_bb (d - l0) b / d is the same in both units up to d = d' / S
                        const float x = sm.pos[ix], y = sm.pos[ix + 4], z = sm.pos[ix + 8];
                        float gx = scale * sm.frc[ix], gy = scale * sm.frc[ix + 4], gz = scale * sm.frc[ix + 8];
                        if (i > 0) {
                            const int im = qidx(i - 1, 0);
                            const float bx = x - sm.pos[im], by = y - sm.pos[im + 4], bz = z - sm.pos[im + 8];
                            const float r2 = fmaf(bz, bz, fmaf(by, by, fmaf(bx, bx, cd.softS)));
                            const float inv = rsqrtf(r2), d = r2 * inv;
                            const float cc = cd.k_bb * fmaf(d, cd.invS, -cd.l0) * inv;
                            gx = fmaf(cc, bx, gx), gy = fmaf(cc, by, gy), gz = fmaf(cc, bz, gz);
                        }
                        if (i < cd.n - 1) {
                            const int ip = qidx(i + 1, 0);
                            const float bx = sm.pos[ip] - x, by = sm.pos[ip + 4] - y, bz = sm.pos[ip + 8] - z;
                            const float r2 = fmaf(bz, bz, fmaf(by, by, fmaf(bx, bx, cd.softS)));
                            const float inv = rsqrtf(r2), d = r2 * inv;
                            const float dl = fmaf(d, cd.invS, -cd.l0);
                            const float cc = cd.k_bb * dl * inv;
                            gx = fmaf(-cc, bx, gx), gy = fmaf(-cc, by, gy), gz = fmaf(-cc, bz, gz);
                            e_prior = fmaf(0.5f * cd.k_bb * dl, dl, e_prior);
                        }
                        if (cd.inv_s2 > 0.f) {
                            const float w = cd.inv_s2 * cd.invS;
                            gx = fmaf(w, x, gx), gy = fmaf(w, y, gy), gz = fmaf(w, z, gz);
                            e_prior = fmaf(0.5f * w * cd.invS, fmaf(x, x, fmaf(y, y, z * z)), e_prior);
                        }
                        if (hmc) {
                            float *pp = cd.pw + off + 3 * i;
                            const float px = fmaf(-kick, gx, pold[uu][0]), py = fmaf(-kick, gy, pold[uu][1]),
                                        pz = fmaf(-kick, gz, pold[uu][2]);
                            __stcg(pp, px), __stcg(pp + 1, py), __stcg(pp + 2, pz);
                            kin = fmaf(px, px, fmaf(py, py, fmaf(pz, pz, kin)));
                            if (k < h.L) {
                                // the next pass's positions, drift included, in the shared-memory layout
                                float *qq = cd.qw + (size_t)c * 3 * cd.n_pad + ix;
                                const float es = eps_c * cd.S;  // the drift in scaled units
                                __stcg(qq, fmaf(es, px, x)), __stcg(qq + 4, fmaf(es, py, y));
                                __stcg(qq + 8, fmaf(es, pz, z));
                            }
                        } else if (call.g.grad) {
                            float *gg = call.g.grad + off + 3 * i;
                            gg[0] = gx, gg[1] = gy, gg[2] = gz;
                        }
                    }
                }
                if (energy) {
                    double ep = chain_sum((double)e_prior, sm, lane, role, R, bar_id);
                    if (EV) ep += (double)cd.ev_k * ((double)cd.invS * cd.invS * cd.invS * cd.invS) *
                                  chain_sum(ev_sum, sm, lane, role, R, bar_id);
                    const double t = (double)tau_c, lt = log(t);
                    const double ga = hmc ? h.gamma_shape : call.g.gamma_shape;
                    const double gb = hmc ? h.gamma_rate : call.g.gamma_rate;
                    const double U = (double)beta_c * (0.5 * t * chi2 - 0.5 * cd.M * lt) + ep -
                                     ((ga - 1.0) * lt - gb * t);
                    if (!hmc) {
                        if (ctid == 0) {
                            if (call.g.logp) call.g.logp[c] = -U;
                            if (call.g.chi2) call.g.chi2[c] = chi2;
                        }
                    } else if (k == 0) {
                        if (ctid == 0) {
                            __stcg(cd.h0 + c, U + 0.5 * kin0);
                            __stcg(cd.chi2_0 + c, chi2);
                        }
                    }
                    if (hmc && k == h.L) {
                        const double h1 = U + 0.5 * chain_sum((double)kin, sm, lane, role, R, bar_id);
                        const double h0 = __ldcg(cd.h0 + c);
                        const double dh = h1 - h0;
                        const uint64_t draw = h.draw + (uint64_t)tr;
                        float uu;
                        if (h.u) uu = h.u[c];
                        else {
                            const u32x4 r = philox4x32_10(h.seed ^ (draw >> 32) * 0x9E3779B97F4A7C15ull,
                                                          h.chain_base + c, (uint32_t)draw,
                                                          (uint32_t)RNG_ACCEPT << 24);
                            uu = u32_to_unit_open0(r.x);
                        }
                        // Metropolis (hmc.py:151); NaN energies reject
                        const bool acc = (dh == dh) && ((double)uu < exp(fmin(709.0, fmax(-308.0, -dh))));
                        const bool last = tr == h.n_traj - 1;
                        if (last && (h.q_end || h.p_end)) {
                            for (int e = ctid; e < D; e += cthreads) {
                                const int bead = e / 3, comp = e - 3 * bead;
                                if (h.q_end) h.q_end[off + e] = sm.pos[qidx(bead, comp)] * cd.invS;
                                if (h.p_end) h.p_end[off + e] = __ldcg(cd.pw + off + e);
                            }
                        }
                        if (acc) {
                            for (int e = ctid; e < D; e += cthreads) {
                                const int bead = e / 3, comp = e - 3 * bead;
                                __stcg(h.q + off + e, sm.pos[qidx(bead, comp)] * cd.invS);
                            }
                        }
                        const double chi2_cur = acc ? chi2 : __ldcg(cd.chi2_0 + c);
                        if (ctid == 0) {
                            __stcg(cd.chi2_state + c, chi2_cur);
                            if (tr < h.n_adapt)  // hmc.py:188-191
                                __stcg(h.eps + c, eps_c * (acc ? h.adapt_up : h.adapt_down));
                            if (h.accepted) h.accepted[c] = acc ? 1 : 0;
                            if (h.e_before) h.e_before[c] = h0;
                            if (h.e_after) h.e_after[c] = h1;
                            if (h.n_accepted) __stcg(h.n_accepted + c, __ldcg(h.n_accepted + c) + (acc ? 1 : 0));
                            if (h.stats) {
                                atomicAdd(h.stats + 0, acc ? 1.0 : 0.0);
                                atomicAdd(h.stats + 1, 1.0);
                                atomicAdd(h.stats + 2, (double)eps_c);
                                atomicAdd(h.stats + 3, (dh == dh) ? exp(fmin(0.0, -dh)) : 0.0);
                            }
                        }
                        if (h.gibbs_mode == BINFB_GIBBS_TAU_LAST) {
                            const double shape = 0.5 * (double)beta_c * cd.M + h.gamma_shape - 1.0;
                            const double rate = 0.5 * (double)beta_c * chi2_cur + h.gamma_rate;
                            const double gd = h.gamma_draws
                                                  ? h.gamma_draws[c]
                                                  : rng_gamma(h.seed, h.chain_base + c, draw, shape);
                            if (ctid == 0) __stcg(h.tau + c, (float)(gd / rate));
                        }
                    }
                }
            }
        }
        // ---- publish the pass ---------------------------------------------------------------
        fence_proxy_async_global();  // writer side: this thread's __stcg stores to qw vs the next pass's bulk copy
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            st_release(cd.pass_done + o, seq + 1);
            int *gate = cd.pass_done + call.n_groups;
            const int gpc = gate[1];  // (written before the launch)
            if (gpc > 0 && seq + 1 == n_seq) {
                // last pass of this group under a pipelined host call: its results may be copied back
                __threadfence_system();
                atomicAdd_system(gate + 2 + o / gpc, 1);
            }
        }
    }
}

// mock contacts for all pairs (AbstractForwardModel.__call__ of the contact model)
__global__ void chrom_forward_kernel(int n, long long M, float A, float B, int algebraic, const float *q,
                                     float *mock) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    const int c = blockIdx.y;
    // invert numpy.triu_indices(n, 1) ordering
    const double nn = (double)n;
    long long i = (long long)(nn - 2.0 - floor(sqrt(-8.0 * (double)idx + 4.0 * nn * (nn - 1.0) - 7.0) / 2.0 - 0.5));
    long long row_start = i * n - i * (i + 1) / 2;
    while (i > 0 && idx < row_start) --i, row_start = i * n - i * (i + 1) / 2;
    while (idx >= row_start + (n - 1 - i)) ++i, row_start = i * n - i * (i + 1) / 2;
    const long long j = idx - row_start + i + 1;
    const float *x = q + (size_t)c * 3 * n;
    const float dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, CHROM_SOFT)));
    const float d = r2 * rsqrtf(r2);
    if (algebraic) {  // A = alpha, B = -alpha d_c: z = -(A d + B), mock = 1/2 (1 + z / sqrt(1 + z^2))
        const float zn = fmaf(d, A, B);
        mock[(size_t)c * M + idx] = 0.5f - 0.5f * zn * rsqrtf(fmaf(zn, zn, 1.0f));
    } else {
        mock[(size_t)c * M + idx] = 1.0f / (1.0f + exp2f(fmaf(d, A, B)));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline long long tri_index(long long n, long long i, long long j) {  // i < j
    return i * n - i * (i + 1) / 2 + (j - i - 1);
}

// stage size in warp-steps per role count (SPR = SS / R steps per role and stage); the ring has
// CHROM_NS stages.  Measured on B200 at n = 1000: 4-step stages x 4 slots beat 8 x 2.
static int chrom_stage_steps(int R) { return R <= BINFB_CHROM_SS ? BINFB_CHROM_SS : R; }
// 16 roles: 32 KiB stages, three of them where a chain leaves room (n <= 5576), else two
static int chrom_ring_depth(int R) { return R >= 16 ? 3 : CHROM_NS; }

// the plan for a fixed role count R; W = 0 if it cannot run.  Two roles of a chain may only touch the
// same partner quad if their offsets differ by <= 31.  Free-running roles drift by up to NS*spr - 1
// slots (the ring holds NS stages of spr slots per role), which needs Lr - drift >= 34; roles in
// LOCKSTEP (a chain barrier after every step) always differ by exactly a multiple of Lr, which needs
// Lr >= 32 only.
static ChromPlan chrom_plan_for(int n, int smem_optin, int R, bool allow_lockstep) {
    ChromPlan pl;
    pl.n_pad = (n + 3) / 4 * 4, pl.Q = pl.n_pad / 4, pl.KS = pl.Q / 2, pl.NRB = (pl.Q + 31) / 32;
    const size_t per_chain = (size_t)6 * pl.n_pad * sizeof(float) + CHAIN_SCRATCH_BYTES;
    pl.R = R;
    pl.SS = chrom_stage_steps(R), pl.NS = chrom_ring_depth(R);
    const int spr = pl.SS / R;  // slots per stage and role
    pl.Lr = ((pl.KS + 1 + R - 1) / R + spr - 1) / spr * spr;  // stages never straddle row blocks
    pl.S_pad = pl.NRB * pl.Lr;
    pl.fixed_smem = (size_t)pl.NS * pl.SS * STEP_BYTES + 128;
    pl.per_chain_smem = per_chain;
    const bool safe_free = R == 1 || pl.Lr - (pl.NS * spr - 1) >= 34;
    const bool safe_lock = pl.Lr >= 32;
    pl.lockstep = !safe_free && safe_lock && allow_lockstep;
    int W = (size_t)smem_optin > pl.fixed_smem ? (int)(((size_t)smem_optin - pl.fixed_smem) / per_chain) : 0;
    if (W > 16 / R) W = 16 / R;
    if (!safe_free && !pl.lockstep) W = 0;
    pl.W = W;
    pl.stream_floats = (long long)pl.S_pad * R * STEP_FLOAT4 * 4;
    if (W < 1 && pl.NS == 3) {  // no room for the third stage of the 16-role ring: two stages
        pl.NS = 2;
        pl.fixed_smem = (size_t)pl.NS * pl.SS * STEP_BYTES + 128;
        W = (size_t)smem_optin > pl.fixed_smem ? (int)(((size_t)smem_optin - pl.fixed_smem) / per_chain) : 0;
        if (W > 16 / R) W = 16 / R;
        if (!(R == 1 || pl.Lr - (pl.NS * spr - 1) >= 34) && !pl.lockstep) W = 0;
        pl.W = W;
    }
    return pl;
}

// Launch geometry: W chains per CTA, R warps ("roles") per chain, W*R <= 16 consumer warps.  More roles
// only pay off while the chains alone cannot fill 16 warps; free-running roles are preferred.
ChromPlan chrom_plan(int n, int smem_optin, int force_roles) {
    if (force_roles > 0) return chrom_plan_for(n, smem_optin, force_roles, true);
    ChromPlan best = chrom_plan_for(n, smem_optin, 1, false);
    for (int R = 2; R <= 16; R *= 2) {
        const ChromPlan pl = chrom_plan_for(n, smem_optin, R, false);
        const size_t per_chain = pl.per_chain_smem;
        if ((size_t)smem_optin < pl.fixed_smem + per_chain) break;
        const int wmax = (int)(((size_t)smem_optin - pl.fixed_smem) / per_chain);
        if (pl.W >= 1 && wmax * (R / 2) < 16) best = pl;
    }
    return best;
}

// Small batches (fewer chains than the primary plan needs to give every SM 16 warps): twice the roles
// per chain, in lockstep if necessary.  W = 0 if there is no such plan.
ChromPlan chrom_plan_small_batch(int n, int smem_optin, const ChromPlan &primary) {
    ChromPlan none;
    none.W = 0;
    const int R = primary.R * 2;
    if (R > 8) return none;
    const ChromPlan pl = chrom_plan_for(n, smem_optin, R, true);
    return pl.W >= 1 ? pl : none;
}

// Contact stream in consumption order: step t' = (rb*Lr + s)*R + role holds, for lane l and row r,
// the float4 y[4a+r][4b..4b+3] with a = 32 rb + l, b = (a + k) mod Q, k = role*Lr + s.
int chrom_build_stream(int n, const float *y_pairs, const ChromPlan &pl, float *out) {
    const int Q = pl.Q, KS = pl.KS;
    const bool q_even = (Q % 2) == 0;
    for (long long i = 0; i < pl.stream_floats; ++i) out[i] = 0.f;
    for (int rb = 0; rb < pl.NRB; ++rb)
        for (int s = 0; s < pl.Lr; ++s)
            for (int role = 0; role < pl.R; ++role) {
                const int k = role * pl.Lr + s;
                if (k > KS) continue;
                const long long t = ((long long)rb * pl.Lr + s) * pl.R + role;
                for (int lane = 0; lane < 32; ++lane) {
                    const int a = rb * 32 + lane;
                    if (a >= Q) continue;
                    int b = a + k;
                    if (b >= Q) b -= Q;
                    if (k > 0 && k == KS && q_even && a >= Q / 2) continue;
                    for (int r = 0; r < 4; ++r)
                        for (int c = 0; c < 4; ++c) {
                            const int i = 4 * a + r, j = 4 * b + c;
                            if (i >= n || j >= n) continue;
                            if (k == 0 && r >= c) continue;
                            const long long idx = i < j ? tri_index(n, i, j) : tri_index(n, j, i);
                            out[((t * 4 + r) * 32 + lane) * 4 + c] = y_pairs[idx];
                        }
                }
            }
    return BINFB_OK;
}

int chrom_reserve(ChromModel &m, int C) {
    const int D = 3 * m.n;
    if (C > m.ws_chains) {
        cudaFree(m.qw), cudaFree(m.pw), cudaFree(m.h0), cudaFree(m.chi2_0), cudaFree(m.chi2_state),
            cudaFree(m.tau_w);
        m.qw = m.pw = m.tau_w = nullptr;
        m.h0 = m.chi2_0 = m.chi2_state = nullptr;
        m.ws_chains = 0, m.chi2_chains = 0;
        BINFB_CUDA(cudaMalloc(&m.qw, (size_t)C * 3 * m.plan.n_pad * sizeof(float)));
        BINFB_CUDA(cudaMalloc(&m.pw, (size_t)C * D * sizeof(float)));
        BINFB_CUDA(cudaMalloc(&m.h0, (size_t)C * sizeof(double)));
        BINFB_CUDA(cudaMalloc(&m.chi2_0, (size_t)C * sizeof(double)));
        BINFB_CUDA(cudaMalloc(&m.chi2_state, (size_t)C * sizeof(double)));
        BINFB_CUDA(cudaMalloc(&m.tau_w, (size_t)C * sizeof(float)));
        m.ws_chains = C;
    }
    const int need = 1 + C + CHROM_GATE_WORDS;  // item counter + one pass counter per chain group (W >= 1) + gate
    if (need > m.sched_len) {
        cudaFree(m.sched);
        m.sched = nullptr;
        m.sched_len = 0;
        BINFB_CUDA(cudaMalloc(&m.sched, (size_t)need * sizeof(int)));
        m.sched_len = need;
    }
    return BINFB_OK;
}

static ChromDev chrom_dev(const ChromModel &m, const ChromPlan &pl, const float *ystream) {
    ChromDev d;
    d.n = m.n, d.n_pad = pl.n_pad, d.Q = pl.Q, d.KS = pl.KS, d.NRB = pl.NRB;
    d.q_even = (pl.Q % 2) == 0;
    d.R = pl.R, d.Lr = pl.Lr, d.SS = pl.SS, d.S_pad = pl.S_pad;
    d.ystream = reinterpret_cast<const float4 *>(ystream);
    const double log2e = 1.4426950408889634;
    if (m.flags & BINFB_FLAG_CONTACT_ALGEBRAIC) {
        // algebraic contact function: positions scaled by alpha, the "B" slot of the pair block carries -alpha d_c,
        // the force scale the -1/2 of the block's coefficient convention (pair_block.cuh, ALG)
        d.A = m.alpha, d.B = (float)(-(double)m.alpha * (double)m.d_c);
        d.S = d.A, d.invS = (float)(1.0 / (double)m.alpha);
        d.softS = (float)((double)d.A * (double)d.A * (double)CHROM_SOFT);
        d.nC = d.B;
        d.alpha = -0.5f * m.alpha;
    } else {
        d.A = (float)((double)m.alpha * log2e);
        d.B = (float)(-(double)m.alpha * (double)m.d_c * log2e);
        d.S = d.A, d.invS = (float)(1.0 / ((double)m.alpha * log2e));
        d.softS = (float)((double)d.A * (double)d.A * (double)CHROM_SOFT);
        d.nC = (float)(-exp2(-(double)m.alpha * (double)m.d_c * log2e));
        d.alpha = m.alpha;
    }
    d.k_bb = m.k_bb, d.l0 = m.l0, d.inv_s2 = m.inv_s2;
    d.ev_k = m.ev_k, d.ev_d = m.ev_d;
    d.M = (double)m.M;
    d.qw = m.qw, d.pw = m.pw, d.h0 = m.h0, d.chi2_0 = m.chi2_0, d.chi2_state = m.chi2_state;
    d.tau_w = m.tau_w;
    d.counter = m.sched, d.pass_done = m.sched + 1;
    return d;
}

// the launch shape chrom_launch will pick for C chains: plan (primary or small-batch) and chains per CTA
static const ChromPlan *chrom_pick_plan(const ChromModel &m, int C, int sm_count, const float **ystream, int *W_out) {
    auto chains_per_cta = [&](const ChromPlan &p) {
        int w = p.W;
        if (w >= 1 && (C + w - 1) / w < sm_count) w = (C + sm_count - 1) / sm_count;
        if (w > p.W) w = p.W;
        if (m.opt_warps > 0 && m.opt_warps < w) w = m.opt_warps;
        if (w > C) w = C;
        return w;
    };
    const ChromPlan *plp = &m.plan;
    *ystream = m.ystream;
    int W = chains_per_cta(m.plan);
    // a batch that leaves the SMs half empty with the primary plan runs with twice the warps per chain
    if (m.ystream_alt && W >= 1 && W * m.plan.R <= 8) {
        const int wa = chains_per_cta(m.plan_alt);
        if (wa >= 1 && wa * m.plan_alt.R > W * m.plan.R) plp = &m.plan_alt, *ystream = m.ystream_alt, W = wa;
    }
    *W_out = W;
    return plp;
}

int chrom_pipe_shape(const ChromModel &m, int C, int sm_count, int max_chunks, ChromPipe *pipe) {
    const float *ys;
    int W = 0;
    chrom_pick_plan(m, C, sm_count, &ys, &W);
    if (W < 1) return BINFB_EUNSUPPORTED;
    const int n_groups = (C + W - 1) / W;
    int n_chunks = max_chunks < 64 ? max_chunks : 64;
    if (n_chunks > n_groups) n_chunks = n_groups;
    pipe->W = W, pipe->n_groups = n_groups;
    pipe->groups_per_chunk = (n_groups + n_chunks - 1) / n_chunks;
    pipe->n_chunks = (n_groups + pipe->groups_per_chunk - 1) / pipe->groups_per_chunk;
    return BINFB_OK;
}

static int chrom_launch(ChromModel &m, ChromCall &call, int C, int sm_count, int smem_optin,
                        cudaStream_t s, ChromPipe *pipe = nullptr) {
    int rc = chrom_reserve(m, C);
    if (rc) return rc;
    const float *ystream = nullptr;
    int W = 0;
    const ChromPlan *plp = chrom_pick_plan(m, C, sm_count, &ystream, &W);
    const ChromPlan &pl = *plp;
    if (W < 1) {
        set_error("chromatin model: one chain does not fit in shared memory (n_beads too large for "
                  "this kernel)");
        return BINFB_EUNSUPPORTED;
    }
    call.W = W;
    call.n_groups = (C + W - 1) / W;
    const int passes = call.mode == CHROM_MODE_HMC ? call.h.L + 1 : 1;
    const int n_traj = call.mode == CHROM_MODE_HMC ? call.h.n_traj : 1;
    const long long total = (long long)call.n_groups * passes * n_traj;
    if (total > 2000000000LL) {
        set_error("chromatin model: too many work items in one launch");
        return BINFB_EUNSUPPORTED;
    }
    call.total_items = (int)total;
    {
        // sets: the hand-off between consecutive passes of a group goes through qw / pw (and q at both ends):
        // 36 bytes per degree of freedom.  With all groups in one pass-major sequence a group's data is evicted
        // from L2 before its next pass when the batch is large (147 MB at 4096 chains x 1000 beads); sets of at
        // most ~48 MB keep it resident.  Sets follow each other in ONE item sequence (no barrier, no tail between
        // sets), and a set never has fewer groups than the GPU has SMs.
        const double bytes_per_group = 36.0 * m.n * W;
        long long gs = (long long)(48e6 / bytes_per_group);
        if (gs < 2LL * sm_count) gs = 2LL * sm_count;
        int n_sets = (int)((call.n_groups + gs - 1) / gs);
        if (n_sets < 1) n_sets = 1;
        call.set_groups = (call.n_groups + n_sets - 1) / n_sets;
        if (m.opt_sets == 0) call.set_groups = call.n_groups;
    }
    size_t smem = pl.fixed_smem + pl.per_chain_smem * W;
    // scratch for the parallel fold of the row sums (chrom_sweep), where it fits next to the chains
    // (F roles at a time, F the largest power of two <= R that fits; the kernel derives F from the size)
    if (pl.R >= 4) {
        int F = pl.R;
        while (F >= 2 && smem + (size_t)W * F * FLUSH_ROLE_BYTES > (size_t)smem_optin) F /= 2;
        if (F >= 2) smem += (size_t)W * F * FLUSH_ROLE_BYTES;
    }
    BINFB_CUDA(cudaMemsetAsync(m.sched, 0, (size_t)(1 + call.n_groups + CHROM_GATE_WORDS) * sizeof(int), s));
    if (pipe) {
        // pipelined host call: nothing has arrived yet (gate = 1 + 0 chains); the caller's copy streams write
        // the gate and wait on the chunk counters (see CHROM_GATE_WORDS)
        if (pipe->W != W || pipe->n_groups != call.n_groups) {
            set_error("chromatin model: pipelined launch shape changed between query and launch");
            return BINFB_EINVAL;
        }
        pipe->gate = m.sched + 1 + call.n_groups;
        pipe->done = pipe->gate + 2;
        pipe->header[0] = 1, pipe->header[1] = pipe->groups_per_chunk;
        BINFB_CUDA(cudaMemcpyAsync(pipe->gate, pipe->header, 2 * sizeof(int), cudaMemcpyHostToDevice, s));
        // the copy streams may touch the gate from here on -- NOT after the kernel, which waits for them
        BINFB_CUDA(cudaEventRecord(pipe->header_written, s));
    }
    const int grid = call.n_groups < sm_count * BINFB_CTAS_PER_SM ? call.n_groups : sm_count * BINFB_CTAS_PER_SM;
    const int threads = W * pl.R * 32;
    const ChromDev dev = chrom_dev(m, pl, ystream);
#define BINFB_CHROM_LAUNCH_E(RR, SPR, LOCK, NSS, EVV, ALGG)                                                    \
    do {                                                                                                       \
        BINFB_CUDA(cudaFuncSetAttribute(chrom_kernel<RR, SPR, LOCK, NSS, EVV, ALGG>,                           \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        if (BINFB_CTAS_PER_SM > 1)                                                                             \
            BINFB_CUDA(cudaFuncSetAttribute(chrom_kernel<RR, SPR, LOCK, NSS, EVV, ALGG>,                       \
                                            cudaFuncAttributePreferredSharedMemoryCarveout, 100));             \
        chrom_kernel<RR, SPR, LOCK, NSS, EVV, ALGG><<<grid, threads, smem, s>>>(dev, call);                    \
    } while (0)
#define BINFB_CHROM_LAUNCH_N(RR, SPR, LOCK, NSS)                                                               \
    do {                                                                                                       \
        const bool alg = (m.flags & BINFB_FLAG_CONTACT_ALGEBRAIC) != 0;                                        \
        if (m.ev_k > 0.f && alg) BINFB_CHROM_LAUNCH_E(RR, SPR, LOCK, NSS, true, true);                         \
        else if (m.ev_k > 0.f) BINFB_CHROM_LAUNCH_E(RR, SPR, LOCK, NSS, true, false);                          \
        else if (alg) BINFB_CHROM_LAUNCH_E(RR, SPR, LOCK, NSS, false, true);                                   \
        else BINFB_CHROM_LAUNCH_E(RR, SPR, LOCK, NSS, false, false);                                           \
    } while (0)
#define BINFB_CHROM_LAUNCH_L(RR, SPR, LOCK) BINFB_CHROM_LAUNCH_N(RR, SPR, LOCK, CHROM_NS)
#define BINFB_CHROM_LAUNCH(RR, SPR) BINFB_CHROM_LAUNCH_L(RR, SPR, false)
    if (pl.R == 16 && pl.SS == 16 && pl.NS == 3 && !pl.lockstep) BINFB_CHROM_LAUNCH_N(16, 1, false, 3);
    else if (pl.R == 16 && pl.SS == 16 && pl.NS == 2 && !pl.lockstep) BINFB_CHROM_LAUNCH_N(16, 1, false, 2);
    else if (pl.lockstep && pl.R == 2 && pl.SS == 4) BINFB_CHROM_LAUNCH_L(2, 2, true);
    else if (pl.lockstep && pl.R == 4 && pl.SS == 4) BINFB_CHROM_LAUNCH_L(4, 1, true);
    else if (pl.lockstep && pl.R == 8 && pl.SS == 8) BINFB_CHROM_LAUNCH_L(8, 1, true);
    else if (pl.lockstep) {
        set_error("chromatin model: unsupported lockstep launch plan");
        return BINFB_EUNSUPPORTED;
    } else if (pl.R == 1 && pl.SS == 4) BINFB_CHROM_LAUNCH(1, 4);
#if BINFB_CHROM_SS == 8
    else if (pl.R == 2 && pl.SS == 8) BINFB_CHROM_LAUNCH(2, 4);
#elif BINFB_CHROM_SS == 2
    else if (pl.R == 2 && pl.SS == 2) BINFB_CHROM_LAUNCH(2, 1);
#elif BINFB_CHROM_SS == 6
    else if (pl.R == 2 && pl.SS == 6) BINFB_CHROM_LAUNCH(2, 3);
#endif
    else if (pl.R == 2 && pl.SS == 4) BINFB_CHROM_LAUNCH(2, 2);
    else if (pl.R == 4 && pl.SS == 4) BINFB_CHROM_LAUNCH(4, 1);
    else if (pl.R == 8 && pl.SS == 8) BINFB_CHROM_LAUNCH(8, 1);
    else {
        set_error("chromatin model: unsupported launch plan");
        return BINFB_EUNSUPPORTED;
    }
#undef BINFB_CHROM_LAUNCH
#undef BINFB_CHROM_LAUNCH_L
#undef BINFB_CHROM_LAUNCH_N
#undef BINFB_CHROM_LAUNCH_E
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

int chrom_hmc_launch(ChromModel &m, const HmcArgs &a, int sm_count, int smem_optin, cudaStream_t s,
                     ChromPipe *pipe) {
    ChromCall call;
    call.mode = CHROM_MODE_HMC;
    call.h = a;
    call.g = GradArgs();
    if (a.n_accepted) BINFB_CUDA(cudaMemsetAsync(a.n_accepted, 0, (size_t)a.C * sizeof(int32_t), s));
    if (m.ev_k > 0.f && a.gibbs_mode == BINFB_GIBBS_TAU_FIRST) {
        // excluded volume + precision-first Gibbs: tau must be drawn before the first sweep, from
        // chi^2 of the incoming state (see chrom_kernel): one energy-only pass fills chi2_state
        int rc = chrom_reserve(m, a.C);
        if (rc) return rc;
        GradArgs g = GradArgs();
        g.q = a.q, g.tau = a.tau, g.beta = a.beta, g.C = a.C, g.logp = nullptr, g.grad = nullptr;
        g.chi2 = m.chi2_state, g.gamma_shape = a.gamma_shape, g.gamma_rate = a.gamma_rate;
        rc = chrom_grad_launch(m, g, sm_count, smem_optin, s);
        if (rc) return rc;
    }
    if (pipe && m.ev_k > 0.f && a.gibbs_mode == BINFB_GIBBS_TAU_FIRST) {
        set_error("chromatin model: the pipelined host call does not cover excluded volume + precision-first Gibbs");
        return BINFB_EUNSUPPORTED;
    }
    const int rc2 = chrom_launch(m, call, a.C, sm_count, smem_optin, s, pipe);
    m.chi2_chains = rc2 ? 0 : a.C;  // chi2_state now holds chi^2 of every chain's current state (binfb_hmc_last_chi2)
    return rc2;
}

int chrom_grad_launch(ChromModel &m, const GradArgs &a, int sm_count, int smem_optin, cudaStream_t s) {
    ChromCall call;
    call.mode = CHROM_MODE_GRAD;
    call.h = HmcArgs();
    call.g = a;
    return chrom_launch(m, call, a.C, sm_count, smem_optin, s);
}

int chrom_forward_launch(const ChromModel &m, const float *q, int C, float *mock, cudaStream_t s) {
    const double log2e = 1.4426950408889634;
    const bool alg = (m.flags & BINFB_FLAG_CONTACT_ALGEBRAIC) != 0;
    const float A = alg ? m.alpha : (float)((double)m.alpha * log2e);
    const float B = alg ? (float)(-(double)m.alpha * (double)m.d_c) : (float)(-(double)m.alpha * (double)m.d_c * log2e);
    dim3 grid((unsigned)((m.M + 255) / 256), (unsigned)C);
    chrom_forward_kernel<<<grid, 256, 0, s>>>(m.n, m.M, A, B, alg ? 1 : 0, q, mock);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

}  // namespace binfb
