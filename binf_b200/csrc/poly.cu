// Fused HMC for the polynomial (Vandermonde) posterior -- sm_100a.
//
// Replaces, per chain and per trajectory, the reference's HMCSampler.sample / _leapfrog
// (binf/samplers/hmc.py:92-125,136-164) together with every pdf.gradient / pdf.log_prob call
// they make: Posterior (binf/pdf/posteriors.py:125-187) -> Likelihood
// (binf/pdf/likelihoods.py:141-155) -> polynomial ForwardModel + GaussianErrorModel
// (binf/example/likelihood.py:24-30,54-61) and the priors (binf/example/priors.py:23-25,49-54).
//
// Mapping: a chain is owned by a group of G lanes (G = 1..32, power of two); each lane
// walks every G-th data row and the group butterfly-reduces the K gradient sums, so all G
// lanes hold bitwise-identical q, p for the whole trajectory (registers; nothing but the final
// state goes back to HBM).  A thread carries J independent chains to reuse each shared-memory
// data row (one LDS.128 feeds 8*J FP32 instructions).  The data rows
// [x, x^2, .., x^(K-1), y] live in shared memory for the whole launch.
#include <math.h>

#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace binfb {

struct PolyDev {
    const float *rows;
    int N;
    unsigned flags;
    float prior_mean[8], prior_inv_var[8];
};

template <int K>
struct PolyRow {
    static constexpr int STRIDE = ((K + 3) / 4) * 4;
};

// One data row for J chains: residual via Horner (K-1 FMA + 1 ADD), gradient sums (1 ADD +
// K-1 FMA), optional chi^2 (1 FMA)  ->  14 flop per chain-datum at K = 4.
template <int K, int J, bool ENERGY>
__device__ __forceinline__ void poly_row(const float *__restrict__ rowp, const float (&c)[J][K],
                                         float (&acc)[J][K], float (&part)[J]) {
    constexpr int S = PolyRow<K>::STRIDE;
    float row[S];
#pragma unroll
    for (int v = 0; v < S / 4; ++v) {
        const float4 t = reinterpret_cast<const float4 *>(rowp)[v];
        row[4 * v + 0] = t.x, row[4 * v + 1] = t.y, row[4 * v + 2] = t.z, row[4 * v + 3] = t.w;
    }
    const float y = row[K - 1];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        float t = c[j][K - 1];
#pragma unroll
        for (int k = K - 2; k >= 0; --k) t = fmaf(t, row[0], c[j][k]);
        const float res = t - y;
        acc[j][0] += res;
#pragma unroll
        for (int k = 1; k < K; ++k) acc[j][k] = fmaf(res, row[k - 1], acc[j][k]);
        if (ENERGY) part[j] = fmaf(res, res, part[j]);
    }
}

// Packed variant for an even number of chains per thread: chains (2*jp, 2*jp+1) share one FFMA2 /
// FADD2 (two FP32 lanes per instruction, the data row enters as a broadcast scalar operand).  Same
// arithmetic per chain as poly_row; half the issue slots, so the FMA pipe -- not the issue port --
// bounds the loop.
template <int K, int JP, bool ENERGY>
__device__ __forceinline__ void poly_row2(const float *__restrict__ rowp, const float2 (&c)[JP][K],
                                          float2 (&acc)[JP][K], float2 (&part)[JP]) {
    constexpr int S = PolyRow<K>::STRIDE;
    float row[S];
#pragma unroll
    for (int v = 0; v < S / 4; ++v) {
        const float4 t = reinterpret_cast<const float4 *>(rowp)[v];
        row[4 * v + 0] = t.x, row[4 * v + 1] = t.y, row[4 * v + 2] = t.z, row[4 * v + 3] = t.w;
    }
    const float2 ny = make_float2(-row[K - 1], -row[K - 1]);
    const float2 x = make_float2(row[0], row[0]);
#pragma unroll
    for (int j = 0; j < JP; ++j) {
        float2 t = c[j][K - 1];
#pragma unroll
        for (int k = K - 2; k >= 0; --k) t = __ffma2_rn(t, x, c[j][k]);
        const float2 res = __fadd2_rn(t, ny);
        acc[j][0] = __fadd2_rn(acc[j][0], res);
#pragma unroll
        for (int k = 1; k < K; ++k) acc[j][k] = __ffma2_rn(res, make_float2(row[k - 1], row[k - 1]), acc[j][k]);
        if (ENERGY) part[j] = __ffma2_rn(res, res, part[j]);
    }
}

template <int K, int G, int J, bool ENERGY>
__device__ __forceinline__ void poly_chunk2(const float *__restrict__ srows, int n_rows, int g,
                                            const float (&cs)[J][K], float (&accs)[J][K],
                                            double (&chi2)[J]) {
    constexpr int S = PolyRow<K>::STRIDE;
    constexpr int U = 8, JP = J / 2;
    float2 c[JP][K], acc[JP][K];
#pragma unroll
    for (int j = 0; j < JP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            c[j][k] = make_float2(cs[2 * j][k], cs[2 * j + 1][k]);
            acc[j][k] = make_float2(accs[2 * j][k], accs[2 * j + 1][k]);
        }
    const int cnt = (n_rows - g + G - 1) / G;
    const float *p = srows + (size_t)g * S;
    int i = 0;
    for (; i + U <= cnt; i += U) {
        float2 part[JP];
#pragma unroll
        for (int j = 0; j < JP; ++j) part[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int uu = 0; uu < U; ++uu) poly_row2<K, JP, ENERGY>(p + (size_t)uu * G * S, c, acc, part);
        p += (size_t)U * G * S;
        if (ENERGY) {
#pragma unroll
            for (int j = 0; j < JP; ++j) chi2[2 * j] += (double)part[j].x, chi2[2 * j + 1] += (double)part[j].y;
        }
    }
    if (i < cnt) {
        float2 part[JP];
#pragma unroll
        for (int j = 0; j < JP; ++j) part[j] = make_float2(0.f, 0.f);
        for (; i < cnt; ++i) {
            poly_row2<K, JP, ENERGY>(p, c, acc, part);
            p += (size_t)G * S;
        }
        if (ENERGY) {
#pragma unroll
            for (int j = 0; j < JP; ++j) chi2[2 * j] += (double)part[j].x, chi2[2 * j + 1] += (double)part[j].y;
        }
    }
#pragma unroll
    for (int j = 0; j < JP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) accs[2 * j][k] = acc[j][k].x, accs[2 * j + 1][k] = acc[j][k].y;
}

// all rows r = g, g+G, ... of a shared-memory chunk
template <int K, int G, int J, bool ENERGY>
__device__ __forceinline__ void poly_chunk(const float *__restrict__ srows, int n_rows, int g,
                                           const float (&c)[J][K], float (&acc)[J][K],
                                           double (&chi2)[J]) {
    constexpr int S = PolyRow<K>::STRIDE;
    constexpr int U = 8;
    const int cnt = (n_rows - g + G - 1) / G;  // rows of this lane
    const float *p = srows + (size_t)g * S;
    int i = 0;
    for (; i + U <= cnt; i += U) {
        float part[J];
#pragma unroll
        for (int j = 0; j < J; ++j) part[j] = 0.f;
#pragma unroll
        for (int uu = 0; uu < U; ++uu) poly_row<K, J, ENERGY>(p + (size_t)uu * G * S, c, acc, part);
        p += (size_t)U * G * S;
        if (ENERGY) {
#pragma unroll
            for (int j = 0; j < J; ++j) chi2[j] += (double)part[j];
        }
    }
    if (i < cnt) {
        float part[J];
#pragma unroll
        for (int j = 0; j < J; ++j) part[j] = 0.f;
        for (; i < cnt; ++i) {
            poly_row<K, J, ENERGY>(p, c, acc, part);
            p += (size_t)G * S;
        }
        if (ENERGY) {
#pragma unroll
            for (int j = 0; j < J; ++j) chi2[j] += (double)part[j];
        }
    }
}

__device__ __forceinline__ void poly_load_rows(float *srows, const float *__restrict__ grows,
                                               int n_floats) {
    const float4 *src = reinterpret_cast<const float4 *>(grows);
    float4 *dst = reinterpret_cast<float4 *>(srows);
    for (int i = threadIdx.x; i < n_floats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
}

// raw gradient sums  graw_k = sum_n (mock_n - y_n) x_n^k  and chi^2 = sum_n (mock_n - y_n)^2,
// identical on all G lanes of the group
template <int K, int G, int J, bool ENERGY>
__device__ __forceinline__ void poly_grad_pass(const PolyDev &pm, float *srows, int rows_per_chunk,
                                               int n_chunks, int g, const float (&c)[J][K],
                                               float (&graw)[J][K], double (&chi2)[J]) {
    constexpr int S = PolyRow<K>::STRIDE;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        chi2[j] = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) graw[j][k] = 0.f;
    }
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int r0 = ch * rows_per_chunk;
        const int n_rows = min(rows_per_chunk, pm.N - r0);
        if (n_chunks > 1) {
            __syncthreads();
            poly_load_rows(srows, pm.rows + (size_t)r0 * S, n_rows * S);
            __syncthreads();
        }
        if constexpr (J % 2 == 0) poly_chunk2<K, G, J, ENERGY>(srows, n_rows, g, c, graw, chi2);
        else poly_chunk<K, G, J, ENERGY>(srows, n_rows, g, c, graw, chi2);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int k = 0; k < K; ++k) graw[j][k] = group_allreduce_sum<G>(graw[j][k]);
        if (ENERGY) chi2[j] = group_allreduce_sum<G>(chi2[j]);
    }
}

// ---------------------------------------------------------------------------------------------
// "uniform-row" mapping (UR).  Measured on B200 (profiles/microbench/ffma2_operands.cu, poly_ur.cu): an
// FFMA2 whose broadcast operand sits in a vector register occupies the FMA pipe 3 cycles, the same FFMA2
// with the broadcast operand in a UNIFORM register 2 cycles.  A data row lands in uniform registers when
// the whole warp reads the same row from constant memory (LDCU).  So here the 32 lanes of a warp own 32
// different chain pairs and all walk the same rows; the W warps of a "set" own the same 32 chain pairs and
// split the rows into W contiguous ranges; the K gradient sums (and chi^2) of a pass are combined through
// shared memory in a fixed order, so that all W warps hold bitwise identical q, p.  No LDS in the row loop.
// ---------------------------------------------------------------------------------------------
constexpr int POLY_CROWS_BYTES = 48 * 1024;
__constant__ float4 c_poly_rows[POLY_CROWS_BYTES / 16];

__device__ __forceinline__ void set_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// scratch of one set: [2 buffers][J*K sums][W warps][32 lanes] float, then [2][J][W][32] double
template <int K, int W, int J>
struct PolyUrScratch {
    static constexpr int F = 2 * J * K * W * 32;  // floats
    static constexpr int D = 2 * J * W * 32;      // doubles
    static constexpr size_t BYTES = (size_t)F * 4 + (size_t)D * 8;
};

template <int K, int W, int J, bool ENERGY>
__device__ __forceinline__ void poly_grad_pass_ur(const PolyDev &pm, unsigned char *set_scratch, int buf, int w,
                                                  int lane, int bar_id, const float (&cs)[J][K],
                                                  float (&graw)[J][K], double (&chi2)[J]) {
    static_assert(J % 2 == 0, "UR mapping packs chain pairs");
    constexpr int S4 = PolyRow<K>::STRIDE / 4, JP = J / 2, U = 8;
    float2 c[JP][K], acc[JP][K];
    double chi[J];
#pragma unroll
    for (int j = 0; j < JP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            c[j][k] = make_float2(cs[2 * j][k], cs[2 * j + 1][k]);
            acc[j][k] = make_float2(0.f, 0.f);
        }
#pragma unroll
    for (int j = 0; j < J; ++j) chi[j] = 0.0;
    const int rpw = (pm.N + W - 1) / W;
    int i = w * rpw;  // w is warp-uniform (broadcast by the caller): the row index stays in the uniform datapath
    const int i_end = min(pm.N, i + rpw);
    for (; i + U <= i_end; i += U) {
        float2 part[JP];
#pragma unroll
        for (int j = 0; j < JP; ++j) part[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int uu = 0; uu < U; ++uu)
            poly_row2<K, JP, ENERGY>(reinterpret_cast<const float *>(c_poly_rows + (size_t)(i + uu) * S4), c, acc, part);
        if (ENERGY) {
#pragma unroll
            for (int j = 0; j < JP; ++j) chi[2 * j] += (double)part[j].x, chi[2 * j + 1] += (double)part[j].y;
        }
    }
    if (i < i_end) {
        float2 part[JP];
#pragma unroll
        for (int j = 0; j < JP; ++j) part[j] = make_float2(0.f, 0.f);
        for (; i < i_end; ++i)
            poly_row2<K, JP, ENERGY>(reinterpret_cast<const float *>(c_poly_rows + (size_t)i * S4), c, acc, part);
        if (ENERGY) {
#pragma unroll
            for (int j = 0; j < JP; ++j) chi[2 * j] += (double)part[j].x, chi[2 * j + 1] += (double)part[j].y;
        }
    }
    if (W == 1) {
#pragma unroll
        for (int j = 0; j < JP; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k) graw[2 * j][k] = acc[j][k].x, graw[2 * j + 1][k] = acc[j][k].y;
#pragma unroll
        for (int j = 0; j < J; ++j) chi2[j] = chi[j];
        return;
    }
    // combine the W row ranges: every warp writes its partial sums, one barrier, every warp adds all W
    // partials in the same order.  Buffer `buf` of pass e is rewritten by pass e + 2, i.e. after the barrier
    // of pass e + 1, which every warp reaches only after it has read pass e.
    using Sc = PolyUrScratch<K, W, J>;
    float *sf = reinterpret_cast<float *>(set_scratch) + (size_t)buf * (Sc::F / 2);
    double *sd = reinterpret_cast<double *>(set_scratch + (size_t)Sc::F * 4) + (size_t)buf * (Sc::D / 2);
#pragma unroll
    for (int j = 0; j < JP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            sf[(((2 * j) * K + k) * W + w) * 32 + lane] = acc[j][k].x;
            sf[(((2 * j + 1) * K + k) * W + w) * 32 + lane] = acc[j][k].y;
        }
    if (ENERGY) {
#pragma unroll
        for (int j = 0; j < J; ++j) sd[(j * W + w) * 32 + lane] = chi[j];
    }
    set_bar(bar_id, W * 32);
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float t = 0.f;
#pragma unroll
            for (int ww = 0; ww < W; ++ww) t += sf[((j * K + k) * W + ww) * 32 + lane];
            graw[j][k] = t;
        }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        double t = 0.0;
        if (ENERGY) {
#pragma unroll
            for (int ww = 0; ww < W; ++ww) t += sd[(j * W + ww) * 32 + lane];
        }
        chi2[j] = t;
    }
}

// U(q) = -log p(q | tau) in float64 (binf/pdf/posteriors.py:141-151 summed components)
template <int K>
__device__ __forceinline__ double poly_potential(const PolyDev &pm, const float (&q)[K], double chi2,
                                                 float tau, float beta, double ga, double gb) {
    const double t = (double)tau;
    double prior = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double d = (double)q[k] - (double)pm.prior_mean[k];
        prior += d * d * (double)pm.prior_inv_var[k];
    }
    const double lt = log(t);
    return (double)beta * (0.5 * t * chi2 - 0.5 * (double)pm.N * lt) + 0.5 * prior -
           ((ga - 1.0) * lt - gb * t);
}

template <int K>
__device__ __forceinline__ void poly_force(const PolyDev &pm, const float (&q)[K],
                                           const float (&graw)[K], float bt, float (&f)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        f[k] = bt * graw[k];
        if (pm.flags & BINFB_FLAG_PRIOR_GRAD)
            f[k] = fmaf(q[k] - pm.prior_mean[k], pm.prior_inv_var[k], f[k]);
    }
}

__device__ __forceinline__ float draw_tau(const HmcArgs &a, double n_data, double chi2, float beta,
                                          uint64_t chain, int cid, uint64_t draw) {
    const double shape = 0.5 * (double)beta * n_data + a.gamma_shape - 1.0;
    const double rate = 0.5 * (double)beta * chi2 + a.gamma_rate;
    const double gdraw =
        a.gamma_draws ? a.gamma_draws[cid] : rng_gamma(a.seed, chain, draw, shape);
    return (float)(gdraw / rate);
}

// threads per CTA (one CTA per SM): 28 warps at <= 72 registers for 2 or 4 chains per thread
// (4 chains with 4 lanes per chain: 16 warps, no register limit that matters)
constexpr int poly_max_block(int G, int J) { return J == 1 ? 1024 : (J == 4 && G <= 4 ? 512 : 896); }

// thread -> (chain tuple, share of the rows).  Regular mapping: G consecutive lanes per tuple, lane g walks
// rows g, g+G, ...  UR mapping: lane l of the G warps of set s owns tuple (s, l); "g" is the warp's index in
// its set, broadcast from lane 0 so that the compiler keeps everything derived from it in uniform registers.
template <int G, bool UR>
struct PolyMap {
    int g, lane, bar_id;
    long long group, n_groups;
    unsigned char *scratch;
    template <int K, int J>
    __device__ __forceinline__ void init(float *smem) {
        if (UR) {
            const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
            const int set = warp / G, sets = (int)(blockDim.x >> 5) / G;
            g = warp % G, lane = threadIdx.x & 31, bar_id = 1 + set;
            group = ((long long)blockIdx.x * sets + set) * 32 + lane;
            n_groups = (long long)gridDim.x * sets * 32;
            scratch = reinterpret_cast<unsigned char *>(smem) + (size_t)set * PolyUrScratch<K, G, J>::BYTES;
        } else {
            g = threadIdx.x % G, lane = 0, bar_id = 0, scratch = nullptr;
            group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
            n_groups = ((long long)gridDim.x * blockDim.x) / G;
        }
    }
};

template <int K, int G, int J, bool UR = false>
__global__ void __launch_bounds__(poly_max_block(G, J), 1)
    poly_hmc_kernel(PolyDev pm, HmcArgs a, int iters, int rows_per_chunk, int n_chunks) {
    extern __shared__ __align__(16) float srows[];
    constexpr int S = PolyRow<K>::STRIDE;
    if (!UR && n_chunks == 1) {
        poly_load_rows(srows, pm.rows, pm.N * S);
        __syncthreads();
    }
    PolyMap<G, UR> map;
    map.template init<K, J>(srows);
    const int g = map.g;
    const long long group = map.group, n_groups = map.n_groups;
    int pass = 0;  // UR: running count of gradient passes (scratch buffer parity)
    auto grad_pass = [&](auto energy_tag, const float (&qq)[J][K], float (&gg)[J][K], double (&cc)[J]) {
        constexpr bool EN = decltype(energy_tag)::value;
        if constexpr (UR) {
            poly_grad_pass_ur<K, G, J, EN>(pm, map.scratch, pass & 1, g, map.lane, map.bar_id, qq, gg, cc);
            ++pass;
        } else {
            poly_grad_pass<K, G, J, EN>(pm, srows, rows_per_chunk, n_chunks, g, qq, gg, cc);
        }
    };
    double st_acc = 0.0, st_prop = 0.0, st_eps = 0.0, st_pacc = 0.0;

    for (int it = 0; it < iters; ++it) {
        const long long tuple = (long long)it * n_groups + group;
        int cid[J];
        bool valid[J];
        float q[J][K], p[J][K], graw[J][K], f[K];
        float tau[J], beta[J], eps[J];
        int nacc[J];
        double chi2[J], chi2_cur[J], h0[J], h1[J];
        bool acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const long long c = tuple * J + j;
            valid[j] = c < a.C;
            cid[j] = valid[j] ? (int)c : a.C - 1;
#pragma unroll
            for (int k = 0; k < K; ++k) q[j][k] = a.q[(size_t)cid[j] * K + k];
            tau[j] = a.tau[cid[j]];
            beta[j] = a.beta ? a.beta[cid[j]] : 1.0f;
            eps[j] = a.eps[cid[j]];
            nacc[j] = 0;
            acc[j] = false;
            h0[j] = h1[j] = 0.0;
        }
        for (int tr = 0; tr < a.n_traj; ++tr) {
            const uint64_t draw = a.draw + (uint64_t)tr;
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int k = 0; k < K; ++k)
                    p[j][k] = a.p0 ? a.p0[(size_t)cid[j] * K + k]
                                   : rng_normal(a.seed, a.chain_base + cid[j], draw, k);
            // ---- force evaluation 0 (+ energy at q0) -----------------------------------
            grad_pass(std::true_type{}, q, graw, chi2);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (a.gibbs_mode == BINFB_GIBBS_TAU_FIRST)
                    tau[j] = draw_tau(a, (double)pm.N, chi2[j], beta[j], a.chain_base + cid[j],
                                      cid[j], draw);
                chi2_cur[j] = chi2[j];
                double kin = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) kin += (double)p[j][k] * (double)p[j][k];
                h0[j] = poly_potential<K>(pm, q[j], chi2[j], tau[j], beta[j], a.gamma_shape,
                                          a.gamma_rate) + 0.5 * kin;
                poly_force<K>(pm, q[j], graw[j], beta[j] * tau[j], f);
                const float he = 0.5f * eps[j];
#pragma unroll
                for (int k = 0; k < K; ++k) p[j][k] = fmaf(-he, f[k], p[j][k]);  // hmc.py:116
            }
            // ---- L-1 full steps (hmc.py:118-120) -----------------------------------------
            for (int s = 1; s < a.L; ++s) {
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int k = 0; k < K; ++k) q[j][k] = fmaf(eps[j], p[j][k], q[j][k]);
                grad_pass(std::false_type{}, q, graw, chi2);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    poly_force<K>(pm, q[j], graw[j], beta[j] * tau[j], f);
#pragma unroll
                    for (int k = 0; k < K; ++k) p[j][k] = fmaf(-eps[j], f[k], p[j][k]);
                }
            }
            // ---- last drift + half kick (hmc.py:122-123) + energy at q_L -------------------
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int k = 0; k < K; ++k) q[j][k] = fmaf(eps[j], p[j][k], q[j][k]);
            grad_pass(std::true_type{}, q, graw, chi2);
            const bool last = tr == a.n_traj - 1;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                poly_force<K>(pm, q[j], graw[j], beta[j] * tau[j], f);
                const float he = 0.5f * eps[j];
                double kin = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    p[j][k] = fmaf(-he, f[k], p[j][k]);
                    kin += (double)p[j][k] * (double)p[j][k];
                }
                h1[j] = poly_potential<K>(pm, q[j], chi2[j], tau[j], beta[j], a.gamma_shape,
                                          a.gamma_rate) + 0.5 * kin;
                // Metropolis (hmc.py:151): NaN energies reject
                const float uu = a.u ? a.u[cid[j]]
                                     : rng_uniform(a.seed, a.chain_base + cid[j], draw, RNG_ACCEPT);
                const double dh = h1[j] - h0[j];
                const double pacc = exp(fmin(0.0, -dh));
                acc[j] = (dh == dh) && ((double)uu < exp(fmin(709.0, fmax(-308.0, -dh))));
                if (last && valid[j] && g == 0) {
                    if (a.q_end)
                        for (int k = 0; k < K; ++k) a.q_end[(size_t)cid[j] * K + k] = q[j][k];
                    if (a.p_end)
                        for (int k = 0; k < K; ++k) a.p_end[(size_t)cid[j] * K + k] = p[j][k];
                }
                if (acc[j]) {
                    chi2_cur[j] = chi2[j];
                    nacc[j]++;
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) q[j][k] = a.q[(size_t)cid[j] * K + k];
                }
                if (valid[j] && g == 0) {
                    st_acc += acc[j] ? 1.0 : 0.0;
                    st_prop += 1.0;
                    st_eps += (double)eps[j];
                    st_pacc += (dh == dh) ? pacc : 0.0;
                }
                if (tr < a.n_adapt) eps[j] *= acc[j] ? a.adapt_up : a.adapt_down;  // hmc.py:188-191
                if (a.gibbs_mode == BINFB_GIBBS_TAU_LAST)
                    tau[j] = draw_tau(a, (double)pm.N, chi2_cur[j], beta[j], a.chain_base + cid[j],
                                      cid[j], draw);
            }
            // the state array is the rejection fallback: commit accepted moves before the
            // next trajectory re-reads it
            if (a.n_traj > 1) {
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if (valid[j] && g == 0 && acc[j])
                        for (int k = 0; k < K; ++k) a.q[(size_t)cid[j] * K + k] = q[j][k];
                if (UR && G > 1) set_bar(map.bar_id, G * 32);  // the tuple's other lanes sit in other warps
                else __syncwarp();
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (valid[j] && g == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) a.q[(size_t)cid[j] * K + k] = q[j][k];
                a.tau[cid[j]] = tau[j];
                a.eps[cid[j]] = eps[j];
                if (a.accepted) a.accepted[cid[j]] = acc[j] ? 1 : 0;
                if (a.e_before) a.e_before[cid[j]] = h0[j];
                if (a.e_after) a.e_after[cid[j]] = h1[j];
                if (a.n_accepted) a.n_accepted[cid[j]] = nacc[j];
            }
        }
    }
    if (a.stats) {
        st_acc = group_allreduce_sum<32>(st_acc);
        st_prop = group_allreduce_sum<32>(st_prop);
        st_eps = group_allreduce_sum<32>(st_eps);
        st_pacc = group_allreduce_sum<32>(st_pacc);
        if ((threadIdx.x & 31) == 0 && st_prop > 0.0) {
            atomicAdd(a.stats + 0, st_acc);
            atomicAdd(a.stats + 1, st_prop);
            atomicAdd(a.stats + 2, st_eps);
            atomicAdd(a.stats + 3, st_pacc);
        }
    }
}

template <int K, int G, int J, bool UR = false>
__global__ void __launch_bounds__(poly_max_block(G, J), 1)
    poly_grad_kernel(PolyDev pm, GradArgs a, int iters, int rows_per_chunk, int n_chunks) {
    extern __shared__ __align__(16) float srows[];
    constexpr int S = PolyRow<K>::STRIDE;
    if (!UR && n_chunks == 1) {
        poly_load_rows(srows, pm.rows, pm.N * S);
        __syncthreads();
    }
    PolyMap<G, UR> map;
    map.template init<K, J>(srows);
    const int g = map.g;
    const long long group = map.group, n_groups = map.n_groups;
    for (int it = 0; it < iters; ++it) {
        const long long tuple = (long long)it * n_groups + group;
        int cid[J];
        bool valid[J];
        float q[J][K], graw[J][K], f[K];
        double chi2[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const long long c = tuple * J + j;
            valid[j] = c < a.C;
            cid[j] = valid[j] ? (int)c : a.C - 1;
#pragma unroll
            for (int k = 0; k < K; ++k) q[j][k] = a.q[(size_t)cid[j] * K + k];
        }
        if constexpr (UR)
            poly_grad_pass_ur<K, G, J, true>(pm, map.scratch, it & 1, g, map.lane, map.bar_id, q, graw, chi2);
        else
            poly_grad_pass<K, G, J, true>(pm, srows, rows_per_chunk, n_chunks, g, q, graw, chi2);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (!valid[j] || g != 0) continue;
            const float tau = a.tau[cid[j]];
            const float beta = a.beta ? a.beta[cid[j]] : 1.0f;
            if (a.logp)
                a.logp[cid[j]] =
                    -poly_potential<K>(pm, q[j], chi2[j], tau, beta, a.gamma_shape, a.gamma_rate);
            if (a.chi2) a.chi2[cid[j]] = chi2[j];
            if (a.grad) {
                poly_force<K>(pm, q[j], graw[j], beta * tau, f);
#pragma unroll
                for (int k = 0; k < K; ++k) a.grad[(size_t)cid[j] * K + k] = f[k];
            }
        }
    }
}

template <int K>
__global__ void poly_forward_kernel(const float *__restrict__ rows, int N, const float *q, int C,
                                    float *mock) {
    constexpr int S = PolyRow<K>::STRIDE;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)C * N) return;
    const int c = (int)(i / N), n = (int)(i % N);
    const float x = K > 1 ? rows[(size_t)n * S] : 0.f;
    float t = q[(size_t)c * K + K - 1];
#pragma unroll
    for (int k = K - 2; k >= 0; --k) t = fmaf(t, x, q[(size_t)c * K + k]);
    mock[i] = t;
}

// ---------------------------------------------------------------------------------------------
// host side: launch-shape heuristics and template dispatch
// ---------------------------------------------------------------------------------------------
struct PolyPlan {
    int G, J, grid, block, iters, rows_per_chunk, n_chunks;
    int ur;  // uniform-row mapping: G = warps per set
    size_t smem;
};

static bool g_allowed(int K, int G) {
    if (K == 4) return true;
    return G == 1 || G == 4 || G == 32;
}

static PolyPlan poly_plan(const PolyModel &m, int C, int sm_count, int smem_optin) {
    PolyPlan pl;
    pl.J = (m.K == 4 && (long long)C >= 64LL * sm_count) ? 2 : 1;
    // (4 chains per thread are instantiated for G = 4, 8 but measured slower than 2 x G=4 at C = 65,536:
    //  3.46 warps per SMSP or register spills at 72 registers; opt in with poly.chains_per_thread = 4)
    if (m.opt_jchains == 1 || m.opt_jchains == 2 || m.opt_jchains == 4) pl.J = (m.K == 4) ? m.opt_jchains : 1;
    int max_block = pl.J == 1 ? 1024 : 896;
    const long long tuples = ((long long)C + pl.J - 1) / pl.J;
    // lanes per chain: fill ~max_block threads per SM, the largest power of two that fits
    int G = 1;
    while (G < 32 && tuples * (G * 2) <= (long long)sm_count * max_block) G *= 2;
    if (m.opt_group > 0) G = m.opt_group;
    while (!g_allowed(m.K, G) && G > 1) G /= 2;
    if (pl.J == 4) {
        G = G >= 8 ? 8 : 4;  // the only two shapes instantiated for 4 chains per thread
        max_block = poly_max_block(G, 4);
    }
    pl.G = G;
    // uniform-row mapping (see poly_grad_pass_ur): K = 4, packed chain pairs, the rows fit the constant bank
    pl.ur = m.K == 4 && pl.J == 2 && G <= 8 && m.opt_ur != 0 &&
            (size_t)m.N * m.stride * sizeof(float) <= (size_t)POLY_CROWS_BYTES;
    if (pl.ur) {
        const int unit = 32 * G;  // threads of one set
        max_block = (max_block / unit) * unit;
        const long long sets = (tuples + 31) / 32;
        long long grid = sets >= sm_count ? sm_count : sets;
        long long spc = (sets + grid - 1) / grid;  // sets per CTA
        if (spc * unit > max_block) spc = max_block / unit;
        if (m.opt_block > 0 && m.opt_block % unit == 0 && m.opt_block <= max_block) spc = m.opt_block / unit;
        pl.grid = (int)grid, pl.block = (int)(spc * unit);
        const long long n_groups = grid * spc * 32;
        pl.iters = (int)((tuples + n_groups - 1) / n_groups);
        pl.rows_per_chunk = m.N, pl.n_chunks = 1;
        pl.smem = G > 1 ? (size_t)spc * PolyUrScratch<4, 8, 2>::BYTES * G / 8 : 0;
        return pl;
    }
    const long long threads = tuples * G;
    long long grid = threads >= 32LL * sm_count ? sm_count : (threads + 31) / 32;
    long long block = (threads + grid - 1) / grid;
    block = ((block + 31) / 32) * 32;
    if (m.opt_block > 0) block = m.opt_block;
    if (block > max_block) block = max_block;
    pl.grid = (int)grid;
    pl.block = (int)block;
    const long long n_groups = grid * block / G;
    pl.iters = (int)((tuples + n_groups - 1) / n_groups);
    const size_t row_bytes = (size_t)m.stride * sizeof(float);
    const size_t cap = (size_t)smem_optin - 1024;
    if ((size_t)m.N * row_bytes <= cap) {
        pl.rows_per_chunk = m.N;
        pl.n_chunks = 1;
    } else {
        pl.rows_per_chunk = (int)(cap / row_bytes) / 256 * 256;
        pl.n_chunks = (m.N + pl.rows_per_chunk - 1) / pl.rows_per_chunk;
    }
    pl.smem = (size_t)pl.rows_per_chunk * row_bytes;
    return pl;
}

static PolyDev poly_dev(const PolyModel &m) {
    PolyDev d;
    d.rows = m.rows;
    d.N = m.N;
    d.flags = m.flags;
    for (int k = 0; k < 8; ++k) d.prior_mean[k] = m.prior_mean[k], d.prior_inv_var[k] = m.prior_inv_var[k];
    return d;
}

// The constant bank holds the rows of ONE polynomial model per device.  A launch of another model rebinds
// it after a device-wide synchronize (kernels of the previous owner may still be reading); the mutex is held
// across the launch so that a concurrent caller cannot rebind between the check and the launch.  A sampler
// that keeps using one model never pays for this.
static std::mutex g_crow_mutex;
static unsigned long long g_crow_owner[64];

template <typename Kern, typename Args>
static int poly_launch_one(Kern kern, const PolyModel &m, const Args &a, const PolyPlan &pl,
                           cudaStream_t s) {
    BINFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    if (pl.ur) {
        int dev = 0;
        BINFB_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(g_crow_mutex);
        if (g_crow_owner[dev & 63] != m.uid) {
            BINFB_CUDA(cudaDeviceSynchronize());
            // device-to-device copies do not synchronize with the host and the legacy stream does not order
            // with the (non-blocking) caller streams: copy on the launch stream and wait for it, so that a
            // later launch of this model on any other stream finds the rows in place
            BINFB_CUDA(cudaMemcpyToSymbolAsync(c_poly_rows, m.rows, (size_t)m.N * m.stride * sizeof(float), 0,
                                               cudaMemcpyDeviceToDevice, s));
            BINFB_CUDA(cudaStreamSynchronize(s));
            g_crow_owner[dev & 63] = m.uid;
        }
        kern<<<pl.grid, pl.block, pl.smem, s>>>(poly_dev(m), a, pl.iters, pl.rows_per_chunk, pl.n_chunks);
        BINFB_CUDA(cudaGetLastError());
        return BINFB_OK;
    }
    kern<<<pl.grid, pl.block, pl.smem, s>>>(poly_dev(m), a, pl.iters, pl.rows_per_chunk, pl.n_chunks);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

#define POLY_DISPATCH_G(KERNEL, K_, J_)                                                     \
    switch (pl.G) {                                                                         \
        case 1: return poly_launch_one(KERNEL<K_, 1, J_>, m, a, pl, s);                      \
        case 4: return poly_launch_one(KERNEL<K_, 4, J_>, m, a, pl, s);                      \
        case 32: return poly_launch_one(KERNEL<K_, 32, J_>, m, a, pl, s);                    \
        default: break;                                                                     \
    }

#define POLY_DISPATCH_G_FULL(KERNEL, K_, J_)                                                \
    switch (pl.G) {                                                                         \
        case 1: return poly_launch_one(KERNEL<K_, 1, J_>, m, a, pl, s);                      \
        case 2: return poly_launch_one(KERNEL<K_, 2, J_>, m, a, pl, s);                      \
        case 4: return poly_launch_one(KERNEL<K_, 4, J_>, m, a, pl, s);                      \
        case 8: return poly_launch_one(KERNEL<K_, 8, J_>, m, a, pl, s);                      \
        case 16: return poly_launch_one(KERNEL<K_, 16, J_>, m, a, pl, s);                    \
        case 32: return poly_launch_one(KERNEL<K_, 32, J_>, m, a, pl, s);                    \
        default: break;                                                                     \
    }

#define POLY_DISPATCH(KERNEL)                                                               \
    switch (m.K) {                                                                          \
        case 1: POLY_DISPATCH_G(KERNEL, 1, 1) break;                                         \
        case 2: POLY_DISPATCH_G(KERNEL, 2, 1) break;                                         \
        case 3: POLY_DISPATCH_G(KERNEL, 3, 1) break;                                         \
        case 4:                                                                             \
            if (pl.J == 4) {                                                                \
                if (pl.G == 8) return poly_launch_one(KERNEL<4, 8, 4>, m, a, pl, s);         \
                return poly_launch_one(KERNEL<4, 4, 4>, m, a, pl, s);                        \
            } else if (pl.J == 2 && pl.ur) {                                                \
                switch (pl.G) {                                                             \
                    case 1: return poly_launch_one(KERNEL<4, 1, 2, true>, m, a, pl, s);      \
                    case 2: return poly_launch_one(KERNEL<4, 2, 2, true>, m, a, pl, s);      \
                    case 4: return poly_launch_one(KERNEL<4, 4, 2, true>, m, a, pl, s);      \
                    case 8: return poly_launch_one(KERNEL<4, 8, 2, true>, m, a, pl, s);      \
                    default: break;                                                         \
                }                                                                           \
            } else if (pl.J == 2) {                                                         \
                POLY_DISPATCH_G_FULL(KERNEL, 4, 2)                                           \
            } else {                                                                        \
                POLY_DISPATCH_G_FULL(KERNEL, 4, 1)                                           \
            }                                                                               \
            break;                                                                          \
        case 5: POLY_DISPATCH_G(KERNEL, 5, 1) break;                                         \
        case 6: POLY_DISPATCH_G(KERNEL, 6, 1) break;                                         \
        case 7: POLY_DISPATCH_G(KERNEL, 7, 1) break;                                         \
        case 8: POLY_DISPATCH_G(KERNEL, 8, 1) break;                                         \
        default: break;                                                                     \
    }

int poly_hmc_launch(const PolyModel &m, const HmcArgs &a, int sm_count, int smem_optin,
                    cudaStream_t s) {
    const PolyPlan pl = poly_plan(m, a.C, sm_count, smem_optin);
    POLY_DISPATCH(poly_hmc_kernel)
    set_error("polynomial model: unsupported (n_coeff, group) combination");
    return BINFB_EUNSUPPORTED;
}

int poly_grad_launch(const PolyModel &m, const GradArgs &a, int sm_count, int smem_optin,
                     cudaStream_t s) {
    const PolyPlan pl = poly_plan(m, a.C, sm_count, smem_optin);
    POLY_DISPATCH(poly_grad_kernel)
    set_error("polynomial model: unsupported (n_coeff, group) combination");
    return BINFB_EUNSUPPORTED;
}

int poly_forward_launch(const PolyModel &m, const float *q, int C, float *mock, cudaStream_t s) {
    const long long total = (long long)C * m.N;
    const int block = 256;
    const int grid = (int)((total + block - 1) / block);
    switch (m.K) {
#define FW(K_) case K_: poly_forward_kernel<K_><<<grid, block, 0, s>>>(m.rows, m.N, q, C, mock); break;
        FW(1) FW(2) FW(3) FW(4) FW(5) FW(6) FW(7) FW(8)
#undef FW
        default:
            set_error("polynomial model: n_coeff must be 1..8");
            return BINFB_EUNSUPPORTED;
    }
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

}  // namespace binfb
