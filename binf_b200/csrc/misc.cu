// Small kernels: conjugate precision update (GammaSampler.sample,
// binf/example/samplers.py:27-51), RNG stream dump for statistical tests, replica-exchange
// decision / apply (build-defined), FP32 / MUFU issue-rate microbenchmarks.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "internal.h"

namespace binfb {

__global__ void gibbs_tau_kernel(const double *chi2, float *tau, const float *beta, int C,
                                 double n_data, double shape0, double rate0, uint64_t seed,
                                 uint64_t draw, uint64_t chain_base, const double *gamma_draws) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double b = beta ? (double)beta[c] : 1.0;
    const double shape = 0.5 * b * n_data + shape0 - 1.0;  // quirk Q3: "- 1" (samplers.py:32)
    const double rate = 0.5 * b * chi2[c] + rate0;         // samplers.py:36-41
    const double g = gamma_draws ? gamma_draws[c] : rng_gamma(seed, chain_base + c, draw, shape);
    tau[c] = (float)(g / rate);                            // samplers.py:47
}

int gibbs_tau_launch(const double *chi2, float *tau, const float *beta, int C, double n_data,
                     double shape, double rate, uint64_t seed, uint64_t draw, uint64_t chain_base,
                     const double *gamma_draws, cudaStream_t s) {
    gibbs_tau_kernel<<<(C + 127) / 128, 128, 0, s>>>(chi2, tau, beta, C, n_data, shape, rate, seed,
                                                     draw, chain_base, gamma_draws);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

__global__ void rng_fill_kernel(uint64_t seed, uint64_t draw, uint64_t chain_base, int C, int D,
                                double gamma_shape, float *normals, float *uniforms, double *gammas) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (normals && i < (long long)C * D) {
        const int c = (int)(i / D), e = (int)(i % D);
        normals[i] = rng_normal(seed, chain_base + c, draw, (uint32_t)e);
    }
    if (i < C) {
        if (uniforms) {
            const u32x4 r = philox4x32_10(seed ^ (draw >> 32) * 0x9E3779B97F4A7C15ull, chain_base + i,
                                          (uint32_t)draw, (uint32_t)RNG_ACCEPT << 24);
            uniforms[i] = u32_to_unit_open0(r.x);
        }
        if (gammas) gammas[i] = rng_gamma(seed, chain_base + i, draw, gamma_shape);
    }
}

int rng_fill_launch(uint64_t seed, uint64_t draw, uint64_t chain_base, int C, int D,
                    double gamma_shape, float *normals, float *uniforms, double *gammas,
                    cudaStream_t s) {
    const long long total = std::max((long long)C * D, (long long)C);
    rng_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(seed, draw, chain_base, C, D,
                                                                   gamma_shape, normals, uniforms, gammas);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

__global__ void swap_decide_kernel(const double *ll_a, const double *ll_b, double beta_a,
                                   double beta_b, int C, uint64_t seed, uint64_t attempt,
                                   uint64_t pair_id, uint64_t chain_base, uint8_t *accept) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    // Delta = (beta_a - beta_b)(l_a - l_b); accept iff u < exp(-Delta)   
    const double delta = (beta_a - beta_b) * (ll_a[c] - ll_b[c]);
    const u32x4 r = philox4x32_10(seed ^ pair_id * 0xD6E8FEB86659FD93ull, chain_base + c,
                                  (uint32_t)attempt, ((uint32_t)RNG_SWAP << 24) | (uint32_t)(attempt >> 32));
    const double u = (double)u32_to_unit_open0(r.x);
    accept[c] = (delta == delta) && (u < exp(fmin(709.0, fmax(-308.0, -delta)))) ? 1 : 0;
}

int swap_decide_launch(const double *ll_a, const double *ll_b, double beta_a, double beta_b, int C,
                       uint64_t seed, uint64_t attempt, uint64_t pair_id, uint64_t chain_base,
                       uint8_t *accept, cudaStream_t s) {
    swap_decide_kernel<<<(C + 127) / 128, 128, 0, s>>>(ll_a, ll_b, beta_a, beta_b, C, seed, attempt,
                                                       pair_id, chain_base, accept);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

__global__ void swap_apply_kernel(float *q_mine, const float *q_theirs, float *eps_mine,
                                  const float *eps_theirs, const uint8_t *accept, int C, int D) {
    const int c = blockIdx.y;
    if (!accept[c]) return;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < D; e += gridDim.x * blockDim.x)
        q_mine[(size_t)c * D + e] = q_theirs[(size_t)c * D + e];
    if (eps_mine && eps_theirs && blockIdx.x == 0 && threadIdx.x == 0) eps_mine[c] = eps_theirs[c];
}

int swap_apply_launch(float *q_mine, const float *q_theirs, float *eps_mine,
                      const float *eps_theirs, const uint8_t *accept, int C, int D,
                      cudaStream_t s) {
    dim3 grid((unsigned)std::min(8, (D + 255) / 256), (unsigned)C);
    swap_apply_kernel<<<grid, 256, 0, s>>>(q_mine, q_theirs, eps_mine, eps_theirs, accept, C, D);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

// ---------------------------------------------------------------------------------------------
// Replica exchange by LABEL swap (build-defined, SURVEY.md A.3 / 8e: "swap beta (and eps) labels so no
// state moves").  The ensemble is a grid [temperature k][column c]; every chain of every rank holds one
// cell, its temperature index tidx, and never moves.  An attempt:
//   1. rex_pack_kernel: one 16-byte record {log L, tidx, eps} per chain (log L from chi^2 of the current
//      state, which the trajectory kernel already left behind -- no extra pair sweep);
//   2. the records of all ranks are all-gathered (world x C x 16 bytes; NCCL over NVLink);
//   3. rex_decide_kernel: every chain looks up the chain of ITS column that holds the neighbouring
//      temperature (even attempts pair (0,1),(2,3).., odd attempts (1,2),(3,4)..), evaluates
//      u < exp(-(beta_k - beta_k')(l_mine - l_theirs)) with a Philox draw keyed by (seed, attempt, lower
//      temperature index, column) -- NOT by rank or chain base, so both partners reach the same decision --
//      and on acceptance adopts the partner's temperature index, beta and step size.
// ---------------------------------------------------------------------------------------------
struct RexRecord {
    double ll;
    int32_t tidx;
    float eps;
};
static_assert(sizeof(RexRecord) == 16, "RexRecord is the 16-byte wire format of binfb_rex_pack");

__global__ void rex_pack_kernel(const double *chi2, const float *tau, const float *eps, const int32_t *tidx,
                                int C, double n_data, RexRecord *rec) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double t = (double)tau[c];
    RexRecord r;
    r.ll = -0.5 * t * chi2[c] + 0.5 * n_data * log(t);  // GaussianErrorModel, binf/example/likelihood.py:54-57
    r.tidx = tidx[c];
    r.eps = eps[c];
    rec[c] = r;
}

int rex_pack_launch(const double *chi2, const float *tau, const float *eps, const int32_t *tidx, int C,
                    double n_data, void *rec, cudaStream_t s) {
    rex_pack_kernel<<<(C + 127) / 128, 128, 0, s>>>(chi2, tau, eps, tidx, C, n_data, (RexRecord *)rec);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

struct RexLadder {
    double beta[BINFB_REX_MAX_TEMPS];
};

__global__ void rex_decide_kernel(const RexRecord *all, int world, int rank, int C, int n_columns,
                                  RexLadder lad, int n_temps, uint64_t seed, uint64_t attempt, double ll_shift,
                                  int32_t *tidx, float *beta, float *eps, uint8_t *accept,
                                  unsigned long long *pair_counts, double *temp_stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    const RexRecord me = all[(size_t)rank * C + i];
    const int k = me.tidx, col = i % n_columns, rows = C / n_columns;
    if (temp_stats && k >= 0 && k < n_temps && me.ll == me.ll) {
        const double d = me.ll - ll_shift;
        atomicAdd(temp_stats + 3 * k, 1.0), atomicAdd(temp_stats + 3 * k + 1, d);
        atomicAdd(temp_stats + 3 * k + 2, d * d);
    }
    const int kp = ((k + (int)(attempt & 1u)) & 1) == 0 ? k + 1 : k - 1;  // pairs (k, k+1) with k + attempt even
    bool acc = false;
    RexRecord other = me;
    if (k >= 0 && k < n_temps && kp >= 0 && kp < n_temps) {
        bool found = false;
        for (int r = 0; r < world && !found; ++r)
            for (int row = 0; row < rows; ++row) {
                const RexRecord o = all[(size_t)r * C + (size_t)row * n_columns + col];
                if (o.tidx == kp) {
                    other = o, found = true;
                    break;
                }
            }
        if (found) {
            const int lo = k < kp ? k : kp;
            // Delta = (beta_k - beta_k')(l_mine - l_theirs) is symmetric in the two partners
            const double delta = (lad.beta[k] - lad.beta[kp]) * (me.ll - other.ll);
            const u32x4 rr = philox4x32_10(seed ^ (uint64_t)(lo + 1) * 0xD6E8FEB86659FD93ull, (uint64_t)col,
                                           (uint32_t)attempt, ((uint32_t)RNG_SWAP << 24) | (uint32_t)(attempt >> 32 & 0xffffffu));
            const double u = (double)u32_to_unit_open0(rr.x);
            acc = (delta == delta) && (u < exp(fmin(709.0, fmax(-308.0, -delta))));
            if (pair_counts && k == lo) {
                atomicAdd(pair_counts + 2 * lo, 1ull);
                if (acc) atomicAdd(pair_counts + 2 * lo + 1, 1ull);
            }
        }
    }
    if (acc) {
        tidx[i] = kp;
        beta[i] = (float)lad.beta[kp];
        eps[i] = other.eps;
    }
    if (accept) accept[i] = acc ? 1 : 0;
}

int rex_decide_launch(const void *all, int world, int rank, int C, int n_columns, const double *betas, int n_temps,
                      uint64_t seed, uint64_t attempt, double ll_shift, int32_t *tidx, float *beta, float *eps,
                      uint8_t *accept, unsigned long long *pair_counts, double *temp_stats, cudaStream_t s) {
    RexLadder lad;
    for (int k = 0; k < BINFB_REX_MAX_TEMPS; ++k) lad.beta[k] = k < n_temps ? betas[k] : 0.0;
    rex_decide_kernel<<<(C + 127) / 128, 128, 0, s>>>((const RexRecord *)all, world, rank, C, n_columns, lad, n_temps,
                                                      seed, attempt, ll_shift, tidx, beta, eps, accept, pair_counts,
                                                      temp_stats);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

// cold[c] = q[c] where tidx[c] == k_sel, else 0: summed over the ranks (exactly one contributor per column)
// this assembles the states of one temperature wherever they currently live
__global__ void rex_select_kernel(const float *q, const float *aux, const int32_t *tidx, int k_sel, int C, int D,
                                  int n_columns, float *out_q, float *out_aux) {
    const int c = blockIdx.y;
    if (tidx[c] != k_sel) return;
    const int col = c % n_columns;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < D; e += gridDim.x * blockDim.x)
        out_q[(size_t)col * D + e] = q[(size_t)c * D + e];
    if (aux && out_aux && blockIdx.x == 0 && threadIdx.x == 0) out_aux[col] = aux[c];
}

int rex_select_launch(const float *q, const float *aux, const int32_t *tidx, int k_sel, int C, int D, int n_columns,
                      float *out_q, float *out_aux, cudaStream_t s) {
    BINFB_CUDA(cudaMemsetAsync(out_q, 0, (size_t)n_columns * D * sizeof(float), s));
    if (out_aux) BINFB_CUDA(cudaMemsetAsync(out_aux, 0, (size_t)n_columns * sizeof(float), s));
    dim3 grid((unsigned)std::min(8, (D + 255) / 256), (unsigned)C);
    rex_select_kernel<<<grid, 256, 0, s>>>(q, aux, tidx, k_sel, C, D, n_columns, out_q, out_aux);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

// ---------------------------------------------------------------------------------------------
// microbenchmarks: the non-tensor FP32 peak (scalar FFMA and packed FFMA2) and the MUFU rate
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) mb_ffma_kernel(float *out, int iters, float a, float b) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(512) mb_ffma2_kernel(float *out, int iters, float a, float b) {
    float2 v[16];
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2((float)(threadIdx.x + i), (float)i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __ffma2_rn(v[i], aa, bb);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(512) mb_mufu_kernel(float *out, int iters) {
    float v[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = 1.0f + 0.01f * (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < 12; i += 3) {
                v[i] = mufu_rsqrt(v[i]);
                v[i + 1] = mufu_ex2(v[i + 1]);
                v[i + 2] = mufu_lg2(v[i + 2]);  // not rcp: ptxas cancels rcp(rcp(x))
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}

template <typename F>
static int time_kernel(F launch, float *ms_out) {
    cudaEvent_t e0, e1;
    BINFB_CUDA(cudaEventCreate(&e0));
    BINFB_CUDA(cudaEventCreate(&e1));
    launch();  // warm-up
    BINFB_CUDA(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        BINFB_CUDA(cudaEventRecord(e0));
        launch();
        BINFB_CUDA(cudaEventRecord(e1));
        BINFB_CUDA(cudaEventSynchronize(e1));
        float ms;
        BINFB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    *ms_out = best;
    return BINFB_OK;
}

int microbench_run(int device, int iters, double *ffma, double *ffma2, double *mufu,
                   double *clock_mhz) {
    BINFB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    BINFB_CUDA(cudaGetDeviceProperties(&prop, device));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, device);
    if (clock_mhz) *clock_mhz = clk_khz / 1000.0;
    float *out;
    BINFB_CUDA(cudaMalloc(&out, 16));
    const int grid = prop.multiProcessorCount * 4, block = 512;
    const double lanes = (double)grid * block;
    float ms;
    int rc;
    if (ffma) {
        rc = time_kernel([&] { mb_ffma_kernel<<<grid, block>>>(out, iters, 1.0001f, 0.5f); }, &ms);
        if (rc) return rc;
        *ffma = lanes * iters * 128.0 * 2.0 / (ms * 1e-3) / 1e12;
    }
    if (ffma2) {
        rc = time_kernel([&] { mb_ffma2_kernel<<<grid, block>>>(out, iters, 1.0001f, 0.5f); }, &ms);
        if (rc) return rc;
        *ffma2 = lanes * iters * 128.0 * 4.0 / (ms * 1e-3) / 1e12;
    }
    if (mufu) {
        rc = time_kernel([&] { mb_mufu_kernel<<<grid, block>>>(out, iters); }, &ms);
        if (rc) return rc;
        *mufu = lanes * iters * 48.0 / (ms * 1e-3) / 1e9;
    }
    BINFB_CUDA(cudaGetLastError());
    cudaFree(out);
    return BINFB_OK;
}

}  // namespace binfb
