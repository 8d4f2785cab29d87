// Random-walk Metropolis sampler and posterior-predictive density of the shipped example, batched
// over chains -- sm_100a.
//
// Reference: RWMCSampler.sample (binf/example/samplers.py:78-92)
//     E_old = -pdf.log_prob(coefficients=state)
//     proposal = state + uniform(-stepsize, stepsize, size=len(state))
//     E_new = -pdf.log_prob(coefficients=proposal)
//     accepted = random() < exp(-(E_new - E_old))
// and predict (binf/example/misc.py:3-16): the mean over posterior samples of the Gaussian density of
// a new datum, N(y; polynomial(x, coefficients), 1/precision).
//
// A move is: propose (one pass over [C, D]), the model's fused log-prob pass on the proposals
// (binfb_logprob_grad without gradient), accept (one pass).  E_old is evaluated once per call and then
// carried (the state does not change between moves of one call, so this is what the reference
// recomputes every time).
#include <math.h>

#include "common.cuh"
#include "internal.h"

namespace binfb {

enum : uint32_t { RNG_RW_STEP = 5, RNG_RW_ACCEPT = 6 };

__global__ void rwmc_propose_kernel(const float *q, const float *stepsize, long long total, int D, uint64_t seed,
                                    uint64_t draw, uint64_t chain_base, const float *change, float *q_prop) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long c = e / D;
        const uint32_t k = (uint32_t)(e - c * D);
        float d;
        if (change) d = change[e];
        else {
            const u32x4 r = philox4x32_10(seed ^ (draw >> 32) * 0x9E3779B97F4A7C15ull, chain_base + c, (uint32_t)draw,
                                          (RNG_RW_STEP << 24) | k);
            d = (2.0f * u32_to_unit(r.x) - 1.0f) * stepsize[c];   // uniform in [-stepsize, stepsize)
        }
        q_prop[e] = q[e] + d;
    }
}

// every thread of a chain reaches the same decision (same Philox draw); lp_in/lp_out ping-pong so that
// no thread reads a value another one has already replaced
__global__ void rwmc_accept_kernel(float *q, const float *q_prop, const double *lp_in, const double *lp_new,
                                   double *lp_out, const float *u_in, long long total, int D, uint64_t seed,
                                   uint64_t draw, uint64_t chain_base, uint8_t *accepted, int32_t *n_accepted) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long c = e / D;
        float u;
        if (u_in) u = u_in[c];
        else {
            const u32x4 r = philox4x32_10(seed ^ (draw >> 32) * 0x9E3779B97F4A7C15ull, chain_base + c, (uint32_t)draw,
                                          RNG_RW_ACCEPT << 24);
            u = u32_to_unit(r.x);
        }
        const double e_old = -lp_in[c], e_new = -lp_new[c];
        const double dE = e_new - e_old;
        // samplers.py:86; NaN energies reject
        const bool acc = (dE == dE) && ((double)u < exp(fmin(709.0, fmax(-745.0, -dE))));
        if (acc) q[e] = q_prop[e];
        if (e - c * D == 0) {
            lp_out[c] = acc ? lp_new[c] : lp_in[c];
            if (accepted) accepted[c] = acc ? 1 : 0;
            if (n_accepted) n_accepted[c] += acc ? 1 : 0;
        }
    }
}

// out[g] = (1/S) sum_s exp(-0.5 tau_s (poly(x_g; c_s) - y_g)^2 + 0.5 log tau_s - 0.5 log 2 pi), evaluated as a
// log-sum-exp like the reference (misc.py:9,16); one block per point, float64
__global__ void __launch_bounds__(256) predictive_kernel(const float *coeffs, const float *tau, long long S, int K,
                                                         const double *xs, const double *ys, double *out) {
    __shared__ double sh_m[8], sh_s[8];
    const double x = xs[blockIdx.x], y = ys[blockIdx.x];
    double m = -INFINITY, s = 0.0;
    for (long long i = threadIdx.x; i < S; i += blockDim.x) {
        const float *c = coeffs + i * K;
        double v = (double)c[K - 1];
        for (int k = K - 2; k >= 0; --k) v = v * x + (double)c[k];      // polyval, ascending coefficients
        const double t = (double)tau[i];
        const double r = v - y;
        const double f = -0.5 * r * r * t + 0.5 * log(t) - 0.9189385332046727;   // 0.5 log(2 pi)
        if (f > m) s = s * exp(m - f) + 1.0, m = f;
        else if (f == f && f > -INFINITY) s += exp(f - m);
    }
    for (int o = 16; o >= 1; o >>= 1) {
        const double m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const double mm = fmax(m, m2);
        s = (mm == -INFINITY) ? 0.0 : s * exp(m - mm) + s2 * exp(m2 - mm);
        m = mm;
    }
    if ((threadIdx.x & 31) == 0) sh_m[threadIdx.x >> 5] = m, sh_s[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double M = -INFINITY, T = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) M = fmax(M, sh_m[w]);
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
            if (sh_m[w] > -INFINITY) T += sh_s[w] * exp(sh_m[w] - M);
        out[blockIdx.x] = (M == -INFINITY) ? 0.0 : exp(M + log(T)) / (double)S;
    }
}

}  // namespace binfb

using namespace binfb;

static int rw_reserve(binfb_model *m, int C) {
    const size_t need = (size_t)C;
    if (need <= m->rw_cap) return BINFB_OK;
    cudaFree(m->rw_prop), cudaFree(m->rw_lp[0]), cudaFree(m->rw_lp[1]), cudaFree(m->rw_lp[2]);
    m->rw_prop = nullptr, m->rw_lp[0] = m->rw_lp[1] = m->rw_lp[2] = nullptr, m->rw_cap = 0;
    BINFB_CUDA(cudaMalloc(&m->rw_prop, need * m->dim * sizeof(float)));
    for (int i = 0; i < 3; ++i) BINFB_CUDA(cudaMalloc(&m->rw_lp[i], need * sizeof(double)));
    m->rw_cap = need;
    return BINFB_OK;
}

extern "C" {

int binfb_rwmc_run(binfb_model *m, float *q_dev, const float *tau_dev, const float *beta_dev,
                   const float *stepsize_dev, int n_chains, int n_moves, uint64_t seed, uint64_t draw,
                   uint64_t chain_base, const float *change_dev, const float *u_dev, uint8_t *accepted_dev,
                   int32_t *n_accepted_dev, double *logp_dev, void *stream) {
    BINFB_TRACE();
    if (!m) {
        set_error("null model handle");
        return BINFB_EINVAL;
    }
    if (!q_dev || !tau_dev || !stepsize_dev || n_chains < 1 || n_moves < 1) {
        set_error("rwmc_run: q, tau, stepsize required; n_chains, n_moves >= 1");
        return BINFB_EINVAL;
    }
    if ((change_dev || u_dev) && n_moves != 1) {
        set_error("rwmc_run: injected proposals / uniforms require n_moves == 1");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(m->device));
    int rc = rw_reserve(m, n_chains);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int C = n_chains, D = m->dim;
    const long long total = (long long)C * D;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)m->sm_count * 8) blocks = (long long)m->sm_count * 8;
    if (n_accepted_dev) BINFB_CUDA(cudaMemsetAsync(n_accepted_dev, 0, (size_t)C * sizeof(int32_t), s));
    // E_old at the incoming state (samplers.py:80)
    rc = binfb_logprob_grad(m, q_dev, tau_dev, beta_dev, C, m->rw_lp[0], nullptr, nullptr, stream);
    if (rc) return rc;
    int cur = 0;
    for (int k = 0; k < n_moves; ++k) {
        rwmc_propose_kernel<<<(unsigned)blocks, 256, 0, s>>>(q_dev, stepsize_dev, total, D, seed, draw + k, chain_base,
                                                             change_dev, m->rw_prop);
        rc = binfb_logprob_grad(m, m->rw_prop, tau_dev, beta_dev, C, m->rw_lp[2], nullptr, nullptr, stream);
        if (rc) return rc;
        rwmc_accept_kernel<<<(unsigned)blocks, 256, 0, s>>>(q_dev, m->rw_prop, m->rw_lp[cur], m->rw_lp[2],
                                                            m->rw_lp[cur ^ 1], u_dev, total, D, seed, draw + k,
                                                            chain_base, accepted_dev, n_accepted_dev);
        cur ^= 1;
    }
    if (logp_dev)
        BINFB_CUDA(cudaMemcpyAsync(logp_dev, m->rw_lp[cur], (size_t)C * sizeof(double), cudaMemcpyDeviceToDevice, s));
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

int binfb_rwmc_run_host(binfb_model *m, float *q, const float *tau, const float *beta, const float *stepsize,
                        int n_chains, int n_moves, uint64_t seed, uint64_t draw, uint64_t chain_base,
                        const float *change, const float *u, uint8_t *accepted, int32_t *n_accepted, double *logp) {
    BINFB_TRACE();
    if (!m) {
        set_error("null model handle");
        return BINFB_EINVAL;
    }
    if (!q || !tau || !stepsize || n_chains < 1) {
        set_error("rwmc_run_host: q, tau, stepsize required; n_chains >= 1");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(m->device));
    const size_t C = n_chains, D = m->dim;
    char *buf = nullptr;
    const size_t oq = 0, och = oq + C * D * 4, ot = och + C * D * 4, ob = ot + C * 4, os = ob + C * 4, ou = os + C * 4,
                 ol = (ou + C * 4 + 7) / 8 * 8, on = ol + C * 8, oa = on + C * 4, end = oa + C;
    BINFB_CUDA(cudaMalloc(&buf, end));
    cudaStream_t s = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    auto up = [&](size_t off, const void *src, size_t bytes) {
        if (e == cudaSuccess && src) e = cudaMemcpyAsync(buf + off, src, bytes, cudaMemcpyHostToDevice, s);
    };
    up(oq, q, C * D * 4), up(och, change, C * D * 4), up(ot, tau, C * 4), up(ob, beta, C * 4);
    up(os, stepsize, C * 4), up(ou, u, C * 4);
    int rc = BINFB_OK;
    if (e == cudaSuccess)
        rc = binfb_rwmc_run(m, (float *)(buf + oq), (const float *)(buf + ot), beta ? (const float *)(buf + ob) : nullptr,
                            (const float *)(buf + os), n_chains, n_moves, seed, draw, chain_base,
                            change ? (const float *)(buf + och) : nullptr, u ? (const float *)(buf + ou) : nullptr,
                            (uint8_t *)(buf + oa), (int32_t *)(buf + on), (double *)(buf + ol), s);
    auto down = [&](void *dst, size_t off, size_t bytes) {
        if (rc == BINFB_OK && e == cudaSuccess && dst) e = cudaMemcpyAsync(dst, buf + off, bytes, cudaMemcpyDeviceToHost, s);
    };
    down(q, oq, C * D * 4), down(accepted, oa, C), down(n_accepted, on, C * 4), down(logp, ol, C * 8);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (s) cudaStreamDestroy(s);
    cudaFree(buf);
    if (e != cudaSuccess) return cuda_fail(e, "rwmc_run_host");
    return rc;
}

int binfb_posterior_predictive_host(const float *coeffs, const float *precision, long long n_samples, int n_coeff,
                                    const double *xs, const double *ys, int n_points, double *out, int device) {
    if (!coeffs || !precision || !xs || !ys || !out || n_samples < 1 || n_coeff < 1 || n_points < 1) {
        set_error("posterior_predictive: all pointers required; n_samples, n_coeff, n_points >= 1");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(device));
    float *dc = nullptr, *dt = nullptr;
    double *dx = nullptr;
    const size_t S = (size_t)n_samples, G = (size_t)n_points;
    cudaError_t e = cudaMalloc(&dc, S * n_coeff * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dt, S * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dx, 3 * G * 8);
    if (e == cudaSuccess) e = cudaMemcpy(dc, coeffs, S * n_coeff * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dt, precision, S * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dx, xs, G * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dx + G, ys, G * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) {
        predictive_kernel<<<(unsigned)G, 256>>>(dc, dt, n_samples, n_coeff, dx, dx + G, dx + 2 * G);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dx + 2 * G, G * 8, cudaMemcpyDeviceToHost);
    cudaFree(dc), cudaFree(dt), cudaFree(dx);
    if (e != cudaSuccess) return cuda_fail(e, "posterior_predictive");
    return BINFB_OK;
}

}  // extern "C"
