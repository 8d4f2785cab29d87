// Sample sink: what the reference's driver loop does with the state after every Gibbs sweep, for
// thousands of chains resident in HBM -- sm_100a.
//
// Reference: `samples.append(deepcopy(gips.sample()))` (example_script.py:32-34), the thinning slice
// `samples[20000::20]` (example_script.py:41), the MAP estimate = kept sample of maximum
// log-probability (binf/example/misc.py:18-22), and the posterior summaries the plots derive from
// the kept samples (binf/example/plots.py).  The reference keeps a Python list of deep copies of
// one chain; here one fused pass over the state per sweep
//   * copies it into a ring buffer of thinned samples        (burn-in / thinning as in the slice),
//   * updates per-chain running moments (Welford, float64)    (every post-burn-in sweep),
//   * keeps each chain's best (maximum log-probability) state (the MAP candidate),
// and a reduction over the chains turns the moments into per-dimension posterior mean, variance,
// split-free R-hat and the effective sample size per chain.
//
// This stage is HBM-bound: 4 B read + 32 B read-modify-write (+ 4 B ring write) per state element
// and sweep, no reuse.  The pass is a flat grid-stride loop of 16-byte accesses sized to the SM
// count; state and ring use streaming (evict-first) accesses, the moments default caching.
#include <math.h>

#include <string>

#include "internal.h"

struct binfb_sink {
    int device = 0;
    int C = 0, D = 0, capacity = 0, burn_in = 0, thin = 1;
    unsigned flags = 0;
    long long n_pushed = 0, n_moment = 0, n_kept = 0;
    int sm_count = 0;
    double *mean = nullptr, *m2 = nullptr;       // [C, D]
    float *ring_q = nullptr, *ring_aux = nullptr;  // [capacity, C, D], [capacity, C]
    float *map_q = nullptr, *map_aux = nullptr;    // [C, D], [C]
    double *best[2] = {nullptr, nullptr};          // [C] ping-pong
    int best_cur = 0;
    double *acc = nullptr;                         // [3, D] summary accumulators + [D] pivot
    cudaStream_t hstream = nullptr;
};

namespace binfb {

struct SinkPush {
    const float *q, *aux;
    const double *logp;
    double *mean, *m2;
    float *ring_q, *ring_aux, *map_q, *map_aux;
    const double *best_in;
    double *best_out;
    long long total;  // C * D
    int C, D;
    double inv_n;
    int do_moments;
};

template <int V>
struct VecF;
template <>
struct VecF<4> {
    typedef float4 T;
};
template <>
struct VecF<1> {
    typedef float T;
};

__device__ __forceinline__ void unpack(const float4 v, float (&x)[4]) { x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w; }
__device__ __forceinline__ void unpack(const float v, float (&x)[1]) { x[0] = v; }

// V = 4 when dim % 4 == 0 (a 16-byte group never straddles two chains), else 1.
// A block walks whole chains (several at a time when a chain is shorter than the block), so the chain
// index and the per-chain scalars cost nothing per element.
template <int V>
__global__ void __launch_bounds__(256) sink_push_kernel(SinkPush a) {
    typedef typename VecF<V>::T vec;
    const int pc = a.D / V;                         // vectors per chain
    const bool wide = pc >= 256;
    const int rows = wide ? 1 : 256 / pc;           // chains per block iteration
    const int tr = wide ? 0 : (int)threadIdx.x / pc;
    const int tj = wide ? (int)threadIdx.x : (int)threadIdx.x - tr * pc;
    const int jstep = wide ? 256 : pc;
    if (tr >= rows) return;
    for (long long c = (long long)blockIdx.x * rows + tr; c < a.C; c += (long long)gridDim.x * rows) {
        double lp = 0.0, old = 0.0;
        bool better = false;
        if (a.map_q) {
            // strictly greater, NaN never wins: numpy.argmax keeps the FIRST maximum (misc.py:20)
            lp = a.logp[c], old = a.best_in[c];
            better = lp > old;
        }
        if (tj == 0) {  // this thread owns the chain's scalars
            if (a.ring_aux && a.aux) a.ring_aux[c] = a.aux[c];
            if (a.map_q) {
                a.best_out[c] = better ? lp : old;
                if (better && a.map_aux && a.aux) a.map_aux[c] = a.aux[c];
            }
        }
        // U groups per thread and trip: all loads are issued before the first dependent instruction, so a
        // thread keeps U x 80 bytes in flight (the pass has no reuse and is bound by HBM latency x bandwidth)
        constexpr int U = 3;
        for (int j0 = tj; j0 < pc; j0 += U * jstep) {
            float x[U][V];
            double mu[U][V], m2[U][V];
            bool on[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * jstep;
                on[u] = j < pc;
                if (!on[u]) continue;
                const long long i = c * pc + j, e = i * V;
                unpack(__ldcs(reinterpret_cast<const vec *>(a.q) + i), x[u]);
                if (a.do_moments) {
                    if constexpr (V == 4) {
                        const double2 *mp = reinterpret_cast<const double2 *>(a.mean + e);
                        const double2 *sp = reinterpret_cast<const double2 *>(a.m2 + e);
                        const double2 ma = mp[0], mb = mp[1], sa = sp[0], sb = sp[1];
                        mu[u][0] = ma.x, mu[u][1] = ma.y, mu[u][2] = mb.x, mu[u][3] = mb.y;
                        m2[u][0] = sa.x, m2[u][1] = sa.y, m2[u][2] = sb.x, m2[u][3] = sb.y;
                    } else {
                        mu[u][0] = a.mean[e], m2[u][0] = a.m2[e];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!on[u]) continue;
                const long long i = c * pc + j0 + u * jstep, e = i * V;
                if (a.do_moments) {
#pragma unroll
                    for (int k = 0; k < V; ++k) {  // Welford
                        const double d0 = (double)x[u][k] - mu[u][k];
                        mu[u][k] += d0 * a.inv_n;
                        m2[u][k] += d0 * ((double)x[u][k] - mu[u][k]);
                    }
                    if constexpr (V == 4) {
                        double2 *mp = reinterpret_cast<double2 *>(a.mean + e), *sp = reinterpret_cast<double2 *>(a.m2 + e);
                        mp[0] = make_double2(mu[u][0], mu[u][1]), mp[1] = make_double2(mu[u][2], mu[u][3]);
                        sp[0] = make_double2(m2[u][0], m2[u][1]), sp[1] = make_double2(m2[u][2], m2[u][3]);
                    } else {
                        a.mean[e] = mu[u][0], a.m2[e] = m2[u][0];
                    }
                }
                if (a.ring_q) __stcs(reinterpret_cast<vec *>(a.ring_q) + i, *reinterpret_cast<const vec *>(x[u]));
                if (better) reinterpret_cast<vec *>(a.map_q)[i] = *reinterpret_cast<const vec *>(x[u]);
            }
        }
    }
}

__global__ void sink_fill_kernel(double *p, long long n, double v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// per-dimension sums over the chains of (mean - pivot), (mean - pivot)^2 and m2;
// block = 32 dims x 8 chain lanes, grid = (dim tiles, chain chunks)
__global__ void __launch_bounds__(256) sink_reduce_kernel(const double *mean, const double *m2, int C, int D,
                                                          int chunk, double *acc) {
    __shared__ double sh[3][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + tx;
    const int c0 = blockIdx.y * chunk, c1 = min(C, c0 + chunk);
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (d < D) {
        const double pivot = mean[d];  // chain 0: keeps the sum of squares well conditioned
        for (int c = c0 + ty; c < c1; c += 8) {
            const double v = mean[(size_t)c * D + d] - pivot;
            s1 += v, s2 += v * v, s3 += m2[(size_t)c * D + d];
        }
    }
    sh[0][ty][tx] = s1, sh[1][ty][tx] = s2, sh[2][ty][tx] = s3;
    __syncthreads();
    if (ty == 0 && d < D) {
#pragma unroll
        for (int k = 1; k < 8; ++k) s1 += sh[0][k][tx], s2 += sh[1][k][tx], s3 += sh[2][k][tx];
        atomicAdd(acc + d, s1), atomicAdd(acc + D + d, s2), atomicAdd(acc + 2 * (size_t)D + d, s3);
    }
}

// mean, pooled within-chain variance W, R-hat (Gelman-Rubin, no chain splitting) and the
// effective sample size per chain n W / B = W / var_c(chain means)
__global__ void sink_finalize_kernel(const double *acc, const double *mean0, int C, int D, double n,
                                     double *o_mean, double *o_var, double *o_rhat, double *o_ess) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const double s1 = acc[d], s2 = acc[D + d], s3 = acc[2 * (size_t)D + d];
    const double mu = s1 / C;
    const double W = s3 / ((n - 1.0) * C);
    const double var_means = C > 1 ? (s2 - C * mu * mu) / (C - 1.0) : 0.0;  // = B / n
    if (o_mean) o_mean[d] = mean0[d] + mu;
    if (o_var) o_var[d] = W;
    if (o_rhat) o_rhat[d] = sqrt(((n - 1.0) / n * W + var_means) / W);
    if (o_ess) o_ess[d] = W / var_means;
}

__global__ void sink_var_kernel(const double *m2, long long total, double inv_nm1, double *var) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        var[i] = m2[i] * inv_nm1;
}

}  // namespace binfb

using namespace binfb;

static int check_sink(const binfb_sink *s) {
    if (!s) {
        set_error("null sink handle");
        return BINFB_EINVAL;
    }
    return BINFB_OK;
}

extern "C" {

int binfb_sink_create(int n_chains, int dim, int capacity, int burn_in, int thin, unsigned flags, int device,
                      binfb_sink **out) {
    if (!out || n_chains < 1 || dim < 1 || capacity < 0 || burn_in < 0 || thin < 1) {
        set_error("sink_create: n_chains, dim >= 1; capacity, burn_in >= 0; thin >= 1");
        return BINFB_EINVAL;
    }
    *out = nullptr;
    BINFB_CUDA(cudaSetDevice(device));
    binfb_sink *s = new binfb_sink();
    s->device = device, s->C = n_chains, s->D = dim, s->capacity = capacity, s->burn_in = burn_in, s->thin = thin;
    s->flags = flags;
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    const size_t CD = (size_t)n_chains * dim;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMalloc(p, bytes);
    };
    alloc((void **)&s->mean, CD * sizeof(double));
    alloc((void **)&s->m2, CD * sizeof(double));
    alloc((void **)&s->ring_q, (size_t)capacity * CD * sizeof(float));
    alloc((void **)&s->ring_aux, (size_t)capacity * n_chains * sizeof(float));
    if (flags & BINFB_SINK_TRACK_MAP) {
        alloc((void **)&s->map_q, CD * sizeof(float));
        alloc((void **)&s->map_aux, (size_t)n_chains * sizeof(float));
        alloc((void **)&s->best[0], (size_t)n_chains * sizeof(double));
        alloc((void **)&s->best[1], (size_t)n_chains * sizeof(double));
    }
    alloc((void **)&s->acc, (size_t)3 * dim * sizeof(double));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->hstream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMemset(s->mean, 0, CD * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(s->m2, 0, CD * sizeof(double));
    if (e == cudaSuccess && s->ring_aux) e = cudaMemset(s->ring_aux, 0, (size_t)capacity * n_chains * sizeof(float));
    if (e == cudaSuccess && s->map_q) {
        e = cudaMemset(s->map_q, 0, CD * sizeof(float));
        if (e == cudaSuccess) e = cudaMemset(s->map_aux, 0, (size_t)n_chains * sizeof(float));
        sink_fill_kernel<<<64, 256>>>(s->best[0], n_chains, -INFINITY);
        sink_fill_kernel<<<64, 256>>>(s->best[1], n_chains, -INFINITY);
    }
    // the memsets above ran on the legacy stream; pushes may use non-blocking streams
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        const int rc = e == cudaErrorMemoryAllocation ? BINFB_ENOMEM : BINFB_ECUDA;
        set_error(std::string("sink_create: ") + cudaGetErrorString(e));
        cudaGetLastError();
        binfb_sink_destroy(s);
        return rc;
    }
    *out = s;
    return BINFB_OK;
}

int binfb_sink_destroy(binfb_sink *s) {
    if (!s) return BINFB_OK;
    cudaSetDevice(s->device);
    cudaFree(s->mean), cudaFree(s->m2), cudaFree(s->ring_q), cudaFree(s->ring_aux), cudaFree(s->map_q);
    cudaFree(s->map_aux), cudaFree(s->best[0]), cudaFree(s->best[1]), cudaFree(s->acc);
    if (s->hstream) cudaStreamDestroy(s->hstream);
    delete s;
    return BINFB_OK;
}

int binfb_sink_info(const binfb_sink *s, long long *n_pushed, long long *n_moment, long long *n_kept) {
    int rc = check_sink(s);
    if (rc) return rc;
    if (n_pushed) *n_pushed = s->n_pushed;
    if (n_moment) *n_moment = s->n_moment;
    if (n_kept) *n_kept = s->n_kept;
    return BINFB_OK;
}

int binfb_sink_push(binfb_sink *s, const float *q_dev, const float *aux_dev, const double *logp_dev, void *stream) {
    BINFB_TRACE();
    int rc = check_sink(s);
    if (rc) return rc;
    if (!q_dev) {
        set_error("sink_push: q is required");
        return BINFB_EINVAL;
    }
    if ((s->flags & BINFB_SINK_TRACK_MAP) && !logp_dev) {
        set_error("sink_push: this sink tracks the MAP state and needs logp");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(s->device));
    const long long t = s->n_pushed;
    const bool post = t >= s->burn_in;
    const bool keep = post && s->capacity > 0 && (t - s->burn_in) % s->thin == 0;
    SinkPush a;
    a.q = q_dev, a.aux = aux_dev, a.logp = logp_dev;
    a.mean = s->mean, a.m2 = s->m2;
    a.total = (long long)s->C * s->D, a.C = s->C, a.D = s->D;
    a.do_moments = post ? 1 : 0;
    a.inv_n = post ? 1.0 / (double)(s->n_moment + 1) : 0.0;
    const long long slot = keep ? s->n_kept % s->capacity : 0;
    a.ring_q = keep ? s->ring_q + (size_t)slot * a.total : nullptr;
    a.ring_aux = keep ? s->ring_aux + (size_t)slot * s->C : nullptr;
    // the reference takes the MAP over the kept (thinned) samples (example_script.py:41-42,51)
    const bool map = (s->flags & BINFB_SINK_TRACK_MAP) && (keep || (post && s->capacity == 0));
    a.map_q = map ? s->map_q : nullptr, a.map_aux = s->map_aux;
    a.best_in = s->best[s->best_cur], a.best_out = s->best[s->best_cur ^ 1];
    if (post || keep || map) {
        const bool vec = s->D % 4 == 0 && ((uintptr_t)q_dev & 15) == 0;
        const int pc = vec ? s->D / 4 : s->D;
        const int rows = pc >= 256 ? 1 : 256 / pc;
        long long blocks = ((long long)s->C + rows - 1) / rows;
        const long long cap = (long long)s->sm_count * 3;  // resident CTAs of 256 threads per SM (80 registers)
        if (blocks > cap) blocks = cap;
        if (vec) sink_push_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
        else sink_push_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
        BINFB_CUDA(cudaGetLastError());
    }
    if (map) s->best_cur ^= 1;
    if (post) s->n_moment += 1;
    if (keep) s->n_kept += 1;
    s->n_pushed += 1;
    return BINFB_OK;
}

int binfb_sink_push_host(binfb_sink *s, const float *q, const float *aux, const double *logp) {
    BINFB_TRACE();
    int rc = check_sink(s);
    if (rc) return rc;
    if (!q) {
        set_error("sink_push: q is required");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(s->device));
    const size_t CD = (size_t)s->C * s->D;
    float *dq = nullptr, *da = nullptr;
    double *dl = nullptr;
    BINFB_CUDA(cudaMalloc(&dq, CD * sizeof(float)));
    // copies and kernel on the same stream: a synchronous cudaMemcpy from pageable memory may return
    // before the DMA has landed, and a non-blocking stream does not wait for the legacy stream
    cudaError_t e = cudaMemcpyAsync(dq, q, CD * sizeof(float), cudaMemcpyHostToDevice, s->hstream);
    if (e == cudaSuccess && aux) {
        e = cudaMalloc(&da, s->C * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpyAsync(da, aux, s->C * sizeof(float), cudaMemcpyHostToDevice, s->hstream);
    }
    if (e == cudaSuccess && logp) {
        e = cudaMalloc(&dl, s->C * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpyAsync(dl, logp, s->C * sizeof(double), cudaMemcpyHostToDevice, s->hstream);
    }
    if (e == cudaSuccess) {
        rc = binfb_sink_push(s, dq, da, dl, s->hstream);
        if (rc == BINFB_OK) e = cudaStreamSynchronize(s->hstream);
    }
    cudaFree(dq), cudaFree(da), cudaFree(dl);
    if (e != cudaSuccess) return cuda_fail(e, "sink_push_host");
    return rc;
}

int binfb_sink_summary(binfb_sink *s, double *mean_dev, double *var_dev, double *rhat_dev, double *ess_dev,
                       void *stream) {
    BINFB_TRACE();
    int rc = check_sink(s);
    if (rc) return rc;
    if (s->n_moment < 2) {
        set_error("sink_summary: needs at least 2 post-burn-in sweeps");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    BINFB_CUDA(cudaMemsetAsync(s->acc, 0, (size_t)3 * s->D * sizeof(double), st));
    const int tiles = (s->D + 31) / 32;
    int chunks = (s->sm_count * 8 + tiles - 1) / tiles;  // enough CTAs to pull full HBM bandwidth
    if (chunks > (s->C + 7) / 8) chunks = (s->C + 7) / 8;
    if (chunks < 1) chunks = 1;
    const int chunk = (s->C + chunks - 1) / chunks;
    dim3 grid(tiles, (s->C + chunk - 1) / chunk);
    sink_reduce_kernel<<<grid, 256, 0, st>>>(s->mean, s->m2, s->C, s->D, chunk, s->acc);
    sink_finalize_kernel<<<(s->D + 127) / 128, 128, 0, st>>>(s->acc, s->mean, s->C, s->D, (double)s->n_moment,
                                                               mean_dev, var_dev, rhat_dev, ess_dev);
    BINFB_CUDA(cudaGetLastError());
    return BINFB_OK;
}

int binfb_sink_sums_host(binfb_sink *s, double *pivot, double *sum_dev, double *sum_dev2, double *sum_m2) {
    int rc = check_sink(s);
    if (rc) return rc;
    if (s->n_moment < 2) {
        set_error("sink_sums: needs at least 2 post-burn-in sweeps");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(s->device));
    // same reduction as binfb_sink_summary, without the finalisation
    rc = binfb_sink_summary(s, nullptr, nullptr, nullptr, nullptr, s->hstream);
    if (rc) return rc;
    BINFB_CUDA(cudaStreamSynchronize(s->hstream));
    const size_t D = s->D;
    if (pivot) BINFB_CUDA(cudaMemcpy(pivot, s->mean, D * sizeof(double), cudaMemcpyDeviceToHost));
    if (sum_dev) BINFB_CUDA(cudaMemcpy(sum_dev, s->acc, D * sizeof(double), cudaMemcpyDeviceToHost));
    if (sum_dev2) BINFB_CUDA(cudaMemcpy(sum_dev2, s->acc + D, D * sizeof(double), cudaMemcpyDeviceToHost));
    if (sum_m2) BINFB_CUDA(cudaMemcpy(sum_m2, s->acc + 2 * D, D * sizeof(double), cudaMemcpyDeviceToHost));
    return BINFB_OK;
}

int binfb_sink_summary_host(binfb_sink *s, double *mean, double *var, double *rhat, double *ess) {
    BINFB_TRACE();
    int rc = check_sink(s);
    if (rc) return rc;
    BINFB_CUDA(cudaSetDevice(s->device));
    double *d = nullptr;
    const size_t D = s->D;
    BINFB_CUDA(cudaMalloc(&d, 4 * D * sizeof(double)));
    rc = binfb_sink_summary(s, d, d + D, d + 2 * D, d + 3 * D, s->hstream);
    cudaError_t e = cudaSuccess;
    if (rc == BINFB_OK) e = cudaStreamSynchronize(s->hstream);
    double *outs[4] = {mean, var, rhat, ess};
    for (int k = 0; k < 4 && rc == BINFB_OK && e == cudaSuccess; ++k)
        if (outs[k]) e = cudaMemcpy(outs[k], d + k * D, D * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "sink_summary_host");
    return rc;
}

int binfb_sink_moments_host(binfb_sink *s, double *mean, double *var) {
    int rc = check_sink(s);
    if (rc) return rc;
    if (s->n_moment < 2) {
        set_error("sink_moments: needs at least 2 post-burn-in sweeps");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(s->device));
    const long long CD = (long long)s->C * s->D;
    BINFB_CUDA(cudaDeviceSynchronize());
    if (mean) BINFB_CUDA(cudaMemcpy(mean, s->mean, CD * sizeof(double), cudaMemcpyDeviceToHost));
    if (var) {
        double *d = nullptr;
        BINFB_CUDA(cudaMalloc(&d, CD * sizeof(double)));
        sink_var_kernel<<<s->sm_count * 8, 256>>>(s->m2, CD, 1.0 / (double)(s->n_moment - 1), d);
        cudaError_t e = cudaMemcpy(var, d, CD * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(d);
        if (e != cudaSuccess) return cuda_fail(e, "sink_moments_host");
    }
    return BINFB_OK;
}

int binfb_sink_read_host(binfb_sink *s, long long first, long long count, float *q_out, float *aux_out) {
    int rc = check_sink(s);
    if (rc) return rc;
    const long long oldest = s->n_kept > s->capacity ? s->n_kept - s->capacity : 0;
    if (count < 0 || first < oldest || first + count > s->n_kept) {
        set_error("sink_read: kept samples [first, first+count) are not (or no longer) in the ring");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(s->device));
    BINFB_CUDA(cudaDeviceSynchronize());
    const size_t CD = (size_t)s->C * s->D;
    for (long long k = 0; k < count; ++k) {
        const long long slot = (first + k) % s->capacity;
        if (q_out)
            BINFB_CUDA(cudaMemcpyAsync(q_out + (size_t)k * CD, s->ring_q + (size_t)slot * CD, CD * sizeof(float),
                                       cudaMemcpyDeviceToHost, s->hstream));
        if (aux_out)
            BINFB_CUDA(cudaMemcpyAsync(aux_out + (size_t)k * s->C, s->ring_aux + (size_t)slot * s->C,
                                       s->C * sizeof(float), cudaMemcpyDeviceToHost, s->hstream));
    }
    BINFB_CUDA(cudaStreamSynchronize(s->hstream));
    return BINFB_OK;
}

int binfb_sink_map_host(binfb_sink *s, double *logp, float *q_map, float *aux_map) {
    int rc = check_sink(s);
    if (rc) return rc;
    if (!(s->flags & BINFB_SINK_TRACK_MAP)) {
        set_error("sink_map: the sink was created without BINFB_SINK_TRACK_MAP");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(s->device));
    BINFB_CUDA(cudaDeviceSynchronize());
    if (logp) BINFB_CUDA(cudaMemcpy(logp, s->best[s->best_cur], s->C * sizeof(double), cudaMemcpyDeviceToHost));
    if (q_map)
        BINFB_CUDA(cudaMemcpy(q_map, s->map_q, (size_t)s->C * s->D * sizeof(float), cudaMemcpyDeviceToHost));
    if (aux_map) BINFB_CUDA(cudaMemcpy(aux_map, s->map_aux, s->C * sizeof(float), cudaMemcpyDeviceToHost));
    return BINFB_OK;
}

}  // extern "C"
