// C ABI of libbinf_b200.so (declared in include/binf_b200.h).
#include <math.h>

#include <algorithm>
#include <atomic>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "internal.h"

namespace binfb {


static thread_local std::string g_error;

void set_error(const std::string &msg) { g_error = msg; }

int cuda_fail(cudaError_t e, const char *what) {
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();  // clear the sticky-less error state
    return BINFB_ECUDA;
}

static int check_model(const binfb_model *m) {
    if (!m) {
        set_error("null model handle");
        return BINFB_EINVAL;
    }
    return BINFB_OK;
}

static HmcArgs make_hmc_args(const binfb_model *m, float *q, float *tau, const float *beta,
                             float *eps, int C, const binfb_hmc_opts *o, const float *p0,
                             const float *u, const double *gamma_draws, uint8_t *accepted,
                             double *e_before, double *e_after, float *q_end, float *p_end,
                             int32_t *n_accepted, double *stats) {
    HmcArgs a;
    a.q = q, a.tau = tau, a.beta = beta, a.eps = eps, a.C = C;
    a.L = o->n_steps, a.n_traj = o->n_traj, a.n_adapt = o->n_adapt, a.gibbs_mode = o->gibbs_mode;
    a.adapt_up = (float)o->adapt_up, a.adapt_down = (float)o->adapt_down;
    a.seed = o->seed, a.draw = o->draw, a.chain_base = o->chain_base;
    a.p0 = p0, a.u = u, a.gamma_draws = gamma_draws;
    a.accepted = accepted, a.e_before = e_before, a.e_after = e_after;
    a.q_end = q_end, a.p_end = p_end, a.n_accepted = n_accepted, a.stats = stats;
    a.gamma_shape = m->gamma_shape, a.gamma_rate = m->gamma_rate;
    return a;
}

static int check_hmc(const binfb_model *m, int C, const binfb_hmc_opts *o, const void *q,
                     const void *tau, const void *eps, const void *p0, const void *u) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!o || !q || !tau || !eps) {
        set_error("hmc_run: q, tau, eps and opts are required");
        return BINFB_EINVAL;
    }
    if (C < 1 || o->n_steps < 1 || o->n_traj < 1) {
        set_error("hmc_run: n_chains, n_steps and n_traj must be >= 1");
        return BINFB_EINVAL;
    }
    if ((p0 || u) && o->n_traj != 1) {
        set_error("hmc_run: injected momenta / uniforms require n_traj == 1");
        return BINFB_EINVAL;
    }
    if (o->gibbs_mode < 0 || o->gibbs_mode > 2) {
        set_error("hmc_run: bad gibbs_mode");
        return BINFB_EINVAL;
    }
    return BINFB_OK;
}

// scratch device memory for the *_host entry points: one growing arena per model
struct Arena {
    binfb_model *m;
    size_t used = 0;
    std::vector<size_t> sizes;
    explicit Arena(binfb_model *mm) : m(mm) {}
    size_t plan(size_t bytes) {
        const size_t off = used;
        used += (bytes + 255) / 256 * 256;
        return off;
    }
    int commit() {
        if (used > m->hb_bytes) {
            if (m->hb) cudaFree(m->hb);
            m->hb = nullptr, m->hb_bytes = 0;
            BINFB_CUDA(cudaMalloc(&m->hb, used));
            m->hb_bytes = used;
        }
        return BINFB_OK;
    }
    template <typename T>
    T *at(size_t off) {
        return reinterpret_cast<T *>(m->hb + off);
    }
};

static int host_stream(binfb_model *m) {
    if (!m->hstream) BINFB_CUDA(cudaStreamCreateWithFlags(&m->hstream, cudaStreamNonBlocking));
    return BINFB_OK;
}

// Stream memory operations of the driver API (cuStreamWriteValue32 / cuStreamWaitValue32), resolved through the
// runtime so that the library keeps no link-time dependency on libcuda.  They let a copy stream tell a RUNNING
// kernel that a chunk of its input has landed, and wait for the kernel to say that a chunk of its output is
// final -- which is what overlaps the PCIe copies of binfb_hmc_run_host with the trajectory kernel.
typedef int (*stream_value32_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
struct StreamMemOps {
    stream_value32_fn write = nullptr, wait = nullptr;
    bool ok = false;
};
static const StreamMemOps &stream_mem_ops() {
    static StreamMemOps ops = [] {
        StreamMemOps o;
        void *w = nullptr, *t = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &qr) == cudaSuccess && w &&
            qr == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuStreamWaitValue32", &t, cudaEnableDefault, &qr) == cudaSuccess && t &&
            qr == cudaDriverEntryPointSuccess) {
            o.write = (stream_value32_fn)w, o.wait = (stream_value32_fn)t, o.ok = true;
        }
        cudaGetLastError();
        return o;
    }();
    return ops;
}
static int pipe_streams(binfb_model *m) {
    if (!m->hstream_in) BINFB_CUDA(cudaStreamCreateWithFlags(&m->hstream_in, cudaStreamNonBlocking));
    if (!m->hstream_out) BINFB_CUDA(cudaStreamCreateWithFlags(&m->hstream_out, cudaStreamNonBlocking));
    if (!m->hevent) BINFB_CUDA(cudaEventCreateWithFlags(&m->hevent, cudaEventDisableTiming));
    return BINFB_OK;
}

}  // namespace binfb

using namespace binfb;

extern "C" {

int binfb_version(void) { return BINFB_VERSION; }

const char *binfb_last_error(void) { return g_error.c_str(); }

int binfb_device_count(int *count) {
    if (!count) return BINFB_EINVAL;
    *count = 0;
    BINFB_CUDA(cudaGetDeviceCount(count));
    return BINFB_OK;
}

int binfb_device_props(int device, int *sm_count, int *smem_optin, int *clock_khz, int *cc) {
    cudaDeviceProp p;
    BINFB_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (smem_optin) *smem_optin = (int)p.sharedMemPerBlockOptin;
    if (clock_khz) cudaDeviceGetAttribute(clock_khz, cudaDevAttrClockRate, device);
    if (cc) *cc = p.major * 10 + p.minor;
    return BINFB_OK;
}

static int model_common_init(binfb_model *m, int device) {
    m->device = device;
    BINFB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    BINFB_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) {
        set_error("binf_b200 kernels are built for sm_100a (Blackwell) only; device is sm_" +
                  std::to_string(p.major * 10 + p.minor));
        return BINFB_EUNSUPPORTED;
    }
    m->sm_count = p.multiProcessorCount;
    m->smem_optin = (int)p.sharedMemPerBlockOptin;
    return BINFB_OK;
}

int binfb_model_create_polynomial(const double *xs, const double *ys, int n_data, int n_coeff,
                                  const double *prior_mean, const double *prior_var,
                                  double gamma_shape, double gamma_rate, unsigned flags, int device,
                                  binfb_model **out) {
    BINFB_TRACE();
    if (!xs || !ys || !out || n_data < 1) {
        set_error("model_create_polynomial: xs, ys, out required, n_data >= 1");
        return BINFB_EINVAL;
    }
    if (n_coeff < 1 || n_coeff > 8) {
        set_error("model_create_polynomial: n_coeff must be in 1..8");
        return BINFB_EUNSUPPORTED;
    }
    binfb_model *m = new binfb_model();
    int rc = model_common_init(m, device);
    if (rc) {
        delete m;
        return rc;
    }
    m->kind = BINFB_MODEL_POLYNOMIAL, m->dim = n_coeff, m->n_data = n_data;
    m->gamma_shape = gamma_shape, m->gamma_rate = gamma_rate;
    PolyModel &pm = m->poly;
    pm.N = n_data, pm.K = n_coeff, pm.stride = (n_coeff + 3) / 4 * 4, pm.flags = flags;
    {
        static std::atomic<unsigned long long> next_uid{1};
        pm.uid = next_uid.fetch_add(1);
    }
    for (int k = 0; k < 8; ++k) {
        pm.prior_mean[k] = (k < n_coeff && prior_mean) ? (float)prior_mean[k] : 0.f;
        pm.prior_inv_var[k] = (k < n_coeff && prior_var) ? (float)(1.0 / prior_var[k]) : 0.f;
    }
    // rows [x, x^2, .., x^(K-1), y, pad]: powers formed in float64 like the reference's
    // xs ** i (binf/example/likelihood.py:30), rounded once to float32
    std::vector<float> rows((size_t)n_data * pm.stride, 0.f);
    for (int n = 0; n < n_data; ++n) {
        double pw = 1.0;
        for (int k = 1; k < n_coeff; ++k) {
            pw *= xs[n];
            rows[(size_t)n * pm.stride + k - 1] = (float)pw;
        }
        rows[(size_t)n * pm.stride + n_coeff - 1] = (float)ys[n];
    }
    cudaError_t e = cudaMalloc(&pm.rows, rows.size() * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMemcpy(pm.rows, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice);
    // a cudaMemcpy from pageable memory may return before the DMA has landed, and the kernels run on
    // non-blocking / caller streams that do not wait for the legacy stream
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        binfb_model_destroy(m);
        return cuda_fail(e, "polynomial data upload");
    }
    *out = m;
    return BINFB_OK;
}

int binfb_model_create_generic(const char *device_code, int n_params, int x_dim, const double *xs,
                               const double *ys, int n_data, const double *prior_mean,
                               const double *prior_var, double gamma_shape, double gamma_rate,
                               unsigned flags, int device, binfb_model **out) {
    BINFB_TRACE();
    if (!device_code || !xs || !ys || !out || n_data < 1 || x_dim < 1) {
        set_error("model_create_generic: device_code, xs, ys, out required, n_data, x_dim >= 1");
        return BINFB_EINVAL;
    }
    if (n_params < 1 || n_params > 16) {
        set_error("model_create_generic: n_params must be in 1..16");
        return BINFB_EUNSUPPORTED;
    }
    binfb_model *m = new binfb_model();
    int rc = model_common_init(m, device);
    if (rc) {
        delete m;
        return rc;
    }
    m->kind = BINFB_MODEL_GENERIC, m->dim = n_params, m->n_data = n_data;
    m->gamma_shape = gamma_shape, m->gamma_rate = gamma_rate;
    rc = gen_create(m->gen, device_code, n_params, x_dim, xs, ys, n_data, prior_mean, prior_var, flags);
    if (rc) {
        binfb_model_destroy(m);
        return rc;
    }
    *out = m;
    return BINFB_OK;
}

int binfb_model_create_chromatin(int n_beads, const float *y_pairs, double alpha, double d_c,
                                 double k_bb, double l0, double conf_s, double gamma_shape,
                                 double gamma_rate, unsigned flags, int device, binfb_model **out) {
    BINFB_TRACE();
    if (!y_pairs || !out || n_beads < 2) {
        set_error("model_create_chromatin: y_pairs, out required, n_beads >= 2");
        return BINFB_EINVAL;
    }
    // the pair loop works on positions scaled by the exponent slope alpha*log2(e) and folds 2^(-alpha d_c log2 e)
    // into a multiplier (pair_block.cuh, SCALED): the slope must be positive and the multiplier a normal float
    const bool algebraic = (flags & BINFB_FLAG_CONTACT_ALGEBRAIC) != 0;  // (no exponential: alpha > 0 is enough)
    if (!(alpha > 0.0) || !(algebraic || fabs(alpha * d_c) <= 80.0) || !(fabs(alpha * d_c) <= 1e6)) {
        set_error("model_create_chromatin: alpha > 0 and |alpha * d_c| <= 80 required");
        return BINFB_EINVAL;
    }
    binfb_model *m = new binfb_model();
    int rc = model_common_init(m, device);
    if (rc) {
        delete m;
        return rc;
    }
    m->kind = BINFB_MODEL_CHROMATIN, m->dim = 3 * n_beads;
    m->n_data = (long long)n_beads * (n_beads - 1) / 2;
    m->gamma_shape = gamma_shape, m->gamma_rate = gamma_rate;
    ChromModel &cm = m->chrom;
    cm.n = n_beads;
    cm.M = m->n_data;
    cm.alpha = (float)alpha, cm.d_c = (float)d_c, cm.k_bb = (float)k_bb, cm.l0 = (float)l0;
    cm.inv_s2 = conf_s > 0.0 ? (float)(1.0 / (conf_s * conf_s)) : 0.f;
    cm.flags = flags;
    const int force_roles = (int)((flags >> 8) & 0x1f);  // test hook: bits 8..12 force R
    cm.plan = chrom_plan(n_beads, m->smem_optin, force_roles);
    if (cm.plan.W < 1) {
        delete m;
        set_error("model_create_chromatin: n_beads too large for the shared-memory resident kernel");
        return BINFB_EUNSUPPORTED;
    }
    std::vector<float> stream((size_t)cm.plan.stream_floats);
    rc = chrom_build_stream(n_beads, y_pairs, cm.plan, stream.data());
    const long long n_floats = cm.plan.stream_floats;
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaMalloc(&cm.ystream, (size_t)n_floats * sizeof(float));
    if (!rc && e == cudaSuccess)
        e = cudaMemcpy(cm.ystream, stream.data(), (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice);
    // small-batch alternative plan with its own stream layout (not when the role count is forced)
    if (!rc && e == cudaSuccess && force_roles == 0) {
        cm.plan_alt = chrom_plan_small_batch(n_beads, m->smem_optin, cm.plan);
        if (cm.plan_alt.W >= 1) {
            std::vector<float> alt((size_t)cm.plan_alt.stream_floats);
            rc = chrom_build_stream(n_beads, y_pairs, cm.plan_alt, alt.data());
            if (!rc) e = cudaMalloc(&cm.ystream_alt, alt.size() * sizeof(float));
            if (!rc && e == cudaSuccess)
                e = cudaMemcpy(cm.ystream_alt, alt.data(), alt.size() * sizeof(float), cudaMemcpyHostToDevice);
        }
    }
    if (!rc && e == cudaSuccess) e = cudaMalloc(&cm.ypairs, (size_t)cm.M * sizeof(float));
    if (!rc && e == cudaSuccess)
        e = cudaMemcpy(cm.ypairs, y_pairs, (size_t)cm.M * sizeof(float), cudaMemcpyHostToDevice);
    if (!rc && e == cudaSuccess) e = cudaDeviceSynchronize();  // see binfb_model_create_polynomial
    if (rc || e != cudaSuccess) {
        binfb_model_destroy(m);
        return rc ? rc : cuda_fail(e, "chromatin data upload");
    }
    *out = m;
    return BINFB_OK;
}

int binfb_model_destroy(binfb_model *m) {
    if (!m) return BINFB_OK;
    cudaSetDevice(m->device);
    cudaFree(m->poly.rows);
    gen_destroy(m->gen);
    cudaFree(m->rw_prop), cudaFree(m->rw_lp[0]), cudaFree(m->rw_lp[1]), cudaFree(m->rw_lp[2]);
    ChromModel &c = m->chrom;
    cudaFree(c.ystream), cudaFree(c.ystream_alt), cudaFree(c.ypairs), cudaFree(c.qw), cudaFree(c.pw), cudaFree(c.h0);
    cudaFree(c.chi2_0), cudaFree(c.chi2_state), cudaFree(c.tau_w), cudaFree(c.sched);
    if (m->hb) cudaFree(m->hb);
    if (m->hstream) cudaStreamDestroy(m->hstream);
    if (m->hstream_in) cudaStreamDestroy(m->hstream_in);
    if (m->hstream_out) cudaStreamDestroy(m->hstream_out);
    if (m->hevent) cudaEventDestroy(m->hevent);
    delete m;
    return BINFB_OK;
}

int binfb_model_info(const binfb_model *m, int *kind, int *dim, long long *n_data, int *device) {
    int rc = check_model(m);
    if (rc) return rc;
    if (kind) *kind = m->kind;
    if (dim) *dim = m->dim;
    if (n_data) *n_data = m->n_data;
    if (device) *device = m->device;
    return BINFB_OK;
}

int binfb_model_set_gamma_prior(binfb_model *m, double shape, double rate) {
    int rc = check_model(m);
    if (rc) return rc;
    m->gamma_shape = shape, m->gamma_rate = rate;
    return BINFB_OK;
}

int binfb_model_set_option(binfb_model *m, const char *key, double value) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!key) return BINFB_EINVAL;
    const int v = value < 0 ? -1 : (int)value;
    if (!strcmp(key, "poly.group")) m->poly.opt_group = v;
    else if (!strcmp(key, "poly.uniform_rows")) m->poly.opt_ur = v;
    else if (!strcmp(key, "poly.chains_per_thread")) m->poly.opt_jchains = v;
    else if (!strcmp(key, "poly.block")) m->poly.opt_block = v;
    else if (!strcmp(key, "chrom.warps")) m->chrom.opt_warps = v;
    else if (!strcmp(key, "chrom.sets")) m->chrom.opt_sets = v;
    else if (!strcmp(key, "host.pipeline")) m->host_pipeline = v != 0;
    else if (!strcmp(key, "host.chunks")) m->host_chunks = v < 2 ? 16 : (v > 64 ? 64 : v);
    else if (!strcmp(key, "generic.split")) m->gen.opt_split = v;
    else if (!strcmp(key, "chrom.ev_k") || !strcmp(key, "chrom.ev_d")) {
        // excluded-volume prior k_ev sum_{i<j} max(0, d_ev - d_ij)^4 of the chromatin model (0 = off)
        if (m->kind != BINFB_MODEL_CHROMATIN || value < 0) {
            set_error("chrom.ev_k / chrom.ev_d: chromatin models only, value >= 0");
            return BINFB_EINVAL;
        }
        (key[9] == 'k' ? m->chrom.ev_k : m->chrom.ev_d) = (float)value;
    } else {
        set_error(std::string("unknown option: ") + key);
        return BINFB_EINVAL;
    }
    return BINFB_OK;
}

int binfb_model_get_option(const binfb_model *m, const char *key, double *value) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!key || !value) return BINFB_EINVAL;
    if (!strcmp(key, "generic.packed")) *value = m->gen.pack;
    else if (!strcmp(key, "generic.uniform_rows")) *value = m->gen.ur;
    else if (!strcmp(key, "generic.warps_per_set")) *value = m->gen.G;
    else if (!strcmp(key, "generic.rows_in_smem")) *value = m->gen.srows;
    else if (!strcmp(key, "generic.split")) *value = m->gen.opt_split;
    else if (!strcmp(key, "poly.group")) *value = m->poly.opt_group;
    else if (!strcmp(key, "poly.uniform_rows")) *value = m->poly.opt_ur;
    else if (!strcmp(key, "poly.chains_per_thread")) *value = m->poly.opt_jchains;
    else if (!strcmp(key, "poly.block")) *value = m->poly.opt_block;
    else if (!strcmp(key, "chrom.warps")) *value = m->chrom.opt_warps;
    else if (!strcmp(key, "chrom.sets")) *value = m->chrom.opt_sets;
    else if (!strcmp(key, "host.pipeline")) *value = m->host_pipeline ? 1 : 0;
    else if (!strcmp(key, "host.chunks")) *value = m->host_chunks;
    else if (!strcmp(key, "chrom.ev_k")) *value = m->chrom.ev_k;
    else if (!strcmp(key, "chrom.ev_d")) *value = m->chrom.ev_d;
    else if (!strcmp(key, "chrom.algebraic")) *value = (m->chrom.flags & BINFB_FLAG_CONTACT_ALGEBRAIC) ? 1 : 0;
    else {
        set_error(std::string("unknown option: ") + key);
        return BINFB_EINVAL;
    }
    return BINFB_OK;
}

int binfb_logprob_grad(binfb_model *m, const float *q, const float *tau, const float *beta, int C,
                       double *logp, float *grad, double *chi2, void *stream) {
    BINFB_TRACE();
    int rc = check_model(m);
    if (rc) return rc;
    if (!q || !tau || C < 1) {
        set_error("logprob_grad: q, tau required, n_chains >= 1");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(m->device));
    GradArgs a;
    a.q = q, a.tau = tau, a.beta = beta, a.C = C, a.logp = logp, a.grad = grad, a.chi2 = chi2;
    a.gamma_shape = m->gamma_shape, a.gamma_rate = m->gamma_rate;
    cudaStream_t s = (cudaStream_t)stream;
    if (m->kind == BINFB_MODEL_POLYNOMIAL) return poly_grad_launch(m->poly, a, m->sm_count, m->smem_optin, s);
    if (m->kind == BINFB_MODEL_GENERIC) return gen_grad_launch(m->gen, a, s);
    return chrom_grad_launch(m->chrom, a, m->sm_count, m->smem_optin, s);
}

int binfb_logprob_grad_host(binfb_model *m, const float *q, const float *tau, const float *beta,
                            int C, double *logp, float *grad, double *chi2) {
    BINFB_TRACE();
    int rc = check_model(m);
    if (rc) return rc;
    if (!q || !tau || C < 1) {
        set_error("logprob_grad_host: q, tau required, n_chains >= 1");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(m->device));
    if ((rc = host_stream(m))) return rc;
    const size_t D = m->dim;
    Arena ar(m);
    const size_t oq = ar.plan(C * D * 4), ot = ar.plan(C * 4), ob = ar.plan(C * 4), ol = ar.plan(C * 8),
                 og = ar.plan(C * D * 4), oc = ar.plan(C * 8);
    if ((rc = ar.commit())) return rc;
    cudaStream_t s = m->hstream;
    BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(oq), q, C * D * 4, cudaMemcpyHostToDevice, s));
    BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ot), tau, C * 4, cudaMemcpyHostToDevice, s));
    if (beta) BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ob), beta, C * 4, cudaMemcpyHostToDevice, s));
    rc = binfb_logprob_grad(m, ar.at<float>(oq), ar.at<float>(ot), beta ? ar.at<float>(ob) : nullptr, C,
                            ar.at<double>(ol), grad ? ar.at<float>(og) : nullptr, ar.at<double>(oc), s);
    if (rc) return rc;
    if (logp) BINFB_CUDA(cudaMemcpyAsync(logp, ar.at<double>(ol), C * 8, cudaMemcpyDeviceToHost, s));
    if (grad) BINFB_CUDA(cudaMemcpyAsync(grad, ar.at<float>(og), C * D * 4, cudaMemcpyDeviceToHost, s));
    if (chi2) BINFB_CUDA(cudaMemcpyAsync(chi2, ar.at<double>(oc), C * 8, cudaMemcpyDeviceToHost, s));
    BINFB_CUDA(cudaStreamSynchronize(s));
    return BINFB_OK;
}

int binfb_forward_host(binfb_model *m, const float *q, int C, float *mock) {
    BINFB_TRACE();
    int rc = check_model(m);
    if (rc) return rc;
    if (!q || !mock || C < 1) {
        set_error("forward_host: q, mock required, n_chains >= 1");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(m->device));
    if ((rc = host_stream(m))) return rc;
    const size_t D = m->dim, N = (size_t)m->n_data;
    Arena ar(m);
    const size_t oq = ar.plan(C * D * 4), om = ar.plan(C * N * 4);
    if ((rc = ar.commit())) return rc;
    cudaStream_t s = m->hstream;
    BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(oq), q, C * D * 4, cudaMemcpyHostToDevice, s));
    if (m->kind == BINFB_MODEL_POLYNOMIAL) rc = poly_forward_launch(m->poly, ar.at<float>(oq), C, ar.at<float>(om), s);
    else if (m->kind == BINFB_MODEL_GENERIC) rc = gen_forward_launch(m->gen, ar.at<float>(oq), C, ar.at<float>(om), s);
    else rc = chrom_forward_launch(m->chrom, ar.at<float>(oq), C, ar.at<float>(om), s);
    if (rc) return rc;
    BINFB_CUDA(cudaMemcpyAsync(mock, ar.at<float>(om), C * N * 4, cudaMemcpyDeviceToHost, s));
    BINFB_CUDA(cudaStreamSynchronize(s));
    return BINFB_OK;
}

int binfb_hmc_run(binfb_model *m, float *q, float *tau, const float *beta, float *eps, int C,
                  const binfb_hmc_opts *opts, const float *p0, const float *u,
                  const double *gamma_draws, uint8_t *accepted, double *e_before, double *e_after,
                  float *q_end, float *p_end, int32_t *n_accepted, double *stats, void *stream) {
    BINFB_TRACE();
    int rc = check_hmc(m, C, opts, q, tau, eps, p0, u);
    if (rc) return rc;
    BINFB_CUDA(cudaSetDevice(m->device));
    const HmcArgs a = make_hmc_args(m, q, tau, beta, eps, C, opts, p0, u, gamma_draws, accepted,
                                    e_before, e_after, q_end, p_end, n_accepted, stats);
    cudaStream_t s = (cudaStream_t)stream;
    if (m->kind == BINFB_MODEL_POLYNOMIAL) return poly_hmc_launch(m->poly, a, m->sm_count, m->smem_optin, s);
    if (m->kind == BINFB_MODEL_GENERIC) return gen_hmc_launch(m->gen, a, s);
    return chrom_hmc_launch(m->chrom, a, m->sm_count, m->smem_optin, s);
}

int binfb_hmc_run_host(binfb_model *m, float *q, float *tau, const float *beta, float *eps, int C,
                       const binfb_hmc_opts *opts, const float *p0, const float *u,
                       const double *gamma_draws, uint8_t *accepted, double *e_before,
                       double *e_after, float *q_end, float *p_end, int32_t *n_accepted,
                       double *stats) {
    BINFB_TRACE();
    int rc = check_hmc(m, C, opts, q, tau, eps, p0, u);
    if (rc) return rc;
    BINFB_CUDA(cudaSetDevice(m->device));
    if ((rc = host_stream(m))) return rc;
    const size_t D = m->dim;
    Arena ar(m);
    const size_t oq = ar.plan(C * D * 4), ot = ar.plan(C * 4), ob = ar.plan(C * 4), oe = ar.plan(C * 4),
                 op = ar.plan(C * D * 4), ou = ar.plan(C * 4), og = ar.plan(C * 8), oa = ar.plan(C),
                 o0 = ar.plan(C * 8), o1 = ar.plan(C * 8), oqe = ar.plan(C * D * 4),
                 ope = ar.plan(C * D * 4), on = ar.plan(C * 4), os = ar.plan(32);
    if ((rc = ar.commit())) return rc;
    cudaStream_t s = m->hstream;
    // ---- pipelined variant (chromatin, batches whose state copy is worth hiding): ONE launch; the state goes
    //      up in chunks on a copy-in stream that opens the kernel's gate chunk by chunk, and comes back in chunks
    //      on a copy-out stream that waits for the kernel's per-chunk completion counters (chromatin.cu,
    //      CHROM_GATE_WORDS).  The small per-chain arrays travel on the kernel's stream as before.
    const bool want_pipe = m->kind == BINFB_MODEL_CHROMATIN && !p0 && !q_end && !p_end && m->host_pipeline &&
                           (size_t)C * D * 4 >= ((size_t)4 << 20) && stream_mem_ops().ok &&
                           !(m->chrom.ev_k > 0.f && opts->gibbs_mode == BINFB_GIBBS_TAU_FIRST);
    ChromPipe pipe;
    if (want_pipe && chrom_pipe_shape(m->chrom, C, m->sm_count, m->host_chunks, &pipe) == BINFB_OK && pipe.n_chunks >= 2 &&
        pipe_streams(m) == BINFB_OK) {
        const StreamMemOps &mo = stream_mem_ops();
        BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ot), tau, C * 4, cudaMemcpyHostToDevice, s));
        BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(oe), eps, C * 4, cudaMemcpyHostToDevice, s));
        if (beta) BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ob), beta, C * 4, cudaMemcpyHostToDevice, s));
        if (u) BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ou), u, C * 4, cudaMemcpyHostToDevice, s));
        if (gamma_draws)
            BINFB_CUDA(cudaMemcpyAsync(ar.at<double>(og), gamma_draws, C * 8, cudaMemcpyHostToDevice, s));
        if (stats) BINFB_CUDA(cudaMemsetAsync(ar.at<double>(os), 0, 32, s));
        const HmcArgs a = make_hmc_args(m, ar.at<float>(oq), ar.at<float>(ot), beta ? ar.at<float>(ob) : nullptr,
                                        ar.at<float>(oe), C, opts, nullptr, u ? ar.at<float>(ou) : nullptr,
                                        gamma_draws ? ar.at<double>(og) : nullptr,
                                        accepted ? ar.at<uint8_t>(oa) : nullptr, e_before ? ar.at<double>(o0) : nullptr,
                                        e_after ? ar.at<double>(o1) : nullptr, nullptr, nullptr,
                                        n_accepted ? ar.at<int32_t>(on) : nullptr, stats ? ar.at<double>(os) : nullptr);
        pipe.header_written = m->hevent;  // recorded on s between the gate header and the kernel
        rc = chrom_hmc_launch(m->chrom, a, m->sm_count, m->smem_optin, s, &pipe);
        if (rc) return rc;
        const size_t chunk_chains = (size_t)pipe.groups_per_chunk * pipe.W;
        BINFB_CUDA(cudaStreamWaitEvent(m->hstream_in, m->hevent, 0));
        int pipe_err = 0;
        for (int j = 0; j < pipe.n_chunks && !pipe_err; ++j) {
            const size_t c0 = (size_t)j * chunk_chains, c1 = std::min((size_t)C, c0 + chunk_chains);
            if (cudaMemcpyAsync(ar.at<float>(oq) + c0 * D, q + c0 * D, (c1 - c0) * D * 4, cudaMemcpyHostToDevice,
                                m->hstream_in) != cudaSuccess)
                pipe_err = 1;
            // even after a failed copy the gate is opened, so that the kernel never waits for data that will not come
            if (mo.write(m->hstream_in, (unsigned long long)(uintptr_t)pipe.gate, (unsigned int)(c1 + 1), 0) != 0)
                pipe_err = 1;
        }
        if (pipe_err) {
            // open the gate from the host side of things and fall through to a plain synchronise
            const int all = C + 1;
            cudaMemcpyAsync(pipe.gate, &all, sizeof(int), cudaMemcpyHostToDevice, m->hstream_in);
            cudaStreamSynchronize(m->hstream_in);
            cudaStreamSynchronize(s);
            return cuda_fail(cudaGetLastError(), "pipelined host call: copy-in");
        }
        for (int j = 0; j < pipe.n_chunks; ++j) {
            const size_t c0 = (size_t)j * chunk_chains, c1 = std::min((size_t)C, c0 + chunk_chains);
            const int g0 = j * pipe.groups_per_chunk;
            const int groups = std::min(pipe.n_groups, g0 + pipe.groups_per_chunk) - g0;
            // CU_STREAM_WAIT_VALUE_GEQ = 0
            if (mo.wait(m->hstream_out, (unsigned long long)(uintptr_t)(pipe.done + j), (unsigned int)groups, 0) != 0) {
                cudaStreamSynchronize(s);  // the kernel still finishes; copy everything after it instead
                BINFB_CUDA(cudaMemcpyAsync(q + c0 * D, ar.at<float>(oq) + c0 * D, ((size_t)C - c0) * D * 4,
                                           cudaMemcpyDeviceToHost, m->hstream_out));
                break;
            }
            BINFB_CUDA(cudaMemcpyAsync(q + c0 * D, ar.at<float>(oq) + c0 * D, (c1 - c0) * D * 4,
                                       cudaMemcpyDeviceToHost, m->hstream_out));
        }
        if (opts->gibbs_mode != BINFB_GIBBS_NONE)
            BINFB_CUDA(cudaMemcpyAsync(tau, ar.at<float>(ot), C * 4, cudaMemcpyDeviceToHost, s));
        if (opts->n_adapt > 0)
            BINFB_CUDA(cudaMemcpyAsync(eps, ar.at<float>(oe), C * 4, cudaMemcpyDeviceToHost, s));
        if (accepted) BINFB_CUDA(cudaMemcpyAsync(accepted, ar.at<uint8_t>(oa), C, cudaMemcpyDeviceToHost, s));
        if (e_before) BINFB_CUDA(cudaMemcpyAsync(e_before, ar.at<double>(o0), C * 8, cudaMemcpyDeviceToHost, s));
        if (e_after) BINFB_CUDA(cudaMemcpyAsync(e_after, ar.at<double>(o1), C * 8, cudaMemcpyDeviceToHost, s));
        if (n_accepted) BINFB_CUDA(cudaMemcpyAsync(n_accepted, ar.at<int32_t>(on), C * 4, cudaMemcpyDeviceToHost, s));
        if (stats) BINFB_CUDA(cudaMemcpyAsync(stats, ar.at<double>(os), 32, cudaMemcpyDeviceToHost, s));
        BINFB_CUDA(cudaStreamSynchronize(m->hstream_in));
        BINFB_CUDA(cudaStreamSynchronize(m->hstream_out));
        BINFB_CUDA(cudaStreamSynchronize(s));
        return BINFB_OK;
    }
    BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(oq), q, C * D * 4, cudaMemcpyHostToDevice, s));
    BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ot), tau, C * 4, cudaMemcpyHostToDevice, s));
    BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(oe), eps, C * 4, cudaMemcpyHostToDevice, s));
    if (beta) BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ob), beta, C * 4, cudaMemcpyHostToDevice, s));
    if (p0) BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(op), p0, C * D * 4, cudaMemcpyHostToDevice, s));
    if (u) BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ou), u, C * 4, cudaMemcpyHostToDevice, s));
    if (gamma_draws)
        BINFB_CUDA(cudaMemcpyAsync(ar.at<double>(og), gamma_draws, C * 8, cudaMemcpyHostToDevice, s));
    if (stats) BINFB_CUDA(cudaMemsetAsync(ar.at<double>(os), 0, 32, s));
    rc = binfb_hmc_run(m, ar.at<float>(oq), ar.at<float>(ot), beta ? ar.at<float>(ob) : nullptr,
                       ar.at<float>(oe), C, opts, p0 ? ar.at<float>(op) : nullptr,
                       u ? ar.at<float>(ou) : nullptr, gamma_draws ? ar.at<double>(og) : nullptr,
                       accepted ? ar.at<uint8_t>(oa) : nullptr, e_before ? ar.at<double>(o0) : nullptr,
                       e_after ? ar.at<double>(o1) : nullptr, q_end ? ar.at<float>(oqe) : nullptr,
                       p_end ? ar.at<float>(ope) : nullptr, n_accepted ? ar.at<int32_t>(on) : nullptr,
                       stats ? ar.at<double>(os) : nullptr, s);
    if (rc) return rc;
    BINFB_CUDA(cudaMemcpyAsync(q, ar.at<float>(oq), C * D * 4, cudaMemcpyDeviceToHost, s));
    // tau only changes under a fused Gibbs update, eps only while the step size adapts
    if (opts->gibbs_mode != BINFB_GIBBS_NONE)
        BINFB_CUDA(cudaMemcpyAsync(tau, ar.at<float>(ot), C * 4, cudaMemcpyDeviceToHost, s));
    if (opts->n_adapt > 0)
        BINFB_CUDA(cudaMemcpyAsync(eps, ar.at<float>(oe), C * 4, cudaMemcpyDeviceToHost, s));
    if (accepted) BINFB_CUDA(cudaMemcpyAsync(accepted, ar.at<uint8_t>(oa), C, cudaMemcpyDeviceToHost, s));
    if (e_before) BINFB_CUDA(cudaMemcpyAsync(e_before, ar.at<double>(o0), C * 8, cudaMemcpyDeviceToHost, s));
    if (e_after) BINFB_CUDA(cudaMemcpyAsync(e_after, ar.at<double>(o1), C * 8, cudaMemcpyDeviceToHost, s));
    if (q_end) BINFB_CUDA(cudaMemcpyAsync(q_end, ar.at<float>(oqe), C * D * 4, cudaMemcpyDeviceToHost, s));
    if (p_end) BINFB_CUDA(cudaMemcpyAsync(p_end, ar.at<float>(ope), C * D * 4, cudaMemcpyDeviceToHost, s));
    if (n_accepted) BINFB_CUDA(cudaMemcpyAsync(n_accepted, ar.at<int32_t>(on), C * 4, cudaMemcpyDeviceToHost, s));
    if (stats) BINFB_CUDA(cudaMemcpyAsync(stats, ar.at<double>(os), 32, cudaMemcpyDeviceToHost, s));
    BINFB_CUDA(cudaStreamSynchronize(s));
    return BINFB_OK;
}

int binfb_gibbs_precision(binfb_model *m, const float *q, float *tau, const float *beta, int C,
                          uint64_t seed, uint64_t draw, uint64_t chain_base,
                          const double *gamma_draws, double *chi2, void *stream) {
    BINFB_TRACE();
    int rc = check_model(m);
    if (rc) return rc;
    if (!q || !tau || !chi2 || C < 1) {
        set_error("gibbs_precision: q, tau, chi2 (device scratch/output [C]) required");
        return BINFB_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // chi^2 at the current state: one fused forward pass (GammaSampler._calculate_rate calls
    // Likelihood.log_prob(precision=1), binf/example/samplers.py:34-38)
    rc = binfb_logprob_grad(m, q, tau, beta, C, nullptr, nullptr, chi2, stream);
    if (rc) return rc;
    return gibbs_tau_launch(chi2, tau, beta, C, (double)m->n_data, m->gamma_shape, m->gamma_rate, seed,
                            draw, chain_base, gamma_draws, s);
}

int binfb_gibbs_precision_host(binfb_model *m, const float *q, float *tau, const float *beta, int C,
                               uint64_t seed, uint64_t draw, uint64_t chain_base,
                               const double *gamma_draws, double *chi2) {
    BINFB_TRACE();
    int rc = check_model(m);
    if (rc) return rc;
    if (!q || !tau || C < 1) {
        set_error("gibbs_precision_host: q, tau required");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(m->device));
    if ((rc = host_stream(m))) return rc;
    const size_t D = m->dim;
    Arena ar(m);
    const size_t oq = ar.plan(C * D * 4), ot = ar.plan(C * 4), ob = ar.plan(C * 4), og = ar.plan(C * 8),
                 oc = ar.plan(C * 8);
    if ((rc = ar.commit())) return rc;
    cudaStream_t s = m->hstream;
    BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(oq), q, C * D * 4, cudaMemcpyHostToDevice, s));
    BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ot), tau, C * 4, cudaMemcpyHostToDevice, s));
    if (beta) BINFB_CUDA(cudaMemcpyAsync(ar.at<float>(ob), beta, C * 4, cudaMemcpyHostToDevice, s));
    if (gamma_draws)
        BINFB_CUDA(cudaMemcpyAsync(ar.at<double>(og), gamma_draws, C * 8, cudaMemcpyHostToDevice, s));
    rc = binfb_gibbs_precision(m, ar.at<float>(oq), ar.at<float>(ot), beta ? ar.at<float>(ob) : nullptr, C,
                               seed, draw, chain_base, gamma_draws ? ar.at<double>(og) : nullptr,
                               ar.at<double>(oc), s);
    if (rc) return rc;
    BINFB_CUDA(cudaMemcpyAsync(tau, ar.at<float>(ot), C * 4, cudaMemcpyDeviceToHost, s));
    if (chi2) BINFB_CUDA(cudaMemcpyAsync(chi2, ar.at<double>(oc), C * 8, cudaMemcpyDeviceToHost, s));
    BINFB_CUDA(cudaStreamSynchronize(s));
    return BINFB_OK;
}

int binfb_swap_decide(const double *ll_a, const double *ll_b, double beta_a, double beta_b, int C,
                      uint64_t seed, uint64_t attempt, uint64_t pair_id, uint64_t chain_base,
                      uint8_t *accept, void *stream) {
    BINFB_TRACE();
    if (!ll_a || !ll_b || !accept || C < 1) {
        set_error("swap_decide: ll_a, ll_b, accept required");
        return BINFB_EINVAL;
    }
    return swap_decide_launch(ll_a, ll_b, beta_a, beta_b, C, seed, attempt, pair_id, chain_base, accept,
                              (cudaStream_t)stream);
}

int binfb_swap_apply(float *q_mine, const float *q_theirs, float *eps_mine, const float *eps_theirs,
                     const uint8_t *accept, int C, int dim, void *stream) {
    BINFB_TRACE();
    if (!q_mine || !q_theirs || !accept || C < 1 || dim < 1) {
        set_error("swap_apply: q_mine, q_theirs, accept required");
        return BINFB_EINVAL;
    }
    return swap_apply_launch(q_mine, q_theirs, eps_mine, eps_theirs, accept, C, dim, (cudaStream_t)stream);
}

int binfb_hmc_last_chi2(binfb_model *m, int C, double *chi2, void *stream) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!chi2 || C < 1) {
        set_error("hmc_last_chi2: chi2 [n_chains] required");
        return BINFB_EINVAL;
    }
    if (m->kind != BINFB_MODEL_CHROMATIN) {
        set_error("hmc_last_chi2: only the chromatin kernel keeps chi^2 of the current state");
        return BINFB_EUNSUPPORTED;
    }
    if (m->chrom.chi2_chains != C) {
        set_error("hmc_last_chi2: the last binfb_hmc_run on this model did not run " + std::to_string(C) + " chains");
        return BINFB_EINVAL;
    }
    BINFB_CUDA(cudaSetDevice(m->device));
    BINFB_CUDA(cudaMemcpyAsync(chi2, m->chrom.chi2_state, (size_t)C * sizeof(double), cudaMemcpyDeviceToDevice,
                               (cudaStream_t)stream));
    return BINFB_OK;
}

int binfb_rex_pack(const double *chi2, const float *tau, const float *eps, const int32_t *tidx, int C,
                   double n_data, void *records, void *stream) {
    BINFB_TRACE();
    if (!chi2 || !tau || !eps || !tidx || !records || C < 1) {
        set_error("rex_pack: chi2, tau, eps, tidx, records required");
        return BINFB_EINVAL;
    }
    return rex_pack_launch(chi2, tau, eps, tidx, C, n_data, records, (cudaStream_t)stream);
}

int binfb_rex_decide(const void *records_all, int world, int rank, int C, int n_columns, const double *betas,
                     int n_temps, uint64_t seed, uint64_t attempt, double ll_shift, int32_t *tidx, float *beta,
                     float *eps, uint8_t *accept, unsigned long long *pair_counts, double *temp_stats,
                     void *stream) {
    BINFB_TRACE();
    if (!records_all || !betas || !tidx || !beta || !eps) {
        set_error("rex_decide: records_all, betas, tidx, beta, eps required");
        return BINFB_EINVAL;
    }
    if (world < 1 || rank < 0 || rank >= world || C < 1 || n_columns < 1 || C % n_columns != 0) {
        set_error("rex_decide: 0 <= rank < world, n_chains a positive multiple of n_columns");
        return BINFB_EINVAL;
    }
    if (n_temps < 1 || n_temps > BINFB_REX_MAX_TEMPS) {
        set_error("rex_decide: n_temps must be in 1.." + std::to_string(BINFB_REX_MAX_TEMPS));
        return BINFB_EINVAL;
    }
    return rex_decide_launch(records_all, world, rank, C, n_columns, betas, n_temps, seed, attempt, ll_shift, tidx,
                             beta, eps, accept, pair_counts, temp_stats, (cudaStream_t)stream);
}

int binfb_rex_select(const float *q, const float *aux, const int32_t *tidx, int k_sel, int C, int dim,
                     int n_columns, float *out_q, float *out_aux, void *stream) {
    BINFB_TRACE();
    if (!q || !tidx || !out_q || C < 1 || dim < 1 || n_columns < 1 || C % n_columns != 0) {
        set_error("rex_select: q, tidx, out_q required, n_chains a positive multiple of n_columns");
        return BINFB_EINVAL;
    }
    return rex_select_launch(q, aux, tidx, k_sel, C, dim, n_columns, out_q, out_aux, (cudaStream_t)stream);
}

int binfb_rng_fill_host(uint64_t seed, uint64_t draw, uint64_t chain_base, int C, int D,
                        double gamma_shape, float *normals, float *uniforms, double *gammas,
                        int device) {
    if (C < 1 || D < 1) return BINFB_EINVAL;
    BINFB_CUDA(cudaSetDevice(device));
    float *dn = nullptr, *du = nullptr;
    double *dg = nullptr;
    if (normals) BINFB_CUDA(cudaMalloc(&dn, (size_t)C * D * 4));
    if (uniforms) BINFB_CUDA(cudaMalloc(&du, (size_t)C * 4));
    if (gammas) BINFB_CUDA(cudaMalloc(&dg, (size_t)C * 8));
    int rc = rng_fill_launch(seed, draw, chain_base, C, D, gamma_shape, dn, du, dg, 0);
    if (!rc) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess && normals) e = cudaMemcpy(normals, dn, (size_t)C * D * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && uniforms) e = cudaMemcpy(uniforms, du, (size_t)C * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && gammas) e = cudaMemcpy(gammas, dg, (size_t)C * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "rng_fill");
    }
    cudaFree(dn), cudaFree(du), cudaFree(dg);
    return rc;
}

int binfb_chromatin_stream_layout(int n_beads, const float *y_pairs, int roles, int smem_bytes,
                                  float *out, long long capacity, long long *n_floats, int *plan6) {
    if (n_beads < 2 || roles < 0 || roles > 16 || (roles & (roles - 1))) {
        set_error("chromatin_stream_layout: n_beads >= 2, roles in {0 (auto),1,2,4,8,16}");
        return BINFB_EINVAL;
    }
    const ChromPlan pl = chrom_plan(n_beads, smem_bytes > 0 ? smem_bytes : 232448, roles);
    if (n_floats) *n_floats = pl.stream_floats;
    if (plan6) {
        plan6[0] = pl.Q, plan6[1] = pl.KS, plan6[2] = pl.NRB, plan6[3] = pl.R, plan6[4] = pl.Lr,
        plan6[5] = pl.W, plan6[6] = pl.SS, plan6[7] = pl.NS;
    }
    if (!out) return BINFB_OK;
    if (!y_pairs || capacity < pl.stream_floats) {
        set_error("chromatin_stream_layout: buffer too small");
        return BINFB_EINVAL;
    }
    return chrom_build_stream(n_beads, y_pairs, pl, out);
}

int binfb_microbench(int device, int iters, double *ffma_tflops, double *ffma2_tflops,
                     double *mufu_gops, double *sm_clock_mhz) {
    if (iters < 1) return BINFB_EINVAL;
    return microbench_run(device, iters, ffma_tflops, ffma2_tflops, mufu_gops, sm_clock_mhz);
}

}  // extern "C"
