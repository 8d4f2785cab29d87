// Host-side model descriptor and launcher declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/binf_b200.h"

namespace binfb {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

#define BINFB_CUDA(call)                                              \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return ::binfb::cuda_fail(e__, #call); \
    } while (0)

// Tracing (the reference has none): every compute entry point of the ABI is an NVTX range named after
// the symbol, so that a profile taken with `ncu --nvtx --nvtx-include "binfb_hmc_run/"` (or any NVTX-aware tool)
// attributes kernels and copies to the call that issued them.  Header-only NVTX 3: without a tool attached a range
// is one load and one branch.
struct TraceRange {
    explicit TraceRange(const char *name) { nvtxRangePushA(name); }
    ~TraceRange() { nvtxRangePop(); }
    TraceRange(const TraceRange &) = delete;
    TraceRange &operator=(const TraceRange &) = delete;
};
#define BINFB_TRACE() ::binfb::TraceRange binfb_trace_range_(__func__)

// BEGIN_KERNEL_ARGS  (this block is also compiled by NVRTC as part of the generic-model source)
// arguments common to every HMC launch (device pointers)
struct HmcArgs {
    float *q;            // [C, D] state, in/out
    float *tau;          // [C] in/out
    const float *beta;   // [C] or null
    float *eps;          // [C] in/out
    int C;
    int L, n_traj, n_adapt, gibbs_mode;
    float adapt_up, adapt_down;
    uint64_t seed, draw, chain_base;
    const float *p0;     // [C, D] or null
    const float *u;      // [C] or null
    const double *gamma_draws;  // [C] or null
    uint8_t *accepted;   // [C] or null
    double *e_before, *e_after;  // [C] or null
    float *q_end, *p_end;        // [C, D] or null
    int32_t *n_accepted; // [C] or null
    double *stats;       // [4] or null
    double gamma_shape, gamma_rate;  // Gamma prior on tau
};

struct GradArgs {
    const float *q;
    const float *tau;
    const float *beta;
    int C;
    double *logp;
    float *grad;
    double *chi2;
    double gamma_shape, gamma_rate;
};

// generic per-datum model (generic.cu / generic_kernel.cuh): data rows [N, stride] = x[0..XD-1], -y, pad
struct GenDev {
    const float *rows;
    int N, stride;
    float prior_mean[16], prior_inv_var[16];
    unsigned flags;
    // workspace of the split trajectory (gen_traj_begin / _mid / _end): proposal and momentum [C, K], energies [C]
    float *w_q, *w_p;
    double *w_h0, *w_chi0, *w_chiL;
};
// END_KERNEL_ARGS

// ---- polynomial --------------------------------------------------------------------------
struct PolyModel {
    int N = 0, K = 0, stride = 0;  // data points, coefficients, floats per smem row
    float *rows = nullptr;         // device [N, stride]: x, x^2, .., x^(K-1), y, pad
    float prior_mean[8], prior_inv_var[8];
    unsigned flags = 0;
    int opt_group = -1, opt_jchains = -1, opt_block = -1;
    int opt_ur = -1;                // 0 disables the uniform-row mapping (poly.cu)
    unsigned long long uid = 0;     // unique per model: owner tag of the constant-bank copy of the rows
};
int poly_hmc_launch(const PolyModel &m, const HmcArgs &a, int sm_count, int smem_optin,
                    cudaStream_t s);
int poly_grad_launch(const PolyModel &m, const GradArgs &a, int sm_count, int smem_optin,
                     cudaStream_t s);
int poly_forward_launch(const PolyModel &m, const float *q, int C, float *mock, cudaStream_t s);

// ---- chromatin -----------------------------------------------------------------------------
struct ChromPlan {
    int n_pad = 0, Q = 0, KS = 0, NRB = 0;  // padded beads, quads, partner steps, row blocks
    int R = 1, W = 1;                       // warps per chain, chains per CTA
    int Lr = 0, SS = 4, S_pad = 0;          // slots per row block, steps per stage, padded slots
    int NS = 4;                             // ring depth
    int lockstep = 0;                       // the roles of a chain advance step by step together
    size_t fixed_smem = 0, per_chain_smem = 0;
    long long stream_floats = 0;
};
struct ChromModel {
    int n = 0;
    ChromPlan plan;
    long long M = 0;
    float *ystream = nullptr;      // device [S_pad * R][4][32] float4
    // small-batch alternative: twice the warps per chain in lockstep (own stream layout), used when the
    // batch cannot fill the SMs with the primary plan
    ChromPlan plan_alt;
    float *ystream_alt = nullptr;
    float *ypairs = nullptr;       // device [M] (triu order; forward/mock kernel only)
    float alpha = 0, d_c = 0, k_bb = 0, l0 = 0, inv_s2 = 0;
    float ev_k = 0, ev_d = 0;      // excluded-volume prior k_ev sum max(0, d_ev - d_ij)^4 (0 = off)
    unsigned flags = 0;
    int opt_warps = -1;
    int opt_sets = -1;             // 0: one set (all groups in one pass-major item sequence)
    // per-launch workspace (grown on demand)
    int ws_chains = 0;
    float *qw = nullptr, *pw = nullptr;   // [C, D] working position / momentum
    double *h0 = nullptr, *chi2_0 = nullptr, *chi2_state = nullptr;  // [C]
    float *tau_w = nullptr;               // [C] precision used by the running trajectory
    int *sched = nullptr;                 // [1 + n_octets]: item counter, per-octet pass counters
    int sched_len = 0;
    int chi2_chains = 0;                  // chains whose chi2_state the last HMC launch left valid (0 = none)
};
ChromPlan chrom_plan(int n, int smem_optin, int force_roles);
ChromPlan chrom_plan_small_batch(int n, int smem_optin, const ChromPlan &primary);
int chrom_build_stream(int n, const float *y_pairs, const ChromPlan &pl, float *out);
int chrom_reserve(ChromModel &m, int C);
// pipelined host call (binfb_hmc_run_host): launch shape and the device words the copy streams talk to
struct ChromPipe {
    int W = 0, n_groups = 0, groups_per_chunk = 0, n_chunks = 0;
    int *gate = nullptr;  // device: 1 + number of leading chains whose positions have arrived
    int *done = nullptr;  // device [n_chunks]: groups of the chunk that have finished their last pass
    int header[2] = {0, 0};  // host staging of {gate, groups_per_chunk} (must outlive the async copy)
    cudaEvent_t header_written = nullptr;  // recorded on the launch stream between the header and the kernel
};
int chrom_pipe_shape(const ChromModel &m, int C, int sm_count, int max_chunks, ChromPipe *pipe);
int chrom_hmc_launch(ChromModel &m, const HmcArgs &a, int sm_count, int smem_optin,
                     cudaStream_t s, ChromPipe *pipe = nullptr);
int chrom_grad_launch(ChromModel &m, const GradArgs &a, int sm_count, int smem_optin,
                      cudaStream_t s);
int chrom_forward_launch(const ChromModel &m, const float *q, int C, float *mock, cudaStream_t s);

// ---- generic per-datum model compiled at run time (NVRTC) ----------------------------------------
struct GenModel {
    int K = 0, XD = 0, G = 8;
    int ur = 0;  // uniform-row mapping: rows in the module's constant bank, G warps per chain set
    int pack = 0;  // two chains per lane, the user's arithmetic compiled over binfb_f2 (generic_pack.cuh)
    int srows = 0;  // uniform-row mapping with the rows in shared memory instead of the constant bank
    size_t smem_bytes = 0;  // dynamic shared memory of the kernels (the rows)
    int ws_chains = 0;      // chains the split-trajectory workspace is sized for
    int opt_split = -1;     // "generic.split": 1 = always run a trajectory as begin / middle / end kernels, 0 = never
    void *k_begin = nullptr, *k_mid = nullptr, *k_end = nullptr;
    GenDev dev;
    float *rows = nullptr;
    void *library = nullptr;                       // cudaLibrary_t
    void *k_hmc = nullptr, *k_grad = nullptr, *k_fwd = nullptr;  // cudaKernel_t
};
int gen_create(GenModel &g, const char *user_code, int n_params, int x_dim, const double *xs, const double *ys,
               int n_data, const double *prior_mean, const double *prior_var, unsigned flags);
void gen_destroy(GenModel &g);
int gen_hmc_launch(GenModel &g, const HmcArgs &a, cudaStream_t s);
int gen_grad_launch(const GenModel &g, const GradArgs &a, cudaStream_t s);
int gen_forward_launch(const GenModel &g, const float *q, int C, float *mock, cudaStream_t s);

// ---- misc ----------------------------------------------------------------------------------
int gibbs_tau_launch(const double *chi2, float *tau, const float *beta, int C, double n_data,
                     double shape, double rate, uint64_t seed, uint64_t draw, uint64_t chain_base,
                     const double *gamma_draws, cudaStream_t s);
int rng_fill_launch(uint64_t seed, uint64_t draw, uint64_t chain_base, int C, int D,
                    double gamma_shape, float *normals, float *uniforms, double *gammas,
                    cudaStream_t s);
int swap_decide_launch(const double *ll_a, const double *ll_b, double beta_a, double beta_b, int C,
                       uint64_t seed, uint64_t attempt, uint64_t pair_id, uint64_t chain_base,
                       uint8_t *accept, cudaStream_t s);
int swap_apply_launch(float *q_mine, const float *q_theirs, float *eps_mine,
                      const float *eps_theirs, const uint8_t *accept, int C, int D,
                      cudaStream_t s);
int rex_pack_launch(const double *chi2, const float *tau, const float *eps, const int32_t *tidx, int C,
                    double n_data, void *rec, cudaStream_t s);
int rex_decide_launch(const void *all, int world, int rank, int C, int n_columns, const double *betas, int n_temps,
                      uint64_t seed, uint64_t attempt, double ll_shift, int32_t *tidx, float *beta, float *eps,
                      uint8_t *accept, unsigned long long *pair_counts, double *temp_stats, cudaStream_t s);
int rex_select_launch(const float *q, const float *aux, const int32_t *tidx, int k_sel, int C, int D, int n_columns,
                      float *out_q, float *out_aux, cudaStream_t s);
int microbench_run(int device, int iters, double *ffma, double *ffma2, double *mufu,
                   double *clock_mhz);

}  // namespace binfb

struct binfb_model {
    int kind = 0;
    int device = 0;
    int dim = 0;
    long long n_data = 0;
    int sm_count = 0, smem_optin = 0;
    double gamma_shape = 1.0, gamma_rate = 1.0;
    binfb::PolyModel poly;
    binfb::ChromModel chrom;
    binfb::GenModel gen;
    // random-walk Metropolis workspace (rwmc.cu): proposals [cap, dim], log-probs 3 x [cap]
    float *rw_prop = nullptr;
    double *rw_lp[3] = {nullptr, nullptr, nullptr};
    size_t rw_cap = 0;
    // cached device buffers for the *_host entry points
    size_t hb_bytes = 0;
    char *hb = nullptr;
    cudaStream_t hstream = nullptr;
    cudaStream_t hstream_in = nullptr, hstream_out = nullptr;  // copy streams of the pipelined host call
    cudaEvent_t hevent = nullptr;
    bool host_pipeline = true;  // option "host.pipeline": overlap the copies of the *_host calls with the kernel
    int host_chunks = 16;       // option "host.chunks": pieces the state travels in (2 .. 64)
};
