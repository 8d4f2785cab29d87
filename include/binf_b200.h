/*
 * binf_b200 -- C ABI of the B200-native HMC hot path of simeoncarstens/binf.
 *
 * The reference is pure Python and has NO FFI boundary; its seams are two
 * duck-types: the pdf seam (`pdf.log_prob(**{name: x})`, `pdf.gradient(**{name: x})`,
 * binf/samplers/hmc.py:114,143) and the sampler seam (`.pdf`, `.state`, `.sample()`,
 * binf/samplers/gibbs.py:52,121-125,148).  Each entry point below names the reference
 * code whose arithmetic it replaces.  Reference paths are relative to the reference root.
 *
 * Conventions
 *   - plain C types only; every function returns 0 (BINFB_OK) or a negative error code and
 *     leaves a message retrievable with binfb_last_error() (thread-local).
 *   - a model handle is bound to one CUDA device; calls on distinct handles are thread-safe.
 *   - "dev" pointers are CUDA device pointers on the model's device; work is enqueued on
 *     `stream` (a cudaStream_t passed as void*, NULL = default stream) and is asynchronous.
 *     "_host" variants take ordinary host pointers, copy in, run, copy out and synchronise.
 *   - positions q are row-major float32 [n_chains, dim], chain c at q + c*dim; energies and
 *     log-probabilities are float64.
 *   - chains are addressed by a global id (chain_base + local index) in the Philox counter,
 *     so results do not depend on how chains are sharded over GPUs.
 */
#ifndef BINF_B200_H
#define BINF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BINFB_VERSION 100 /* 0.1.0 */

#define BINFB_OK 0
#define BINFB_EINVAL (-1)       /* bad argument */
#define BINFB_ECUDA (-2)        /* CUDA runtime error (message has the CUDA string) */
#define BINFB_EUNSUPPORTED (-3) /* shape outside what the kernels implement */
#define BINFB_ENOMEM (-4)

/* model kinds */
#define BINFB_MODEL_POLYNOMIAL 1
#define BINFB_MODEL_CHROMATIN 2
#define BINFB_MODEL_GENERIC 3

/* model flags */
#define BINFB_FLAG_PRIOR_GRAD 1u /* polynomial: add the Gaussian-prior force (c-mu)/v that the
                                    reference's Posterior.gradient silently drops (quirk Q1,
                                    binf/pdf/posteriors.py:182-185, binf/example/priors.py:45) */
#define BINFB_FLAG_CONTACT_ALGEBRAIC 4u /* chromatin: the algebraic contact function mock = 1/2 (1 + z / sqrt(1 + z^2)),
                                          z = alpha (d_c - d), instead of the logistic 1 / (1 + exp(-z)) (SURVEY.md A.2:
                                          two rsqrt and no exponential per bead pair) */
#define BINFB_FLAG_GENERIC_SCALAR 2u /* generic models: one chain per lane even where the device code compiles over
                                        chain pairs (packed FP32, see binfb_model_create_generic) */

/* Gibbs coupling of the precision update with a trajectory (binf/samplers/gibbs.py:146-149
 * sweeps variables in sorted-name order: 'precision' < 'structure', 'coefficients' < 'precision') */
#define BINFB_GIBBS_NONE 0
#define BINFB_GIBBS_TAU_FIRST 1 /* draw tau | q, then run the trajectory with the new tau */
#define BINFB_GIBBS_TAU_LAST 2  /* run the trajectory, then draw tau | q_new */

typedef struct binfb_model binfb_model;

typedef struct binfb_hmc_opts {
    int32_t n_steps;    /* L, leapfrog steps per trajectory      (HMCSampler.nsteps, hmc.py:54)      */
    int32_t n_traj;     /* trajectories (HMCSampler.sample calls) fused into this launch             */
    int32_t n_adapt;    /* the first n_adapt trajectories adapt the step size (hmc.py:153-157)       */
    int32_t gibbs_mode; /* BINFB_GIBBS_*                                                             */
    double adapt_up;    /* eps *= adapt_up   after an accepted move (hmc.py:188-189)                 */
    double adapt_down;  /* eps *= adapt_down after a rejected move  (hmc.py:190-191)                 */
    uint64_t seed;      /* Philox key                                                                */
    uint64_t draw;      /* index of the first trajectory of this call in the sampler's history       */
    uint64_t chain_base;/* global id of local chain 0                                                */
} binfb_hmc_opts;

/* ---- library ------------------------------------------------------------------------------- */
int binfb_version(void);
const char *binfb_last_error(void);
/* number of visible CUDA devices; SM count, opt-in shared memory per block (bytes), SM clock
 * (kHz) and compute capability (major*10+minor) of one of them */
int binfb_device_count(int *count);
int binfb_device_props(int device, int *sm_count, int *smem_optin, int *clock_khz, int *cc);

/* ---- models -------------------------------------------------------------------------------- */
/* Polynomial forward model + Gaussian error model + Gaussian/Gamma priors:
 * binf/example/likelihood.py:11-68, binf/example/priors.py:10-64.  xs, ys: host float64
 * [n_data]; prior_mean/prior_var: host float64 [n_coeff]. */
int binfb_model_create_polynomial(const double *xs, const double *ys, int n_data, int n_coeff,
                                  const double *prior_mean, const double *prior_var,
                                  double gamma_shape, double gamma_rate, unsigned flags,
                                  int device, binfb_model **out);

/* Chromatin bead chain with a logistic (or, with BINFB_FLAG_CONTACT_ALGEBRAIC, algebraic) contact forward model behind the reference's
 * AbstractForwardModel / GaussianErrorModel / AbstractPrior API (build-defined, SURVEY.md A.2;
 * the reference's Likelihood._evaluate_gradient, binf/pdf/likelihoods.py:148-155, is what the
 * pair kernel fuses).  y_pairs: host float32 [n(n-1)/2] in numpy.triu_indices(n, 1) order.
 * alpha > 0 and |alpha * d_c| <= 80 (BINFB_EINVAL otherwise): the pair loop works on positions scaled by
 * alpha*log2(e) and folds 2^(-alpha d_c log2 e) into a multiplier that has to stay a normal float (the
 * algebraic form only needs alpha > 0).  flags bits 8..12: force the warps per chain (tests). */
int binfb_model_create_chromatin(int n_beads, const float *y_pairs, double alpha, double d_c,
                                 double k_bb, double l0, double conf_s, double gamma_shape,
                                 double gamma_rate, unsigned flags, int device,
                                 binfb_model **out);
/* A USER-DEFINED per-datum forward model: the reference's extension point
 * AbstractForwardModel._evaluate / _evaluate_jacobi_matrix (binf/model/forwardmodels.py:30-38),
 * given as CUDA device code and compiled at run time (NVRTC, sm_100a) into the same fused
 * HMC / log-prob kernels the polynomial model has.  device_code must define
 *     __device__ float binfb_mock(const float *theta, const float *x, float *dmock);
 * returning f_n(theta) for one datum with abscissae x[x_dim] and writing dmock[k] = d f_n / d theta_k,
 * k < n_params (<= 16).  Error model: GaussianErrorModel (binf/example/likelihood.py:54-61); priors:
 * independent Gaussians on theta, Gamma on the precision.  xs: host f64 [n_data, x_dim], ys [n_data].
 * Where the data rows fit the constant bank (48 KiB) and n_params <= 8 the code is first compiled with `float`
 * standing for a pair of chains (all arithmetic packed: FFMA2 / FADD2 / FMUL2, two chains per lane; the results are
 * bit-identical to the scalar build); code that does not compile that way (branches on values, casts to double,
 * functions outside the usual float math set) is compiled as written, one chain per lane, which
 * BINFB_FLAG_GENERIC_SCALAR also forces.  Launches with enough work run a trajectory as three kernels (begin /
 * middle / end, option "generic.split"), whose middle kernel gets the data rows as uniform-register operands.
 * Compile errors return BINFB_EINVAL with the NVRTC log in binfb_last_error(). */
int binfb_model_create_generic(const char *device_code, int n_params, int x_dim, const double *xs,
                               const double *ys, int n_data, const double *prior_mean,
                               const double *prior_var, double gamma_shape, double gamma_rate,
                               unsigned flags, int device, binfb_model **out);
/* compile device_code only (needs libnvrtc, no GPU): 0 if it compiles; the NVRTC log (warnings or
 * errors) is copied to log_out[log_capacity] if non-NULL */
int binfb_generic_compile_check(const char *device_code, int n_params, int x_dim, char *log_out,
                                int log_capacity);
int binfb_model_destroy(binfb_model *m);
int binfb_model_info(const binfb_model *m, int *kind, int *dim, long long *n_data, int *device);
/* update the Gamma prior on the precision (it enters log_prob and the Gibbs update only) */
int binfb_model_set_gamma_prior(binfb_model *m, double shape, double rate);
/* tuning knobs ("poly.group", "poly.chains_per_thread", "poly.block", "chrom.warps"; value < 0 restores
 * the heuristic; "host.pipeline" = 0 makes binfb_hmc_run_host copy in, run, copy out one after the other instead of
 * overlapping the state copies with the kernel; "host.chunks" (2 .. 64, default 16) is the number of pieces the state travels in; "poly.uniform_rows" = 0 keeps the data rows in shared memory instead of the constant bank; "generic.split" = 1 / 0 runs every / no trajectory of a user-defined model as three launches) and model extras: "chrom.ev_k", "chrom.ev_d" switch on the excluded-volume prior
 * -k_ev sum_{i<j} max(0, d_ev - d_ij)^4 of the chromatin model, whose force is
 * fused into the pair loop */
int binfb_model_set_option(binfb_model *m, const char *key, double value);
/* what a model was built with: "generic.packed" (1 = two chains per lane, packed FP32), "generic.uniform_rows",
 * "generic.warps_per_set", and the current value of every knob binfb_model_set_option takes */
int binfb_model_get_option(const binfb_model *m, const char *key, double *value);

/* ---- pdf seam: AbstractBinfPDF.log_prob / gradient ------------------------------------------ */
/* log_prob of the conditional posterior over the sampled variable at per-chain precision tau
 * (binf/pdf/posteriors.py:125-151) and the gradient of the ENERGY -log p
 * (binf/pdf/posteriors.py:173-187 -> binf/pdf/likelihoods.py:148-155).  beta (NULL = 1) tempers
 * the likelihood term.  Outputs may be NULL: logp [C] f64, grad [C, dim] f32, chi2 [C] f64
 * (sum of squared residuals, what GammaSampler needs: binf/example/samplers.py:34-41). */
int binfb_logprob_grad(binfb_model *m, const float *q_dev, const float *tau_dev,
                       const float *beta_dev, int n_chains, double *logp_dev, float *grad_dev,
                       double *chi2_dev, void *stream);
int binfb_logprob_grad_host(binfb_model *m, const float *q, const float *tau, const float *beta,
                            int n_chains, double *logp, float *grad, double *chi2);
/* forward model mock data (AbstractForwardModel.__call__, binf/__init__.py:105-120 ->
 * binf/example/likelihood.py:24-26): mock [C, n_data] f32 */
int binfb_forward_host(binfb_model *m, const float *q, int n_chains, float *mock);

/* ---- sampler seam: HMCSampler.sample / _leapfrog -------------------------------------------- */
/* n_traj fused HMC transitions per chain: momentum draw, L-step leapfrog with L+1 force
 * evaluations, Metropolis accept, per-chain step-size adaption, optional conjugate precision
 * update (binf/samplers/hmc.py:92-125,136-164,183-191; binf/example/samplers.py:27-51).
 *   q [C, dim] in/out; tau [C] in/out; beta [C] or NULL; eps [C] in/out.
 *   p0 [C, dim] / u [C]: injected momenta / uniforms (parity tests; require n_traj == 1) or NULL
 *   for Philox draws.  gamma_draws [C]: injected standard-Gamma variates or NULL.
 *   accepted [C] u8, e_before/e_after [C] f64, q_end/p_end [C, dim]: last trajectory's accept
 *   flag, Hamiltonians and leapfrog end point (any may be NULL); n_accepted [C] i32 counts over
 *   the call; stats_dev [4] f64 is atomically incremented by {accepted, proposed, sum eps,
 *   sum exp(min(0,-dH))} for the diagnostics all-reduce (NULL to skip). */
int binfb_hmc_run(binfb_model *m, float *q_dev, float *tau_dev, const float *beta_dev,
                  float *eps_dev, int n_chains, const binfb_hmc_opts *opts, const float *p0_dev,
                  const float *u_dev, const double *gamma_draws_dev, uint8_t *accepted_dev,
                  double *e_before_dev, double *e_after_dev, float *q_end_dev, float *p_end_dev,
                  int32_t *n_accepted_dev, double *stats_dev, void *stream);
int binfb_hmc_run_host(binfb_model *m, float *q, float *tau, const float *beta, float *eps,
                       int n_chains, const binfb_hmc_opts *opts, const float *p0, const float *u,
                       const double *gamma_draws, uint8_t *accepted, double *e_before,
                       double *e_after, float *q_end, float *p_end, int32_t *n_accepted,
                       double *stats);

/* ---- GammaSampler.sample (binf/example/samplers.py:27-51) ----------------------------------- */
/* tau[c] = Gamma(beta*n_data/2 + a - 1, 1) / (beta*chi2(q[c])/2 + b)   (shape quirk Q3 kept) */
int binfb_gibbs_precision(binfb_model *m, const float *q_dev, float *tau_dev,
                          const float *beta_dev, int n_chains, uint64_t seed, uint64_t draw,
                          uint64_t chain_base, const double *gamma_draws_dev, double *chi2_dev,
                          void *stream);
int binfb_gibbs_precision_host(binfb_model *m, const float *q, float *tau, const float *beta,
                               int n_chains, uint64_t seed, uint64_t draw, uint64_t chain_base,
                               const double *gamma_draws, double *chi2);

/* ---- replica exchange (build-defined; the reference only alludes to it at
 *      binf/samplers/hmc.py:171-177) --------------------------------------------------------- */
/* State-swap variant (two fixed-temperature ranks exchange states; kept for callers that want the cold
 * replica pinned to one rank): accept[c] = u < exp(-(beta_a - beta_b)(ll_a[c] - ll_b[c])), u from
 * Philox(seed; attempt, pair_id, chain_base + c).  BOTH partners must pass the same chain_base (a swap
 * stream id, not the ranks' HMC chain bases) so that they reach the same decision without talking. */
int binfb_swap_decide(const double *ll_a_dev, const double *ll_b_dev, double beta_a,
                      double beta_b, int n_chains, uint64_t seed, uint64_t attempt,
                      uint64_t pair_id, uint64_t chain_base, uint8_t *accept_dev, void *stream);
/* q_mine[c] <- q_theirs[c] (and eps likewise if non-NULL) where accept[c] */
int binfb_swap_apply(float *q_mine_dev, const float *q_theirs_dev, float *eps_mine_dev,
                     const float *eps_theirs_dev, const uint8_t *accept_dev, int n_chains,
                     int dim, void *stream);

/* Replica exchange by LABEL swap ("swap beta (and eps) labels so no state moves").  The
 * ensemble is a grid [temperature][column]; a chain never moves, it carries a temperature index.  An attempt
 * is: binfb_hmc_last_chi2 (chi^2 of every chain's current state, left behind by the trajectory kernel -- no
 * extra pair sweep), binfb_rex_pack (one 16-byte record {f64 log L, i32 tidx, f32 eps} per chain), an
 * all-gather of the records of all ranks (NCCL; world x n_chains x 16 bytes), binfb_rex_decide. */
#define BINFB_REX_MAX_TEMPS 64
#define BINFB_REX_RECORD_BYTES 16
/* chi2_dev [n_chains] f64 <- sum of squared residuals of every chain's CURRENT state as of the last
 * binfb_hmc_run on this model with the same n_chains (chromatin models; BINFB_EUNSUPPORTED otherwise) */
int binfb_hmc_last_chi2(binfb_model *m, int n_chains, double *chi2_dev, void *stream);
/* records_dev [n_chains] <- {-tau chi2/2 + n_data log(tau)/2, tidx, eps}  (binf/example/likelihood.py:54-57) */
int binfb_rex_pack(const double *chi2_dev, const float *tau_dev, const float *eps_dev,
                   const int32_t *tidx_dev, int n_chains, double n_data, void *records_dev, void *stream);
/* records_all_dev [world, n_chains] (rank-major, the all-gather of binfb_rex_pack).  Local chain i sits in
 * column i % n_columns.  Attempt a pairs temperature indices (k, k+1) with k + a even.  A chain at index k
 * finds the chain of its column that holds the partner index k', accepts iff
 * u < exp(-(beta_k - beta_k')(l_mine - l_theirs)) with u = Philox(seed; a, min(k, k'), column) -- the same
 * on both sides, independent of rank and chain base -- and then takes over k', betas[k'] and the partner's
 * step size: tidx_dev, beta_dev, eps_dev [n_chains] are updated in place.  betas: HOST f64 [n_temps].
 * Optional: accept_dev [n_chains] u8; pair_counts_dev [n_temps-1][2] u64 += {attempted, accepted} (counted by
 * the lower index of a pair); temp_stats_dev [n_temps][3] f64 += {1, l - ll_shift, (l - ll_shift)^2} of every
 * chain at its temperature BEFORE the swap (for the ladder adaption). */
int binfb_rex_decide(const void *records_all_dev, int world, int rank, int n_chains, int n_columns,
                     const double *betas, int n_temps, uint64_t seed, uint64_t attempt, double ll_shift,
                     int32_t *tidx_dev, float *beta_dev, float *eps_dev, uint8_t *accept_dev,
                     unsigned long long *pair_counts_dev, double *temp_stats_dev, void *stream);
/* out_q_dev [n_columns, dim] <- q[c] for the local chains with tidx[c] == k_sel (row = their column), zero
 * elsewhere; out_aux_dev [n_columns] likewise from aux_dev (either may be NULL).  Summed over the ranks this
 * assembles the replicas of one temperature (k_sel = 0: the posterior samples) wherever they live. */
int binfb_rex_select(const float *q_dev, const float *aux_dev, const int32_t *tidx_dev, int k_sel,
                     int n_chains, int dim, int n_columns, float *out_q_dev, float *out_aux_dev, void *stream);

/* ---- sample sink ------------------------------------------------------- */
/* What the reference's driver loop does with every Gibbs sweep, for chains resident in HBM:
 * `samples.append(deepcopy(gips.sample()))` (example_script.py:32-34), the burn-in/thinning slice
 * `samples[20000::20]` (example_script.py:41), the MAP estimate = kept sample of maximum
 * log-probability (binf/example/misc.py:18-22, example_script.py:51) and the posterior summaries
 * the plots derive from the kept samples (binf/example/plots.py).
 *   sweep t = 0, 1, ... (one push each).  Sweeps t >= burn_in enter the per-chain running moments
 *   (Welford, float64); sweeps with (t - burn_in) % thin == 0 are also copied into a ring of
 *   `capacity` kept samples (the oldest is overwritten); with BINFB_SINK_TRACK_MAP every chain keeps
 *   its kept sample of maximum logp (ties: the first, like numpy.argmax). */
#define BINFB_SINK_TRACK_MAP 1u
typedef struct binfb_sink binfb_sink;
int binfb_sink_create(int n_chains, int dim, int capacity, int burn_in, int thin, unsigned flags,
                      int device, binfb_sink **out);
int binfb_sink_destroy(binfb_sink *s);
/* sweeps pushed, sweeps in the moments, samples kept so far (the ring holds the last `capacity`) */
int binfb_sink_info(const binfb_sink *s, long long *n_pushed, long long *n_moment, long long *n_kept);
/* one sweep: q [C, dim] f32, aux [C] f32 or NULL (e.g. the precision), logp [C] f64 (NULL unless
 * the sink tracks the MAP).  One fused HBM-bound pass; asynchronous on `stream`. */
int binfb_sink_push(binfb_sink *s, const float *q_dev, const float *aux_dev, const double *logp_dev,
                    void *stream);
int binfb_sink_push_host(binfb_sink *s, const float *q, const float *aux, const double *logp);
/* per-dimension summaries over all chains, each [dim] f64 or NULL: posterior mean, pooled
 * within-chain variance W, Gelman-Rubin R-hat sqrt(((n-1)/n W + B/n) / W), and the effective sample
 * size per chain W / var_c(chain means) (independent chains). */
int binfb_sink_summary(binfb_sink *s, double *mean_dev, double *var_dev, double *rhat_dev,
                       double *ess_dev, void *stream);
int binfb_sink_summary_host(binfb_sink *s, double *mean, double *var, double *rhat, double *ess);
/* the raw per-dimension sums behind the summary, for merging across GPUs (binf_b200.distributed.
 * merge_sink_sums): pivot[d] = running mean of chain 0, sum_c (mean_c - pivot), sum_c (mean_c - pivot)^2,
 * sum_c M2_c; each [dim] f64, any may be NULL */
int binfb_sink_sums_host(binfb_sink *s, double *pivot, double *sum_dev, double *sum_dev2, double *sum_m2);
/* per-chain running mean and unbiased variance, [C, dim] f64 each (either may be NULL) */
int binfb_sink_moments_host(binfb_sink *s, double *mean, double *var);
/* kept samples number first .. first+count-1 (0 = the first ever kept) -> q_out [count, C, dim],
 * aux_out [count, C]; they must still be in the ring */
int binfb_sink_read_host(binfb_sink *s, long long first, long long count, float *q_out, float *aux_out);
/* per-chain MAP candidate: logp [C] f64, state [C, dim] f32, aux [C] f32 (any may be NULL) */
int binfb_sink_map_host(binfb_sink *s, double *logp, float *q_map, float *aux_map);

/* ---- the shipped example's RWMC sampler and predictive density --------- */
/* n_moves random-walk Metropolis moves per chain: proposal = state + U(-stepsize, stepsize)^dim,
 * accept iff u < exp(-(E_new - E_old)), E = -log_prob (RWMCSampler.sample,
 * binf/example/samplers.py:78-92).  q [C, dim] in/out; stepsize [C]; change [C, dim] / u [C]:
 * injected proposal displacements / uniforms (parity tests, n_moves == 1) or NULL for Philox draws;
 * accepted [C] u8 (last move), n_accepted [C] i32 (over the call), logp [C] f64 = log_prob of the
 * final state; any output may be NULL. */
int binfb_rwmc_run(binfb_model *m, float *q_dev, const float *tau_dev, const float *beta_dev,
                   const float *stepsize_dev, int n_chains, int n_moves, uint64_t seed, uint64_t draw,
                   uint64_t chain_base, const float *change_dev, const float *u_dev,
                   uint8_t *accepted_dev, int32_t *n_accepted_dev, double *logp_dev, void *stream);
int binfb_rwmc_run_host(binfb_model *m, float *q, const float *tau, const float *beta,
                        const float *stepsize, int n_chains, int n_moves, uint64_t seed, uint64_t draw,
                        uint64_t chain_base, const float *change, const float *u, uint8_t *accepted,
                        int32_t *n_accepted, double *logp);
/* posterior-predictive density of new data (x_g, y_g) under the polynomial model: the mean over
 * the samples of N(y_g; polyval(x_g, coeffs_s), 1/precision_s), summed as a log-sum-exp in
 * float64 (predict, binf/example/misc.py:3-16).  coeffs [S, n_coeff] f32 (ascending powers),
 * precision [S] f32, xs/ys [n_points] f64 -> out [n_points] f64.  Host pointers. */
int binfb_posterior_predictive_host(const float *coeffs, const float *precision, long long n_samples,
                                    int n_coeff, const double *xs, const double *ys, int n_points,
                                    double *out, int device);

/* ---- test / measurement helpers ------------------------------------------------------------- */
/* the device RNG streams, for statistical tests: normals [C, dim] f32 exactly as the momentum
 * draw of trajectory `draw`; uniforms [C]; standard gammas of the given shape [C] f64 */
int binfb_rng_fill_host(uint64_t seed, uint64_t draw, uint64_t chain_base, int n_chains, int dim,
                        double gamma_shape, float *normals, float *uniforms, double *gammas,
                        int device);
/* host-side layout pass of the chromatin contact stream (no GPU needed).  roles = warps per
 * chain (0 = the heuristic for smem_bytes of opt-in shared memory, 0 = 227 KiB; else 1..16, a power of 2).  Writes the number
 * of float32 the kernel streams per force evaluation and plan6[8] = {quads, partner steps, row
 * blocks, roles, slots per row block, chains per CTA, warp-steps per ring stage, ring depth}; if
 * out != NULL (capacity floats) fills it. */
int binfb_chromatin_stream_layout(int n_beads, const float *y_pairs, int roles, int smem_bytes,
                                  float *out, long long capacity, long long *n_floats, int *plan6);
/* FP32 pipe microbenchmarks used as roofline denominators: dependent-chain-free FFMA, packed
 * FFMA2 and MUFU (rsqrt/ex2/rcp mix) issue loops over the whole device.  Results in
 * TFLOP/s (2 flop per FMA lane-op) and Gop/s. */
int binfb_microbench(int device, int iters, double *ffma_tflops, double *ffma2_tflops,
                     double *mufu_gops, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* BINF_B200_H */
