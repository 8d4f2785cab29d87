// Fused HMC / log-prob+gradient kernels for a USER-DEFINED per-datum forward model -- compiled at run
// time with NVRTC for sm_100a (not compiled by nvcc).
//
// The reference's extension point is AbstractForwardModel._evaluate / _evaluate_jacobi_matrix
// (binf/model/forwardmodels.py:30-38): mock data f(theta) [N] and its Jacobian [K, N], which
// Likelihood._evaluate_gradient contracts with the error model's gradient
// (binf/pdf/likelihoods.py:148-155).  Here the user supplies the same two things as CUDA device code,
// one datum at a time,
//
//     __device__ float binfb_mock(const float *theta, const float *x, float *dmock);
//         theta[GEN_K]  the sampled variable of one chain
//         x[GEN_XD]     the abscissae of one datum
//         returns f_n(theta) and writes dmock[k] = d f_n / d theta_k
//
// and this file wraps it into the same fused transition the built-in polynomial model gets: the chains
// walk the data (mappings below), reduce the GEN_K gradient sums and chi^2 in a fixed order, and keep q, p
// in registers for the whole trajectory (leapfrog, energies,
// Metropolis test, step-size adaption, conjugate precision update: binf/samplers/hmc.py:92-125,136-164,
// 183-191; binf/example/samplers.py:27-51).  Error model: GaussianErrorModel
// (binf/example/likelihood.py:54-61); priors: Gaussian on theta, Gamma on the precision.
//
// Compile-time parameters (-D): GEN_K, GEN_XD, GEN_G (power of two <= 32), GEN_UR, GEN_PACK, GEN_SROWS.
//
// GEN_UR = 1, the uniform-row mapping (like poly.cu): lane l of a warp owns chain l (chains 2l, 2l + 1 with
// GEN_PACK) of its set and every lane walks the same rows; the G warps of a set own the same chains and split the
// rows into G contiguous ranges, and the K gradient sums (+ chi^2) of a pass are combined through shared memory in
// a fixed order, so that all G warps keep bitwise identical q, p.  The rows sit in the module's own constant bank
// and, with GEN_SROWS = 1, also in shared memory (see gen_row_load).  GEN_UR = 0: a chain is owned by G lanes that
// stride over rows in global memory (data sets that do not fit the constant bank).
//
// GEN_PACK = 1 (with GEN_UR = 1; the default where K <= 8): every lane owns TWO chains and the user's function is
// compiled over the pack type binfb_f2 (generic_pack.cuh; the model source is compiled with `float` standing for
// binfb_f2), so that its arithmetic comes out as FFMA2 / FADD2 / FMUL2 -- two chains per instruction like the
// built-in polynomial kernel.  Each component of a pack sees exactly the operations the scalar build performs (the
// two builds agree bit for bit).  A model whose code does not compile over packs (data-dependent branches,
// unsupported functions) is built with GEN_PACK = 0.  Measured (profiles/r2_generic_vs_builtin.txt, the cubic,
// 65,536 chains x 1000 rows): 0.41-0.51 ms per trajectory as pairs, 0.64 with one chain per lane, 0.36 built-in.
constexpr int K = GEN_K, XD = GEN_XD, G = GEN_G;
#ifndef GEN_UR
#define GEN_UR 0
#endif
#ifndef GEN_PACK
#define GEN_PACK 0
#endif
static_assert(!GEN_PACK || GEN_UR, "chain pairs are packed in the uniform-row mapping only");
constexpr bool UR = GEN_UR != 0;
constexpr int J = GEN_PACK ? 2 : 1;  // chains per lane
// threads per block (= gen_launch's): 128 -- one chain set of four warps per block in the uniform-row mapping, so that
// the sets spread evenly over the SMs (65,536 chains as pairs = 1024 sets = 6.9 per SM: 7 or 6; with two sets per
// block it is 8 or 6 and the middle kernel of the split trajectory runs 12 % longer)
constexpr int GEN_BLOCK = 128;
constexpr int GEN_CROW_FLOATS = 12288;  // 48 KiB of the constant bank
constexpr int GEN_STRIDE = (XD + 1 + 3) / 4 * 4;  // floats per data row: x[XD], -y, padding (= gen_create's)
// slot of -y in a row.  With one abscissa it sits at slot 2, not next to x: side by side the two are fetched with
// ONE 64-bit load, and if the user's code needs x in a vector register as well (x * x does: a packed multiply
// takes one uniform operand) ptxas then loads both with a per-lane LDC -- every FFMA2 fed from that load takes 3
// FMA-pipe cycles instead of 2 (measured: 0.63 against 0.55 ms for the cubic, chain pairs, split trajectory).
constexpr int GEN_YSLOT = XD == 1 ? 2 : XD;
#ifndef GEN_SROWS
#define GEN_SROWS 0
#endif
// Where the rows of the uniform-row mapping sit.  The constant bank of the module always holds them (rows up to
// 48 KiB); GEN_SROWS = 1 additionally keeps a copy in (dynamic) shared memory, made by every block of the
// one-launch kernels at their start, which those kernels read instead (one broadcast LDS.128 per 4 floats):
// in the fused trajectory kernel ptxas fetches constant-bank rows with per-lane LDC, which caps it at about one
// load per 2.4 SM-cycles.  The split trajectory's middle kernel reads the constant bank: there the loads go
// through the uniform datapath (LDCU) and the values are uniform-register operands.
#if GEN_UR
__constant__ float4 gen_crows[GEN_CROW_FLOATS / 4];
extern __shared__ float4 gen_srows[];
template <bool SROWS>
__device__ __forceinline__ float4 gen_row_load(int i) {
    if (SROWS) return gen_srows[i];
    return gen_crows[i];
}
template <bool SROWS>
__device__ __forceinline__ void gen_load_rows(const GenDev &gm) {
    if (!SROWS) return;
    const float4 *src = reinterpret_cast<const float4 *>(gm.rows);
    for (int i = threadIdx.x; i < gm.N * (GEN_STRIDE / 4); i += GEN_BLOCK) gen_srows[i] = __ldg(src + i);
    __syncthreads();
}
#else
template <bool SROWS>
__device__ __forceinline__ void gen_load_rows(const GenDev &) {}
#endif
constexpr bool SROWS_ONE_LAUNCH = GEN_UR && GEN_SROWS;  // the fused trajectory kernel and the log-prob kernel
constexpr int GEN_SETS = UR ? (GEN_BLOCK / 32) / G : 1;  // chain sets per block

#if GEN_PACK
typedef binfb_f2 gen_real;
__device__ __forceinline__ gen_real gen_join(const float (&a)[J]) { return binfb_f2(a[0], a[1]); }
__device__ __forceinline__ float gen_part(gen_real v, int j) { return j ? v.v.y : v.v.x; }
#else
typedef float gen_real;
__device__ __forceinline__ gen_real gen_join(const float (&a)[J]) { return a[0]; }
__device__ __forceinline__ float gen_part(gen_real v, int) { return v; }
#endif

// who am I: g = share of the rows (lane in the group, or warp in the set), c = first of this thread's J chains
struct GenMap {
    int g, lane, bar_id, set;
    long long c;
    __device__ __forceinline__ void init() {
        if (UR) {
            // broadcast from lane 0: the compiler keeps everything derived from it in uniform registers
            const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
            set = warp / G, g = warp % G, lane = threadIdx.x & 31, bar_id = 1 + set;
            c = (((long long)blockIdx.x * GEN_SETS + set) * 32 + lane) * J;
        } else {
            g = threadIdx.x % G, lane = 0, bar_id = 0, set = 0;
            c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
        }
    }
};

__device__ __forceinline__ void gen_set_bar(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(G * 32) : "memory");
}

#if GEN_UR
// one pass over this warp's share of the rows; `buf` alternates between passes (one barrier per pass: a
// buffer is rewritten two passes later, after the barrier of the pass in between)
// the scratch of the cross-warp reduction: ONE instance per kernel (static shared memory of a plain function,
// not of the ENERGY template)
__device__ __forceinline__ float (*gen_s_sum())[2][J * K][G][32] {
    __shared__ float s[GEN_SETS][2][J * K][G][32];
    return s;
}
__device__ __forceinline__ double (*gen_s_chi())[2][J][G][32] {
    __shared__ double s[GEN_SETS][2][J][G][32];
    return s;
}
// the user's function on one row, for the J chains of this lane at once.  A row is [x_0 .. x_{XD-1}, -y, pad] ([x, 0, -y, 0] with one abscissa)
// with a compile-time stride (GEN_STRIDE floats, a multiple of 4), fetched as float4 with a warp-uniform index.
template <bool ENERGY, bool SROWS>
__device__ __forceinline__ void gen_row(int n, const gen_real (&th)[K], gen_real (&gacc)[K], gen_real &c32) {
    float row[GEN_STRIDE];
#pragma unroll
    for (int v = 0; v < GEN_STRIDE / 4; ++v) {
        const float4 t = gen_row_load<SROWS>(n * (GEN_STRIDE / 4) + v);
        row[4 * v + 0] = t.x, row[4 * v + 1] = t.y, row[4 * v + 2] = t.z, row[4 * v + 3] = t.w;
    }
    gen_real xr[XD], dm[K];
#pragma unroll
    for (int j = 0; j < XD; ++j) xr[j] = gen_real(row[j]);
    const gen_real m = binfb_mock(th, xr, dm);
    const gen_real r = m + gen_real(row[GEN_YSLOT]);   // mock - y (the rows hold -y)
#pragma unroll
    for (int k = 0; k < K; ++k) gacc[k] = fmaf(r, dm[k], gacc[k]);   // J . (mock - y), likelihoods.py:155
    if (ENERGY) c32 = fmaf(r, r, c32);
}
// (ENERGY = false: chi^2 is not needed -- the L - 1 interior leapfrog steps -- and is returned as 0)
template <bool ENERGY, bool SROWS>
__device__ __forceinline__ void gen_pass(const GenDev &gm, const GenMap &mp, int buf, const float (&q)[J][K],
                                         float (&graw)[J][K], double (&chi2)[J]) {
    float (*s_sum)[2][J * K][G][32] = gen_s_sum();
    double (*s_chi)[2][J][G][32] = gen_s_chi();
    gen_real th[K], gacc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t[J];
#pragma unroll
        for (int j = 0; j < J; ++j) t[j] = q[j][k];
        th[k] = gen_join(t), gacc[k] = gen_real(0.f);
    }
    gen_real c32 = gen_real(0.f);
    double c64[J];
#pragma unroll
    for (int j = 0; j < J; ++j) c64[j] = 0.0;
    const int rpw = (gm.N + G - 1) / G;
    const int n0 = mp.g * rpw, n1 = min(gm.N, n0 + rpw);
    int n = n0;
    for (; n + 8 <= n1; n += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) gen_row<ENERGY, SROWS>(n + u, th, gacc, c32);
        if (ENERGY) {
#pragma unroll
            for (int j = 0; j < J; ++j) c64[j] += (double)gen_part(c32, j);
            c32 = gen_real(0.f);
        }
    }
    for (; n < n1; ++n) gen_row<ENERGY, SROWS>(n, th, gacc, c32);
    if (ENERGY) {
#pragma unroll
        for (int j = 0; j < J; ++j) c64[j] += (double)gen_part(c32, j);
    }
    if (G == 1) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
#pragma unroll
            for (int k = 0; k < K; ++k) graw[j][k] = gen_part(gacc[k], j);
            chi2[j] = c64[j];
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int k = 0; k < K; ++k) s_sum[mp.set][buf][j * K + k][mp.g][mp.lane] = gen_part(gacc[k], j);
        if (ENERGY) s_chi[mp.set][buf][j][mp.g][mp.lane] = c64[j];
    }
    gen_set_bar(mp.bar_id);
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < G; ++w) t += s_sum[mp.set][buf][j * K + k][w][mp.lane];
            graw[j][k] = t;
        }
        double t = 0.0;
        if (ENERGY) {
#pragma unroll
            for (int w = 0; w < G; ++w) t += s_chi[mp.set][buf][j][w][mp.lane];
        }
        chi2[j] = t;
    }
}
#else
template <bool ENERGY, bool SROWS>
__device__ __forceinline__ void gen_pass(const GenDev &gm, const GenMap &mp, int, const float (&q)[J][K],
                                         float (&graw)[J][K], double (&chi2)[J]) {
    const int g = mp.g;
    float gacc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) gacc[k] = 0.f;
    float c32 = 0.f;
    double c64 = 0.0;
    int cnt = 0;
    for (int n = g; n < gm.N; n += G) {
        const float *row = gm.rows + (size_t)n * gm.stride;
        float dm[K];
        const float m = binfb_mock(q[0], row, dm);
        const float r = m + __ldg(row + GEN_YSLOT);   // the rows hold -y
#pragma unroll
        for (int k = 0; k < K; ++k) gacc[k] = fmaf(r, dm[k], gacc[k]);   // J . (mock - y), likelihoods.py:155
        c32 = fmaf(r, r, c32);
        if (++cnt == 8) c64 += (double)c32, c32 = 0.f, cnt = 0;
    }
    c64 += (double)c32;
#pragma unroll
    for (int k = 0; k < K; ++k) graw[0][k] = group_allreduce_sum<G>(gacc[k]);
    chi2[0] = group_allreduce_sum<G>(c64);
}
#endif

// U(q) = -log p(q | tau) in float64 (binf/pdf/posteriors.py:141-151 summed components)
__device__ __forceinline__ double gen_potential(const GenDev &gm, const float (&q)[K], double chi2, float tau,
                                                float beta, double ga, double gb) {
    const double t = (double)tau;
    double prior = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double d = (double)q[k] - (double)gm.prior_mean[k];
        prior += d * d * (double)gm.prior_inv_var[k];
    }
    const double lt = log(t);
    return (double)beta * (0.5 * t * chi2 - 0.5 * (double)gm.N * lt) + 0.5 * prior - ((ga - 1.0) * lt - gb * t);
}

__device__ __forceinline__ void gen_force(const GenDev &gm, const float (&q)[K], const float (&graw)[K], float bt,
                                          float (&f)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        f[k] = bt * graw[k];
        if (gm.flags & BINFB_FLAG_PRIOR_GRAD) f[k] = fmaf(q[k] - gm.prior_mean[k], gm.prior_inv_var[k], f[k]);
    }
}

__device__ __forceinline__ float gen_draw_tau(const HmcArgs &a, double n_data, double chi2, float beta,
                                              uint64_t chain, int cid, uint64_t draw) {
    const double shape = 0.5 * (double)beta * n_data + a.gamma_shape - 1.0;   // samplers.py:32 (quirk Q3)
    const double rate = 0.5 * (double)beta * chi2 + a.gamma_rate;
    const double gdraw = a.gamma_draws ? a.gamma_draws[cid] : rng_gamma(a.seed, chain, draw, shape);
    return (float)(gdraw / rate);
}

extern "C" __global__ void __launch_bounds__(GEN_BLOCK) gen_hmc_kernel(GenDev gm, HmcArgs a) {
    constexpr bool S = SROWS_ONE_LAUNCH;
    gen_load_rows<S>(gm);
    GenMap mp;
    mp.init();
    const int g = mp.g;
    int pass = 0;
    bool valid[J];
    int cid[J];
    float q[J][K], p[J][K], graw[J][K], tau[J], eps[J], beta[J];
    int nacc[J];
    bool acc[J];
    double h0[J], h1[J], chi2[J], chi2_cur[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        valid[j] = mp.c + j < a.C;
        cid[j] = valid[j] ? (int)(mp.c + j) : a.C - 1;
#pragma unroll
        for (int k = 0; k < K; ++k) q[j][k] = a.q[(size_t)cid[j] * K + k];
        tau[j] = a.tau[cid[j]], eps[j] = a.eps[cid[j]];
        beta[j] = a.beta ? a.beta[cid[j]] : 1.0f;
        nacc[j] = 0, acc[j] = false, h0[j] = h1[j] = 0.0;
    }
    double st_acc = 0.0, st_prop = 0.0, st_eps = 0.0, st_pacc = 0.0;
    for (int tr = 0; tr < a.n_traj; ++tr) {
        const uint64_t draw = a.draw + (uint64_t)tr;
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k)
                p[j][k] = a.p0 ? a.p0[(size_t)cid[j] * K + k] : rng_normal(a.seed, a.chain_base + cid[j], draw, k);   // hmc.py:146
        gen_pass<true, S>(gm, mp, pass++ & 1, q, graw, chi2);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            if (a.gibbs_mode == BINFB_GIBBS_TAU_FIRST)
                tau[j] = gen_draw_tau(a, (double)gm.N, chi2[j], beta[j], a.chain_base + cid[j], cid[j], draw);
            chi2_cur[j] = chi2[j];
            double kin = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) kin += (double)p[j][k] * (double)p[j][k];
            h0[j] = gen_potential(gm, q[j], chi2[j], tau[j], beta[j], a.gamma_shape, a.gamma_rate) + 0.5 * kin;
            float f[K];
            gen_force(gm, q[j], graw[j], beta[j] * tau[j], f);
#pragma unroll
            for (int k = 0; k < K; ++k) p[j][k] = fmaf(-0.5f * eps[j], f[k], p[j][k]);             // hmc.py:116
        }
        for (int s = 1; s < a.L; ++s) {                                              // hmc.py:118-120
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int k = 0; k < K; ++k) q[j][k] = fmaf(eps[j], p[j][k], q[j][k]);
            gen_pass<false, S>(gm, mp, pass++ & 1, q, graw, chi2);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float f[K];
                gen_force(gm, q[j], graw[j], beta[j] * tau[j], f);
#pragma unroll
                for (int k = 0; k < K; ++k) p[j][k] = fmaf(-eps[j], f[k], p[j][k]);
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k) q[j][k] = fmaf(eps[j], p[j][k], q[j][k]);                 // hmc.py:122
        gen_pass<true, S>(gm, mp, pass++ & 1, q, graw, chi2);
        const bool last = tr == a.n_traj - 1;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            float f[K];
            gen_force(gm, q[j], graw[j], beta[j] * tau[j], f);
            double kin = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                p[j][k] = fmaf(-0.5f * eps[j], f[k], p[j][k]);                                 // hmc.py:123
                kin += (double)p[j][k] * (double)p[j][k];
            }
            h1[j] = gen_potential(gm, q[j], chi2[j], tau[j], beta[j], a.gamma_shape, a.gamma_rate) + 0.5 * kin;
            const float uu = a.u ? a.u[cid[j]] : rng_uniform(a.seed, a.chain_base + cid[j], draw, RNG_ACCEPT);
            const double dh = h1[j] - h0[j];
            acc[j] = (dh == dh) && ((double)uu < exp(fmin(709.0, fmax(-308.0, -dh))));    // hmc.py:151; NaN rejects
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
                st_acc += acc[j] ? 1.0 : 0.0, st_prop += 1.0, st_eps += (double)eps[j];
                st_pacc += (dh == dh) ? exp(fmin(0.0, -dh)) : 0.0;
            }
            if (tr < a.n_adapt) eps[j] *= acc[j] ? a.adapt_up : a.adapt_down;               // hmc.py:188-191
            if (a.gibbs_mode == BINFB_GIBBS_TAU_LAST)
                tau[j] = gen_draw_tau(a, (double)gm.N, chi2_cur[j], beta[j], a.chain_base + cid[j], cid[j], draw);
        }
        if (a.n_traj > 1) {
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (valid[j] && g == 0 && acc[j])
                    for (int k = 0; k < K; ++k) a.q[(size_t)cid[j] * K + k] = q[j][k];
            if (UR && G > 1) gen_set_bar(mp.bar_id);  // the chain's other owners sit in other warps
            else __syncwarp();
        }
    }
#pragma unroll
    for (int j = 0; j < J; ++j)
        if (valid[j] && g == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) a.q[(size_t)cid[j] * K + k] = q[j][k];
            a.tau[cid[j]] = tau[j], a.eps[cid[j]] = eps[j];
            if (a.accepted) a.accepted[cid[j]] = acc[j] ? 1 : 0;
            if (a.e_before) a.e_before[cid[j]] = h0[j];
            if (a.e_after) a.e_after[cid[j]] = h1[j];
            if (a.n_accepted) a.n_accepted[cid[j]] = nacc[j];
        }
    if (a.stats) {
        st_acc = group_allreduce_sum<32>(st_acc), st_prop = group_allreduce_sum<32>(st_prop);
        st_eps = group_allreduce_sum<32>(st_eps), st_pacc = group_allreduce_sum<32>(st_pacc);
        if ((threadIdx.x & 31) == 0 && st_prop > 0.0) {
            atomicAdd(a.stats + 0, st_acc), atomicAdd(a.stats + 1, st_prop);
            atomicAdd(a.stats + 2, st_eps), atomicAdd(a.stats + 3, st_pacc);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The split trajectory: the same transition as gen_hmc_kernel in three launches per trajectory,
//   gen_traj_begin   momenta, pass 0 (+ energy), (Gibbs: precision | q), H at the start, first half kick
//   gen_traj_mid     passes 1 .. L with drift and kicks -- nothing but the row loop and FMAs
//   gen_traj_end     H at the end, Metropolis test, step-size adaption, (Gibbs: precision | q_new), write-back
// with proposal, momentum and energies handed over through the model's workspace (gm.w_*).  The middle kernel is
// what the split is for: free of random draws, float64 energies and the accept branch, it is compiled with the
// rows arriving through the uniform datapath (LDCU) as uniform-register operands, which the fused kernel never
// gets.  Per chain the three kernels perform exactly the operations of the fused one (bit-identical results);
// the host launcher picks the split for launches with enough work to hide the two extra launches per trajectory.
// ---------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(GEN_BLOCK) gen_traj_begin(GenDev gm, HmcArgs a, int tr) {
    GenMap mp;
    mp.init();
    const bool leader = mp.g == 0;
    const uint64_t draw = a.draw + (uint64_t)tr;
    bool valid[J];
    int cid[J];
    float q[J][K], p[J][K], graw[J][K];
    double chi2[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        valid[j] = mp.c + j < a.C;
        cid[j] = valid[j] ? (int)(mp.c + j) : a.C - 1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            q[j][k] = a.q[(size_t)cid[j] * K + k];
            p[j][k] = a.p0 ? a.p0[(size_t)cid[j] * K + k] : rng_normal(a.seed, a.chain_base + cid[j], draw, k);   // hmc.py:146
        }
    }
    gen_pass<true, false>(gm, mp, 0, q, graw, chi2);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        float tau = a.tau[cid[j]];
        const float eps = a.eps[cid[j]], beta = a.beta ? a.beta[cid[j]] : 1.0f;
        if (a.gibbs_mode == BINFB_GIBBS_TAU_FIRST)
            tau = gen_draw_tau(a, (double)gm.N, chi2[j], beta, a.chain_base + cid[j], cid[j], draw);
        double kin = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) kin += (double)p[j][k] * (double)p[j][k];
        const double h0 = gen_potential(gm, q[j], chi2[j], tau, beta, a.gamma_shape, a.gamma_rate) + 0.5 * kin;
        float f[K];
        gen_force(gm, q[j], graw[j], beta * tau, f);
        if (valid[j] && leader) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                gm.w_p[(size_t)cid[j] * K + k] = fmaf(-0.5f * eps, f[k], p[j][k]);   // hmc.py:116
                gm.w_q[(size_t)cid[j] * K + k] = q[j][k];
            }
            gm.w_h0[cid[j]] = h0, gm.w_chi0[cid[j]] = chi2[j];
            if (a.gibbs_mode == BINFB_GIBBS_TAU_FIRST) a.tau[cid[j]] = tau;
        }
    }
}

extern "C" __global__ void __launch_bounds__(GEN_BLOCK) gen_traj_mid(GenDev gm, HmcArgs a) {
    GenMap mp;
    mp.init();
    const bool leader = mp.g == 0;
    bool valid[J];
    int cid[J];
    float q[J][K], p[J][K], graw[J][K], bt[J], eps[J];
    double chi2[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        valid[j] = mp.c + j < a.C;
        cid[j] = valid[j] ? (int)(mp.c + j) : a.C - 1;
#pragma unroll
        for (int k = 0; k < K; ++k) q[j][k] = gm.w_q[(size_t)cid[j] * K + k], p[j][k] = gm.w_p[(size_t)cid[j] * K + k];
        eps[j] = a.eps[cid[j]];
        bt[j] = (a.beta ? a.beta[cid[j]] : 1.0f) * a.tau[cid[j]];
    }
    const int L = a.L;
    for (int s = 1; s <= L; ++s) {
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k) q[j][k] = fmaf(eps[j], p[j][k], q[j][k]);   // hmc.py:119,122
        if (s == L) gen_pass<true, false>(gm, mp, s & 1, q, graw, chi2);
        else gen_pass<false, false>(gm, mp, s & 1, q, graw, chi2);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            float f[K];
            gen_force(gm, q[j], graw[j], bt[j], f);
            const float kick = s == L ? -0.5f * eps[j] : -eps[j];                   // hmc.py:120,123
#pragma unroll
            for (int k = 0; k < K; ++k) p[j][k] = fmaf(kick, f[k], p[j][k]);
        }
    }
#pragma unroll
    for (int j = 0; j < J; ++j)
        if (valid[j] && leader) {
#pragma unroll
            for (int k = 0; k < K; ++k) gm.w_q[(size_t)cid[j] * K + k] = q[j][k], gm.w_p[(size_t)cid[j] * K + k] = p[j][k];
            gm.w_chiL[cid[j]] = chi2[j];
        }
}

// one thread per chain
extern "C" __global__ void __launch_bounds__(GEN_BLOCK) gen_traj_end(GenDev gm, HmcArgs a, int tr) {
    const int c = blockIdx.x * GEN_BLOCK + threadIdx.x;
    double st[4] = {0.0, 0.0, 0.0, 0.0};
    if (c < a.C) {
        const uint64_t draw = a.draw + (uint64_t)tr;
        float q[K], p[K];
        double kin = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            q[k] = gm.w_q[(size_t)c * K + k], p[k] = gm.w_p[(size_t)c * K + k];
            kin += (double)p[k] * (double)p[k];
        }
        float tau = a.tau[c], eps = a.eps[c];
        const float beta = a.beta ? a.beta[c] : 1.0f;
        const double chiL = gm.w_chiL[c], h0 = gm.w_h0[c];
        const double h1 = gen_potential(gm, q, chiL, tau, beta, a.gamma_shape, a.gamma_rate) + 0.5 * kin;
        const float uu = a.u ? a.u[c] : rng_uniform(a.seed, a.chain_base + c, draw, RNG_ACCEPT);
        const double dh = h1 - h0;
        const bool acc = (dh == dh) && ((double)uu < exp(fmin(709.0, fmax(-308.0, -dh))));    // hmc.py:151; NaN rejects
        if (tr == a.n_traj - 1) {
            if (a.q_end)
                for (int k = 0; k < K; ++k) a.q_end[(size_t)c * K + k] = q[k];
            if (a.p_end)
                for (int k = 0; k < K; ++k) a.p_end[(size_t)c * K + k] = p[k];
        }
        if (acc) {
#pragma unroll
            for (int k = 0; k < K; ++k) a.q[(size_t)c * K + k] = q[k];
        }
        st[0] = acc ? 1.0 : 0.0, st[1] = 1.0, st[2] = (double)eps;
        st[3] = (dh == dh) ? exp(fmin(0.0, -dh)) : 0.0;
        if (tr < a.n_adapt) a.eps[c] = eps * (acc ? a.adapt_up : a.adapt_down);                // hmc.py:188-191
        if (a.gibbs_mode == BINFB_GIBBS_TAU_LAST)
            a.tau[c] = gen_draw_tau(a, (double)gm.N, acc ? chiL : gm.w_chi0[c], beta, a.chain_base + c, c, draw);
        if (a.accepted) a.accepted[c] = acc ? 1 : 0;
        if (a.e_before) a.e_before[c] = h0;
        if (a.e_after) a.e_after[c] = h1;
        if (a.n_accepted) a.n_accepted[c] += acc ? 1 : 0;
    }
    if (a.stats) {
#pragma unroll
        for (int i = 0; i < 4; ++i) st[i] = group_allreduce_sum<32>(st[i]);
        if ((threadIdx.x & 31) == 0 && st[1] > 0.0) {
            atomicAdd(a.stats + 0, st[0]), atomicAdd(a.stats + 1, st[1]);
            atomicAdd(a.stats + 2, st[2]), atomicAdd(a.stats + 3, st[3]);
        }
    }
}

extern "C" __global__ void __launch_bounds__(GEN_BLOCK) gen_grad_kernel(GenDev gm, GradArgs a) {
    constexpr bool S = SROWS_ONE_LAUNCH;
    gen_load_rows<S>(gm);
    GenMap mp;
    mp.init();
    const int g = mp.g;
    bool valid[J];
    int cid[J];
    float q[J][K], graw[J][K];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        valid[j] = mp.c + j < a.C;
        cid[j] = valid[j] ? (int)(mp.c + j) : a.C - 1;
#pragma unroll
        for (int k = 0; k < K; ++k) q[j][k] = a.q[(size_t)cid[j] * K + k];
    }
    double chi2[J];
    gen_pass<true, S>(gm, mp, 0, q, graw, chi2);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        if (!valid[j] || g != 0) continue;
        const float tau = a.tau[cid[j]];
        const float beta = a.beta ? a.beta[cid[j]] : 1.0f;
        if (a.logp) a.logp[cid[j]] = -gen_potential(gm, q[j], chi2[j], tau, beta, a.gamma_shape, a.gamma_rate);
        if (a.chi2) a.chi2[cid[j]] = chi2[j];
        if (a.grad) {
            float f[K];
            gen_force(gm, q[j], graw[j], beta * tau, f);
#pragma unroll
            for (int k = 0; k < K; ++k) a.grad[(size_t)cid[j] * K + k] = f[k];
        }
    }
}

// mock data of every chain (AbstractForwardModel.__call__): mock [C, N]
extern "C" __global__ void gen_forward_kernel(GenDev gm, const float *q, int C, float *mock) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)C * gm.N) return;
    const int c = (int)(i / gm.N), n = (int)(i - (long long)c * gm.N);
    gen_real th[K], dm[K], xr[XD];
#pragma unroll
    for (int k = 0; k < K; ++k) th[k] = gen_real(q[(size_t)c * K + k]);
#pragma unroll
    for (int j = 0; j < XD; ++j) xr[j] = gen_real(gm.rows[(size_t)n * gm.stride + j]);
    mock[i] = gen_part(binfb_mock(th, xr, dm), 0);
}
