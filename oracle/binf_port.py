"""CPU restatement (numpy, float64) of the reference's HMC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `binf_b200/` imports this file; it is the
checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
--impl reference legs).  It restates, function by function, what simeoncarstens/binf
computes; every function cites the reference file:line it follows.  Citations are
relative to /root/reference.

Parity status: the reference's own tests pin only plumbing.  This port
is therefore pinned against *outputs of the reference itself run in the build
container* (oracle/make_golden.py imports the unmodified reference through
oracle/ref_import.py and writes tests/golden/*.npz); tests/test_oracle_golden.py
re-checks the port against those vectors everywhere, and against the live reference
whenever /root/reference is present.

All functions take an optional leading chain axis: coefficients of shape (K,) or
(C, K); tau scalar or (C,).
"""
import numpy as np

polyval = np.polynomial.polynomial.polyval


# ----------------------------------------------------------------------------------
# polynomial forward model + Gaussian error model          binf/example/likelihood.py
# ----------------------------------------------------------------------------------
def poly_forward(xs, c):
    """mock_n = sum_k c_k x_n^k                         (example/likelihood.py:24-26)"""
    c = np.asarray(c, dtype=np.float64)
    if c.ndim == 1:
        return polyval(xs, c)
    return polyval(xs, c.T)  # (C, N)


def poly_jacobian(xs, n_coeff):
    """J[k, n] = x_n^k, rebuilt on every call like the reference
    (example/likelihood.py:28-30)."""
    return np.vstack([xs ** i for i in range(n_coeff)])


def gauss_em_log_prob(mock, ys, tau):
    """-1/2 tau sum (mock-y)^2 + 1/2 N log tau  -- no 2*pi term
    (example/likelihood.py:54-57)."""
    tau = np.asarray(tau, dtype=np.float64)
    log_z = len(ys) * 0.5 * np.log(tau)
    return -0.5 * np.sum((mock - ys) ** 2, axis=-1) * tau + log_z


def gauss_em_gradient(mock, ys, tau):
    """dE/dmock = tau (mock - y)                        (example/likelihood.py:59-61)"""
    tau = np.asarray(tau, dtype=np.float64)
    return (mock - ys) * tau[..., None] if tau.ndim else (mock - ys) * tau


# ----------------------------------------------------------------------------------
# priors                                                    binf/example/priors.py
# ----------------------------------------------------------------------------------
def gamma_prior_log_prob(tau, shape, rate):
    """(a-1) log tau - b tau                               (example/priors.py:23-25)"""
    tau = np.asarray(tau, dtype=np.float64)
    return (shape - 1.0) * np.log(tau) - tau * rate


def gaussian_prior_log_prob(c, means, variances):
    """-1/2 sum (c-mu)^2 / v                               (example/priors.py:49-54)"""
    c = np.asarray(c, dtype=np.float64)
    return -0.5 * np.sum((c - means) ** 2 / variances, axis=-1)


class PolynomialPosterior(object):
    """Posterior over `coefficients` at fixed `precision`, as the reference's
    conditional pdf evaluates it.

    log_prob = log L + log pi(c) + log pi(tau)          (pdf/posteriors.py:125-151)
    gradient = likelihood term only -- the Gaussian prior registers `coefficients`
      as non-differentiable and is skipped (quirk Q1; pdf/posteriors.py:182-185,
      example/priors.py:45).  `prior_grad=True` adds (c-mu)/v (the corrected force).
    gamma_rate: the reference's conditional pdfs carry rate == shape because
      GammaPrior.clone passes shape twice (quirk Q2; example/priors.py:29); callers
      pass whichever rate the pdf they mirror actually holds.
    """

    def __init__(self, xs, ys, prior_means, prior_variances, gamma_shape, gamma_rate,
                 prior_grad=False):
        self.xs = np.asarray(xs, dtype=np.float64)
        self.ys = np.asarray(ys, dtype=np.float64)
        self.means = np.asarray(prior_means, dtype=np.float64)
        self.variances = np.asarray(prior_variances, dtype=np.float64)
        self.gamma_shape = float(gamma_shape)
        self.gamma_rate = float(gamma_rate)
        self.prior_grad = prior_grad

    def chi2(self, c):
        return np.sum((poly_forward(self.xs, c) - self.ys) ** 2, axis=-1)

    def likelihood_log_prob(self, c, tau):
        """Likelihood._evaluate_log_prob              (pdf/likelihoods.py:141-146)"""
        return gauss_em_log_prob(poly_forward(self.xs, c), self.ys, tau)

    def log_prob(self, c, tau):
        return (self.likelihood_log_prob(c, tau)
                + gaussian_prior_log_prob(c, self.means, self.variances)
                + gamma_prior_log_prob(tau, self.gamma_shape, self.gamma_rate))

    def gradient(self, c, tau):
        """Likelihood._evaluate_gradient: J(theta) . dE/dmock, J rebuilt per call
        (pdf/likelihoods.py:148-155).  Returns the gradient of the ENERGY -log p."""
        c = np.asarray(c, dtype=np.float64)
        mock = poly_forward(self.xs, c)
        jac = poly_jacobian(self.xs, c.shape[-1])
        g = gauss_em_gradient(mock, self.ys, tau)
        res = g.dot(jac.T) if c.ndim > 1 else jac.dot(g)
        if self.prior_grad:
            res = res + (c - self.means) / self.variances
        return res


# ----------------------------------------------------------------------------------
# HMC                                                        binf/samplers/hmc.py
# ----------------------------------------------------------------------------------
def leapfrog(gradient, q, p, timestep, nsteps):
    """Half kick, (L-1) x [drift, kick], drift, half kick; L+1 gradient calls
    (samplers/hmc.py:116-123).  `timestep` scalar or per-chain (C,)."""
    q = np.array(q, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    dt = np.asarray(timestep, dtype=np.float64)
    if dt.ndim:
        dt = dt[:, None]
    p -= 0.5 * dt * gradient(q)
    for _ in range(nsteps - 1):
        q += p * dt
        p -= dt * gradient(q)
    q += p * dt
    p -= 0.5 * dt * gradient(q)
    return q, p


def hmc_sample(log_prob, gradient, q0, timestep, nsteps, p0, u):
    """One HMC transition with injected momenta p0 and uniforms u
    (samplers/hmc.py:136-164).  Returns a dict with the new state, the accept flag,
    the energies, and the end point of the trajectory.  NaN energies reject, as
    `u < exp(nan)` is False in the reference."""
    q0 = np.asarray(q0, dtype=np.float64)
    p0 = np.asarray(p0, dtype=np.float64)
    e_before = -log_prob(q0) + 0.5 * np.sum(p0 ** 2, axis=-1)
    q_l, p_l = leapfrog(gradient, q0, p0, timestep, nsteps)
    e_after = -log_prob(q_l) + 0.5 * np.sum(p_l ** 2, axis=-1)
    with np.errstate(over="ignore", invalid="ignore"):
        # csb.numeric.exp clips its argument to [-308, 709] (samplers/hmc.py:10,151)
        acc = np.asarray(u) < np.exp(np.clip(-(e_after - e_before), -308.0, 709.0))
    acc = np.asarray(acc)
    if q0.ndim == 1:
        q_new = q_l if bool(acc) else q0
    else:
        q_new = np.where(acc[:, None], q_l, q0)
    return dict(q=q_new, accepted=acc, e_before=e_before, e_after=e_after,
                q_end=q_l, p_end=p_l)


def adapt_timestep(timestep, accepted, counter_after, limit, uprate=1.05, downrate=0.95):
    """`counter += 1; if counter < limit: dt *= uprate if accepted else downrate`
    (samplers/hmc.py:153-157,183-191; quirk Q4: strict `<` after the increment, and
    accepted => uprate although the docstring says the opposite)."""
    timestep = np.asarray(timestep, dtype=np.float64)
    if counter_after < limit:
        return np.where(accepted, timestep * uprate, timestep * downrate)
    return timestep


# ----------------------------------------------------------------------------------
# conjugate precision update                               binf/example/samplers.py
# ----------------------------------------------------------------------------------
def gamma_precision_params(chi2, n_data, prior_shape, prior_rate, beta=1.0):
    """shape = 1/2 N + a - 1 (quirk Q3: the conjugate result is 1/2 N + a),
    rate = 1/2 chi^2 + b                              (example/samplers.py:27-41).
    `beta` tempers the likelihood term (build-defined)."""
    shape = 0.5 * beta * n_data + prior_shape - 1.0
    rate = 0.5 * beta * np.asarray(chi2, dtype=np.float64) + prior_rate
    return shape, rate


def gamma_precision_sample(chi2, n_data, prior_shape, prior_rate, rng, beta=1.0):
    """tau = Gamma(shape, 1) / rate                        (example/samplers.py:43-47)"""
    shape, rate = gamma_precision_params(chi2, n_data, prior_shape, prior_rate, beta)
    return rng.gamma(shape, size=np.shape(rate)) / rate


# ----------------------------------------------------------------------------------
# Gibbs sweep                                               binf/samplers/gibbs.py
# ----------------------------------------------------------------------------------
def gibbs_sweep(state, samplers):
    """Sweep the variables in sorted-name order, each sub-sampler seeing the state as
    updated so far (samplers/gibbs.py:146-149; quirk Q5).  `samplers[name]` is a
    callable state_dict -> new value."""
    for var in sorted(state):
        state[var] = samplers[var](state)
    return state


def rwmc_sample(log_prob, state, change, u):
    """RWMCSampler.sample with the randomness injected (binf/example/samplers.py:78-92):
    `change` replaces np.random.uniform(-stepsize, stepsize, len(state)), `u` np.random.random()."""
    state = np.asarray(state, dtype=np.float64)
    e_old = -log_prob(state)                                    # samplers.py:80
    proposal = state + np.asarray(change, dtype=np.float64)     # samplers.py:81-83
    e_new = -log_prob(proposal)                                 # samplers.py:84
    with np.errstate(over="ignore", invalid="ignore"):
        accepted = bool(u < np.exp(-(e_new - e_old)))           # samplers.py:86
    return dict(state=proposal if accepted else state, accepted=accepted, e_old=e_old, e_new=e_new)


def predict(x, y, coefficients, precisions):
    """predict (binf/example/misc.py:3-16): exp(log_sum_exp(integrands)) / len(samples) with
    integrand = -0.5 (polyval(x, c) - y)^2 tau + 0.5 log tau - 0.5 log 2 pi"""
    c = np.asarray(coefficients, dtype=np.float64)
    t = np.asarray(precisions, dtype=np.float64)
    mock = np.polynomial.polynomial.polyval(x, c.T)
    f = -0.5 * (mock - y) ** 2 * t + 0.5 * np.log(t) - 0.5 * np.log(2.0 * np.pi)
    m = np.max(f)
    return np.exp(m + np.log(np.sum(np.exp(f - m)))) / len(t)


class UserModelPosterior(object):
    """Posterior over the parameters of a user-defined per-datum forward model, evaluated the way the
    reference composes it: mock = f(theta) (AbstractForwardModel._evaluate), gradient =
    J(theta) . tau (mock - y) (binf/pdf/likelihoods.py:148-155 with binf/example/likelihood.py:61),
    Gaussian prior in log_prob only (quirk Q1), Gamma prior on the precision.
    f(theta [..., K], x [N]) -> [..., N];  jac(theta, x) -> [..., K, N]."""

    def __init__(self, xs, ys, f, jac, prior_means, prior_variances, gamma_shape, gamma_rate, prior_grad=False):
        self.xs, self.ys, self.f, self.jac = np.asarray(xs, float), np.asarray(ys, float), f, jac
        self.mu, self.var = np.asarray(prior_means, float), np.asarray(prior_variances, float)
        self.a, self.b, self.prior_grad = gamma_shape, gamma_rate, prior_grad

    def chi2(self, q):
        r = self.f(np.asarray(q, float), self.xs) - self.ys
        return np.sum(r * r, axis=-1)

    def log_prob(self, q, tau):
        q = np.asarray(q, float)
        lik = -0.5 * tau * self.chi2(q) + 0.5 * len(self.ys) * np.log(tau)     # example/likelihood.py:56-57
        prior = -0.5 * np.sum((q - self.mu) ** 2 / self.var, axis=-1)           # example/priors.py:54
        return lik + prior + (self.a - 1.0) * np.log(tau) - self.b * tau         # example/priors.py:25

    def gradient(self, q, tau):
        q = np.asarray(q, float)
        r = self.f(q, self.xs) - self.ys
        g = np.einsum("...kn,...n->...k", self.jac(q, self.xs), tau * r)        # likelihoods.py:155
        if self.prior_grad:
            g = g + (q - self.mu) / self.var
        return g
