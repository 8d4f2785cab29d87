"""CPU restatement (numpy, float64) of the chromatin bead-chain posterior.

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA pair kernel.

The reference ships no chromatin model: this model is build-defined
 and expressed *behind the reference's own API*:
`reference_classes(binf)` below returns an `AbstractForwardModel` subclass and an
`AbstractPrior` subclass built on the live reference's base classes, so that the
reference's unmodified `Likelihood._evaluate_gradient` (dense J.dot(g),
pdf/likelihoods.py:148-155), `Posterior` (pdf/posteriors.py:125-187) and `HMCSampler`
(samplers/hmc.py:92-164) drive it.  oracle/make_golden.py uses that route at small n to
pin the matrix-free formulas in this file, which are what is compared with the GPU at
n = 1000.

Model (float64 here, float32 on the device):
  structure q in R^{3n} = X.flatten(), X of shape (n, 3)
  pairs (i<j) in np.triu_indices(n, 1) order, M = n(n-1)/2
  d_ij   = sqrt(|x_i - x_j|^2 + SOFT)                       SOFT = 1e-12 guards d -> 0
  mock_ij = 1 / (1 + exp(alpha (d_ij - d_c)))               logistic contact function (default), or
          = 1/2 (1 + z / sqrt(1 + z^2)), z = alpha (d_c - d_ij)   the algebraic one (contact="algebraic",
                                                                  SURVEY.md A.2: two rsqrt, no exponential)
  error model: the reference's GaussianErrorModel (example/likelihood.py:40-68)
  prior on structure: backbone  -1/2 k_bb sum_i (|x_{i+1}-x_i|_soft - l0)^2
                      optional confinement  -1/2 |X|^2 / s^2   (conf_s = 0 disables)
                      optional excluded volume  -k_ev sum_{i<j} max(0, d_ev - d_ij)^4  (ev_k = 0 disables;
                      the quartic repulsion of the ISD chromatin models the README's paper describes)
  tempering: log p_beta = beta * log L + log prior            
"""
import numpy as np

SOFT = 1e-12


def contact_function(d, alpha, d_c, contact="logistic"):
    """(mock, d mock / d d) of the contact model at distances d"""
    if contact == "algebraic":
        z = alpha * (d_c - d)
        s = 1.0 / np.sqrt(1.0 + z * z)
        return 0.5 * (1.0 + z * s), -alpha * 0.5 * s ** 3
    assert contact == "logistic"
    with np.errstate(over="ignore"):
        m = 1.0 / (1.0 + np.exp(alpha * (d - d_c)))
    return m, -alpha * m * (1.0 - m)


class ChromatinModel(object):
    def __init__(self, n_beads, y_pairs, alpha, d_c, k_bb, l0, conf_s=0.0,
                 gamma_shape=1.0, gamma_rate=1.0, ev_k=0.0, ev_d=0.0, contact="logistic"):
        self.contact = contact
        self.n = int(n_beads)
        self.iu = np.triu_indices(self.n, 1)
        self.y = np.asarray(y_pairs, dtype=np.float64)
        assert self.y.shape == (self.n * (self.n - 1) // 2,)
        self.alpha, self.d_c = float(alpha), float(d_c)
        self.k_bb, self.l0, self.conf_s = float(k_bb), float(l0), float(conf_s)
        self.gamma_shape, self.gamma_rate = float(gamma_shape), float(gamma_rate)
        self.ev_k, self.ev_d = float(ev_k), float(ev_d)
        self.Y = np.zeros((self.n, self.n))
        self.Y[self.iu] = self.y
        self.Y += self.Y.T

    @property
    def n_pairs(self):
        return self.y.shape[0]

    # ---- forward model ----------------------------------------------------------
    def _dist_full(self, X):
        diff = X[:, None, :] - X[None, :, :]
        return diff, np.sqrt(np.sum(diff * diff, axis=-1) + SOFT)

    def forward(self, q):
        """mock data for all pairs, triu order."""
        X = np.asarray(q, dtype=np.float64).reshape(self.n, 3)
        d = np.sqrt(np.sum((X[self.iu[0]] - X[self.iu[1]]) ** 2, axis=-1) + SOFT)
        return contact_function(d, self.alpha, self.d_c, self.contact)[0]

    def jacobian_dense(self, q):
        """d mock_ij / d q as a (3n, M) matrix -- what the reference's
        `jacobi_matrix` contract asks for (model/forwardmodels.py:23-28).  Small n only."""
        X = np.asarray(q, dtype=np.float64).reshape(self.n, 3)
        i, j = self.iu
        diff = X[i] - X[j]
        d = np.sqrt(np.sum(diff * diff, axis=-1) + SOFT)
        m, dmdd = contact_function(d, self.alpha, self.d_c, self.contact)
        dm = (dmdd / d)[:, None] * diff                                  # d mock / d x_i
        J = np.zeros((self.n, 3, self.n_pairs))
        cols = np.arange(self.n_pairs)
        for a in range(3):
            J[i, a, cols] = dm[:, a]
            J[j, a, cols] = -dm[:, a]
        return J.reshape(3 * self.n, self.n_pairs)

    def chi2(self, q):
        return np.sum((self.forward(q) - self.y) ** 2)

    # ---- likelihood -------------------------------------------------------------
    def likelihood_log_prob(self, q, tau, beta=1.0):
        """GaussianErrorModel._evaluate_log_prob on the mock contacts
        (example/likelihood.py:54-57), times beta."""
        return beta * (-0.5 * tau * self.chi2(q) + 0.5 * self.n_pairs * np.log(tau))

    def likelihood_gradient(self, q, tau, beta=1.0):
        """Matrix-free J . (tau (mock - y)) (pdf/likelihoods.py:152-155 without the
        12 GB Jacobian): dE/dx_i = sum_j tau (m-y) (-alpha) m (1-m) (x_i-x_j)/d."""
        X = np.asarray(q, dtype=np.float64).reshape(self.n, 3)
        diff, d = self._dist_full(X)
        m, dmdd = contact_function(d, self.alpha, self.d_c, self.contact)
        w = (m - self.Y) * dmdd / d
        np.fill_diagonal(w, 0.0)
        g = np.einsum("ij,ija->ia", w, diff)
        return (beta * tau) * g.reshape(-1)

    # ---- prior on the structure ---------------------------------------------------
    def prior_log_prob(self, q):
        X = np.asarray(q, dtype=np.float64).reshape(self.n, 3)
        b = X[1:] - X[:-1]
        d = np.sqrt(np.sum(b * b, axis=-1) + SOFT)
        lp = -0.5 * self.k_bb * np.sum((d - self.l0) ** 2)
        if self.conf_s > 0.0:
            lp += -0.5 * np.sum(X * X) / self.conf_s ** 2
        if self.ev_k > 0.0:
            dp = np.sqrt(np.sum((X[self.iu[0]] - X[self.iu[1]]) ** 2, axis=-1) + SOFT)
            lp += -self.ev_k * np.sum(np.maximum(self.ev_d - dp, 0.0) ** 4)
        return lp

    def prior_gradient(self, q):
        """gradient of the prior ENERGY (-log prior)."""
        X = np.asarray(q, dtype=np.float64).reshape(self.n, 3)
        b = X[1:] - X[:-1]
        d = np.sqrt(np.sum(b * b, axis=-1) + SOFT)
        c = (self.k_bb * (d - self.l0) / d)[:, None] * b
        g = np.zeros_like(X)
        g[1:] += c
        g[:-1] -= c
        if self.conf_s > 0.0:
            g += X / self.conf_s ** 2
        if self.ev_k > 0.0:
            diff, dd = self._dist_full(X)
            w = -4.0 * self.ev_k * np.maximum(self.ev_d - dd, 0.0) ** 3 / dd
            np.fill_diagonal(w, 0.0)
            g += np.einsum("ij,ija->ia", w, diff)
        return g.reshape(-1)

    # ---- posterior over the structure at fixed precision --------------------------
    def gamma_prior_log_prob(self, tau):
        return (self.gamma_shape - 1.0) * np.log(tau) - tau * self.gamma_rate

    def log_prob(self, q, tau, beta=1.0):
        """sum of component log-probs (pdf/posteriors.py:141-151)."""
        return (self.likelihood_log_prob(q, tau, beta) + self.prior_log_prob(q)
                + self.gamma_prior_log_prob(tau))

    def gradient(self, q, tau, beta=1.0):
        """sum over components with a differentiable variable
        (pdf/posteriors.py:173-187): likelihood + structure prior."""
        return self.likelihood_gradient(q, tau, beta) + self.prior_gradient(q)


# --------------------------------------------------------------------------------------
# synthetic data generator shared by the tests and bench.py (SURVEY.md 8d, C3 input)
# --------------------------------------------------------------------------------------
def synthetic_chromatin(n_beads, alpha=2.0, d_c=2.5, l0=1.0, noise=0.05, seed=0, contact="logistic"):
    """Ground truth = 3-D random walk with N(0,1)*l0 steps; y = sigma(alpha(d_c-d*)) + N(0, noise^2)."""
    rng = np.random.RandomState(seed)
    X = np.cumsum(rng.normal(size=(n_beads, 3)) * l0, axis=0)
    X -= X.mean(axis=0)
    i, j = np.triu_indices(n_beads, 1)
    d = np.sqrt(np.sum((X[i] - X[j]) ** 2, axis=-1) + SOFT)
    y = contact_function(d, alpha, d_c, contact)[0] + rng.normal(size=d.shape) * noise
    return X, y


# --------------------------------------------------------------------------------------
# the same model as subclasses of the LIVE reference's base classes
# --------------------------------------------------------------------------------------
def reference_classes(binf):
    """Build chromatin forward-model / prior classes on the reference's own bases
    (`binf` = the reference package imported by oracle/ref_import.py)."""
    import importlib
    fwm_mod = importlib.import_module("binf.model.forwardmodels")
    priors_mod = importlib.import_module("binf.pdf.priors")
    ArrayParameter = binf.ArrayParameter

    class ContactForwardModel(fwm_mod.AbstractForwardModel):
        def __init__(self, model):
            super(ContactForwardModel, self).__init__("contacts")
            self.model = model
            self._register_variable("structure", differentiable=True)
            self.update_var_param_types(structure=ArrayParameter)
            self._set_original_variables()

        def _evaluate(self, structure):
            return self.model.forward(structure)

        def _evaluate_jacobi_matrix(self, structure):
            return self.model.jacobian_dense(structure)

        def clone(self):
            copy = self.__class__(self.model)
            self._set_parameters(copy)
            return copy

    class BackbonePrior(priors_mod.AbstractPrior):
        def __init__(self, model):
            super(BackbonePrior, self).__init__("structure_prior")
            self.model = model
            self._register_variable("structure", differentiable=True)
            self.update_var_param_types(structure=ArrayParameter)
            self._set_original_variables()

        def _evaluate_log_prob(self, structure):
            return self.model.prior_log_prob(structure)

        def _evaluate_gradient(self, structure):
            return self.model.prior_gradient(structure)

        def clone(self):
            copy = self.__class__(self.model)
            copy.set_fixed_variables_from_pdf(self)
            return copy

    return ContactForwardModel, BackbonePrior


def reference_posterior(binf, model):
    """Posterior({contacts likelihood}, {structure prior, precision prior}) assembled from
    the live reference's Likelihood / Posterior / GaussianErrorModel / GammaPrior."""
    import importlib
    Likelihood = importlib.import_module("binf.pdf.likelihoods").Likelihood
    Posterior = importlib.import_module("binf.pdf.posteriors").Posterior
    GaussianErrorModel = importlib.import_module("binf.example.likelihood").GaussianErrorModel
    GammaPrior = importlib.import_module("binf.example.priors").GammaPrior
    Fwm, Backbone = reference_classes(binf)
    lik = Likelihood("points", Fwm(model), GaussianErrorModel(model.y))
    priors = {"structure_prior": Backbone(model),
              "precision_prior": GammaPrior(model.gamma_shape, model.gamma_rate)}
    return Posterior({lik.name: lik}, priors)


def acceptance_inputs(n_beads, n_chains, seed):
    """the seeded inputs of `chromatin_acceptance_case`, regenerated (not stored) by the GPU test"""
    alpha, d_c, k_bb, l0 = 2.0, 2.5, 4.0, 1.0
    X, y = synthetic_chromatin(n_beads, alpha, d_c, l0, 0.05, seed)
    rng = np.random.RandomState(seed + 1)
    q0 = X.reshape(-1)[None, :] + 0.05 * rng.normal(size=(n_chains, 3 * n_beads))
    p0 = rng.normal(size=q0.shape)
    u = rng.uniform(size=n_chains)
    return (alpha, d_c, k_bb, l0), y, q0, p0, u
