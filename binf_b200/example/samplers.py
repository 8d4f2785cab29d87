"""Conjugate precision sampler, random-walk Metropolis sampler and the example's Gibbs wiring
(reference: binf/example/samplers.py:7-111).  `make_sampler(posterior, rwmc_stepsize, start_state)` wires
the reference's RWMCSampler + GammaSampler exactly as the reference does (both batched on the device);
`make_sampler(posterior, timestep, start_state, nsteps=L)` wires the fused device HMCSampler instead."""
from collections import namedtuple

import numpy as np

RWMCSampleStats = namedtuple("RWMCSampleStats", "acceptance_rate")


class RWMCSampler(object):
    """Random-walk Metropolis on the coefficients (samplers.py:54-92): proposal = state +
    U(-stepsize, stepsize), accept iff u < exp(-(E_new - E_old)).  `state` (D,) is one chain,
    (C, D) are C chains; a CUDA tensor keeps the chains in HBM."""

    def __init__(self, pdf, state, stepsize, seed=None, chain_base=0):
        self.pdf = pdf
        self.state = state
        self.stepsize = stepsize
        self._n_moves = 0
        self._n_accepted_moves = 0
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.chain_base = int(chain_base)
        self._draw = 0
        self._lowered = None
        self._dev = None

    @property
    def last_draw_stats(self):
        return {"coefficients": RWMCSampleStats(self.acceptance_rate)}

    @property
    def acceptance_rate(self):
        if self._n_moves > 0:
            n = self._n_accepted_moves
            mean = float(n.double().mean()) if hasattr(n, "data_ptr") else float(np.mean(n))
            return mean / float(self._n_moves)
        return 0.0

    def _lower(self):
        from binf_b200.lowering import lower
        if self._lowered is None or self._lowered[0] is not self.pdf:
            low = lower(self.pdf, n_coeff=int(self.state.shape[-1]))
            if low is None:
                raise NotImplementedError("RWMCSampler: the pdf is not lowered to the device (no CPU fallback)")
            self._lowered = (self.pdf, low)
        self._lowered[1].refresh()
        return self._lowered[1]

    def sample(self, change=None, u=None, n_moves=1):
        """`change` / `u` inject the proposal displacement and the uniform (parity tests)."""
        from binf_b200.lowering import _is_tensor
        low = self._lower()
        if _is_tensor(self.state):
            import torch
            q = self.state
            n, dev = q.shape[0], q.device
            if self._dev is None:
                self._dev = (torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int32, device=dev))
            tau = low.tau(n)
            tau = tau if _is_tensor(tau) else torch.as_tensor(tau, dtype=torch.float32, device=dev)
            beta = low.beta(n)
            beta = None if beta is None else torch.as_tensor(beta, dtype=torch.float32, device=dev)
            step = torch.as_tensor(np.broadcast_to(np.asarray(self.stepsize, dtype=np.float32), (n,)).copy(), device=dev)
            low.model.rwmc_run_device(q, tau, step, n_moves, beta=beta, seed=self.seed, draw=self._draw,
                                      chain_base=self.chain_base, accepted=self._dev[0], n_accepted=self._dev[1],
                                      stream=torch.cuda.current_stream().cuda_stream)
            self._n_accepted_moves = self._n_accepted_moves + self._dev[1].to(torch.int64)
        else:
            single = np.ndim(self.state) == 1
            q = np.asarray(self.state, dtype=np.float64).reshape(-1, low.dim)
            r = low.model.rwmc_run(q, low.tau(len(q)), self.stepsize, n_moves, beta=low.beta(len(q)),
                                   change=change, u=u, seed=self.seed, draw=self._draw, chain_base=self.chain_base)
            new = r["q"].astype(np.float64)
            self.state = new[0] if single else new
            self._n_accepted_moves = self._n_accepted_moves + (
                int(r["n_accepted"][0]) if single else r["n_accepted"].astype(np.int64))
        self._n_moves += n_moves
        self._draw += n_moves
        return self.state


class GammaSampler(object):
    """tau ~ Gamma(N/2 + a - 1, 1) / (chi^2/2 + b) per chain (samplers.py:27-47; the "- 1" is the
    reference's, quirk Q3).  chi^2 comes from one fused forward pass on the device."""

    def __init__(self, pdf, state, seed=None, chain_base=0):
        self.pdf = pdf
        self.state = state
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.chain_base = int(chain_base)
        self._draw = 0

    def _get_prior(self):
        from binf_b200.example.priors import GammaPrior
        prior = [p for p in self.pdf.priors.values() if "precision" in p.variables][0]
        if not isinstance(prior, GammaPrior):
            raise NotImplementedError("Prior for precision is not a Gamma distribution")
        return prior

    def _likelihood(self):
        liks = list(self.pdf.likelihoods.values())
        if len(liks) != 1:
            raise NotImplementedError("GammaSampler needs exactly one likelihood")
        return liks[0]

    def _sampled_variable(self):
        lik = self._likelihood()
        names = [p for p in lik.parameters if p != "precision" and p in lik._original_variables]
        return names[0]

    def _calculate_shape(self):
        return 0.5 * len(self._likelihood().error_model.ys) + self._get_prior().shape - 1

    def sample(self, state=42, gamma_draws=None):
        from binf_b200.lowering import lower
        lik = self._likelihood()
        var = self._sampled_variable()
        q = np.asarray(self.pdf[var].value, dtype=np.float64)
        low = lower(lik, n_coeff=int(q.shape[-1]))
        if low is None:
            raise NotImplementedError("GammaSampler: likelihood is not lowered to the device")
        prior = self._get_prior()
        low.model.set_gamma_prior(prior.shape, prior.rate)
        single = q.ndim == 1
        q2 = q.reshape(-1, low.dim)
        tau, _ = low.model.gibbs_precision(q2, np.ones(len(q2)), beta=low.beta(len(q2)),
                                           gamma_draws=gamma_draws, seed=self.seed, draw=self._draw,
                                           chain_base=self.chain_base)
        self._draw += 1
        self.state = float(tau[0]) if single else tau.astype(np.float64)
        return self.state


def make_sampler(posterior, rwmc_stepsize, start_state, nsteps=None, timestep_adaption_limit=0, seed=None):
    """GibbsSampler over (coefficients, precision) exactly as the reference wires it
    (samplers.py:94-111): random-walk Metropolis with `rwmc_stepsize` on the coefficients, the conjugate
    Gamma sampler on the precision.  With `nsteps` given, the coefficients are sampled by HMC instead
    (leapfrog time step = the second argument, `nsteps` steps per trajectory) and a sweep is ONE launch
    of the fused kernel."""
    from binf_b200.samplers.gibbs import GibbsSampler
    from binf_b200.samplers.hmc import HMCSampler
    coeffs = start_state.variables["coefficients"]
    precision = start_state.variables["precision"]
    if nsteps is None:
        sub = RWMCSampler(posterior.conditional_factory(precision=precision), coeffs, rwmc_stepsize, seed=seed)
    else:
        sub = HMCSampler(posterior.conditional_factory(precision=precision), coeffs, rwmc_stepsize, nsteps,
                         timestep_adaption_limit=timestep_adaption_limit, variable_name="coefficients",
                         seed=seed)
    gam = GammaSampler(posterior.conditional_factory(coefficients=coeffs), precision,
                       seed=None if seed is None else seed + 1)
    return GibbsSampler(posterior, start_state, {"coefficients": sub, "precision": gam})
