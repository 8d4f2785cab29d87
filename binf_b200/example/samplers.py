"""Conjugate precision sampler and the example's Gibbs wiring
(reference: binf/example/samplers.py:7-51,94-111).  The reference wires a random-walk Metropolis
sampler for the coefficients; here `make_sampler` wires the batched device HMCSampler instead (the
RWMC sampler is outside the hot path)."""
import numpy as np


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


def make_sampler(posterior, timestep, start_state, nsteps=20, timestep_adaption_limit=0, seed=None):
    """GibbsSampler(HMC on the coefficients, conjugate Gamma on the precision)."""
    from binf_b200.samplers.gibbs import GibbsSampler
    from binf_b200.samplers.hmc import HMCSampler
    coeffs = start_state.variables["coefficients"]
    precision = start_state.variables["precision"]
    hmc = HMCSampler(posterior.conditional_factory(precision=precision), coeffs, timestep, nsteps,
                     timestep_adaption_limit=timestep_adaption_limit, variable_name="coefficients",
                     seed=seed)
    gam = GammaSampler(posterior.conditional_factory(coefficients=coeffs), precision,
                       seed=None if seed is None else seed + 1)
    return GibbsSampler(posterior, start_state, {"coefficients": hmc, "precision": gam})
