"""Priors of the polynomial example: Gamma on the precision, Gaussian on the coefficients
(binf/example/priors.py:10-73).  Closed forms of O(K) scalars; in a lowered Posterior they are model
constants of the CUDA kernels (binfb_model_create_polynomial / binfb_model_set_gamma_prior)."""
import numpy as np

from binf_b200 import ArrayParameter
from binf_b200.params import Parameter as ScalarParameter
from binf_b200.pdf.priors import AbstractPrior

# Quirk Q2: the reference's GammaPrior.clone hands `shape` to both constructor arguments
# (priors.py:29), so every conditional pdf -- they are all made through clone() -- has rate == shape.
# Reproduced unless this switch is set.
FIX_GAMMA_CLONE = False


def _declare(prior, variable, param_type, differentiable=False):
    prior._register_variable(variable, differentiable=differentiable)
    prior.update_var_param_types(**{variable: param_type})
    prior._set_original_variables()


class GammaPrior(AbstractPrior):
    """log p(tau) = (shape - 1) log tau - rate * tau   (priors.py:23-25)"""

    def __init__(self, shape, rate):
        AbstractPrior.__init__(self, "precision_prior")
        self.shape, self.rate = shape, rate
        _declare(self, "precision", ScalarParameter)

    def _evaluate_log_prob(self, precision):
        return (self.shape - 1.0) * np.log(precision) - self.rate * precision

    def clone(self):
        twin = type(self)(self.shape, self.rate if FIX_GAMMA_CLONE else self.shape)
        twin.set_fixed_variables_from_pdf(self)
        return twin


class GaussianPrior(AbstractPrior):
    """log p(c) = -1/2 sum (c - means)^2 / variances   (priors.py:49-54).

    Registered non-differentiable like the reference (priors.py:45), so Posterior.gradient leaves
    its force out (quirk Q1); `differentiable=True` includes (c - means)/variances
    (BINFB_FLAG_PRIOR_GRAD on the device)."""

    def __init__(self, means, variances, differentiable=False, variable="coefficients"):
        AbstractPrior.__init__(self, variable + "_prior")
        for key, value in (("means", means), ("variances", variances)):
            self._register(key)
            self[key] = ArrayParameter(value, key)
        self._differentiable = bool(differentiable)
        self._variable = variable   # "coefficients" in the reference; user-defined models name their own
        _declare(self, variable, ArrayParameter, self._differentiable)

    def _z(self, value):
        return np.asarray(value, dtype=np.float64) - self["means"].value

    def _evaluate_log_prob(self, **variables):
        return -0.5 * np.sum(self._z(variables[self._variable]) ** 2 / self["variances"].value, axis=-1)

    def _evaluate_gradient(self, **variables):
        return self._z(variables[self._variable]) / self["variances"].value

    def clone(self):
        return type(self)(self["means"].value, self["variances"].value, self._differentiable, self._variable)


def make_priors():
    """the priors example_script.py uses: Gamma(1, 0.2) and N(0, 5 I_4)  (priors.py:67-73)"""
    priors = (GammaPrior(1.0, 0.2), GaussianPrior(means=np.zeros(4), variances=np.full(4, 5.0)))
    return {p.name: p for p in priors}
