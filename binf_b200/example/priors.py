"""Gamma prior on the precision, Gaussian prior on the coefficients
(reference: binf/example/priors.py:10-73).  Scalar closed forms; they enter the device log_prob as
model constants (binfb_model_create_polynomial / binfb_model_set_gamma_prior)."""
import numpy as np

from binf_b200 import ArrayParameter
from binf_b200.params import Parameter as ScalarParameter
from binf_b200.pdf.priors import AbstractPrior

# The reference's GammaPrior.clone passes `shape` twice (priors.py:29), so every conditional pdf
# carries rate == shape (quirk Q2).  Reproduced by default; set to True for the intended behaviour.
FIX_GAMMA_CLONE = False


class GammaPrior(AbstractPrior):
    def __init__(self, shape, rate):
        super(GammaPrior, self).__init__("precision_prior")
        self.shape = shape
        self.rate = rate
        self._register_variable("precision")
        self.update_var_param_types(precision=ScalarParameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, precision):
        return (self.shape - 1.0) * np.log(precision) - precision * self.rate

    def clone(self):
        copy = self.__class__(self.shape, self.rate if FIX_GAMMA_CLONE else self.shape)
        copy.set_fixed_variables_from_pdf(self)
        return copy


class GaussianPrior(AbstractPrior):
    """-1/2 sum (c - mu)^2 / v.  `differentiable=False` by default like the reference
    (priors.py:45), which makes Posterior.gradient skip it (quirk Q1); pass differentiable=True
    to include the force (c - mu)/v (device flag BINFB_FLAG_PRIOR_GRAD)."""

    def __init__(self, means, variances, differentiable=False):
        super(GaussianPrior, self).__init__("coefficients_prior")
        self._register("means")
        self._register("variances")
        self["means"] = ArrayParameter(means, "means")
        self["variances"] = ArrayParameter(variances, "variances")
        self._differentiable = differentiable
        self._register_variable("coefficients", differentiable=differentiable)
        self.update_var_param_types(coefficients=ArrayParameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, coefficients):
        c = np.asarray(coefficients, dtype=np.float64)
        return -0.5 * np.sum((c - self["means"].value) ** 2 / self["variances"].value, axis=-1)

    def _evaluate_gradient(self, coefficients):
        return (np.asarray(coefficients, dtype=np.float64) - self["means"].value) / self["variances"].value

    def clone(self):
        return self.__class__(self["means"].value, self["variances"].value, self._differentiable)


def make_priors():
    pp = GammaPrior(1.0, 0.2)
    cp = GaussianPrior(means=np.zeros(4), variances=np.ones(4) * 5)
    return {pp.name: pp, cp.name: cp}
