"""Polynomial forward model and Gaussian error model (reference: binf/example/likelihood.py:11-79).

These classes are *descriptors*: they hold the data and declare variables exactly like the
reference's, and Posterior / Likelihood lower them to the fused CUDA kernels.  Their arithmetic
(polyval, the Vandermonde Jacobian, the residual reduction) exists only on the device."""
import numpy as np

from binf_b200 import ArrayParameter
from binf_b200.params import Parameter as ScalarParameter
from binf_b200.model.forwardmodels import AbstractForwardModel
from binf_b200.model.errormodels import AbstractErrorModel


class ForwardModel(AbstractForwardModel):
    """mock_n = sum_k c_k x_n^k  (likelihood.py:24-26)"""

    def __init__(self, xses, polynomial):
        AbstractForwardModel.__init__(self, "polynomial")
        self.xses, self.polynomial = xses, polynomial
        self._register_variable("coefficients", differentiable=True)
        self.update_var_param_types(coefficients=ArrayParameter)
        self._set_original_variables()

    def _device_model(self, n_coeff):
        from binf_b200 import _cabi
        from binf_b200.lowering import _cached_model, get_device
        if getattr(self.polynomial, "__name__", "") != "polyval":
            raise NotImplementedError("only numpy.polynomial.polynomial.polyval is lowered to the device")
        key = ("poly-fwd", id(self.xses), n_coeff, get_device())
        return _cached_model(key, (self.xses,), lambda: _cabi.Model.polynomial(
            self.xses, np.zeros(len(self.xses)), n_coeff, device=get_device()))

    def _evaluate(self, coefficients):
        c = np.asarray(coefficients, dtype=np.float64)
        mock = self._device_model(c.shape[-1]).forward(c.reshape(-1, c.shape[-1])).astype(np.float64)
        return mock[0] if c.ndim == 1 else mock

    def _evaluate_jacobi_matrix(self, coefficients):
        raise NotImplementedError(
            "the dense Jacobian is never formed on the B200 path: Likelihood.gradient applies it "
            "inside the fused kernel (reference: binf/pdf/likelihoods.py:148-155)")

    def clone(self):
        twin = type(self)(self.xses, self.polynomial)
        self._set_parameters(twin)
        return twin


class GaussianErrorModel(AbstractErrorModel):
    """log p = -1/2 tau sum (mock - y)^2 + 1/2 N log tau  (likelihood.py:54-57); evaluated fused
    with the forward model on the device."""

    def __init__(self, ys):
        AbstractErrorModel.__init__(self, "error_model")
        self.ys = ys
        for variable, kind in (("mock_data", ArrayParameter), ("precision", ScalarParameter)):
            self._register_variable(variable)
            self.update_var_param_types(**{variable: kind})
        self._set_original_variables()

    def _evaluate_log_prob(self, mock_data, precision):
        raise NotImplementedError(
            "GaussianErrorModel is evaluated fused with a built-in forward model on the device; "
            "stand-alone evaluation on host mock data is not part of the B200 path")

    _evaluate_gradient = _evaluate_log_prob

    def clone(self):
        twin = type(self)(self.ys)
        twin.set_fixed_variables_from_pdf(self)
        return twin


def make_likelihood(xses, ys, polynomial):
    from binf_b200.pdf.likelihoods import Likelihood
    return Likelihood("points", ForwardModel(xses, polynomial), GaussianErrorModel(ys))
