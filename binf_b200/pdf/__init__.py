"""PDF base class: named variables + bindable parameters + log_prob / gradient / conditioning
(the role of binf/pdf/__init__.py:14-160 in the reference)."""
import numpy as np

from binf_b200 import AbstractBinfNamedCallable
from binf_b200.params import AbstractParameter, ParameterNotFoundError, ParameterRegistry  # noqa: F401


class AbstractBinfPDF(ParameterRegistry, AbstractBinfNamedCallable):
    def __init__(self, name="", **args):
        AbstractBinfNamedCallable.__init__(self, name)
        self._init_registry()
        self._var_param_types = {}

    def _accepts(self, name, obj):
        return isinstance(obj, AbstractParameter)

    def set_params(self, *values, **named):
        for key, obj in list(zip(self.parameters, values)) + list(named.items()):
            self[key] = obj

    @property
    def estimator(self):
        raise NotImplementedError

    def estimate(self, data):
        raise NotImplementedError

    # ---- evaluation: fixed parameters are merged into the keyword arguments first -------------
    def _complete_variables(self, variables):
        variables.update(self._fixed_values(self._original_variables))

    def _evaluate_log_prob(self, **variables):
        raise NotImplementedError

    def log_prob(self, **variables):
        self._complete_variables(variables)
        return self._evaluate_log_prob(**variables)

    def gradient(self, **variables):
        """gradient of the ENERGY -log p, the sign HMCSampler._leapfrog subtracts (hmc.py:116)"""
        self._complete_variables(variables)
        return self._evaluate_gradient(**variables)

    def _evaluate(self, **variables):
        return np.exp(np.clip(self.log_prob(**variables), -308.0, 709.0))

    # ---- conditioning ------------------------------------------------------------------------
    def clone(self):
        raise NotImplementedError

    def conditional_factory(self, **fixed_vars):
        """copy of this pdf with the given variables frozen (pdf/__init__.py:49-70)"""
        conditioned = self.clone()
        conditioned.fix_variables(**self._get_variables_intersection(fixed_vars))
        return conditioned

    def set_fixed_variables_from_pdf(self, pdf):
        """freeze here whatever `pdf` has frozen and this object has not (pdf/__init__.py:142-151)"""
        theirs = {n: pdf[n].value for n in pdf.parameters if n not in self.parameters}
        self.fix_variables(**self._get_variables_intersection(theirs))
