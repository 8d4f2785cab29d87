"""PDF base class (reference: binf/pdf/__init__.py:14-160)."""
from collections import OrderedDict

import numpy as np

from binf_b200 import AbstractBinfNamedCallable
from binf_b200.params import AbstractParameter


class ParameterNotFoundError(AttributeError):
    pass


class AbstractBinfPDF(AbstractBinfNamedCallable):
    """A density with named variables and a registry of bindable parameters
    (CSB's ParameterizedDensity + AbstractBinfNamedCallable in the reference)."""

    def __init__(self, name="", **args):
        AbstractBinfNamedCallable.__init__(self, name)
        self._params = OrderedDict()
        self._var_param_types = {}

    # -- parameter registry ------------------------------------------------------------------
    def _register(self, name):
        if name not in self._params:
            self._params[name] = None

    def __getitem__(self, param):
        if param in self._params:
            return self._params[param]
        raise ParameterNotFoundError(param)

    def __setitem__(self, param, value):
        if param not in self._params:
            raise ParameterNotFoundError(param)
        if not isinstance(value, AbstractParameter):
            raise TypeError(value)
        self._params[param] = value

    @property
    def parameters(self):
        return tuple(self._params)

    def get_params(self):
        return [self._params[n] for n in self.parameters]

    def set_params(self, *values, **named):
        for p, v in zip(self.parameters, values):
            self[p] = v
        for p, v in named.items():
            self[p] = v

    @property
    def estimator(self):
        raise NotImplementedError

    def estimate(self, data):
        raise NotImplementedError

    # -- evaluation ------------------------------------------------------------------------
    def _evaluate_log_prob(self, **variables):
        raise NotImplementedError

    def _evaluate(self, **variables):
        return np.exp(np.clip(self.log_prob(**variables), -308.0, 709.0))

    def log_prob(self, **variables):
        self._complete_variables(variables)
        return self._evaluate_log_prob(**variables)

    def gradient(self, **variables):
        """Gradient of the ENERGY -log p (the sign HMCSampler._leapfrog expects, hmc.py:116)."""
        self._complete_variables(variables)
        return self._evaluate_gradient(**variables)

    def _complete_variables(self, variables):
        variables.update({p: self[p].value for p in self.parameters if p in self._original_variables})

    # -- conditioning -------------------------------------------------------------------------
    def clone(self):
        raise NotImplementedError

    def conditional_factory(self, **fixed_vars):
        """A copy with some variables frozen to the given values (pdf/__init__.py:49-70)."""
        result = self.clone()
        result.fix_variables(**self._get_variables_intersection(fixed_vars))
        return result

    def set_fixed_variables_from_pdf(self, pdf):
        values = {p: pdf[p].value for p in pdf.parameters if p not in self.parameters}
        self.fix_variables(**self._get_variables_intersection(values))
