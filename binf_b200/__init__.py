"""binf_b200 -- B200-native HMC hot path behind the Python API of simeoncarstens/binf.

Module layout mirrors the reference package (`binf/__init__.py`, `binf/pdf`, `binf/model`,
`binf/samplers`, `binf/example`) so that existing scripts keep working after

    import binf_b200; binf_b200.install_as_binf()      # `import binf...` now resolves here

This top-level module holds what `binf/__init__.py` holds in the reference: the named-callable
core (binf/__init__.py:16-226) and `ArrayParameter` (binf/__init__.py:238-244).  The CSB
parameter classes the reference imports (setup.py:25) are re-stated in `binf_b200.params`, so
there is no CSB dependency.  All numerics run in libbinf_b200.so (hand-written sm_100a CUDA,
loaded by `binf_b200._cabi`); there is no CPU fallback for the lowered models.
"""
import sys

import numpy as np

from .params import (AbstractParameter, Parameter, ArrayParameter, ParameterValueError,  # noqa: F401
                     ParameterizationError)

__version__ = "0.1.0"


class AbstractBinfNamedCallable(object):
    """Something that is called with named arguments ("variables"), some of which may later be
    frozen into bound parameters (reference: binf/__init__.py:16-226)."""

    def __init__(self, name):
        self._name = name
        self._variables = set()
        self._differentiable_variables = set()
        self._var_param_types = {}
        self._original_variables = set()

    # -- variable registry ---------------------------------------------------------------
    def _set_original_variables(self):
        self._original_variables.update(self.variables)

    def _register_variable(self, name, differentiable=False):
        if not isinstance(name, str):
            raise ValueError("Variable name must be a string, not %r" % type(name))
        if name in self._variables:
            raise ValueError('Variable name "%s" must be unique' % name)
        self._variables.add(name)
        if differentiable:
            self._differentiable_variables.add(name)

    def _delete_variable(self, name):
        if name not in self._variables:
            raise ValueError('"%s": unknown variable name' % name)
        self._variables.remove(name)
        self._differentiable_variables.discard(name)

    @property
    def variables(self):
        return self._variables

    @property
    def differentiable_variables(self):
        return self._differentiable_variables

    @property
    def name(self):
        return self._name

    @property
    def var_param_types(self):
        return self._var_param_types.copy()

    def update_var_param_types(self, **values):
        self._var_param_types.update(**values)

    def _get_variables_intersection(self, test_variables):
        return {k: v for k, v in test_variables.items() if k in self.variables}

    # -- evaluation -----------------------------------------------------------------------
    def _check_arity(self, variables):
        if len(variables) != len(self.variables):
            raise ValueError("Function called with %d arguments instead of %d!"
                             % (len(variables), len(self.variables)))

    def __call__(self, **variables):
        self._check_arity(variables)
        self._complete_variables(variables)
        return self._evaluate(**variables)

    def _evaluate(self, **variables):
        raise NotImplementedError

    def _check_differentiability(self, **variables):
        if not (set(variables) & set(self._differentiable_variables)):
            raise ValueError("Function cannot be differentiated w.r.t. any of the variables %s"
                             % sorted(variables))

    def _evaluate_gradient(self, **variables):
        raise NotImplementedError

    def gradient(self, **variables):
        self._check_arity(variables)
        self._complete_variables(variables)
        return self._evaluate_gradient(**variables)

    def _complete_variables(self, variables):
        raise NotImplementedError

    # -- freezing variables into parameters --------------------------------------------------
    def fix_variables(self, **fixed_vars):
        for v, value in fixed_vars.items():
            if v not in self.variables:
                raise ValueError("%r is not a variable of %r" % (v, self))
            if v not in self.var_param_types:
                raise ValueError('Parameter type for variable "%s" not defined' % v)
            self._delete_variable(v)
            self._register(v)
            self[v] = self.var_param_types[v](value, v)


def install_as_binf(alias_csb=True):
    """Register this package under the name `binf` so that scripts written against the reference
    (`from binf.samplers.hmc import HMCSampler`, ...) import the B200 implementation.

    alias_csb: also register the slice of CSB such scripts import (`csb.statistics.pdf.parameterized.Parameter`,
    `csb.numeric.exp`, `csb.statistics.samplers.State`, ...: binf/example/likelihood.py:3, binf/samplers/hmc.py:10,
    binf/samplers/gibbs.py:7-8) backed by binf_b200.params -- unless a real CSB is importable, whose parameter
    objects are wrapped when they are assigned to a pdf of this package (binf_b200.params.ForeignParameter)."""
    import importlib
    if alias_csb:
        from . import csbshim
        csbshim.install()
    names = ["pdf", "pdf.posteriors", "pdf.likelihoods", "pdf.priors", "model", "model.forwardmodels",
             "model.errormodels", "samplers", "samplers.hmc", "samplers.gibbs", "example",
             "example.likelihood", "example.priors", "example.samplers", "example.misc"]
    sys.modules["binf"] = sys.modules[__name__]
    for n in names:
        sys.modules["binf." + n] = importlib.import_module(__name__ + "." + n)
    return sys.modules["binf"]
