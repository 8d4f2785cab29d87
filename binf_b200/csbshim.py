"""The slice of CSB the reference and scripts written against it import, backed by binf_b200.params.

The reference depends on CSB (setup.py:25) for its parameter objects and a few helpers:
`csb.statistics.pdf.parameterized.{AbstractParameter, Parameter, ParameterizedDensity}` (binf/__init__.py:13,
binf/pdf/__init__.py:11, binf/example/likelihood.py:3, binf/example/priors.py:3, binf/tests/pdf/__init__.py:6),
`csb.numeric.{exp, log, log_sum_exp}` (binf/samplers/hmc.py:10, binf/example/misc.py:5),
`csb.statistics.samplers.State` and the `AbstractMC` / `AbstractSingleChainMC` marker classes
(binf/samplers/gibbs.py:7-8,119), `csb.core.OrderedDict` (binf/model/__init__.py:8).  binf_b200 re-states the
parameter classes in `binf_b200.params` and does not need CSB; `install()` registers module objects under the
`csb...` names that hand out THOSE classes, so that a user script doing

    from csb.statistics.pdf.parameterized import Parameter

gets objects the mirror's pdfs accept and can bind.  Nothing is registered when a real CSB is importable
(its parameter objects are then wrapped on assignment, see binf_b200.params.adopt)."""
import importlib.util
import sys
import types
from collections import OrderedDict

import numpy as np

from binf_b200 import params
from binf_b200.samplers import AbstractMC, State

EXP_MIN, EXP_MAX = -308.0, 709.0
LOG_MIN, LOG_MAX = 1e-308, 1e308


def exp(x, x_min=EXP_MIN, x_max=EXP_MAX):
    """csb.numeric.exp: exponential of the argument clipped to the representable range (hmc.py:151)"""
    return np.exp(np.clip(x, x_min, x_max))


def log(x, x_min=LOG_MIN, x_max=LOG_MAX):
    return np.log(np.clip(x, x_min, x_max))


def log_sum_exp(x, axis=0):
    x = np.asarray(x)
    xmax = x.max(axis)
    return np.log(np.exp(x - xmax).sum(axis)) + xmax


class AbstractSingleChainMC(AbstractMC):
    """marker base class (binf/samplers/gibbs.py:8,11: GibbsSampler only inherits the type)"""


class AbstractDensity(params.ParameterRegistry):
    """csb.statistics.pdf.AbstractDensity: an ordered table of named parameter objects"""

    def __init__(self):
        self._init_registry()

    def _accepts(self, name, obj):
        return isinstance(obj, params.AbstractParameter)

    def set_params(self, *values, **named):
        for key, obj in list(zip(self.parameters, values)) + list(named.items()):
            self[key] = obj


class ParameterizedDensity(AbstractDensity):
    pass


def iterable(obj):
    try:
        iter(obj)
        return True
    except TypeError:
        return False


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__binf_b200_shim__ = True
    return m


def real_csb_present():
    if "csb" in sys.modules:
        return not getattr(sys.modules["csb"], "__binf_b200_shim__", False)
    try:
        return importlib.util.find_spec("csb") is not None
    except (ImportError, ValueError):
        return False


def install(force=False):
    """Register the shim modules under the `csb...` names.  Returns True if they were installed, False if a
    real CSB is importable (left alone unless force=True)."""
    if real_csb_present() and not force:
        return False
    parameterized = _module("csb.statistics.pdf.parameterized", AbstractParameter=params.AbstractParameter,
                            Parameter=params.Parameter, ParameterizedDensity=ParameterizedDensity,
                            ParameterValueError=params.ParameterValueError,
                            ParameterizationError=params.ParameterizationError,
                            NonVirtualParameter=params.Parameter)
    pdf = _module("csb.statistics.pdf", AbstractDensity=AbstractDensity, parameterized=parameterized,
                  ParameterNotFoundError=params.ParameterNotFoundError,
                  ParameterValueError=params.ParameterValueError)
    singlechain = _module("csb.statistics.samplers.mc.singlechain", AbstractSingleChainMC=AbstractSingleChainMC)
    mc = _module("csb.statistics.samplers.mc", AbstractMC=AbstractMC, singlechain=singlechain)
    samplers = _module("csb.statistics.samplers", State=State, mc=mc)
    statistics = _module("csb.statistics", pdf=pdf, samplers=samplers)
    numeric = _module("csb.numeric", exp=exp, log=log, log_sum_exp=log_sum_exp, EXP_MIN=EXP_MIN, EXP_MAX=EXP_MAX,
                      LOG_MIN=LOG_MIN, LOG_MAX=LOG_MAX)
    core = _module("csb.core", OrderedDict=OrderedDict, iterable=iterable)
    csb = _module("csb", statistics=statistics, numeric=numeric, core=core)
    for m in (csb, statistics, pdf, parameterized, samplers, mc, singlechain, numeric, core):
        m.__path__ = []       # importable as packages: `import csb.statistics.pdf.parameterized` works
        sys.modules[m.__name__] = m
    return True
