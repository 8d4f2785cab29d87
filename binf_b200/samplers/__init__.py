"""Sampler state container (reference: binf/samplers/__init__.py:9-57)."""


class State(object):
    """position/momentum attribute bag (CSB's csb.statistics.samplers.State, which the reference's
    GibbsSampler unwraps at gibbs.py:131-134)."""

    def __init__(self, position, momentum=None):
        self.position = position
        self.momentum = momentum


class AbstractMC(object):
    """marker base class of samplers that expect a `State` (gibbs.py:119-123)"""


class BinfState(object):
    def __init__(self, variables={}, momenta={}):
        self._variables = {}
        self._momenta = {}
        self.update_variables(**variables)
        self.update_momenta(**momenta)

    @property
    def variables(self):
        """a shallow copy of the name -> value mapping"""
        return self._variables.copy()

    def update_variables(self, **variables):
        self._variables.update(variables)

    @property
    def momenta(self):
        return self._momenta.copy()

    def update_momenta(self, **momenta):
        self._momenta.update(momenta)
