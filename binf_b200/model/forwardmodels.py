"""Forward models map model variables to idealised ("mock") data (binf/model/forwardmodels.py:10-66).
On the B200 path `jacobi_matrix` exists for user-defined models only: the built-in models never form
their Jacobian, the fused kernels apply it on the fly."""
from binf_b200.model import AbstractModel


class AbstractForwardModel(AbstractModel):
    def __init__(self, name, parameters=()):
        AbstractModel.__init__(self, name, parameters)

    @property
    def data(self):
        return self._data

    def _evaluate_jacobi_matrix(self, **model_parameters):
        self._check_differentiability(**model_parameters)

    def jacobi_matrix(self, **variables):
        """d mock / d variables with the fixed parameters filled in (forwardmodels.py:23-28)"""
        self._complete_variables(variables)
        return self._evaluate_jacobi_matrix(**variables)

    def clone(self):
        raise NotImplementedError

    def _set_parameters(self, copy):
        """give a fresh clone the variables this model has already frozen (forwardmodels.py:59-66)"""
        for key in (k for k in self.parameters if k not in copy.parameters):
            frozen = self[key]
            copy._register(key)
            copy[key] = type(frozen)(frozen.value, key)
            if key in copy.variables:
                copy._delete_variable(key)
