"""Forward-model interface (reference: binf/model/forwardmodels.py:10-66)."""
from binf_b200.model import AbstractModel


class AbstractForwardModel(AbstractModel):
    def __init__(self, name, parameters=()):
        super(AbstractForwardModel, self).__init__(name, parameters)

    @property
    def data(self):
        return self._data

    def jacobi_matrix(self, **variables):
        self._complete_variables(variables)
        return self._evaluate_jacobi_matrix(**variables)

    def _evaluate_jacobi_matrix(self, **model_parameters):
        self._check_differentiability(**model_parameters)

    def clone(self):
        raise NotImplementedError

    def _set_parameters(self, copy):
        """carry frozen variables over to a clone (forwardmodels.py:59-66)"""
        for p in self.parameters:
            if p not in copy.parameters:
                copy._register(p)
                copy[p] = self[p].__class__(self[p].value, p)
                if p in copy.variables:
                    copy._delete_variable(p)
