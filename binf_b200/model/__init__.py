"""Model base class: a named callable with its own parameter registry
(reference: binf/model/__init__.py:13-90)."""
from collections import OrderedDict

from binf_b200 import AbstractBinfNamedCallable
from binf_b200.pdf import ParameterNotFoundError


class AbstractModel(AbstractBinfNamedCallable):
    def __init__(self, name, parameters=()):
        super(AbstractModel, self).__init__(name)
        self._params = OrderedDict()
        for p in parameters:
            self._register(p.name)
            self[p.name] = p

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
        self._validate(param, value)
        self._params[param] = value

    def _validate(self, param, value):
        pass

    def set_params(self, *values, **named_params):
        for p, v in zip(self.parameters, values):
            self[p].set(v.value)
        for p, v in named_params.items():
            self[p].set(v.value)

    @property
    def parameters(self):
        return tuple(self._params)

    def get_params(self):
        return [self._params[n] for n in self.parameters]

    def _complete_variables(self, variables):
        variables.update({p: self[p].value for p in self.parameters if p in self._original_variables})

    def _reduce_variables(self, **variables):
        for p in self.parameters:
            variables.pop(p, None)
