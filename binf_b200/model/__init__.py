"""Models (forward models) are named callables that carry bindable parameters but are not pdfs
(the role of binf/model/__init__.py:13-90)."""
from binf_b200 import AbstractBinfNamedCallable
from binf_b200.params import ParameterNotFoundError, ParameterRegistry  # noqa: F401


class AbstractModel(ParameterRegistry, AbstractBinfNamedCallable):
    def __init__(self, name, parameters=()):
        AbstractBinfNamedCallable.__init__(self, name)
        self._init_registry()
        for param in parameters:
            self._register(param.name)
            self[param.name] = param

    def _accepts(self, name, obj):
        self._validate(name, obj)
        return True

    def _validate(self, param, value):
        """hook: raise to reject a parameter object"""

    def set_params(self, *values, **named_params):
        for key, src in list(zip(self.parameters, values)) + list(named_params.items()):
            self[key].set(src.value)

    def _complete_variables(self, variables):
        variables.update(self._fixed_values(self._original_variables))

    def _reduce_variables(self, **variables):
        for key in self.parameters:
            variables.pop(key, None)
