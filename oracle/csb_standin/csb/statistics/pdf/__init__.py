"""csb.statistics.pdf stand-in: the parameter-holding density base class."""
from collections import OrderedDict


class ParameterNotFoundError(AttributeError):
    pass


class ParameterValueError(ValueError):
    def __init__(self, param, value):
        self.param = param
        self.value = value
        super().__init__("{0} = {1}".format(param, value))


class AbstractDensity(object):
    """Ordered name -> parameter-object registry."""

    def __init__(self):
        self._params = OrderedDict()

    def _register(self, name):
        if name not in self._params:
            self._params[name] = None

    def __getitem__(self, param):
        if param in self._params:
            return self._params[param]
        raise ParameterNotFoundError(param)

    def __setitem__(self, param, value):
        if param in self._params:
            self._validate(param, value)
            self._params[param] = value
        else:
            raise ParameterNotFoundError(param)

    def _validate(self, param, value):
        pass

    @property
    def parameters(self):
        return tuple(self._params)

    def get_params(self):
        return [self._params[name] for name in self.parameters]

    def set_params(self, *values, **named_params):
        for p, v in zip(self.parameters, values):
            self[p] = v
        for p in named_params:
            self[p] = named_params[p]
