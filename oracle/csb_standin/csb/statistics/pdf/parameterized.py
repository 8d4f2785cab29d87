"""csb.statistics.pdf.parameterized stand-in: bindable parameter objects."""
from csb.statistics.pdf import (AbstractDensity, ParameterNotFoundError,
                                ParameterValueError)


class ParameterizationError(ValueError):
    pass


class AbstractParameter(object):
    """A value holder that can be bound to a base parameter; a bound parameter
    recomputes its value lazily from the base whenever the base changes."""

    def __init__(self, value=None, name=None, base=None):
        self._derivatives = set()
        self._base = None
        self._consistent = True
        self._name = str(name)
        self._value = None
        self._update(value)
        if base is not None:
            self.bind_to(base)

    @property
    def name(self):
        return self._name

    @property
    def value(self):
        if not self._consistent:
            self._value = self._validate(self._compute(self._base.value))
            self._consistent = True
        return self._value

    @property
    def is_virtual(self):
        return self._base is not None

    def _validate(self, value):
        return value

    def _compute(self, base_value):
        return base_value

    def _update(self, value):
        self._value = self._validate(value)
        self._consistent = True
        self._invalidate_derivatives()

    def _invalidate_derivatives(self):
        for d in self._derivatives:
            d._consistent = False
            d._invalidate_derivatives()

    def set(self, value):
        if self.is_virtual:
            raise ParameterizationError("Virtual parameters can't be updated explicitly")
        self._update(value)

    def bind_to(self, base):
        if base is self:
            raise ParameterizationError("circular binding")
        if self._base is not None:
            self._base._derivatives.discard(self)
        self._base = base
        base._derivatives.add(self)
        self._consistent = False
        self._invalidate_derivatives()


class Parameter(AbstractParameter):
    def _validate(self, value):
        return float(value)

    def _compute(self, base_value):
        return base_value


class NonVirtualParameter(Parameter):
    def bind_to(self, parameter):
        raise ParameterizationError("cannot bind")


class ParameterizedDensity(AbstractDensity):
    def _validate(self, param, value):
        if not isinstance(value, AbstractParameter):
            raise TypeError(value)
