"""Bindable parameter objects (what the reference takes from CSB's
csb.statistics.pdf.parameterized: setup.py:25, binf/__init__.py:13, binf/example/likelihood.py:3).

A parameter holds a value; it can be *bound* to a base parameter, after which it follows the base
lazily (the base marks its dependants stale on `set`; they recompute on the next read).  This is
how a value set on a Posterior reaches the error model inside a Likelihood
(binf/pdf/posteriors.py:55, binf/pdf/likelihoods.py:87-88, binf/samplers/gibbs.py:54-62)."""
import numpy as np


class ParameterizationError(ValueError):
    pass


class ParameterValueError(ValueError):
    def __init__(self, name, value):
        super().__init__("%s = %r" % (name, value))
        self.name, self.value = name, value


class AbstractParameter(object):
    def __init__(self, value=None, name=None, base=None):
        self._name = str(name)
        self._base = None
        self._followers = []
        self._stale = False
        self._value = self._validate(value)
        if base is not None:
            self.bind_to(base)

    # hooks -----------------------------------------------------------------------------
    def _validate(self, value):
        return value

    def _compute(self, base_value):
        return base_value

    # API ---------------------------------------------------------------------------------
    @property
    def name(self):
        return self._name

    @property
    def is_virtual(self):
        return self._base is not None

    @property
    def _volatile(self):
        """True if a base up the chain changes without telling its followers (a foreign parameter object)"""
        return self._base is not None and self._base._volatile

    @property
    def value(self):
        if self._stale or self._volatile:
            self._value = self._validate(self._compute(self._base.value))
            self._stale = False
        return self._value

    def set(self, value):
        if self.is_virtual:
            raise ParameterizationError("a bound parameter cannot be set directly: " + self._name)
        self._value = self._validate(value)
        self._stale = False
        self._mark_followers()

    def bind_to(self, base):
        node = base
        while node is not None:
            if node is self:
                raise ParameterizationError("circular parameter binding: " + self._name)
            node = node._base
        if self._base is not None and self in self._base._followers:
            self._base._followers.remove(self)
        self._base = base
        base._followers.append(self)
        self._stale = True
        self._mark_followers()

    def _mark_followers(self):
        for f in self._followers:
            f._stale = True
            f._mark_followers()

    def __repr__(self):
        return "<%s %s=%r>" % (type(self).__name__, self._name, self.value)


class Parameter(AbstractParameter):
    """Scalar parameter (CSB's `Parameter`).  Batched chains may carry one value per chain: a
    1-d array is kept as float64 array instead of being coerced with float()."""

    def _validate(self, value):
        if hasattr(value, "detach"):      # torch tensor: per-chain values living on the device
            return value
        try:
            arr = np.asarray(value, dtype=np.float64)
        except (TypeError, ValueError):
            raise ParameterValueError(self._name, value)
        return float(arr) if arr.ndim == 0 else arr


class ArrayParameter(AbstractParameter):
    """Array-valued parameter (reference: binf/__init__.py:238-244)."""

    def _validate(self, value):
        if hasattr(value, "detach"):
            return value
        try:
            return np.array(value)
        except (TypeError, ValueError):
            raise ParameterValueError(self._name, value)


class ForeignParameter(AbstractParameter):
    """Adapter around a parameter object of ANOTHER library that quacks like CSB's (`.value`, `.set`, `.name`) --
    e.g. a real `csb.statistics.pdf.parameterized.Parameter` handed to a pdf of this package.  Reads and writes
    go through to the wrapped object; because that object cannot notify this package's followers when it
    changes, everything bound to the adapter re-reads it on every access."""

    def __init__(self, foreign):
        self._foreign = foreign
        super().__init__(None, getattr(foreign, "name", None))

    _volatile = True

    @property
    def value(self):
        return self._foreign.value

    def set(self, value):
        self._foreign.set(value)
        self._mark_followers()

    def bind_to(self, base):
        raise ParameterizationError("a foreign parameter object cannot be bound here: " + self._name)


def is_parameter_like(obj):
    return all(hasattr(obj, a) for a in ("value", "set", "name"))


def adopt(obj):
    """a parameter object this package can store and bind: itself, or a ForeignParameter around a duck-typed one"""
    if isinstance(obj, AbstractParameter):
        return obj
    if is_parameter_like(obj):
        return ForeignParameter(obj)
    return None


class ParameterNotFoundError(AttributeError):
    """raised when a parameter name is not registered (binf/pdf/__init__.py:14)"""


class ParameterRegistry(object):
    """Ordered name -> parameter-object table shared by pdfs and models.  A name has to be
    registered before an object can be stored under it; `_accepts` lets a subclass restrict what
    may be stored."""

    def _init_registry(self):
        self._table = {}

    def _register(self, name):
        self._table.setdefault(name, None)

    def _accepts(self, name, obj):
        return True

    def __getitem__(self, name):
        try:
            return self._table[name]
        except KeyError:
            raise ParameterNotFoundError(name)

    def __setitem__(self, name, obj):
        if name not in self._table:
            raise ParameterNotFoundError(name)
        if not isinstance(obj, AbstractParameter) and is_parameter_like(obj):
            obj = ForeignParameter(obj)   # duck-typed parameter of another library (a real CSB install)
        if not self._accepts(name, obj):
            raise TypeError(obj)
        self._table[name] = obj

    @property
    def parameters(self):
        return tuple(self._table)

    def get_params(self):
        return list(self._table.values())

    def _fixed_values(self, original):
        """values of the parameters that started life as variables"""
        return {n: self._table[n].value for n in self._table if n in original}
