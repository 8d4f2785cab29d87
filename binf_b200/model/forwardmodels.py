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


class DeviceForwardModel(AbstractForwardModel):
    """A USER-DEFINED per-datum forward model lowered to the device.

    The reference's extension point is a subclass of AbstractForwardModel with `_evaluate` (mock data
    [N]) and `_evaluate_jacobi_matrix` ([K, N]) in numpy (binf/model/forwardmodels.py:30-38).  Here the
    same two things are given as CUDA device code for ONE datum,

        __device__ float binfb_mock(const float *theta, const float *x, float *dmock);

    (return f_n(theta), write dmock[k] = d f_n / d theta_k), and `Likelihood(name, model,
    GaussianErrorModel(ys))` / `Posterior` / `HMCSampler` run it through kernels compiled at run time
    with NVRTC -- the same fused transition the built-in polynomial model gets.

        class Decay(DeviceForwardModel):
            device_code = '''
            __device__ float binfb_mock(const float *theta, const float *x, float *dmock) {
                const float e = __expf(-theta[1] * x[0]);
                dmock[0] = e; dmock[1] = -theta[0] * x[0] * e; dmock[2] = 1.0f;
                return theta[0] * e + theta[2];
            }'''
        fwm = Decay("decay", xs, variable="rates", n_params=3)
    """
    device_code = None

    def __init__(self, name, xses, variable, n_params, device_code=None):
        from binf_b200 import ArrayParameter
        AbstractForwardModel.__init__(self, name)
        self.xses = xses
        self.variable = variable
        self.n_params = int(n_params)
        if device_code is not None:
            self.device_code = device_code
        if not self.device_code or "binfb_mock" not in self.device_code:
            raise ValueError("DeviceForwardModel: device_code must define __device__ float binfb_mock(...)")
        self._register_variable(variable, differentiable=True)
        self.update_var_param_types(**{variable: ArrayParameter})
        self._set_original_variables()

    def check_device_code(self):
        """compile the device code (NVRTC, no GPU needed); raises ValueError with the compiler log"""
        import numpy as np
        from binf_b200 import _cabi
        x_dim = int(np.asarray(self.xses).reshape(len(self.xses), -1).shape[1])
        ok, log = _cabi.generic_compile_check(self.device_code, self.n_params, x_dim)
        if not ok:
            raise ValueError("device code of %r does not compile:\n%s" % (self.name, log))
        return log

    def _device_model(self):
        import numpy as np
        from binf_b200 import _cabi
        from binf_b200.lowering import _cached_model, get_device
        key = ("generic-fwd", id(self.xses), self.device_code, self.n_params, get_device())
        return _cached_model(key, (self.xses,), lambda: _cabi.Model.generic(
            self.device_code, self.n_params, self.xses, np.zeros(len(self.xses)), device=get_device()))

    def _evaluate(self, **variables):
        import numpy as np
        c = np.asarray(variables[self.variable], dtype=np.float64)
        mock = self._device_model().forward(c.reshape(-1, c.shape[-1])).astype(np.float64)
        return mock[0] if c.ndim == 1 else mock

    def _evaluate_jacobi_matrix(self, **variables):
        raise NotImplementedError(
            "the dense Jacobian is never formed on the B200 path: Likelihood.gradient applies it "
            "inside the fused kernel (reference: binf/pdf/likelihoods.py:148-155)")

    def clone(self):
        twin = type(self)(self.name, self.xses, self.variable, self.n_params, self.device_code)
        self._set_parameters(twin)
        return twin
