"""Lowering of a pdf object graph to a device model descriptor.

The reference evaluates `pdf.log_prob` / `pdf.gradient` by walking Posterior -> Likelihood ->
forward model + error model through dict plumbing on every call.  Here the graph
is inspected once: if it is made of the model classes the CUDA kernels implement, it is lowered
to a `binfb_model` handle (cached per data set) and every later evaluation is one C-ABI call.
Values that may change between Gibbs sweeps (the precision, the Gamma prior's numbers; reference
binf/samplers/gibbs.py:54-62) are re-read from the bound parameters at every call.

Unknown component classes are not lowered (`lower()` returns None) and the generic Python
plumbing of Posterior / Likelihood runs the user's own `_evaluate*` code instead; the built-in
model classes themselves have no host implementation, so nothing silently falls back to the CPU.
"""
import numpy as np

from binf_b200 import _cabi

_DEVICE = [0]
_CACHE = {}


def set_device(index):
    _DEVICE[0] = int(index)


def get_device():
    return _DEVICE[0]


def _is_tensor(x):
    return hasattr(x, "detach") and hasattr(x, "data_ptr")


class Lowered(object):
    """A pdf lowered to the device: `model` (owning _cabi.Model), the sampled variable and how to
    read the current precision / temperature."""

    def __init__(self, model, variable, precision_getter, beta_getter, likelihood_only, gamma):
        self.model = model
        self.variable = variable
        self.dim = model.dim
        self._precision_getter = precision_getter
        self._beta_getter = beta_getter
        self.likelihood_only = likelihood_only
        self._gamma = gamma

    def refresh(self):
        shape, rate = self._gamma()
        self.model.set_gamma_prior(shape, rate)

    def tau(self, n_chains, variables=None):
        v = self._precision_getter(variables)
        if _is_tensor(v):
            return v
        return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float32), (n_chains,)))

    def beta(self, n_chains):
        b = self._beta_getter()
        if b is None or _is_tensor(b):
            return b
        return np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float32), (n_chains,)))

    # ---- host-array evaluation used by Posterior / Likelihood.log_prob / gradient ------------
    def _shape(self, q):
        q = np.asarray(q, dtype=np.float64)
        single = q.ndim == 1
        return q.reshape(-1, self.dim), single

    def log_prob(self, q, variables=None):
        q2, single = self._shape(q)
        self.refresh()
        tau = self.tau(len(q2), variables)
        logp, _, chi2 = self.model.logprob_grad(q2, tau, self.beta(len(q2)), want_grad=False)
        if self.likelihood_only:
            # error-model log-prob from the device-reduced chi^2 (example/likelihood.py:54-57)
            t = np.asarray(tau, dtype=np.float64)
            b = self.beta(len(q2))
            b = 1.0 if b is None else np.asarray(b, dtype=np.float64)
            logp = b * (-0.5 * t * chi2 + 0.5 * self.model.n_data * np.log(t))
        return float(logp[0]) if single else logp

    def gradient(self, q, variables=None):
        q2, single = self._shape(q)
        self.refresh()
        _, grad, _ = self.model.logprob_grad(q2, self.tau(len(q2), variables), self.beta(len(q2)))
        grad = grad.astype(np.float64)
        return grad[0] if single else grad

    def chi2(self, q):
        q2, single = self._shape(q)
        _, _, chi2 = self.model.logprob_grad(q2, np.ones(len(q2), dtype=np.float32), None, want_grad=False)
        return float(chi2[0]) if single else chi2


def _cached_model(key, keep_alive, factory):
    entry = _CACHE.get(key)
    if entry is None:
        entry = (factory(), keep_alive)
        _CACHE[key] = entry
    return entry[0]


def _value_of(pdf, name, variables):
    if variables is not None and name in variables:
        return variables[name]
    if name in pdf.parameters:
        return pdf[name].value
    raise ValueError('no value for "%s": it is neither passed nor a fixed parameter' % name)


def lower(pdf, n_coeff=None):
    """Return a `Lowered` for `pdf` (a Posterior or a Likelihood built from the supported model
    classes) or None when the graph contains components the kernels do not implement."""
    from binf_b200.pdf.likelihoods import Likelihood
    from binf_b200.pdf.posteriors import Posterior
    from binf_b200.example.likelihood import ForwardModel as PolyForward, GaussianErrorModel
    from binf_b200.example.priors import GammaPrior, GaussianPrior
    from binf_b200.chromatin import ContactForwardModel, BackbonePrior, ExcludedVolumePrior
    from binf_b200.model.forwardmodels import DeviceForwardModel

    if isinstance(pdf, Posterior):
        liks = list(pdf.likelihoods.values())
        priors = list(pdf.priors.values())
        likelihood_only = False
    elif isinstance(pdf, Likelihood):
        liks, priors, likelihood_only = [pdf], [], True
    else:
        return None
    if len(liks) != 1 or type(liks[0]) is not Likelihood:
        return None
    lik = liks[0]
    fwm, em = lik.forward_model, lik.error_model
    if type(em) is not GaussianErrorModel:
        return None
    gamma_prior = None
    gauss_prior = None
    backbone = None
    exvol = None
    for pr in priors:
        if type(pr) is GammaPrior:
            gamma_prior = pr
        elif type(pr) is GaussianPrior:
            gauss_prior = pr
        elif type(pr) is BackbonePrior:
            backbone = pr
        elif type(pr) is ExcludedVolumePrior:
            exvol = pr
        else:
            return None

    def gamma():
        return (gamma_prior.shape, gamma_prior.rate) if gamma_prior is not None else (1.0, 0.0)

    def precision(variables):
        return _value_of(pdf, "precision", variables)

    def beta():
        return getattr(pdf, "beta", None)

    dev = get_device()
    if exvol is not None and type(fwm) is not ContactForwardModel:
        return None
    if type(fwm) is PolyForward:
        if backbone is not None:
            return None
        if getattr(fwm.polynomial, "__name__", "") != "polyval":
            return None  # only numpy.polynomial.polynomial.polyval ordering is implemented
        if n_coeff is None:
            if "coefficients" in pdf.parameters:
                n_coeff = len(np.atleast_1d(pdf["coefficients"].value))
            elif gauss_prior is not None:
                n_coeff = len(np.atleast_1d(gauss_prior["means"].value))
            else:
                raise ValueError("cannot infer the number of polynomial coefficients")
        if gauss_prior is not None:
            means = np.broadcast_to(np.asarray(gauss_prior["means"].value, dtype=np.float64), (n_coeff,))
            varis = np.broadcast_to(np.asarray(gauss_prior["variances"].value, dtype=np.float64), (n_coeff,))
            # quirk Q1: the reference's Posterior.gradient only sums components that declare a
            # differentiable variable (posteriors.py:183); GaussianPrior does not (priors.py:45)
            flags = _cabi.FLAG_PRIOR_GRAD if "coefficients" in gauss_prior.differentiable_variables else 0
        else:
            means, varis, flags = None, None, 0
        key = ("poly", id(fwm.xses), id(em.ys), n_coeff,
               None if means is None else (tuple(means), tuple(varis)), flags, dev)
        model = _cached_model(key, (fwm.xses, em.ys), lambda: _cabi.Model.polynomial(
            fwm.xses, em.ys, n_coeff, means, varis, *gamma(), flags=flags, device=dev))
        return Lowered(model, "coefficients", precision, beta, likelihood_only, gamma)
    if isinstance(fwm, DeviceForwardModel):
        # a user-defined per-datum model: kernels compiled at run time from its device code
        if backbone is not None:
            return None
        K = fwm.n_params
        if gauss_prior is not None:
            means = np.broadcast_to(np.asarray(gauss_prior["means"].value, dtype=np.float64), (K,))
            varis = np.broadcast_to(np.asarray(gauss_prior["variances"].value, dtype=np.float64), (K,))
            flags = _cabi.FLAG_PRIOR_GRAD if fwm.variable in gauss_prior.differentiable_variables else 0
        else:
            means, varis, flags = None, None, 0
        key = ("generic", id(fwm.xses), id(em.ys), fwm.device_code, K,
               None if means is None else (tuple(means), tuple(varis)), flags, dev)
        model = _cached_model(key, (fwm.xses, em.ys), lambda: _cabi.Model.generic(
            fwm.device_code, K, fwm.xses, em.ys, means, varis, *gamma(), flags=flags, device=dev))
        return Lowered(model, fwm.variable, precision, beta, likelihood_only, gamma)
    if type(fwm) is ContactForwardModel:
        if gauss_prior is not None:
            return None
        k_bb, l0, conf = (backbone.k_bb, backbone.l0, backbone.conf_s) if backbone is not None else (0.0, 1.0, 0.0)
        ev_k, ev_d = (exvol.k_ev, exvol.d_ev) if exvol is not None else (0.0, 0.0)
        contact = getattr(fwm, "contact", "logistic")
        key = ("chrom", id(em.ys), fwm.n_beads, fwm.alpha, fwm.d_c, contact, k_bb, l0, conf, ev_k, ev_d, dev)
        model = _cached_model(key, (em.ys,), lambda: _cabi.Model.chromatin(
            fwm.n_beads, em.ys, fwm.alpha, fwm.d_c, k_bb, l0, conf, *gamma(), device=dev, ev_k=ev_k, ev_d=ev_d,
            contact=contact))
        return Lowered(model, "structure", precision, beta, likelihood_only, gamma)
    return None
