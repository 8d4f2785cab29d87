"""Posterior wiring, MAP estimate and predictive density of the example
(reference: binf/example/misc.py:3-33)."""
import numpy as np


def predict(x, y, samples, polynomial=None):
    """Posterior-predictive density of new data (x, y): the mean over `samples` of the Gaussian
    density N(y; polynomial(x, coefficients), 1/precision) (misc.py:3-16), evaluated on the device
    for all points at once.  `samples`: a list of BinfState like the reference's, or a pair
    (coefficients [S, K], precision [S]) -- e.g. the flattened output of SampleSink.samples().
    x, y: scalars or arrays of equal shape."""
    from binf_b200 import _cabi
    if isinstance(samples, tuple):
        coeffs, precision = samples
    else:
        coeffs = np.array([s.variables["coefficients"] for s in samples])
        precision = np.array([s.variables["precision"] for s in samples])
    out = _cabi.posterior_predictive(coeffs, precision, x, y)
    return float(out.reshape(-1)[0]) if np.ndim(x) == 0 else out


def get_MAP(samples, log_probs):
    """the sample of maximum log-probability (misc.py:18-22)"""
    map_sample = samples[int(np.argmax(log_probs))]
    return map_sample.variables["coefficients"], map_sample.variables["precision"]


def make_posterior(xses, ys, polynomial):
    from binf_b200.pdf.posteriors import Posterior
    from binf_b200.example.likelihood import make_likelihood
    from binf_b200.example.priors import make_priors
    lik = make_likelihood(xses, ys, polynomial)
    return Posterior({lik.name: lik}, make_priors())
