"""Chromatin bead-chain model behind the reference's forward-model / prior API.

Not in the reference; specified in SURVEY.md Appendix A.2 and restated for the
checker in oracle/chromatin_port.py.  `structure` is X.flatten() for X of shape (n_beads, 3);
contact data follow numpy.triu_indices(n_beads, 1)."""
import numpy as np

from binf_b200 import ArrayParameter
from binf_b200.model.forwardmodels import AbstractForwardModel
from binf_b200.pdf.priors import AbstractPrior


class ContactForwardModel(AbstractForwardModel):
    """mock_ij = s(alpha (d_c - |x_i - x_j|)) with the logistic s(z) = 1 / (1 + exp(-z)) (default) or the
    algebraic s(z) = 1/2 (1 + z / sqrt(1 + z^2)) (contact="algebraic")"""

    def __init__(self, n_beads, alpha, d_c, contact="logistic"):
        super(ContactForwardModel, self).__init__("contacts")
        if contact not in ("logistic", "algebraic"):
            raise ValueError("contact: 'logistic' or 'algebraic', got %r" % (contact,))
        self.n_beads, self.alpha, self.d_c, self.contact = int(n_beads), float(alpha), float(d_c), contact
        self._register_variable("structure", differentiable=True)
        self.update_var_param_types(structure=ArrayParameter)
        self._set_original_variables()

    def _evaluate(self, structure):
        from binf_b200 import _cabi
        from binf_b200.lowering import _cached_model, get_device
        n = self.n_beads
        key = ("chrom-fwd", n, self.alpha, self.d_c, self.contact, get_device())
        model = _cached_model(key, (), lambda: _cabi.Model.chromatin(
            n, np.zeros(n * (n - 1) // 2, dtype=np.float32), self.alpha, self.d_c, 0.0, 1.0,
            device=get_device(), contact=self.contact))
        q = np.asarray(structure, dtype=np.float64)
        mock = model.forward(q.reshape(-1, 3 * n)).astype(np.float64)
        return mock[0] if q.ndim == 1 else mock

    def _evaluate_jacobi_matrix(self, structure):
        raise NotImplementedError(
            "the 3n x n(n-1)/2 Jacobian is never formed: Likelihood.gradient applies it inside the "
            "fused pair kernel")

    def clone(self):
        copy = self.__class__(self.n_beads, self.alpha, self.d_c, self.contact)
        self._set_parameters(copy)
        return copy


class BackbonePrior(AbstractPrior):
    """-1/2 k_bb sum_i (|x_{i+1} - x_i| - l0)^2  [- 1/2 |X|^2 / conf_s^2 if conf_s > 0]; its force is
    part of the fused kernel (differentiable=True, so Posterior.gradient includes it)."""

    def __init__(self, n_beads, k_bb, l0, conf_s=0.0):
        super(BackbonePrior, self).__init__("structure_prior")
        self.n_beads, self.k_bb, self.l0, self.conf_s = int(n_beads), float(k_bb), float(l0), float(conf_s)
        self._register_variable("structure", differentiable=True)
        self.update_var_param_types(structure=ArrayParameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, structure):
        raise NotImplementedError("BackbonePrior is evaluated inside the fused chromatin kernel")

    _evaluate_gradient = _evaluate_log_prob

    def clone(self):
        copy = self.__class__(self.n_beads, self.k_bb, self.l0, self.conf_s)
        copy.set_fixed_variables_from_pdf(self)
        return copy


class ExcludedVolumePrior(AbstractPrior):
    """-k_ev sum_{i<j} max(0, d_ev - |x_i - x_j|)^4: a quartic repulsion between beads closer than d_ev
    (SURVEY.md 8f rank 2).  Its force is fused into the pair loop of the chromatin kernel."""

    def __init__(self, n_beads, k_ev, d_ev):
        super(ExcludedVolumePrior, self).__init__("excluded_volume")
        self.n_beads, self.k_ev, self.d_ev = int(n_beads), float(k_ev), float(d_ev)
        self._register_variable("structure", differentiable=True)
        self.update_var_param_types(structure=ArrayParameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, structure):
        raise NotImplementedError("ExcludedVolumePrior is evaluated inside the fused chromatin kernel")

    _evaluate_gradient = _evaluate_log_prob

    def clone(self):
        copy = self.__class__(self.n_beads, self.k_ev, self.d_ev)
        copy.set_fixed_variables_from_pdf(self)
        return copy


def make_chromatin_posterior(n_beads, y_pairs, alpha=2.0, d_c=2.5, k_bb=4.0, l0=1.0, conf_s=0.0,
                             gamma_shape=1.0, gamma_rate=1.0, ev_k=0.0, ev_d=0.0, contact="logistic"):
    """Posterior({contacts likelihood}, {backbone prior, Gamma precision prior[, excluded volume]}) over
    the variables `structure` and `precision`."""
    from binf_b200.pdf.likelihoods import Likelihood
    from binf_b200.pdf.posteriors import Posterior
    from binf_b200.example.likelihood import GaussianErrorModel
    from binf_b200.example.priors import GammaPrior
    y = np.ascontiguousarray(y_pairs, dtype=np.float32)
    lik = Likelihood("points", ContactForwardModel(n_beads, alpha, d_c, contact), GaussianErrorModel(y))
    priors = {"structure_prior": BackbonePrior(n_beads, k_bb, l0, conf_s),
              "precision_prior": GammaPrior(gamma_shape, gamma_rate)}
    if ev_k > 0.0:
        priors["excluded_volume"] = ExcludedVolumePrior(n_beads, ev_k, ev_d)
    return Posterior({lik.name: lik}, priors)
