"""Posterior wiring of the example (reference: binf/example/misc.py:24-33)."""


def make_posterior(xses, ys, polynomial):
    from binf_b200.pdf.posteriors import Posterior
    from binf_b200.example.likelihood import make_likelihood
    from binf_b200.example.priors import make_priors
    lik = make_likelihood(xses, ys, polynomial)
    return Posterior({lik.name: lik}, make_priors())
