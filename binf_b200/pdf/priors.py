"""Priors are ordinary pdfs over (a subset of) the model variables; the class only marks the role a
component plays inside a Posterior (`Posterior(likelihoods, priors)`), as binf/pdf/priors.py:10 does.
A prior takes part in `Posterior.gradient` only if it registers its variable with
`differentiable=True` (binf/pdf/posteriors.py:182-185)."""
from binf_b200.pdf import AbstractBinfPDF


class AbstractPrior(AbstractBinfPDF):
    is_prior = True
