"""Prior base class (reference: binf/pdf/priors.py:10-12)."""
from binf_b200.pdf import AbstractBinfPDF


class AbstractPrior(AbstractBinfPDF):
    pass
