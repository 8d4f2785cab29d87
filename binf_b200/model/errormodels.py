"""Error-model interface (reference: binf/model/errormodels.py:15-17)."""
from binf_b200.pdf import AbstractBinfPDF


class AbstractErrorModel(AbstractBinfPDF):
    pass
