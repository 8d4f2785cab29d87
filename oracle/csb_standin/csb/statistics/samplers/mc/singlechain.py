from csb.statistics.samplers.mc import AbstractMC


class AbstractSingleChainMC(AbstractMC):
    pass
