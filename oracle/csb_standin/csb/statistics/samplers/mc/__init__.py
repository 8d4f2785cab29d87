class AbstractMC(object):
    pass
