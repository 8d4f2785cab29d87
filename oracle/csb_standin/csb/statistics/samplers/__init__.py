"""csb.statistics.samplers stand-in: the State attribute bag."""


class State(object):
    def __init__(self, position, momentum=None):
        self.position = position
        self.momentum = momentum
