"""csb.numeric stand-in: clipped exp/log and log-sum-exp."""
import numpy

EXP_MIN, EXP_MAX = -308.0, 709.0
LOG_MIN, LOG_MAX = 1e-308, 1e308


def exp(x, x_min=EXP_MIN, x_max=EXP_MAX):
    return numpy.exp(numpy.clip(x, x_min, x_max))


def log(x, x_min=LOG_MIN, x_max=LOG_MAX):
    return numpy.log(numpy.clip(x, x_min, x_max))


def log_sum_exp(x, axis=0):
    x = numpy.asarray(x)
    xmax = x.max(axis)
    return numpy.log(numpy.exp(x - xmax).sum(axis)) + xmax
