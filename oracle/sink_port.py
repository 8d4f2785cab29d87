"""ORACLE (test infrastructure, not product code): the reference's sample bookkeeping restated for a
batch of chains, with plain Python lists and numpy slices exactly as the reference's driver does it.

Follows example_script.py:32-34 (`samples.append(deepcopy(gips.sample()))`), :41
(`samples_thin = samples[burn_in::thin]`), :42 (log-probs of the kept samples) and
binf/example/misc.py:18-22 (`get_MAP`: `samples[np.argmax(log_probs)]`).  The posterior summaries
(mean/variance, Gelman-Rubin R-hat, effective sample size from the spread of the chain means) are the
textbook formulas in float64; the reference itself only plots histograms of the kept samples
(binf/example/plots.py)."""
from copy import deepcopy

import numpy as np


class ListSink(object):
    def __init__(self, burn_in=0, thin=1, capacity=None):
        self.samples, self.aux, self.logp = [], [], []
        self.burn_in, self.thin, self.capacity = burn_in, thin, capacity

    def append(self, q, aux=None, logp=None):
        self.samples.append(deepcopy(np.asarray(q, dtype=np.float64)))   # example_script.py:33
        self.aux.append(None if aux is None else deepcopy(np.asarray(aux, dtype=np.float64)))
        self.logp.append(None if logp is None else deepcopy(np.asarray(logp, dtype=np.float64)))

    def thinned(self):
        """samples[burn_in::thin] (example_script.py:41), restricted to the last `capacity`"""
        sl = slice(self.burn_in, None, self.thin)
        q, a, l = self.samples[sl], self.aux[sl], self.logp[sl]
        if self.capacity is not None:
            q, a, l = q[-self.capacity:], a[-self.capacity:], l[-self.capacity:]
        return q, a, l

    def moments(self):
        x = np.array(self.samples[self.burn_in:])                      # [n, C, D]
        return x.mean(axis=0), x.var(axis=0, ddof=1)

    def summary(self):
        x = np.array(self.samples[self.burn_in:])
        n, C = x.shape[0], x.shape[1]
        chain_mean, chain_var = x.mean(axis=0), x.var(axis=0, ddof=1)
        W = chain_var.mean(axis=0)
        var_means = chain_mean.var(axis=0, ddof=1) if C > 1 else np.zeros(x.shape[2])   # B / n
        with np.errstate(divide="ignore", invalid="ignore"):
            rhat = np.sqrt(((n - 1.0) / n * W + var_means) / W)
            ess = W / var_means
        return dict(mean=chain_mean.mean(axis=0), var=W, rhat=rhat, ess_per_chain=ess)

    def get_MAP(self):
        """per chain: kept sample of maximum log-probability (misc.py:18-22; argmax = first maximum)"""
        sl = slice(self.burn_in, None, self.thin)
        q, a, l = np.array(self.samples[sl]), self.aux[sl], np.array(self.logp[sl])   # no capacity cut: running max
        idx = np.argmax(np.where(np.isnan(l), -np.inf, l), axis=0)                    # [C]
        C = q.shape[1]
        qm = q[idx, np.arange(C)]
        am = None if a[0] is None else np.array(a)[idx, np.arange(C)]
        return qm, am, l[idx, np.arange(C)]
