"""Sample sink for batched chains.

The reference's driver keeps every Gibbs sweep of ONE chain in a Python list and post-processes it
with slices (example_script.py:32-34,41-42; binf/example/misc.py:18-22):

    samples = []
    for i in range(30000):
        samples.append(deepcopy(gips.sample()))
    samples_thin = samples[20000::20]
    MAP_coeffs, _ = get_MAP(samples_thin, log_probs)

With thousands of chains resident in HBM that list would be 50 MB per sweep; `SampleSink` does the
same bookkeeping on the device in one fused HBM-bound pass per sweep (binfb_sink_push): burn-in and
thinning as in the slice, a ring of the kept samples, per-chain running moments, the per-chain MAP
candidate, and cross-chain R-hat / effective sample size from the moments.

    sink = SampleSink(n_chains, dim, capacity=500, burn_in=20000, thin=20, track_map=True)
    for i in range(30000):
        state = gips.sample()
        sink.append(state, variable="structure", aux="precision", log_prob=logp)
    samples_thin, precisions = sink.samples()
"""
import numpy as np

from binf_b200 import _cabi


class SampleSink(object):
    def __init__(self, n_chains, dim, capacity=0, burn_in=0, thin=1, track_map=False, device=0):
        self._sink = _cabi.Sink(n_chains, dim, capacity, burn_in, thin, track_map, device)
        self.n_chains, self.dim = n_chains, dim

    def __len__(self):
        """number of kept samples, like len(samples[burn_in::thin])"""
        return self._sink.info()["n_kept"]

    @property
    def n_sweeps(self):
        return self._sink.info()["n_pushed"]

    def append(self, state, variable=None, aux=None, log_prob=None, stream=None):
        """state: a BinfState (give `variable`, optionally the name `aux` of a per-chain scalar
        variable) or the [C, dim] array / CUDA tensor itself (then aux is the [C] array)."""
        if hasattr(state, "variables"):
            v = state.variables
            q = v[variable]
            a = v[aux] if isinstance(aux, str) else aux
        else:
            q, a = state, aux
        self._sink.push(q, a, log_prob, stream)

    def samples(self, first=None, count=None):
        return self._sink.read(first, count)

    def moments(self):
        """per-chain (mean, unbiased variance), each [C, dim]"""
        return self._sink.moments()

    def summary(self):
        """dict of per-dimension mean, var, rhat, ess_per_chain over all chains"""
        return self._sink.summary()

    def get_MAP(self):
        """per-chain (state, aux, log_prob) of the kept sample with maximum log-probability
        (binf/example/misc.py:18-22, applied to every chain)"""
        logp, q, aux = self._sink.map_estimate()
        return q, aux, logp

    def close(self):
        self._sink.close()
