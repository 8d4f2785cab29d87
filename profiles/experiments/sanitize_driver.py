"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).
    compute-sanitizer --tool racecheck python profiles/experiments/sanitize_driver.py
Shapes are tiny on purpose (the tools slow kernels down by 10-100x) but cover: the chromatin kernel
with 1, 2 and 4 roles per chain (free-running and LOCKSTEP), the small-batch alternative plan, the
polynomial kernel, the NVRTC generic kernel, the sink, RWMC and the predictive density."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from binf_b200 import _cabi  # noqa: E402


def contacts(n, seed):
    rng = np.random.RandomState(seed)
    X = np.cumsum(rng.normal(size=(n, 3)), axis=0)
    iu = np.triu_indices(n, 1)
    d = np.sqrt(((X[iu[0]] - X[iu[1]]) ** 2).sum(-1))
    y = 1.0 / (1.0 + np.exp(2.0 * (d - 2.5))) + 0.05 * rng.normal(size=len(d))
    return X, y.astype(np.float32)


def main():
    rng = np.random.RandomState(0)
    for n, roles, C in [(64, 0, 3), (300, 0, 2), (500, 2, 2), (1000, 4, 1), (1000, 0, 20)]:
        X, y = contacts(n, n)
        m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0, roles=roles)
        q = X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))
        m.logprob_grad(q, 50.0)
        r = m.hmc_run(q, 50.0, 0.002, 2, n_traj=2, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=1)
        assert np.all(np.isfinite(r["q"]))
        print("chromatin n=%d roles=%d chains=%d ok" % (n, roles, C), flush=True)
        m.close()
    xs = np.linspace(-2, 2, 100)
    ys = np.polynomial.polynomial.polyval(xs, [2.0, -4.0, 1.0, 1.5]) + 0.3 * rng.normal(size=100)
    pm = _cabi.Model.polynomial(xs, ys, 4, np.zeros(4), 5 * np.ones(4))
    c0 = np.ones((70, 4)) + 0.1 * rng.normal(size=(70, 4))
    pm.hmc_run(c0, 2.5, 0.01, 5, n_traj=2, gibbs_mode=_cabi.GIBBS_TAU_LAST, seed=2)
    pm.rwmc_run(c0, 2.5, 0.05, n_moves=3, seed=3)
    print("polynomial + rwmc ok", flush=True)
    code = """__device__ float binfb_mock(const float *t, const float *x, float *d) {
        float v = t[GEN_K - 1], pw = 1.0f;
        for (int k = GEN_K - 2; k >= 0; --k) v = fmaf(v, x[0], t[k]);
        for (int k = 0; k < GEN_K; ++k) { d[k] = pw; pw *= x[0]; }
        return v; }"""
    gm = _cabi.Model.generic(code, 4, xs, ys, np.zeros(4), 5 * np.ones(4))
    gm.hmc_run(c0, 2.5, 0.01, 5, n_traj=2, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=2)
    print("generic ok", flush=True)
    sink = _cabi.Sink(33, 12, capacity=3, burn_in=1, thin=2, track_map=True)
    sink2 = _cabi.Sink(5, 7, capacity=2, track_map=True)
    for t in range(6):
        sink.push(rng.normal(size=(33, 12)).astype(np.float32), rng.uniform(size=33), rng.normal(size=33))
        sink2.push(rng.normal(size=(5, 7)).astype(np.float32), None, rng.normal(size=5))
    sink.summary(), sink.read(), sink.map_estimate(), sink2.moments()
    _cabi.posterior_predictive(c0, np.full(70, 2.5), xs[:5], ys[:5])
    print("sink + predictive ok", flush=True)


if __name__ == "__main__":
    main()
