"""Checkpoint / resume: the random streams are counter-based (Philox keyed by seed, chain, draw), so a
run restored from a checkpoint continues bit-for-bit like the uninterrupted one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(seed=7, rwmc=False):
    from binf_b200.example.misc import make_posterior
    from binf_b200.example.samplers import make_sampler
    from binf_b200.samplers import BinfState
    rng = np.random.RandomState(0)
    xs = np.linspace(-2, 2, 60)
    ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2.0, -4.0, 1.0, 1.5]), 1 / np.sqrt(2.5))
    post = make_posterior(xs, ys, np.polynomial.polynomial.polyval)
    C = 200
    start = BinfState(dict(coefficients=np.array([2.0, -4.0, 1.0, 1.5]) + 0.05 * rng.normal(size=(C, 4)),
                           precision=np.full(C, 2.0)))
    if rwmc:
        return make_sampler(post, 0.03, start, seed=seed)
    return make_sampler(post, 0.01, start, nsteps=8, timestep_adaption_limit=6, seed=seed)


@pytest.mark.parametrize("rwmc", [False, True])
def test_resumed_gibbs_run_is_bit_identical(gpu, tmp_path, rwmc):
    from binf_b200 import checkpoint
    a = _make(rwmc=rwmc)
    for _ in range(5):
        a.sample()
    path = str(tmp_path / "run.npz")
    checkpoint.save(path, a)
    ref = [a.sample() for _ in range(4)][-1]
    b = _make(rwmc=rwmc)                      # fresh construction, then restore
    checkpoint.load(path, b)
    out = [b.sample() for _ in range(4)][-1]
    for name in ("coefficients", "precision"):
        np.testing.assert_array_equal(out.variables[name], ref.variables[name])
    sa, sb = a.subsamplers["coefficients"], b.subsamplers["coefficients"]
    assert sa.acceptance_rate == sb.acceptance_rate
    if not rwmc:
        np.testing.assert_array_equal(sa.timestep, sb.timestep)      # adapted step sizes travelled too
        assert sa.counter == sb.counter == 9


def test_single_sampler_and_mismatch(gpu, tmp_path):
    from binf_b200 import checkpoint
    g = _make()
    hmc = g.subsamplers["coefficients"]
    hmc.sample()
    path = str(tmp_path / "hmc.npz")
    checkpoint.save(path, hmc)
    g2 = _make()
    h2 = checkpoint.load(path, g2.subsamplers["coefficients"])
    np.testing.assert_array_equal(h2.state, hmc.state)
    assert h2._draw == hmc._draw == 1
    with pytest.raises(ValueError):
        checkpoint.load(path, g2)             # a single-sampler checkpoint is not a Gibbs checkpoint
