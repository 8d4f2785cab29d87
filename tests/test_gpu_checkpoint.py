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


def _make_device(seed=7):
    """the same Gibbs sampler with the chains resident in HBM (CUDA-tensor state)"""
    import torch
    from binf_b200.example.misc import make_posterior
    from binf_b200.example.samplers import make_sampler
    from binf_b200.samplers import BinfState
    rng = np.random.RandomState(0)
    xs = np.linspace(-2, 2, 60)
    ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2.0, -4.0, 1.0, 1.5]), 1 / np.sqrt(2.5))
    post = make_posterior(xs, ys, np.polynomial.polynomial.polyval)
    C = 200
    dev = torch.device("cuda")
    q = torch.as_tensor((np.array([2.0, -4.0, 1.0, 1.5]) + 0.05 * rng.normal(size=(C, 4))).astype(np.float32), device=dev)
    start = BinfState(dict(coefficients=q, precision=torch.full((C,), 2.0, device=dev)))
    return make_sampler(post, 0.01, start, nsteps=8, timestep_adaption_limit=6, seed=seed)


def test_resume_of_a_device_resident_run(gpu, tmp_path):
    """the restored acceptance counters must be of the kind the device path adds to
    (an int64 CUDA tensor), otherwise the first sample() after load() raises"""
    import torch
    from binf_b200 import checkpoint
    a = _make_device()
    for _ in range(5):
        a.sample()
    path = str(tmp_path / "dev.npz")
    checkpoint.save(path, a)
    ref = [a.sample() for _ in range(3)][-1]
    b = _make_device()
    checkpoint.load(path, b)
    out = [b.sample() for _ in range(3)][-1]
    torch.cuda.synchronize()
    for name in ("coefficients", "precision"):
        assert torch.equal(out.variables[name], ref.variables[name])
    sa, sb = a.subsamplers["coefficients"], b.subsamplers["coefficients"]
    assert torch.equal(sa.n_accepted, sb.n_accepted) and sa.acceptance_rate == sb.acceptance_rate
    np.testing.assert_array_equal(sa.timestep, sb.timestep)


def test_device_entry_points_reject_bad_tensors(gpu):
    """dtype / contiguity / device / shape are checked before a raw pointer reaches a kernel"""
    import torch
    from binf_b200 import _cabi
    rng = np.random.RandomState(0)
    xs = np.linspace(-2, 2, 50)
    ys = rng.normal(size=50)
    m = _cabi.Model.polynomial(xs, ys, 4)
    dev = torch.device("cuda")
    C = 16
    q = torch.zeros(C, 4, device=dev)
    tau, eps = torch.ones(C, device=dev), torch.full((C,), 0.01, device=dev)
    opts = _cabi.HmcOpts(3, 1, 0, 0, 1.05, 0.95, 1, 0, 0)
    m.hmc_run_device(q, tau, eps, opts)                       # the well-formed call goes through
    for bad_q in (q.double(), torch.zeros(C, 8, device=dev)[:, ::2], torch.zeros(4, device=dev), q.cpu(),
                  torch.zeros(C, 5, device=dev)):
        with pytest.raises(ValueError):
            m.hmc_run_device(bad_q, tau, eps, opts)
    with pytest.raises(ValueError):
        m.hmc_run_device(q, tau.double(), eps, opts)
    with pytest.raises(ValueError):
        m.hmc_run_device(q, tau, eps[:8], opts)
    with pytest.raises(ValueError):
        m.logprob_grad_device(q, tau, logp=torch.zeros(C, device=dev))          # logp must be float64
    with pytest.raises(ValueError):
        m.logprob_grad_device(q, tau, grad=torch.zeros(C, 4, device=dev).t().contiguous().t())
    sink = _cabi.Sink(C, 4, capacity=2)
    with pytest.raises(ValueError):
        sink.push(q.double())
    torch.cuda.synchronize()
