"""GPU: replica-exchange decision / apply kernels (binfb_swap_decide, binfb_swap_apply)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_swap_decide_and_apply(gpu):
    import torch
    from binf_b200.distributed import _device_decide, _device_apply
    C, D = 20000, 7
    dev = torch.device("cuda")
    ll_a = torch.zeros(C, dtype=torch.float64, device=dev)
    ll_b = torch.zeros(C, dtype=torch.float64, device=dev)
    # Delta = (beta_a - beta_b)(l_a - l_b) = 0.5 * 2 ln2 = ln 2  =>  accept with probability 1/2
    ll_a += 2.0 * np.log(2.0)
    m_low = _device_decide(ll_a, ll_b, 1.0, 0.5, True, 3, 9, 4, 0)
    m_high = _device_decide(ll_b, ll_a, 0.5, 1.0, False, 3, 9, 4, 0)      # the partner's view
    torch.cuda.synchronize()
    assert torch.equal(m_low, m_high)                                     # same decision on both sides
    assert abs(m_low.float().mean().item() - 0.5) < 0.02
    assert not torch.equal(m_low, _device_decide(ll_a, ll_b, 1.0, 0.5, True, 3, 10, 4, 0))  # new attempt
    assert _device_decide(ll_b, ll_a, 1.0, 0.5, True, 3, 9, 4, 0).all()   # Delta < 0: always accept
    ll_big = ll_a * 1000
    assert not _device_decide(ll_big, ll_b, 1.0, 0.5, True, 3, 9, 4, 0).any()
    nan = torch.full((C,), float("nan"), dtype=torch.float64, device=dev)
    assert not _device_decide(nan, ll_b, 1.0, 0.5, True, 3, 9, 4, 0).any()
    q_mine = torch.zeros(C, D, device=dev)
    q_theirs = torch.ones(C, D, device=dev)
    _device_apply(q_mine, q_theirs, m_low)
    torch.cuda.synchronize()
    assert torch.equal(q_mine[:, 0].bool(), m_low.bool()) and torch.equal(q_mine[:, 0], q_mine[:, D - 1])


def test_chain_shard_log_likelihood(gpu):
    import torch
    import chromatin_port as chrom
    from binf_b200 import _cabi
    from binf_b200.distributed import ChainShard
    n, C = 48, 10
    X, y = chrom.synthetic_chromatin(n, seed=2)
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0)
    rng = np.random.RandomState(0)
    q0 = (X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))).astype(np.float32)
    dev = torch.device("cuda")
    q = torch.as_tensor(q0, device=dev)
    tau = torch.full((C,), 80.0, device=dev)
    eps = torch.full((C,), 0.003, device=dev)
    beta = torch.full((C,), 0.5, device=dev)
    sh = ChainShard(m, q, tau, eps, 5, gibbs_mode=_cabi.GIBBS_TAU_FIRST, beta=beta, seed=1)
    ll = sh.log_likelihood().cpu().numpy()
    for c in range(C):
        assert ll[c] == pytest.approx(o.likelihood_log_prob(q0[c].astype(np.float64), 80.0), rel=1e-5)
    sh.sweep(n_traj=3)
    torch.cuda.synchronize()
    assert sh.stats[1].item() == 3 * C and not torch.equal(sh.q.cpu(), torch.as_tensor(q0))
    # tempered conjugate update: shape beta*M/2 + a - 1, rate beta*chi2/2 + b
    assert 5.0 < sh.tau.mean().item() < 2000.0


def test_driver_for_shard_with_sink(gpu):
    """single rank: the driver steps the fused kernel at its beta and feeds the sample sink"""
    import torch
    import chromatin_port as chrom
    from binf_b200 import _cabi
    from binf_b200.distributed import ChainShard, ReplicaExchangeDriver
    n, C = 40, 12
    X, y = chrom.synthetic_chromatin(n, seed=5)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0)
    dev = torch.device("cuda")
    rng = np.random.RandomState(1)
    q = torch.as_tensor((X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))).astype(np.float32), device=dev)
    tau = torch.full((C,), 60.0, device=dev)
    eps = torch.full((C,), 0.003, device=dev)
    sh = ChainShard(m, q, tau, eps, 4, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=2)
    drv = ReplicaExchangeDriver.for_shard(sh, 0, 1, [0.25])
    assert torch.all(sh.beta == 0.25)
    sink = _cabi.Sink(C, 3 * n, capacity=4, thin=2)
    drv.run(6, sink=sink)
    torch.cuda.synchronize()
    info = sink.info()
    assert info["n_pushed"] == 6 and info["n_kept"] == 3 and drv.n_sweeps == 6
    kept, aux = sink.read()
    assert kept.shape == (3, C, 3 * n) and not np.array_equal(kept[0], kept[-1])
    assert drv.swap_rates() == [] and drv.last_draw_stats["swap"] is None
    mean, _ = sink.moments()
    assert np.all(np.isfinite(mean)) and np.all(aux > 0)
