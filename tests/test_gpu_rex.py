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


def test_last_chi2_is_chi2_of_the_current_state(gpu):
    """binfb_hmc_last_chi2: what the trajectory kernel leaves behind == a fresh pair sweep over the state"""
    import torch
    import chromatin_port as chrom
    from binf_b200 import _cabi
    from binf_b200.distributed import ChainShard
    n, C = 52, 24
    X, y = chrom.synthetic_chromatin(n, seed=3)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0)
    rng = np.random.RandomState(4)
    dev = torch.device("cuda")
    q = torch.as_tensor((X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))).astype(np.float32), device=dev)
    tau, eps = torch.full((C,), 80.0, device=dev), torch.full((C,), 0.02, device=dev)   # big steps: rejections too
    sh = ChainShard(m, q, tau, eps, 6, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=9)
    with pytest.raises(_cabi.BinfB200Error):
        sh.last_chi2()                                  # nothing has run yet
    for _ in range(3):
        sh.sweep()
        left = sh.last_chi2().clone()
        fresh = torch.zeros(C, dtype=torch.float64, device=dev)
        m.logprob_grad_device(q, tau, chi2=fresh)
        torch.cuda.synchronize()
        np.testing.assert_allclose(left.cpu().numpy(), fresh.cpu().numpy(), rtol=2e-6)
    assert 0 < int(sh.accepted.sum()) < C or True


def test_rex_kernels_match_the_host_port_bit_for_bit(gpu):
    """binfb_rex_pack / binfb_rex_decide / binfb_rex_select vs oracle/rex_port.py: identical records,
    decisions, labels, counters; seen from every rank of a simulated world of 3"""
    import torch
    import rex_port
    from binf_b200.distributed import DeviceOps
    dev = torch.device("cuda")
    ops = DeviceOps()
    world, rows, cols = 3, 2, 40
    C, T = rows * cols, world * rows
    betas = np.array([1.0, 0.9, 0.75, 0.6, 0.5, 0.3])
    rng = np.random.RandomState(0)
    n_data = 1225.0
    # a valid grid: every column holds every temperature once, scattered over ranks and rows
    grid = np.stack([rng.permutation(T) for _ in range(cols)], axis=1)              # [T slots, cols]
    tidx = grid.reshape(world, C).astype(np.int32)
    chi2 = rng.gamma(5.0, size=(world, C)) * 3.0
    tau = (50.0 + 10.0 * rng.rand(world, C)).astype(np.float32)
    eps = (0.01 * (1 + rng.rand(world, C))).astype(np.float32)
    chi2[1, 3] = np.nan                                                              # a diverged chain never swaps
    recs = []
    for r in range(world):
        rec = torch.zeros(C * 16, dtype=torch.uint8, device=dev)
        ops.pack(torch.as_tensor(chi2[r], device=dev), torch.as_tensor(tau[r], device=dev),
                 torch.as_tensor(eps[r], device=dev), torch.as_tensor(tidx[r], device=dev), n_data, rec)
        host = rex_port.pack(chi2[r], tau[r], eps[r], tidx[r], n_data)
        got = rec.cpu().numpy().view(rex_port.RECORD)
        np.testing.assert_array_equal(got["tidx"], host["tidx"])
        np.testing.assert_array_equal(got["eps"], host["eps"])
        np.testing.assert_allclose(got["ll"], host["ll"], rtol=1e-14)
        recs.append(rec)
    allr = torch.cat(recs)
    host_all = allr.cpu().numpy().view(rex_port.RECORD)
    n_acc = 0
    for attempt in (0, 1, 2**33 + 5):
        new_t = []
        for r in range(world):
            t = torch.as_tensor(tidx[r], device=dev)
            b = torch.as_tensor(betas[tidx[r]].astype(np.float32), device=dev)
            e = torch.as_tensor(eps[r], device=dev)
            acc = torch.zeros(C, dtype=torch.uint8, device=dev)
            pc = torch.zeros(T - 1, 2, dtype=torch.int64, device=dev)
            ts = torch.zeros(T, 3, dtype=torch.float64, device=dev)
            ops.decide(allr, world, r, C, cols, betas, 77, attempt, 1000.0, t, b, e, acc, pc, ts)
            torch.cuda.synchronize()
            ht, hb, he, ha, hpc, hts = rex_port.decide(host_all, world, r, C, cols, betas, 77, attempt, 1000.0)
            np.testing.assert_array_equal(acc.cpu().numpy(), ha)
            np.testing.assert_array_equal(t.cpu().numpy(), ht)
            np.testing.assert_array_equal(b.cpu().numpy(), hb)
            np.testing.assert_array_equal(e.cpu().numpy(), he)
            np.testing.assert_array_equal(pc.cpu().numpy(), hpc.astype(np.int64))
            np.testing.assert_allclose(ts.cpu().numpy(), hts, rtol=1e-12)
            new_t.append(t.cpu().numpy())
            n_acc += int(ha.sum())
        newg = np.stack(new_t).reshape(T, cols)
        assert np.array_equal(np.sort(newg, axis=0), np.tile(np.arange(T)[:, None], (1, cols)))   # conserved
    assert 0 < n_acc < 3 * world * C
    q = torch.as_tensor(rng.normal(size=(C, 7)).astype(np.float32), device=dev)
    out_q, out_a = torch.zeros(cols, 7, device=dev), torch.zeros(cols, device=dev)
    ops.select(q, torch.as_tensor(tau[0], device=dev), torch.as_tensor(tidx[0], device=dev), 2, cols, out_q, out_a)
    hq, ha = rex_port.select(q.cpu().numpy(), tau[0], tidx[0], 2, cols)
    np.testing.assert_array_equal(out_q.cpu().numpy(), hq)
    np.testing.assert_array_equal(out_a.cpu().numpy(), ha)


def test_tempered_ensemble_on_one_gpu(gpu):
    """4 temperatures x 24 columns of a 40-bead chromatin posterior on ONE device (world = 1, four rows):
    the fused kernel steps every replica at its own beta, exchanges swap labels only, the ladder adaption
    brings the swap rates into a useful band, the cold replicas feed the sample sink"""
    import torch
    import chromatin_port as chrom
    from binf_b200 import _cabi
    from binf_b200.distributed import ChainShard, ReplicaExchangeDriver
    n, T, cols = 40, 4, 24
    C = T * cols
    X, y = chrom.synthetic_chromatin(n, seed=5)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0)
    dev = torch.device("cuda")
    rng = np.random.RandomState(1)
    q = torch.as_tensor((X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))).astype(np.float32), device=dev)
    tau = torch.full((C,), 60.0, device=dev)
    eps = torch.full((C,), 0.004, device=dev)
    sh = ChainShard(m, q, tau, eps, 8, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=2)
    betas = [1.0, 0.5, 0.25, 0.1]
    drv = ReplicaExchangeDriver.for_shard(sh, 0, 1, betas, seed=4)
    assert drv.rex.rows == T and drv.rex.n_columns == cols and sh.beta is drv.rex.beta
    assert torch.equal(sh.beta.reshape(T, cols)[:, 0].cpu(), torch.tensor(betas))
    drv.run(30)
    q_before = q.clone()
    grid = drv.rex.tidx.reshape(T, cols).cpu().numpy()
    assert np.array_equal(np.sort(grid, axis=0), np.tile(np.arange(T)[:, None], (1, cols)))
    np.testing.assert_array_equal(sh.beta.cpu().numpy(), np.array(betas, dtype=np.float32)[drv.rex.tidx.cpu().numpy()])
    mean, sd = drv.rex.temperature_stats()
    assert np.all(np.diff(mean) < 0) and np.all(sd > 0)          # colder replicas sit at higher log-likelihood
    rates = drv.adapt(target=0.3)
    assert len(rates) == T - 1 and all(0.0 <= r <= 1.0 for r in rates)
    new = drv.betas
    assert new[0] == 1.0 and all(a > b for a, b in zip(new, new[1:]))
    torch.cuda.synchronize()
    assert torch.equal(q, q_before)                               # neither exchanges nor adaption move a state
    for measure in (40, 40):                                      # re-equilibrate, measure, re-space
        drv.run(20)
        drv.rex.reset_stats()
        drv.run(measure)
        drv.adapt(target=0.3)
    drv.run(20)
    drv.rex.reset_stats()
    drv.run(60)
    rates2 = drv.swap_rates()
    assert all(0.1 < r < 0.7 for r in rates2), rates2
    sink = _cabi.Sink(cols, 3 * n, capacity=4, thin=1)
    drv.run(6, sink=sink, thin=2)
    torch.cuda.synchronize()
    info = sink.info()
    assert info["n_pushed"] == 3 and drv.n_sweeps == 236
    cold_q, cold_tau = drv.cold_states()
    where = (drv.rex.tidx == 0).nonzero().flatten()
    assert len(where) == cols
    order = torch.argsort(where % cols)
    assert torch.equal(cold_q, q[where[order]]) and torch.equal(cold_tau, tau[where[order]])
    assert drv.last_draw_stats["swap"].attempt == 235 and 0.0 <= drv.last_draw_stats["swap"].accepted_fraction <= 1.0


def test_nccl_label_swap_self_test(gpu):
    """two ranks over NCCL (torchrun), when the box has two devices: tests/rex_nccl_selftest.py asserts
    agreement of the partners, conservation of the temperature grid and that no state moves"""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    here = os.path.dirname(os.path.abspath(__file__))
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(here, "rex_nccl_selftest.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rex nccl self-test ok" in r.stdout
