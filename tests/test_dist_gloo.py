"""Multi-rank host logic on CPU (gloo, world sizes 2 and 3): chain sharding, the diagnostics
all-reduce and the replica-exchange neighbour-swap protocol (pairing, agreement of both partners
on every decision, conservation of the states).  The kernels' host restatement (oracle/rex_port.py)
stands in for the C-ABI device kernels (binfb_rex_pack / binfb_rex_decide / binfb_rex_select), which the GPU
tests compare with it bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from binf_b200.distributed import allreduce_stats, shard_range, swap_partner


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class ToyReplica(object):
    """exact tempered draws of log L(x) = -|x|^2 / 2 in D dimensions (chi^2 = |x|^2, tau = 1, no data term):
    x ~ N(0, 1/beta) per replica, so mean log L = -D / (2 beta) -- the equipartition law the ladder
    adaption assumes between its measurements"""

    def __init__(self, C, D, seed):
        self.C, self.D, self.rng = C, D, np.random.RandomState(seed)
        self.q = torch.zeros(C, D, dtype=torch.float32)
        self.tau = torch.ones(C)
        self.eps = torch.full((C,), 0.1)
        self.beta = None            # set by the driver: the exchange's per-chain tensor
        self.n_data = 0.0

    def sweep(self):
        self.q.copy_(torch.from_numpy(self.rng.normal(size=(self.C, self.D))).float() / torch.sqrt(self.beta)[:, None])

    def last_chi2(self):
        return (self.q.double() ** 2).sum(dim=1)


def _gather(obj, world):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def _worker(rank, world, port, out_dir):
    import rex_port
    from binf_b200.distributed import ReplicaExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        C, D = 64, 6
        # --- sharding + stats reduction ------------------------------------------------------
        lo, hi = shard_range(1000, rank, world)
        stats = torch.tensor([hi - lo, 1.0, float(rank), 0.5], dtype=torch.float64)
        allreduce_stats(stats)
        assert stats[0].item() == 1000 and stats[1].item() == world
        assert stats[2].item() == sum(range(world))
        # --- label-swap replica exchange: one temperature per rank to begin with -------------------
        betas = [1.0 / (1 + r) for r in range(world)]
        rex = ReplicaExchange(rank, world, betas, C, seed=11, ops=rex_port.HostOps())
        assert rex.n_columns == C and rex.rows == 1 and torch.all(rex.tidx == rank)
        rng = np.random.RandomState(100 + rank)
        q = torch.full((C, D), float(rank)) + torch.arange(C, dtype=torch.float32)[:, None] * 1e-3
        tau = torch.ones(C)
        eps = torch.full((C,), 0.1 * (1 + rank))              # tuned per temperature: travels with the label
        q0 = q.clone()
        for attempt in range(6):
            chi2 = torch.from_numpy(rng.gamma(3.0, size=C) * 4.0)
            before = rex.tidx.clone()
            acc = rex.swap(chi2, tau, eps, 0.0).clone()
            moved = rex.tidx != before
            assert torch.equal(moved, acc.bool())
            assert torch.all((rex.tidx - before).abs()[moved] == 1)
            assert torch.equal(q, q0)                                                  # no state ever moves
            np.testing.assert_allclose(eps.numpy(), 0.1 * (1 + rex.tidx.numpy()), rtol=1e-6)
            np.testing.assert_allclose(rex.beta.numpy(), np.array(betas, dtype=np.float32)[rex.tidx.numpy()])
            # every column still holds every temperature exactly once; partners agreed
            allt = np.array(_gather(rex.tidx.tolist(), world))                            # [world, C]
            assert np.array_equal(np.sort(allt, axis=0), np.tile(np.arange(world)[:, None], (1, C)))
            alla = np.array(_gather(acc.tolist(), world))
            assert alla.sum(axis=0).max() <= 2 * (world // 2) and np.all(alla.sum(axis=0) % 2 == 0)
        rates = rex.swap_rates()
        assert len(rates) == world - 1 and all(0.0 < r < 1.0 for r in rates)
        assert all(g == rates for g in _gather(rates, world))
        # the cold replicas, assembled from wherever they live
        cold, cold_eps = rex.select(q, eps, 0)
        allq = np.array(_gather(q.numpy(), world))
        allt = np.array(_gather(rex.tidx.tolist(), world))
        want = np.stack([allq[np.argmin(allt[:, c]), c] for c in range(C)])
        np.testing.assert_array_equal(cold.numpy(), want)
        np.testing.assert_allclose(cold_eps.numpy(), 0.1, rtol=1e-6)
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharding_stats_and_replica_exchange(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_pairing_and_sharding_arithmetic():
    assert [swap_partner(r, 4, 0) for r in range(4)] == [1, 0, 3, 2]
    assert [swap_partner(r, 4, 1) for r in range(4)] == [None, 2, 1, None]
    assert [swap_partner(r, 3, 0) for r in range(3)] == [1, 0, None]
    for n, w in [(4096, 8), (1000, 3), (5, 8)]:
        rs = [shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1


def test_partners_decide_alike_whatever_their_chain_base():
    """the swap uniform is keyed by (seed, attempt, lower temperature index, column) and by
    nothing rank-specific, so the two partners of a pair cannot disagree; several rows per rank, one process"""
    import rex_port
    from binf_b200.distributed import ReplicaExchange
    T, cols = 4, 16
    betas = [1.0, 0.8, 0.6, 0.4]
    rex = ReplicaExchange(0, 1, betas, T * cols, seed=5, ops=rex_port.HostOps())
    assert rex.rows == T and rex.n_columns == cols
    assert rex.tidx.tolist() == [r for r in range(T) for _ in range(cols)]
    rng = np.random.RandomState(2)
    tau, eps = torch.ones(T * cols), torch.rand(T * cols) + 0.5
    eps_by_temp0 = {}
    for attempt in range(8):
        chi2 = torch.from_numpy(rng.gamma(2.0, size=T * cols) * 3.0)
        before, eps_before = rex.tidx.clone(), eps.clone()
        acc = rex.swap(chi2, tau, eps, 0.0)
        grid_b, grid_a = before.reshape(T, cols).numpy(), rex.tidx.reshape(T, cols).numpy()
        assert np.array_equal(np.sort(grid_a, axis=0), np.tile(np.arange(T)[:, None], (1, cols)))
        accg = acc.reshape(T, cols).numpy().astype(bool)
        for c in range(cols):
            for r in range(T):
                if accg[r, c]:                      # its partner accepted too and they traded places
                    r2 = int(np.where(grid_b[:, c] == grid_a[r, c])[0][0])
                    assert accg[r2, c] and grid_a[r2, c] == grid_b[r, c]
                    assert eps[r * cols + c] == eps_before[r2 * cols + c]
    assert all(0.0 < r < 1.0 for r in rex.swap_rates())
    mean, sd = rex.temperature_stats()
    assert mean.shape == (T,) and np.all(sd > 0)


# ------------------------------------------------------------------------------------------------
# full replica-exchange driver: statistics and ladder adaption
# ------------------------------------------------------------------------------------------------
def _driver_worker(rank, world, port, out_dir):
    import rex_port
    from binf_b200.distributed import ReplicaExchangeDriver
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        C, D = 128, 300
        betas = [float(b) for b in np.geomspace(1.0, 0.05, world)]
        drv = ReplicaExchangeDriver(ToyReplica(C, D, 7 + rank), rank, world, betas, seed=3, ops=rex_port.HostOps())
        drv.run(12)
        assert drv.n_sweeps == 12 and drv.rex.attempt == 12 and drv.last_draw_stats["swap"].attempt == 11
        rates = drv.swap_rates()                       # identical on every rank
        assert all(g == rates for g in _gather(rates, world))
        # a geometric ladder down to 0.05 with 300 degrees of freedom never swaps ...
        assert len(rates) == world - 1 and max(rates) < 0.02
        mean, _ = drv.rex.temperature_stats()
        np.testing.assert_allclose(mean, -0.5 * D / np.array(betas), rtol=0.05)
        # ... and the ladder adaption recovers from all-zero rates: it is driven by the measured log-likelihoods
        old_eps = drv.replica.eps.clone()
        drv.adapt(target=0.3)
        new = drv.betas
        assert new[0] == 1.0 and all(a > b for a, b in zip(new, new[1:])) and new[-1] > 0.5
        np.testing.assert_allclose(drv.replica.beta.numpy(), np.array(new, dtype=np.float32)[drv.rex.tidx.numpy()])
        ratio = np.sqrt(np.array(betas) / np.array(new))[drv.rex.tidx.numpy()]
        np.testing.assert_allclose(drv.replica.eps.numpy(), old_eps.numpy() * ratio, rtol=1e-5)
        drv.run(60)
        rates2 = drv.swap_rates()
        assert all(0.15 < r < 0.5 for r in rates2), rates2
        # the cold replicas, wherever they live, follow the beta = 1 law: chi^2 ~ D
        cold, _ = drv.cold_states()
        assert cold.shape == (C, D) and abs(float((cold.double() ** 2).sum(dim=1).mean()) / D - 1.0) < 0.05
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_replica_exchange_driver(tmp_path, world):
    mp.spawn(_driver_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_adapt_ladder_is_pure_and_hits_its_target():
    from binf_b200.distributed import adapt_ladder, expected_swap_rate
    d = 3000.0                                       # equipartition: L(beta) = -d / (2 beta)
    betas = np.geomspace(1.0, 0.05, 8)
    L = -0.5 * d / betas
    new = adapt_ladder(betas, L, target=0.3)
    assert new[0] == 1.0 and np.all(np.diff(new) < 0) and len(new) == 8
    mu = (new[:-1] - new[1:]) * (-0.5 * d) * (1.0 / new[:-1] - 1.0 / new[1:])
    np.testing.assert_allclose([expected_swap_rate(m) for m in mu], 0.3, atol=1e-3)
    np.testing.assert_array_equal(new, adapt_ladder(betas, L, target=0.3))     # pure
    kept = adapt_ladder(betas, L, keep_ends=True)
    assert kept[0] == 1.0 and kept[-1] == betas[-1] and np.all(np.diff(kept) < 0)
    mu = (kept[:-1] - kept[1:]) * (-0.5 * d) * (1.0 / kept[:-1] - 1.0 / kept[1:])
    np.testing.assert_allclose(mu, mu[0], rtol=1e-3)                          # one common rate
    assert list(adapt_ladder([1.0], [3.0])) == [1.0]
    assert list(adapt_ladder([1.0, 0.5], [np.nan, 1.0])) == [1.0, 0.5]         # nothing measured: unchanged


def test_merge_sink_sums_equals_one_big_sink():
    """the cross-rank merge of per-GPU sink sums == the summary over all chains at once (numpy stand-in
    for Sink.sums(): same definitions as binfb_sink_sums_host)"""
    from binf_b200.distributed import merge_sink_sums
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sink_port
    rng = np.random.RandomState(5)
    n, D = 40, 6
    shards = [30.0 + rng.normal(size=(n, C, D)) + rng.normal(size=(1, C, 1)) for C in (7, 12, 5)]

    def sums(x):
        mean, m2 = x.mean(axis=0), x.var(axis=0, ddof=0) * x.shape[0]
        pivot = mean[0]
        dev = mean - pivot
        return dict(pivot=pivot, s1=dev.sum(0), s2=(dev * dev).sum(0), s3=m2.sum(0), n_chains=x.shape[1], n=n)

    merged = merge_sink_sums([sums(x) for x in shards])
    ref = sink_port.ListSink()
    for t in range(n):
        ref.append(np.concatenate([x[t] for x in shards], axis=0))
    r = ref.summary()
    assert merged["n_chains"] == 24
    for k in ("mean", "var", "rhat", "ess_per_chain"):
        np.testing.assert_allclose(merged[k], r[k], rtol=1e-9, err_msg=k)
    with pytest.raises(ValueError):
        merge_sink_sums([sums(shards[0]), dict(sums(shards[1]), n=n + 1)])
