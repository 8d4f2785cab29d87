"""Multi-rank host logic on CPU (gloo, world sizes 2 and 3): chain sharding, the diagnostics
all-reduce and the replica-exchange neighbour-swap protocol (pairing, agreement of both partners
on every decision, conservation of the states).  The decision / apply hooks are host stand-ins for
the C-ABI device kernels (binfb_swap_decide / binfb_swap_apply), which the GPU tests cover."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from binf_b200.distributed import (ReplicaExchange, allreduce_stats, shard_range, swap_partner)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def host_decide(ll_mine, ll_theirs, beta_mine, beta_theirs, i_am_low, seed, attempt, pair_id, chain_base):
    a, b = (ll_mine, ll_theirs) if i_am_low else (ll_theirs, ll_mine)
    ba, bb = (beta_mine, beta_theirs) if i_am_low else (beta_theirs, beta_mine)
    u = np.random.RandomState([seed, attempt, pair_id, chain_base]).uniform(size=len(a))
    delta = (ba - bb) * (a.numpy() - b.numpy())
    return torch.from_numpy((u < np.exp(np.clip(-delta, -308, 709))).astype(np.uint8))


def host_apply(q_mine, q_theirs, mask):
    q_mine.copy_(torch.where(mask.bool()[:, None], q_theirs, q_mine))


def _worker(rank, world, port, out_dir):
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
        # --- replica exchange ------------------------------------------------------------------
        betas = [1.0 / (1 + r) for r in range(world)]
        rex = ReplicaExchange(rank, world, betas[rank], seed=11, decide=host_decide, apply=host_apply)
        rng = np.random.RandomState(100 + rank)
        q = torch.full((C, D), float(rank)) + torch.arange(C, dtype=torch.float32)[:, None] * 1e-3
        tau = torch.full((C,), 10.0 + rank)
        history = []
        for attempt in range(4):
            ll = torch.from_numpy(rng.normal(size=C) * 3.0)
            before = q.clone()
            mask = rex.swap(q, tau, ll, betas)
            partner = swap_partner(rank, world, attempt)
            if partner is None:
                assert mask is None and torch.equal(q, before)
            else:
                m = mask.bool()
                assert torch.equal(q[~m], before[~m])
                # structure and precision travel together: tag of the state == tag of its tau
                assert torch.equal(torch.floor(q[:, 0] + 1e-6), tau - 10.0)
            history.append(None if mask is None else mask.clone())
        # both partners reached the same decisions; every state is still held exactly once
        gathered = [None] * world
        dist.all_gather_object(gathered, (rank, [None if h is None else h.tolist() for h in history],
                                          q[:, 0].tolist(), tau.tolist()))
        if rank == 0:
            for attempt in range(4):
                for r in range(world):
                    p = swap_partner(r, world, attempt)
                    if p is not None:
                        assert gathered[r][1][attempt] == gathered[p][1][attempt]
                        assert any(gathered[r][1][attempt]) or True
            owners = np.array([g[2] for g in gathered])            # [world, C]
            tags = np.sort(np.floor(owners + 1e-6), axis=0)        # per chain slot: which replicas' states
            assert np.array_equal(tags, np.tile(np.arange(world)[:, None], (1, C)))
            assert sum(1 for h in gathered[0][1] if h is not None and any(h)) >= 1
            with open(os.path.join(out_dir, "ok"), "w") as fh:
                fh.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharding_stats_and_replica_exchange(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


def test_pairing_and_sharding_arithmetic():
    assert [swap_partner(r, 4, 0) for r in range(4)] == [1, 0, 3, 2]
    assert [swap_partner(r, 4, 1) for r in range(4)] == [None, 2, 1, None]
    assert [swap_partner(r, 3, 0) for r in range(3)] == [1, 0, None]
    for n, w in [(4096, 8), (1000, 3), (5, 8)]:
        rs = [shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1


# ------------------------------------------------------------------------------------------------
# full replica-exchange driver: statistics and ladder adaption
# ------------------------------------------------------------------------------------------------
def _driver_worker(rank, world, port, out_dir):
    from binf_b200.distributed import ReplicaExchangeDriver
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # toy tempered target: log L(x) = -0.5 * 40 * |x|^2 in D = 2, exact tempered draws N(0, 1/(40 beta))
        C, D, k = 256, 2, 40.0
        betas = [1.0, 0.7, 0.1][:world] if world == 3 else [1.0, 0.5]
        rng = np.random.RandomState(7 + rank)
        q = torch.zeros(C, D, dtype=torch.float32)
        tau = torch.full((C,), float(rank))
        state = {"beta": betas[rank]}

        def sweep():
            q.copy_(torch.from_numpy(rng.normal(size=(C, D)) / np.sqrt(k * state["beta"])).float())

        def log_likelihood():
            return -0.5 * k * (q.double() ** 2).sum(dim=1)

        drv = ReplicaExchangeDriver(rank, world, betas, q, tau, sweep, log_likelihood,
                                    set_beta=lambda b: state.update(beta=b), seed=3,
                                    decide=host_decide, apply=host_apply)
        drv.run(40)
        assert drv.n_sweeps == 40 and drv.rex.attempt == 40
        assert drv.last_draw_stats["swap"].attempt == 39
        rates = drv.swap_rates()                       # identical on every rank
        gathered = [None] * world
        dist.all_gather_object(gathered, rates)
        assert all(g == gathered[0] for g in gathered)
        assert len(rates) == world - 1 and all(0.0 < r <= 1.0 for r in rates)
        if world == 3:
            # the wide gap (0.7 -> 0.1) swaps less often than the narrow one (1.0 -> 0.7)
            assert rates[1] < rates[0]
            old = list(drv.betas)
            drv.adapt(gain=2.0)
            new = drv.betas
            assert new[0] == old[0] and new[-1] == old[-1] and old[1] > new[1] > new[2]
            assert state["beta"] == new[rank] and drv.rex.beta == new[rank]
            assert drv.pair_attempted == 0
            drv.run(40)
            rates2 = drv.swap_rates()
            assert abs(rates2[0] - rates2[1]) < abs(rates[0] - rates[1])    # more even after adaption
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_replica_exchange_driver(tmp_path, world):
    mp.spawn(_driver_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_adapt_ladder_is_pure_and_keeps_end_points():
    from binf_b200.distributed import adapt_ladder
    betas = np.geomspace(1.0, 0.05, 6)
    even = adapt_ladder(betas, [0.3] * 5)
    np.testing.assert_allclose(even, betas, rtol=1e-12)          # equal rates: nothing moves
    new = adapt_ladder(betas, [0.9, 0.9, 0.1, 0.1, 0.5])
    assert new[0] == betas[0] and new[-1] == betas[-1] and np.all(np.diff(new) < 0)
    gaps_old, gaps_new = -np.diff(np.log(betas)), -np.diff(np.log(new))
    assert gaps_new[0] > gaps_old[0] and gaps_new[2] < gaps_old[2]
    np.testing.assert_allclose(gaps_new.sum(), gaps_old.sum(), rtol=1e-12)
    assert list(adapt_ladder([1.0, 0.5], [0.2])) == [1.0, 0.5]


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
