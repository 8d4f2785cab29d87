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
