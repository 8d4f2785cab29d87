"""Multi-rank GPU self-test of the label-swap replica exchange over NCCL (run under torchrun; launched by
tests/test_gpu_rex.py::test_nccl_label_swap_self_test when two devices are visible, and by hand with
`gpurun --gpus N`).  One temperature per rank to begin with; checks after every attempt that the partners
agreed, that every column still holds every temperature exactly once, that no state moved, that beta and eps
follow the labels; then that the ladder adaption brings the swap rates into (0.1, 0.7) and that the cold
replicas can be assembled on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import chromatin_port as chrom
    from binf_b200 import _cabi
    from binf_b200.distributed import ChainShard, ReplicaExchangeDriver, init_from_env
    rank, world, local = init_from_env("nccl")
    dev = torch.device("cuda", local)
    n, C = 48, 96
    X, y = chrom.synthetic_chromatin(n, seed=5)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0, device=local)
    rng = np.random.RandomState(10 + rank)
    q = torch.as_tensor((X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))).astype(np.float32), device=dev)
    tau = torch.full((C,), 60.0, device=dev)
    betas = [float(b) for b in np.geomspace(1.0, 0.3, world)]
    eps = torch.full((C,), 0.004 / np.sqrt(betas[rank]), device=dev, dtype=torch.float32)
    # distinct HMC chain bases per rank (independent momentum streams): the swap draw must not depend on them
    sh = ChainShard(m, q, tau, eps, 8, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=2, chain_base=rank * C)
    drv = ReplicaExchangeDriver.for_shard(sh, rank, world, betas, seed=4)

    def gathered(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return torch.stack(out).cpu().numpy()

    for k in range(40):
        drv.replica.sweep()
        drv.n_sweeps += 1
        q_before, t_before, e_before = q.clone(), drv.rex.tidx.clone(), eps.clone()
        acc = drv.rex.swap(sh.last_chi2(), sh.tau, sh.eps, sh.n_data)
        torch.cuda.synchronize()
        assert torch.equal(q, q_before)
        allt, alla, alle = gathered(drv.rex.tidx), gathered(acc), gathered(eps)
        allt0, alle0 = gathered(t_before), gathered(e_before)
        assert np.array_equal(np.sort(allt, axis=0), np.tile(np.arange(world)[:, None], (1, C))), "grid not conserved"
        for c in range(C):
            for r in range(world):
                if alla[r, c]:
                    r2 = int(np.where(allt0[:, c] == allt[r, c])[0][0])     # who held my new temperature before
                    assert alla[r2, c] and allt[r2, c] == allt0[r, c], "partners disagree"
                    assert alle[r, c] == alle0[r2, c], "step size did not travel with the label"
                else:
                    assert allt[r, c] == allt0[r, c]
        np.testing.assert_array_equal(sh.beta.cpu().numpy(),
                                      np.array(drv.betas, dtype=np.float32)[drv.rex.tidx.cpu().numpy()])
    drv.adapt(target=0.3)
    for measure in (40, 40):                                      # re-equilibrate, measure, re-space
        drv.run(20)
        drv.rex.reset_stats()
        drv.run(measure)
        drv.adapt(target=0.3)
    drv.run(20)
    drv.rex.reset_stats()
    drv.run(80)
    rates = drv.swap_rates()
    assert all(0.1 < r < 0.7 for r in rates), rates
    cold_q, cold_tau = drv.cold_states(dst=0)
    allq, allt = gathered(q), gathered(drv.rex.tidx)
    if rank == 0:
        want = np.stack([allq[np.argmin(allt[:, c]), c] for c in range(C)])
        np.testing.assert_array_equal(cold_q.cpu().numpy(), want)
        print("rex nccl self-test ok: world %d, swap rates %s, ladder %s" % (
            world, ["%.2f" % r for r in rates], ["%.3f" % b for b in drv.betas]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
