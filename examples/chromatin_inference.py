"""Chromatin structure inference end to end on the B200 path: a bead-chain posterior with a logistic
contact forward model (binf_b200.chromatin, behind the reference's Posterior / Likelihood / prior
API), thousands of chains per GPU sampled by the fused Gibbs/HMC kernel, samples kept by the on-device
sink, chains sharded over the GPUs of one box.

    python examples/chromatin_inference.py --beads 200 --chains 512 --sweeps 60
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        examples/chromatin_inference.py --beads 1000 --chains 4096 --sweeps 200            # 8 x 4096 chains
    ... --tempered          # one inverse temperature per rank to begin with + replica exchange (label swaps)

Every rank owns a contiguous range of chains (Philox streams keyed by the global chain id); the only
exchanges are the diagnostics reductions at the end and, with --tempered, the 16-byte-per-chain all-gather of
an exchange attempt; with --tempered only the cold replicas (temperature index 0) are posterior samples: they
are assembled from wherever they live and recorded on rank 0.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synthetic(n, alpha, d_c, noise, seed, contact="logistic"):
    rng = np.random.RandomState(seed)
    X = np.cumsum(rng.normal(size=(n, 3)), axis=0)
    X -= X.mean(axis=0)
    i, j = np.triu_indices(n, 1)
    d = np.sqrt(np.sum((X[i] - X[j]) ** 2, axis=-1))
    if contact == "algebraic":
        z = alpha * (d_c - d)
        y = 0.5 * (1.0 + z / np.sqrt(1.0 + z * z)) + rng.normal(size=d.shape) * noise
    else:
        with np.errstate(over="ignore"):
            y = 1.0 / (1.0 + np.exp(alpha * (d - d_c))) + rng.normal(size=d.shape) * noise
    return X, y.astype(np.float32)


def pair_distances(Z):
    i, j = np.triu_indices(len(Z), 1)
    return np.sqrt(np.sum((Z[i] - Z[j]) ** 2, axis=-1))


def main(argv=None):
    import torch
    from binf_b200 import _cabi
    from binf_b200.chromatin import make_chromatin_posterior
    from binf_b200.distributed import (ChainShard, ReplicaExchangeDriver, allreduce_stats, init_from_env,
                                       sink_summary_all_ranks)
    from binf_b200.lowering import lower, set_device
    from binf_b200.samplers.sink import SampleSink

    ap = argparse.ArgumentParser()
    ap.add_argument("--beads", type=int, default=200)
    ap.add_argument("--chains", type=int, default=512, help="chains per GPU")
    ap.add_argument("--sweeps", type=int, default=60)
    ap.add_argument("--leapfrog", type=int, default=20)
    ap.add_argument("--timestep", type=float, default=2e-3)
    ap.add_argument("--excluded-volume", type=float, default=0.0, help="k_ev of the quartic repulsion (0 = off)")
    ap.add_argument("--tempered", action="store_true")
    ap.add_argument("--contact", default="logistic", choices=["logistic", "algebraic"],
                    help="contact function of the forward model: 1/(1+exp(-z)) or 1/2 (1 + z/sqrt(1+z^2))")
    args = ap.parse_args(argv)

    rank, world, local = init_from_env()
    set_device(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n, C = args.beads, args.chains
    X, y = synthetic(n, 2.0, 2.5, 0.05, seed=0, contact=args.contact)  # the same data on every rank
    posterior = make_chromatin_posterior(n, y, alpha=2.0, d_c=2.5, k_bb=4.0, l0=1.0,
                                         ev_k=args.excluded_volume, ev_d=1.5, contact=args.contact)
    model = lower(posterior.conditional_factory(precision=1.0)).model   # the lowered device model
    rng = np.random.RandomState(1 + rank)
    q = torch.as_tensor((X.reshape(-1)[None] + 0.3 * rng.normal(size=(C, 3 * n))).astype(np.float32), device=dev)
    tau = torch.full((C,), 50.0, device=dev)
    eps = torch.full((C,), args.timestep, device=dev)
    shard = ChainShard(model, q, tau, eps, args.leapfrog, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=7,
                       chain_base=rank * C)
    betas = [float(b) for b in np.geomspace(1.0, 0.8, world)] if args.tempered else [1.0] * world
    driver = ReplicaExchangeDriver.for_shard(shard, rank, world, betas, seed=11) if args.tempered else None
    burn = args.sweeps // 2
    sink = SampleSink(C, 3 * n, capacity=4, burn_in=burn, thin=max(1, (args.sweeps - burn) // 4), device=local)
    for sweep in range(args.sweeps):
        if driver is not None:
            driver.step()
            if sweep == burn // 2:
                driver.adapt(target=0.3)                                 # re-space the ladder once, early in burn-in
            cold_q, cold_tau = driver.cold_states(dst=0)                 # hot replicas are not posterior samples
            if rank == 0:
                sink.append(cold_q, aux=cold_tau)
        else:
            shard.sweep()
            sink.append(q, aux=tau)
    torch.cuda.synchronize()
    stats = allreduce_stats(shard.stats.clone())                         # accepted, proposed, sum eps, sum p_acc
    if driver is not None:                                               # the cold replicas live in rank 0's sink
        parts = [sink._sink.sums() if rank == 0 else None]
        if world > 1:
            torch.distributed.broadcast_object_list(parts, src=0)
        from binf_b200.distributed import merge_sink_sums
        summary = merge_sink_sums(parts)
    else:
        summary = sink_summary_all_ranks(sink)                           # over the chains of all ranks
    # distance RMSD of the posterior-mean structure to the truth, over the pairs the data constrain
    d_true = pair_distances(X)
    near = d_true < 4.0
    d_mean = pair_distances(summary["mean"].reshape(n, 3))
    drmsd = float(np.sqrt(np.mean((d_mean[near] - d_true[near]) ** 2)))
    out = dict(chains=summary["n_chains"], acceptance=float(stats[0] / stats[1]), precision=float(tau.mean()),
               max_rhat=float(np.nanmax(summary["rhat"])), contact_drmsd=drmsd,
               swap_rates=driver.swap_rates() if driver is not None else None)
    if rank == 0:
        print("chains %(chains)d  HMC acceptance %(acceptance).3f  precision %(precision).1f  "
              "max R-hat %(max_rhat).3f  dRMSD of contacting pairs %(contact_drmsd).3f  swaps %(swap_rates)s" % out)
    return out


if __name__ == "__main__":
    main()
