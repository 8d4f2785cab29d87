"""Chain sharding over the GPUs of one box, diagnostics reduction and replica exchange.

The reference is single-process and single-chain; chains never interact inside
HMCSampler / GibbsSampler (binf/samplers/hmc.py:136-164, binf/samplers/gibbs.py:136-151), so the
batch shards with NO data-path collective: rank r owns the contiguous chain range
[r*C/R, (r+1)*C/R), the data set is replicated, and the Philox streams are keyed by the global
chain id so results do not depend on the number of ranks.  One process per GPU (torchrun),
`torch.distributed` with the NCCL backend over NVLink for the only two exchanges there are:

  * `allreduce_stats`: a handful of float64 accumulators (accepted, proposed, sum eps, sum p_acc,
    optional moments) -- once per reporting interval, never per step;
  * `ReplicaExchange`: neighbour swaps between ranks r and r+1 holding inverse temperatures
    beta_r > beta_{r+1} (the reference only alludes to this: binf/samplers/hmc.py:171-177).
    Even/odd pairs alternate per attempt; partners exchange the untempered log-likelihoods and the
    states with batched isend/irecv (NCCL grouped send/recv), both evaluate
    u < exp(-(beta_a - beta_b)(l_a - l_b)) with the SAME Philox draw and keep or
    adopt the partner's (structure, precision) per chain.

Everything here is host-side plumbing; the numerics are C-ABI kernels (binfb_hmc_run,
binfb_logprob_grad, binfb_swap_decide, binfb_swap_apply).
"""
import os

import numpy as np


def shard_range(n_total, rank, world):
    """contiguous chain range [lo, hi) of `rank`; sizes differ by at most one"""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """(rank, world, local_rank); initialises torch.distributed when launched by torchrun"""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def allreduce_stats(stats, group=None):
    """SUM-reduce a small float64 tensor of sampler diagnostics over all ranks (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def swap_partner(rank, world, attempt):
    """Neighbour of `rank` in attempt `attempt` (even attempts pair (0,1),(2,3),..; odd attempts
    pair (1,2),(3,4),..), or None when the rank sits out."""
    if (rank + attempt) % 2 == 0:
        partner = rank + 1
    else:
        partner = rank - 1
    return partner if 0 <= partner < world else None


class ChainShard(object):
    """The chains of one rank, resident in HBM, stepped with the fused Gibbs/HMC kernel."""

    def __init__(self, model, q, tau, eps, n_steps, gibbs_mode=0, beta=None, seed=0, chain_base=0):
        import torch
        self.model = model
        self.q, self.tau, self.eps, self.beta = q, tau, eps, beta
        n = q.shape[0]
        dev = q.device
        self.n_steps, self.gibbs_mode, self.seed, self.chain_base = n_steps, gibbs_mode, seed, chain_base
        self.accepted = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.n_accepted = torch.zeros(n, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(4, dtype=torch.float64, device=dev)
        self.chi2 = torch.zeros(n, dtype=torch.float64, device=dev)
        self.draw = 0

    def sweep(self, n_traj=1, n_adapt=0):
        import torch
        from binf_b200 import _cabi
        opts = _cabi.HmcOpts(self.n_steps, n_traj, n_adapt, self.gibbs_mode, 1.05, 0.95, self.seed,
                             self.draw, self.chain_base)
        self.model.hmc_run_device(self.q, self.tau, self.eps, opts, beta=self.beta,
                                  accepted=self.accepted, n_accepted=self.n_accepted, stats=self.stats,
                                  stream=torch.cuda.current_stream().cuda_stream)
        self.draw += n_traj

    def log_likelihood(self):
        """untempered log L per chain at the current (structure, precision): one fused pass"""
        import torch
        self.model.logprob_grad_device(self.q, self.tau, chi2=self.chi2,
                                       stream=torch.cuda.current_stream().cuda_stream)
        t = self.tau.double()
        return -0.5 * t * self.chi2 + 0.5 * float(self.model.n_data) * torch.log(t)


def _device_decide(ll_mine, ll_theirs, beta_mine, beta_theirs, i_am_low, seed, attempt, pair_id,
                   chain_base):
    """accept mask from binfb_swap_decide; the lower rank of the pair is 'a' on both sides"""
    import torch
    from binf_b200 import _cabi
    a, b = (ll_mine, ll_theirs) if i_am_low else (ll_theirs, ll_mine)
    ba, bb = (beta_mine, beta_theirs) if i_am_low else (beta_theirs, beta_mine)
    mask = torch.zeros(ll_mine.shape[0], dtype=torch.uint8, device=ll_mine.device)
    _cabi.check(_cabi.lib().binfb_swap_decide(_cabi.ptr(a), _cabi.ptr(b), float(ba), float(bb),
                                              int(ll_mine.shape[0]), seed, attempt, pair_id, chain_base,
                                              _cabi.ptr(mask),
                                              _cabi.ptr(torch.cuda.current_stream().cuda_stream)))
    return mask


def _device_apply(q_mine, q_theirs, mask):
    import torch
    from binf_b200 import _cabi
    _cabi.check(_cabi.lib().binfb_swap_apply(_cabi.ptr(q_mine), _cabi.ptr(q_theirs), None, None,
                                             _cabi.ptr(mask), int(q_mine.shape[0]), int(q_mine.shape[1]),
                                             _cabi.ptr(torch.cuda.current_stream().cuda_stream)))


class ReplicaExchange(object):
    """Neighbour swaps between temperature-ordered ranks.

    `decide(ll_mine, ll_theirs, beta_mine, beta_theirs, i_am_low, seed, attempt, pair_id,
    chain_base) -> uint8 mask` and `apply(q_mine, q_theirs, mask)` default to the C-ABI device
    kernels; the CPU (gloo) tests of the protocol substitute host implementations."""

    def __init__(self, rank, world, beta, seed=0, chain_base=0, decide=None, apply=None, group=None):
        self.rank, self.world, self.beta = rank, world, float(beta)
        self.seed, self.chain_base, self.group = int(seed), int(chain_base), group
        self.decide = decide or _device_decide
        self.apply = apply or _device_apply
        self.attempt = 0
        self.n_attempted = 0
        self.n_swapped = 0

    def swap(self, q, tau, ll, betas):
        """One attempt.  q [C, D] and tau [C] are updated in place where the swap is accepted;
        `ll` [C] float64 are this rank's untempered log-likelihoods, `betas` the inverse
        temperature of every rank.  Returns the accept mask (or None if this rank sat out)."""
        import torch
        import torch.distributed as dist
        partner = swap_partner(self.rank, self.world, self.attempt)
        attempt = self.attempt
        self.attempt += 1
        if partner is None:
            return None
        ll_theirs, q_theirs, tau_theirs = torch.empty_like(ll), torch.empty_like(q), torch.empty_like(tau)
        ops = [dist.P2POp(dist.isend, ll, partner, self.group), dist.P2POp(dist.irecv, ll_theirs, partner, self.group),
               dist.P2POp(dist.isend, q, partner, self.group), dist.P2POp(dist.irecv, q_theirs, partner, self.group),
               dist.P2POp(dist.isend, tau, partner, self.group), dist.P2POp(dist.irecv, tau_theirs, partner, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        low = min(self.rank, partner)
        mask = self.decide(ll, ll_theirs, betas[self.rank], betas[partner], self.rank == low, self.seed,
                           attempt, low, self.chain_base)
        self.apply(q, q_theirs, mask)
        tau.copy_(torch.where(mask.bool(), tau_theirs, tau))
        self.n_attempted += int(mask.numel())
        self.n_swapped += int(mask.sum().item())
        return mask


# ------------------------------------------------------------------------------------------------
# full replica-exchange driver
# ------------------------------------------------------------------------------------------------
class RESwapStats(object):
    """what a replica-exchange scheme logs per attempt (the reference only hints at it in the
    docstrings of `last_draw_stats`, binf/samplers/hmc.py:171-177, binf/samplers/gibbs.py:117,143)"""

    def __init__(self, attempt, partner, accepted_fraction):
        self.attempt, self.partner, self.accepted_fraction = attempt, partner, accepted_fraction

    def __repr__(self):
        return "RESwapStats(attempt=%d, partner=%s, accepted_fraction=%s)" % (
            self.attempt, self.partner, self.accepted_fraction)


def adapt_ladder(betas, pair_rates, gain=1.0, floor=1e-3):
    """New inverse-temperature ladder with the end points kept: the log-gap between neighbours i and
    i+1 grows where the measured swap rate is above the mean rate and shrinks where it is below
    (gap_i *= exp(gain * (p_i - mean p))), then the gaps are rescaled to the original total range.
    Pure function of its inputs, so every rank that holds the all-gathered rates computes the same
    ladder."""
    betas = np.asarray(betas, dtype=np.float64)
    p = np.clip(np.asarray(pair_rates, dtype=np.float64), 0.0, 1.0)
    if len(betas) < 3:
        return betas.copy()
    gaps = np.log(betas[:-1]) - np.log(betas[1:])
    total = gaps.sum()
    gaps = np.maximum(gaps * np.exp(gain * (p - p.mean())), floor * total / len(gaps))
    gaps *= total / gaps.sum()
    out = betas.copy()
    out[1:-1] = np.exp(np.log(betas[0]) - np.cumsum(gaps)[:-1])
    return out


class ReplicaExchangeDriver(object):
    """Sweeps + neighbour swaps + statistics + ladder adaption for one rank of a tempered ensemble.

    The sampling itself is injected so that the protocol runs on CPU (gloo) in the tests:
      sweep()              -- advance this rank's chains by one Gibbs/HMC sweep at self.beta
      log_likelihood()     -- untempered log L per chain, float64 [C]
      set_beta(beta)       -- called when the ladder changes
    `ChainShard` provides all three on the GPU (see `for_shard`)."""

    def __init__(self, rank, world, betas, q, tau, sweep, log_likelihood, set_beta=None, seed=0,
                 chain_base=0, swap_interval=1, decide=None, apply=None, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.betas = [float(b) for b in betas]
        assert len(self.betas) == world
        self.q, self.tau = q, tau
        self._sweep, self._ll, self._set_beta = sweep, log_likelihood, set_beta
        self.swap_interval = int(swap_interval)
        self.rex = ReplicaExchange(rank, world, self.betas[rank], seed=seed, chain_base=chain_base,
                                   decide=decide, apply=apply, group=group)
        self.n_sweeps = 0
        # swap bookkeeping of the pair (rank, rank+1), kept on the lower rank
        self.pair_attempted = 0
        self.pair_swapped = 0
        self._last = None

    @classmethod
    def for_shard(cls, shard, rank, world, betas, seed=0, swap_interval=1, group=None):
        import torch

        def set_beta(b):
            shard.beta = torch.full_like(shard.tau, float(b))
        set_beta(betas[rank])
        return cls(rank, world, betas, shard.q, shard.tau, shard.sweep, shard.log_likelihood, set_beta,
                   seed=seed, chain_base=shard.chain_base, swap_interval=swap_interval, group=group)

    @property
    def beta(self):
        return self.betas[self.rank]

    def step(self):
        """one sweep of every chain, then (every swap_interval sweeps) one swap attempt"""
        self._sweep()
        self.n_sweeps += 1
        if self.world > 1 and self.n_sweeps % self.swap_interval == 0:
            attempt = self.rex.attempt
            partner = swap_partner(self.rank, self.world, attempt)
            mask = self.rex.swap(self.q, self.tau, self._ll(), self.betas)
            frac = None
            if mask is not None:
                n, k = int(mask.numel()), int(mask.sum().item())
                frac = k / float(n)
                if partner > self.rank:
                    self.pair_attempted += n
                    self.pair_swapped += k
            self._last = RESwapStats(attempt, partner, frac)

    def run(self, n_sweeps, sink=None, log_prob=None):
        """n_sweeps steps; after each one the (cold or any) replica's state goes to `sink`
        (a SampleSink / _cabi.Sink) if given"""
        for _ in range(n_sweeps):
            self.step()
            if sink is not None:
                sink.push(self.q, self.tau, None if log_prob is None else log_prob())

    @property
    def last_draw_stats(self):
        return {"swap": self._last}

    def swap_rates(self):
        """acceptance rate of every neighbour pair (r, r+1), r = 0..world-2, identical on all ranks
        (one all-gather of two int64 per rank); NaN where nothing has been attempted yet"""
        import torch
        import torch.distributed as dist
        mine = torch.tensor([self.pair_attempted, self.pair_swapped], dtype=torch.int64, device=self.tau.device)
        if self.world > 1:
            allv = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allv, mine, group=self.group)
        else:
            allv = [mine]
        rates = []
        for r in range(self.world - 1):
            a, s = int(allv[r][0].item()), int(allv[r][1].item())
            rates.append(s / float(a) if a else float("nan"))
        return rates

    def adapt(self, gain=1.0):
        """re-space the ladder from the swap rates measured since the last call (all ranks must call)"""
        rates = self.swap_rates()
        if self.world > 2 and not any(np.isnan(rates)):
            self.betas = [float(b) for b in adapt_ladder(self.betas, rates, gain)]
            self.rex.beta = self.betas[self.rank]
            if self._set_beta is not None:
                self._set_beta(self.betas[self.rank])
        self.pair_attempted = self.pair_swapped = 0
        return rates


# ------------------------------------------------------------------------------------------------
# posterior summaries over the chains of ALL ranks
# ------------------------------------------------------------------------------------------------
def merge_sink_sums(parts):
    """Combine the per-rank sums of `Sink.sums()` into the global per-dimension summary (mean, pooled
    within-chain variance, Gelman-Rubin R-hat, effective sample size per chain) -- what `Sink.summary()`
    returns for one GPU, over the chains of every rank.  Each part carries its own pivot (the running
    mean of its chain 0); the sums are re-centred on the first part's pivot before they are added:
        sum (m - p0)^2 = sum (m - p)^2 + 2 (p - p0) sum (m - p) + C (p - p0)^2."""
    p0 = np.asarray(parts[0]["pivot"], dtype=np.float64)
    n = parts[0]["n"]
    C, s1, s2, s3 = 0, 0.0, 0.0, 0.0
    for part in parts:
        if part["n"] != n:
            raise ValueError("ranks hold different numbers of sweeps (%d vs %d)" % (part["n"], n))
        d = np.asarray(part["pivot"], dtype=np.float64) - p0
        c = part["n_chains"]
        s2 = s2 + part["s2"] + 2.0 * d * part["s1"] + c * d * d
        s1 = s1 + part["s1"] + c * d
        s3 = s3 + part["s3"]
        C += c
    mu = s1 / C
    W = s3 / ((n - 1.0) * C)
    var_means = (s2 - C * mu * mu) / (C - 1.0) if C > 1 else np.zeros_like(W)
    with np.errstate(divide="ignore", invalid="ignore"):
        rhat = np.sqrt(((n - 1.0) / n * W + var_means) / W)
        ess = W / var_means
    return dict(mean=p0 + mu, var=W, rhat=rhat, ess_per_chain=ess, n_chains=C)


def sink_summary_all_ranks(sink, group=None):
    """`Sink.summary()` over the chains of all ranks: one all-gather of 4 x dim doubles per rank
    (a diagnostics exchange, once per reporting interval)."""
    import torch.distributed as dist
    mine = getattr(sink, "_sink", sink).sums()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_sink_sums([mine])
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, mine, group=group)
    return merge_sink_sums(parts)
