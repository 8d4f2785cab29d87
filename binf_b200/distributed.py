"""Chain sharding over the GPUs of one box, diagnostics reduction and replica exchange.

The reference is single-process and single-chain; chains never interact inside
HMCSampler / GibbsSampler (binf/samplers/hmc.py:136-164, binf/samplers/gibbs.py:136-151), so the
batch shards with NO data-path collective: rank r owns the contiguous chain range
[r*C/R, (r+1)*C/R), the data set is replicated, and the Philox streams are keyed by the global
chain id so results do not depend on the number of ranks.  One process per GPU (torchrun),
`torch.distributed` with the NCCL backend over NVLink for the only two exchanges there are:

  * `allreduce_stats`: a handful of float64 accumulators (accepted, proposed, sum eps, sum p_acc,
    optional moments) -- once per reporting interval, never per step;
  * `ReplicaExchange`: exchanges between neighbouring inverse temperatures of a tempered ensemble (the
    reference only alludes to this: binf/samplers/hmc.py:171-177) by LABEL swap: replicas never move, they
    swap temperature indices.  Per attempt the ranks all-gather 16 bytes per chain (log L from the chi^2 the
    trajectory kernel left behind, the temperature index, the step size); even/odd temperature pairs
    alternate; both partners evaluate u < exp(-(beta_a - beta_b)(l_a - l_b)) with the SAME Philox draw
    (SURVEY.md A.3).  No state crosses NVLink, nothing synchronises the host.

Everything here is host-side plumbing; the numerics are C-ABI kernels (binfb_hmc_run, binfb_hmc_last_chi2,
binfb_rex_pack, binfb_rex_decide, binfb_rex_select).
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
    """Neighbour of temperature index `rank` in attempt `attempt` (even attempts pair (0,1),(2,3),..; odd
    attempts pair (1,2),(3,4),..), or None when the index sits out."""
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
        self.n_data = float(model.n_data)
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

    def last_chi2(self):
        """chi^2 of every chain's current state, left behind by the last sweep (no extra pair sweep)"""
        import torch
        self.model.hmc_last_chi2(self.chi2, stream=torch.cuda.current_stream().cuda_stream)
        return self.chi2

    def log_likelihood(self):
        """untempered log L per chain at the current (structure, precision): one fused pass"""
        import torch
        self.model.logprob_grad_device(self.q, self.tau, chi2=self.chi2,
                                       stream=torch.cuda.current_stream().cuda_stream)
        t = self.tau.double()
        return -0.5 * t * self.chi2 + 0.5 * self.n_data * torch.log(t)


def _device_decide(ll_mine, ll_theirs, beta_mine, beta_theirs, i_am_low, seed, attempt, pair_id,
                   swap_stream=0):
    """State-swap variant (binfb_swap_decide): accept mask for two fixed-temperature partners; the lower
    rank of the pair is 'a' on both sides.  `swap_stream` keys the Philox draw and MUST be the same number
    on both partners (it is not the ranks' HMC chain base)."""
    import torch
    from binf_b200 import _cabi
    a, b = (ll_mine, ll_theirs) if i_am_low else (ll_theirs, ll_mine)
    ba, bb = (beta_mine, beta_theirs) if i_am_low else (beta_theirs, beta_mine)
    mask = torch.zeros(ll_mine.shape[0], dtype=torch.uint8, device=ll_mine.device)
    _cabi.check(_cabi.lib().binfb_swap_decide(_cabi.ptr(a), _cabi.ptr(b), float(ba), float(bb),
                                              int(ll_mine.shape[0]), seed, attempt, pair_id, swap_stream,
                                              _cabi.ptr(mask),
                                              _cabi.ptr(torch.cuda.current_stream().cuda_stream)))
    return mask


def _device_apply(q_mine, q_theirs, mask):
    import torch
    from binf_b200 import _cabi
    _cabi.check(_cabi.lib().binfb_swap_apply(_cabi.ptr(q_mine), _cabi.ptr(q_theirs), None, None,
                                             _cabi.ptr(mask), int(q_mine.shape[0]), int(q_mine.shape[1]),
                                             _cabi.ptr(torch.cuda.current_stream().cuda_stream)))


class DeviceOps(object):
    """The three replica-exchange kernels behind the C ABI (csrc/misc.cu), on the current CUDA stream."""

    @staticmethod
    def _stream():
        import torch
        return _ptr(torch.cuda.current_stream().cuda_stream)

    def pack(self, chi2, tau, eps, tidx, n_data, records):
        from binf_b200 import _cabi
        _cabi.check(_cabi.lib().binfb_rex_pack(_ptr(chi2), _ptr(tau), _ptr(eps), _ptr(tidx), int(chi2.shape[0]),
                                               float(n_data), _ptr(records), self._stream()))

    def decide(self, records_all, world, rank, n_chains, n_columns, betas, seed, attempt, ll_shift, tidx, beta, eps,
               accept, pair_counts, temp_stats):
        from binf_b200 import _cabi
        b = np.ascontiguousarray(betas, dtype=np.float64)
        _cabi.check(_cabi.lib().binfb_rex_decide(_ptr(records_all), world, rank, n_chains, n_columns, _ptr(b), len(b),
                                                 int(seed), int(attempt), float(ll_shift), _ptr(tidx), _ptr(beta),
                                                 _ptr(eps), _ptr(accept), _ptr(pair_counts), _ptr(temp_stats),
                                                 self._stream()))

    def select(self, q, aux, tidx, k_sel, n_columns, out_q, out_aux):
        from binf_b200 import _cabi
        _cabi.check(_cabi.lib().binfb_rex_select(_ptr(q), _ptr(aux), _ptr(tidx), int(k_sel), int(q.shape[0]),
                                                 int(q.shape[1]), int(n_columns), _ptr(out_q), _ptr(out_aux),
                                                 self._stream()))


def _ptr(a):
    from binf_b200 import _cabi
    return _cabi.ptr(a)


class ReplicaExchange(object):
    """Replica exchange by LABEL swap over a grid [temperature k][column c] of replicas.

    Every rank holds `n_chains` = rows x `n_columns` replicas (local chain i sits in column i % n_columns);
    together the ranks hold len(betas) = world x rows temperatures of every column.  A replica never moves:
    it carries a temperature index `tidx`, and an accepted exchange swaps the indices (and with them the
    inverse temperatures `beta` and the step sizes `eps`, which are tuned per temperature) of the two
    replicas.  Per attempt the ranks all-gather one 16-byte record per chain (log L, tidx, eps) -- 8 KB per
    rank at 512 chains -- and every rank decides for its own chains; both partners draw the same Philox
    uniform, keyed by (seed, attempt, lower temperature index, column) and by nothing rank-specific.
    Nothing on this path synchronises the host: counters and statistics stay on the device until asked for.

    `ops` defaults to the C-ABI kernels (`DeviceOps`); the CPU (gloo) tests of the protocol pass the host
    restatement of oracle/rex_port.py."""

    def __init__(self, rank, world, betas, n_chains, n_columns=None, seed=0, device=None, group=None, ops=None):
        import torch
        self.rank, self.world, self.group = int(rank), int(world), group
        self.betas = np.array(betas, dtype=np.float64)
        T = len(self.betas)
        total = self.world * int(n_chains)
        if total % T != 0:
            raise ValueError("world x n_chains = %d replicas do not divide into %d temperatures" % (total, T))
        self.n_chains = int(n_chains)
        self.n_columns = total // T if n_columns is None else int(n_columns)
        if self.n_chains % self.n_columns != 0 or self.n_columns * T != total:
            raise ValueError("n_chains = %d must be rows x n_columns with world x rows = %d temperatures"
                             % (self.n_chains, T))
        if T > 64:
            raise ValueError("at most 64 temperatures")
        self.rows = self.n_chains // self.n_columns
        self.seed = int(seed)
        self.ops = ops or DeviceOps()
        dev = device
        # rank r starts with the temperatures r*rows .. (r+1)*rows - 1, one per row of its chains
        row = torch.arange(self.n_chains, dtype=torch.int32) // self.n_columns
        self.tidx = (row + self.rank * self.rows).to(torch.int32).to(dev)
        self.beta = torch.as_tensor(self.betas[self.tidx.cpu().numpy()], dtype=torch.float32).to(dev)
        self.records = torch.zeros(self.n_chains * 16, dtype=torch.uint8, device=dev)
        self.records_all = torch.zeros(self.world * self.n_chains * 16, dtype=torch.uint8, device=dev)
        self.accept = torch.zeros(self.n_chains, dtype=torch.uint8, device=dev)
        self.pair_counts = torch.zeros(max(T - 1, 1), 2, dtype=torch.int64, device=dev)
        self.temp_stats = torch.zeros(T, 3, dtype=torch.float64, device=dev)
        self.ll_shift = 0.0
        self.attempt = 0

    @property
    def n_temps(self):
        return len(self.betas)

    def swap(self, chi2, tau, eps, n_data):
        """One attempt.  chi2 [C] f64 (chi^2 of the current states), tau [C], eps [C]: this rank's chains;
        self.tidx, self.beta and eps are updated in place where an exchange is accepted.  Stream-ordered,
        no host synchronisation; self.accept holds the decisions of this attempt afterwards."""
        import torch.distributed as dist
        self.ops.pack(chi2, tau, eps, self.tidx, n_data, self.records)
        if self.world > 1:
            dist.all_gather_into_tensor(self.records_all, self.records, group=self.group)
            allr = self.records_all
        else:
            allr = self.records
        self.ops.decide(allr, self.world, self.rank, self.n_chains, self.n_columns, self.betas, self.seed,
                        self.attempt, self.ll_shift, self.tidx, self.beta, eps, self.accept, self.pair_counts,
                        self.temp_stats)
        self.attempt += 1
        return self.accept

    def set_betas(self, betas, eps=None):
        """New ladder (same length): every replica keeps its temperature index; step sizes follow
        eps ~ 1/sqrt(beta) (the likelihood force scales with beta)."""
        import torch
        new = np.array(betas, dtype=np.float64)
        assert len(new) == len(self.betas)
        if eps is not None:
            ratio = torch.as_tensor(np.sqrt(self.betas / new), dtype=torch.float32).to(self.tidx.device)
            eps.mul_(ratio[self.tidx.long()])
        self.betas = new
        self.beta.copy_(torch.as_tensor(new, dtype=torch.float32).to(self.tidx.device)[self.tidx.long()])

    def _reduced(self, t):
        import torch.distributed as dist
        t = t.clone()
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def swap_rates(self):
        """acceptance rate of every temperature pair (k, k+1) since the last reset, identical on all ranks
        (one all-reduce of the counters); NaN where nothing has been attempted"""
        c = self._reduced(self.pair_counts)[: self.n_temps - 1].astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            return [float(x) for x in np.where(c[:, 0] > 0, c[:, 1] / c[:, 0], np.nan)]

    def temperature_stats(self):
        """(mean, std) of the untempered log-likelihood at every temperature since the last reset"""
        s = self._reduced(self.temp_stats)
        n = np.maximum(s[:, 0], 1.0)
        mean = s[:, 1] / n
        var = np.maximum(s[:, 2] / n - mean * mean, 0.0)
        return mean + self.ll_shift, np.sqrt(var)

    def reset_stats(self, ll_shift=None):
        self.pair_counts.zero_()
        self.temp_stats.zero_()
        if ll_shift is not None:
            self.ll_shift = float(ll_shift)

    def select(self, q, aux=None, k_sel=0, dst=None):
        """The replicas of temperature index `k_sel` (0 = the posterior samples), one per column, wherever
        they currently live: every rank contributes its own (zeros elsewhere) and the contributions are
        summed over the ranks (exactly one non-zero contributor per column, so the sum is exact).
        dst = None: all-reduce, every rank gets them; dst = r: reduce onto rank r.  Returns (q [n_columns, D],
        aux [n_columns] or None)."""
        import torch
        import torch.distributed as dist
        out_q = torch.zeros(self.n_columns, q.shape[1], dtype=q.dtype, device=q.device)
        out_aux = None if aux is None else torch.zeros(self.n_columns, dtype=aux.dtype, device=q.device)
        self.ops.select(q, aux, self.tidx, k_sel, self.n_columns, out_q, out_aux)
        if self.world > 1:
            for t in (out_q, out_aux):
                if t is None:
                    continue
                if dst is None:
                    dist.all_reduce(t, group=self.group)
                else:
                    dist.reduce(t, dst=dst, group=self.group)
        return out_q, out_aux


# ------------------------------------------------------------------------------------------------
# full replica-exchange driver
# ------------------------------------------------------------------------------------------------
class RESwapStats(object):
    """what a replica-exchange scheme logs per attempt (the reference only hints at it in the
    docstrings of `last_draw_stats`, binf/samplers/hmc.py:171-177, binf/samplers/gibbs.py:117,143)"""

    def __init__(self, attempt, accepted):
        self.attempt, self._accepted = attempt, accepted

    @property
    def accepted_fraction(self):
        """fraction of this rank's replicas that exchanged in that attempt (synchronises)"""
        return float(self._accepted.float().mean().item())

    def __repr__(self):
        return "RESwapStats(attempt=%d)" % self.attempt


def expected_swap_rate(mu):
    """acceptance rate of an exchange whose log-ratio Delta is Gaussian with mean mu = var/2 (what detailed
    balance implies for overlapping Gaussian energy distributions): erfc(sqrt(mu) / 2)"""
    import math
    return math.erfc(math.sqrt(max(mu, 0.0)) / 2.0)


def adapt_ladder(betas, mean_ll, target=0.3, keep_ends=False):
    """New inverse-temperature ladder from the MEASURED mean untempered log-likelihoods at the current one.

    The mean of the exchange statistic between neighbours a, b is mu = (beta_a - beta_b)(L(beta_a) - L(beta_b))
    with L(beta) the mean log-likelihood at beta; the expected acceptance is erfc(sqrt(mu)/2).  L is
    interpolated through the measurements linearly in 1/beta (equipartition: L(beta) = L_max - d/(2 beta) for
    d effective degrees of freedom) and extrapolated with the nearest segment.  Starting from the cold end
    betas[0], every next temperature is placed where the expected acceptance equals `target` -- this also
    works when every measured swap rate is zero, the case a rate-driven rule cannot leave.  With
    keep_ends=True the hot end stays where it is and `target` is replaced by the one common rate that makes
    the ladder span [betas[-1], betas[0]].  Pure function of its inputs: every rank computes the same ladder."""
    betas = np.asarray(betas, dtype=np.float64)
    L = np.asarray(mean_ll, dtype=np.float64)
    T = len(betas)
    if T < 2 or not np.all(np.isfinite(L)):
        return betas.copy()
    x = 1.0 / betas                                   # increasing
    Lm = np.minimum.accumulate(L)                     # colder = higher log-likelihood (monotone fix)

    def Lof(b):
        xi = 1.0 / b
        j = int(np.clip(np.searchsorted(x, xi) - 1, 0, T - 2))
        dx = x[j + 1] - x[j]
        return Lm[j] + (Lm[j + 1] - Lm[j]) * (xi - x[j]) / dx if dx > 0 else Lm[j]

    def mu(ba, bb):
        return (ba - bb) * (Lof(ba) - Lof(bb))

    def build(mu_t):
        out = [betas[0]]
        for _ in range(T - 1):
            ba = out[-1]
            lo, hi = ba * 1e-6, ba                    # mu(ba, b) grows as b falls below ba
            if mu(ba, lo) <= mu_t:
                out.append(lo)
                continue
            for _ in range(200):
                mid = 0.5 * (lo + hi)
                if mu(ba, mid) > mu_t:
                    lo = mid
                else:
                    hi = mid
            out.append(0.5 * (lo + hi))
        return np.array(out)

    def mu_for(rate):
        lo, hi = 0.0, 1e4
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if expected_swap_rate(mid) > rate:
                lo = mid
            else:
                hi = mid
        return 0.5 * (lo + hi)

    if not keep_ends:
        return build(mu_for(float(np.clip(target, 1e-6, 0.999999))))
    lo, hi = 1e-9, 1e7                                # larger mu per gap = wider ladder
    for _ in range(200):
        mid = np.sqrt(lo * hi)
        if build(mid)[-1] > betas[-1]:
            lo = mid
        else:
            hi = mid
    out = build(np.sqrt(lo * hi))
    out[-1] = betas[-1]
    return out


class ReplicaExchangeDriver(object):
    """Sweeps + label swaps + statistics + ladder adaption for one rank of a tempered ensemble.

    `replica` provides the sampling: attributes `tau`, `eps` (per-chain tensors, updated in place), `beta`
    (set by the driver to the exchange's per-chain tensor), `n_data`, and methods `sweep()` (one Gibbs/HMC
    sweep of every chain at its beta) and `last_chi2()` (chi^2 of the current states, float64 [C]).
    `ChainShard` is that on the GPU; the CPU (gloo) tests pass a toy replica."""

    def __init__(self, replica, rank, world, betas, n_columns=None, seed=0, swap_interval=1, group=None, ops=None):
        self.replica, self.rank, self.world, self.group = replica, rank, world, group
        n = int(replica.tau.shape[0])
        self.rex = ReplicaExchange(rank, world, betas, n, n_columns=n_columns, seed=seed,
                                   device=replica.tau.device, group=group, ops=ops)
        replica.beta = self.rex.beta
        self.swap_interval = int(swap_interval)
        self.n_sweeps = 0
        self._last = None

    @classmethod
    def for_shard(cls, shard, rank, world, betas, n_columns=None, seed=0, swap_interval=1, group=None):
        return cls(shard, rank, world, betas, n_columns=n_columns, seed=seed, swap_interval=swap_interval,
                   group=group)

    @property
    def betas(self):
        return [float(b) for b in self.rex.betas]

    def step(self):
        """one sweep of every chain, then (every swap_interval sweeps) one exchange attempt"""
        self.replica.sweep()
        self.n_sweeps += 1
        if self.rex.n_temps > 1 and self.n_sweeps % self.swap_interval == 0:
            attempt = self.rex.attempt
            acc = self.rex.swap(self.replica.last_chi2(), self.replica.tau, self.replica.eps, self.replica.n_data)
            self._last = RESwapStats(attempt, acc)

    def run(self, n_sweeps, sink=None, thin=1):
        """n_sweeps steps; every `thin`-th one the cold replicas (temperature index 0: the posterior samples,
        assembled from wherever they live) go to `sink` on rank 0"""
        for k in range(n_sweeps):
            self.step()
            if sink is not None and (k + 1) % thin == 0:
                q, tau = self.cold_states(dst=0)
                if self.rank == 0:
                    sink.push(q, tau)

    def cold_states(self, dst=None):
        return self.rex.select(self.replica.q, self.replica.tau, 0, dst=dst)

    @property
    def last_draw_stats(self):
        return {"swap": self._last}

    def swap_rates(self):
        return self.rex.swap_rates()

    def adapt(self, target=0.3, keep_ends=False):
        """re-space the ladder from the log-likelihood statistics gathered since the last call (all ranks
        must call); returns the swap rates measured over the same period"""
        rates = self.rex.swap_rates()
        mean, _ = self.rex.temperature_stats()
        if self.rex.n_temps > 1:
            self.rex.set_betas(adapt_ladder(self.rex.betas, mean, target=target, keep_ends=keep_ends),
                               eps=self.replica.eps)
        self.rex.reset_stats(ll_shift=float(mean[0]))
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
