#!/usr/bin/env python
"""Benchmark of the B200-native HMC hot path (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload chromatin|poly]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port)

The default run measures BASELINE.json configs[2] (the headline line) and then, on the same box in the same
process, short legs of configs[1] (poly), configs[3] (chromatin5k), configs[4] (rex) and the sample sink; they
are attached to the line as extra.{poly,chromatin_algebraic,generic,chromatin5k,sink,rex}, each with ms_per_step,
roofline and clocks (generic = configs[1] with the cubic as a user-defined NVRTC model; chromatin_algebraic =
configs[2] with the algebraic contact function of SURVEY.md A.2 instead of the logistic one).

A "step" is one Gibbs sweep over every chain of the batch: the conjugate precision update
followed by one HMC trajectory of L leapfrog steps (L+1 fused force evaluations) and the
Metropolis test -- one launch of the fused kernel.  metric = leapfrog steps/s =
chains x L x sweeps / device time.  Default workload = BASELINE.json configs[2], the
1000-bead chromatin posterior the north star's target is quoted on (4,096 chains per GPU,
weak scaling: every rank owns its own 4,096 chains, no data-path collective; the only
collective is the 4-double diagnostics all-reduce at the end of the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints exactly ONE line on stdout.  Libraries write there too (NCCL's "NCCL version ..." banner at
# communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes
# to a saved copy of the real stdout.
_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def _capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)

USER_CUBIC = """
__device__ float binfb_mock(const float *theta, const float *x, float *dmock) {
    dmock[0] = 1.0f; dmock[1] = x[0]; dmock[2] = x[1]; dmock[3] = x[2];
    return fmaf(theta[3], x[2], fmaf(theta[2], x[1], fmaf(theta[1], x[0], theta[0])));
}
"""
FLOP_PER_PAIR = 31.0          # per unordered bead pair and force evaluation
FLOP_PER_DATUM = 14.0         # per chain, datum and force evaluation (K = 4)

WORKLOADS = {
    # eps: the step size at which the EQUILIBRATED chains accept 0.8-0.9 of the proposals (SURVEY.md 8d;
    # profiles/experiments/eps_scan.py); equilibrate: untimed sweeps in front of the warm-up, because the synthetic
    # starting points (truth + noise) relax downhill and would accept everything at any step size
    "chromatin": dict(name="chromatin_n1000_c4096_L20_gibbs", n_beads=1000, chains=4096, L=20,
                      eps=9e-3, equilibrate=75, alpha=2.0, d_c=2.5, k_bb=4.0, l0=1.0, noise=0.05),
    "poly": dict(name="poly_K4_N1000_c65536_L20", n_data=1000, chains=65536, L=20, eps=0.010, equilibrate=300,
                 tau=2.5),
    # configs[1] again with the cubic written as a USER-DEFINED forward model: CUDA device code
    # for one datum, compiled at run time with NVRTC into the fused kernels.  The powers of x are passed as abscissae,
    # which is what the built-in model precomputes on the host
    "generic": dict(name="poly_K4_N1000_c65536_L20_user_model", n_data=1000, chains=65536, L=20, eps=0.010,
                    equilibrate=300, tau=2.5),
    # BASELINE.json configs[3]: 5000 beads, chains sharded across the GPUs (4 chains per SM and GPU)
    "chromatin5k": dict(name="chromatin_n5000_c592_L20_gibbs", n_beads=5000, chains=592, L=20,
                        eps=8e-3, equilibrate=30, alpha=2.0, d_c=2.5, k_bb=4.0, l0=1.0, noise=0.05),
    # BASELINE.json configs[4]: one inverse temperature per rank (geometric in [0.05, 1]), 512 chains
    # per rank, a neighbour swap attempt (NCCL send/recv over NVLink) after every sweep
    "rex": dict(name="chromatin_n1000_rex_512_per_rank_L20", n_beads=1000, chains=512, L=20,
                eps=9e-3, alpha=2.0, d_c=2.5, k_bb=4.0, l0=1.0, noise=0.05),
}


def working_set_bytes(name, C):
    """state a sweep streams: q, and for the chromatin kernels the working copies of q and p"""
    w = WORKLOADS[name]
    return 3 * 4 * C * 3 * w["n_beads"] if "n_beads" in w else 2 * 4 * C * 4


def needs_l2_flush(name, C):
    return working_set_bytes(name, C) < 130e6


def workload_config(name, world, chains=None, eps=None, equilibrate=True):
    """the `config` object of a bench line: the same for the product arm and the reference arm of a workload"""
    w = WORKLOADS[name]
    C = chains or w["chains"]
    gibbs = name not in ("poly", "generic")
    return dict(workload=w["name"], chains_per_gpu=C, leapfrog_steps=w["L"], timestep=eps or w["eps"],
                parallelism="chain-sharded x%d (no data-path collective)" % world,
                l2="L2 flushed (256 MiB fill) between timed steps" if needs_l2_flush(name, C)
                else "working set %.0f MB per step > 126 MB L2" % (working_set_bytes(name, C) / 1e6),
                gibbs="precision update fused in front of each trajectory" if gibbs else "none",
                equilibration_sweeps=int(w.get("equilibrate", 0)) if equilibrate else 0)


# --------------------------------------------------------------------------------------------
# synthetic inputs; generated without touching oracle/ so the product arm
# never imports the checker
# --------------------------------------------------------------------------------------------
def chromatin_inputs(w, chains, seed, contact="logistic"):
    n = w["n_beads"]
    rng = np.random.RandomState(0)
    X = np.cumsum(rng.normal(size=(n, 3)) * w["l0"], axis=0)
    X -= X.mean(axis=0)
    i, j = np.triu_indices(n, 1)
    d = np.sqrt(np.sum((X[i] - X[j]) ** 2, axis=-1) + 1e-12)
    if contact == "algebraic":
        z = w["alpha"] * (w["d_c"] - d)
        y = 0.5 * (1.0 + z / np.sqrt(1.0 + z * z)) + rng.normal(size=d.shape) * w["noise"]
    else:
        with np.errstate(over="ignore"):
            y = 1.0 / (1.0 + np.exp(w["alpha"] * (d - w["d_c"]))) + rng.normal(size=d.shape) * w["noise"]
    rng = np.random.RandomState(1 + seed)
    q = (X.reshape(-1)[None, :] + 0.1 * rng.normal(size=(chains, 3 * n))).astype(np.float32)
    return y.astype(np.float32), q


def poly_inputs(w, chains, seed):
    rng = np.random.RandomState(0)
    xs = np.linspace(-2, 2, w["n_data"])
    ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2.0, -4.0, 1.0, 1.5]), 1 / np.sqrt(2.5))
    rng = np.random.RandomState(1 + seed)
    q = (np.ones((chains, 4)) + 0.1 * rng.normal(size=(chains, 4))).astype(np.float32)
    return xs, ys, q


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4)
                          if r[3 + k].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


# --------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the oracle port timed on the host cores
# --------------------------------------------------------------------------------------------
def _cpu_worker(args):
    workload, seed, budget_s = args
    workload = "chromatin" if workload in ("rex", "chromatin5k") else ("poly" if workload == "generic" else workload)
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binf_port as port
    import chromatin_port as chrom
    rng = np.random.RandomState(100 + seed)
    if workload == "chromatin":
        w = WORKLOADS["chromatin"]
        X, y = chrom.synthetic_chromatin(w["n_beads"], w["alpha"], w["d_c"], w["l0"], w["noise"], 0)
        model = chrom.ChromatinModel(w["n_beads"], y, w["alpha"], w["d_c"], w["k_bb"], w["l0"])
        q = X.reshape(-1) + 0.1 * rng.normal(size=3 * w["n_beads"])
        tau, n_pairs = 100.0, model.n_pairs
        steps, t0 = 0, time.perf_counter()
        while True:
            # one Gibbs sweep of the reference: precision update, then one HMC transition
            tau = float(port.gamma_precision_sample(model.chi2(q), n_pairs, 1.0, 1.0, rng))
            r = port.hmc_sample(lambda x: model.log_prob(x, tau), lambda x: model.gradient(x, tau), q,
                                w["eps"], w["L"], rng.normal(size=q.shape), rng.uniform())
            q = r["q"]
            steps += w["L"]
            if time.perf_counter() - t0 > budget_s:
                break
        return steps, time.perf_counter() - t0
    w = WORKLOADS["poly"]
    xs = np.linspace(-2, 2, w["n_data"])
    ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2.0, -4.0, 1.0, 1.5]), 1 / np.sqrt(2.5))
    pp = port.PolynomialPosterior(xs, ys, np.zeros(4), 5 * np.ones(4), 1.0, 1.0)
    q, tau = np.ones(4), w["tau"]
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        r = port.hmc_sample(lambda c: pp.log_prob(c, tau), lambda c: pp.gradient(c, tau), q, w["eps"],
                            w["L"], rng.normal(size=4), rng.uniform())
        q = r["q"]
        steps += w["L"]
    return steps, time.perf_counter() - t0


def cpu_baseline(workload, budget_s, cores=None):
    """P independent single-chain processes of the oracle port (float64 numpy, the reference's
    arithmetic), aggregate leapfrog steps/s."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(workload, s, budget_s) for s in range(cores)])
    value = sum(st / el for st, el in res)
    sample = "%d single-chain processes x %.0f s of Gibbs sweeps (%d trajectories total)" % (
        cores, budget_s, sum(st for st, _ in res) // WORKLOADS[workload]["L"])
    return dict(value=value, unit="leapfrog steps/s", cores=cores, kind="port", sample=sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    t0 = time.perf_counter()
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    cpu_baseline(args.workload, 1.0)  # warm-up (imports, page-in)
    vals = [cpu_baseline(args.workload, per_step) for _ in range(max(1, min(args.steps, 3)))]
    best = max(vals, key=lambda v: v["value"])
    line = dict(impl="reference", metric="HMC leapfrog steps/s (chains x steps)", value=best["value"],
                unit="leapfrog steps/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=None, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                # the product arm's config of this workload; the reference runs a bounded sample of it (same model,
                # L, step size and sweep definition on `cores` single chains: cpu_baseline.sample)
                config=workload_config(args.workload, max(1, args.gpus), args.chains or None, args.eps or None,
                                       not args.no_equilibrate),
                note="oracle port of the reference's numpy path (float64, one chain per host core); the reference "
                     "itself is Python 2 + CSB and cannot travel to the box",
                cpu_baseline=best,
                e2e=dict(value=best["value"], unit="leapfrog steps/s", h2d_bytes_per_step=0,
                         d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t0)
    emit(json.dumps(line))


# --------------------------------------------------------------------------------------------
# the product arm: one process per GPU; every leg is timed on the device, max over ranks
# --------------------------------------------------------------------------------------------
class Ctx(object):
    """rank / device / process group of this process (torch.distributed over NCCL when launched by torchrun)"""

    def __init__(self):
        import torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
        self._mb = None

    def sync_all(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, x):
        import torch
        if self.world == 1:
            return float(x)
        import torch.distributed as dist
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def microbench(self):
        """FFMA / FFMA2 / MUFU issue-rate microbenchmarks of this device (roofline denominators), run once"""
        if self._mb is None:
            from binf_b200 import _cabi
            self._mb = _cabi.microbench(self.local, 3000)
        return self._mb

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def recorded_traffic(key, ok=True):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json)"""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return tj[key]["dram_bytes_per_launch"] if ok and key in tj else None
    except Exception:
        return None


def make_hmc_workload(ctx, args, name, chains=None):
    """model + device state of one HMC workload; returns a dict the legs below step"""
    import torch
    from binf_b200 import _cabi
    w = WORKLOADS[name]
    C = chains or w["chains"]
    L = w["L"]
    if name in ("chromatin", "rex", "chromatin5k"):
        contact = getattr(args, "contact", "logistic")
        y, q_host = chromatin_inputs(w, C, ctx.rank, contact)
        model = _cabi.Model.chromatin(w["n_beads"], y, w["alpha"], w["d_c"], w["k_bb"], w["l0"], 0.0,
                                      1.0, 1.0, device=ctx.local, roles=args.roles, ev_k=args.ev_k, ev_d=1.5,
                                      contact=contact)
        if args.chrom_sets >= 0:
            model.set_option("chrom.sets", args.chrom_sets)
        if getattr(args, "chrom_warps", 0) > 0:
            model.set_option("chrom.warps", args.chrom_warps)
        if getattr(args, "host_chunks", 0) > 0:
            model.set_option("host.chunks", args.host_chunks)
        tau0, gibbs = 100.0, _cabi.GIBBS_TAU_FIRST
        units = float(model.n_data)            # pairs per force evaluation
        # (the algebraic contact function: 32 flop + 2 special-function ops per pair by the same count)
        flop_per_launch = (32.0 if contact == "algebraic" else FLOP_PER_PAIR) * units * (L + 1) * C
        bytes_per_launch = 2.0 * 4 * q_host.shape[1] * C + 4.0 * units
    else:
        xs, ys, q_host = poly_inputs(w, C, ctx.rank)
        if name == "generic":
            model = _cabi.Model.generic(USER_CUBIC, 4, np.stack([xs, xs ** 2, xs ** 3], 1), ys, np.zeros(4),
                                        5 * np.ones(4), 1.0, 1.0, device=ctx.local)
        else:
            model = _cabi.Model.polynomial(xs, ys, 4, np.zeros(4), 5 * np.ones(4), 1.0, 1.0, device=ctx.local)
        tau0, gibbs = w["tau"], _cabi.GIBBS_NONE
        units = float(w["n_data"])
        flop_per_launch = (FLOP_PER_DATUM * (L + 1) + 4) * units * C
        bytes_per_launch = 2.0 * 4 * 4 * C
    dev = ctx.dev
    return dict(name=name, w=w, C=C, L=L, D=q_host.shape[1], model=model,
                contact=getattr(args, "contact", "logistic") if "n_beads" in w else None, q_host=q_host, tau0=tau0, gibbs=gibbs,
                units=units, flop_per_launch=flop_per_launch, bytes_per_launch=bytes_per_launch,
                q=torch.from_numpy(q_host).to(dev), tau=torch.full((C,), tau0, device=dev, dtype=torch.float32),
                accepted=torch.zeros(C, device=dev, dtype=torch.uint8),
                nacc=torch.zeros(C, device=dev, dtype=torch.int32),
                stats=torch.zeros(4, device=dev, dtype=torch.float64))


def chromatin_roofline(ctx, wl, ms_kernel, traffic):
    """FP32-FMA roofline of one sweep launch + the special-function unit as the binding pipe (chromatin)"""
    mb = ctx.microbench()
    peaks = measured_peaks()
    achieved = wl["flop_per_launch"] / (ms_kernel * 1e-3) / 1e12
    sfu = None
    if wl["name"] not in ("poly", "generic"):
        # the binding pipe of the pair kernel: 3 MUFU (rsqrt, ex2, rcp) per bead pair
        n_sfu = 2 if wl.get("contact") == "algebraic" else 3
        sfu_gops = n_sfu * wl["units"] * (wl["L"] + 1) * wl["C"] / (ms_kernel * 1e-3) / 1e9
        sfu = dict(ops_per_pair=n_sfu, achieved_gops=sfu_gops, peak_gops=mb["mufu_gops"],
                   frac=sfu_gops / mb["mufu_gops"],
                   note="MUFU issues at 16 lanes/clk/SM = 24 SMSP-cycles per warp-pair (64.6 % of the FP32-FMA "
                        "peak if it were the only limit).  It is the busiest single pipe but not the wall: with all "
                        "three MUFU replaced by ALU ops the isolated pair block still takes 26.1 of 29.9 cycles "
                        "(profiles/r2_microbench_pairbench.txt) -- 18 FMA-pipe ops per pair at 1.07-1.5 "
                        "issue cycles each (three-operand FFMA2 / FFMA pay for register-file bandwidth) plus the "
                        "shared-memory staging of a step")
    return dict(
        bound="fp32_fma", achieved=achieved, peak=mb["ffma_tflops"], unit="TFLOP/s",
        frac=achieved / mb["ffma_tflops"], traffic=traffic, sfu=sfu,
        peak_source="live FFMA issue microbenchmark in this run (binfb_microbench); "
                    "MEASURED_PEAKS.json has no non-tensor FP32 figure",
        peak_formula_tflops=148 * 128 * 2 * 1.965e9 / 1e12,
        ffma2_tflops=mb["ffma2_tflops"], mufu_gops=mb["mufu_gops"],
        flop_per_launch=wl["flop_per_launch"], kernel_ms=ms_kernel,
        hbm_sanity=dict(algorithmic_gb_per_launch=wl["bytes_per_launch"] / 1e9,
                        achieved_gbs=wl["bytes_per_launch"] / (ms_kernel * 1e-3) / 1e9,
                        peak_gbs=peaks.get("hbm_gbs"),
                        note="compulsory HBM bytes are O(chains x dim) per trajectory: not the bound"))


def hmc_leg(ctx, args, name, steps, warmup, chains=None, eps=None, with_e2e=True):
    """K timed Gibbs/HMC sweeps of one workload (one fused launch each), chains resident in HBM; then the same
    sweep end to end through the host-buffer C-ABI call.  Returns the measurement (complete on rank 0)."""
    import torch
    from binf_b200 import _cabi
    wl = make_hmc_workload(ctx, args, name, chains)
    w, C, L, D, model = wl["w"], wl["C"], wl["L"], wl["D"], wl["model"]
    dev, world = ctx.dev, ctx.world
    eps0 = eps or w["eps"]
    chain_base = ctx.rank * C
    eps_t = torch.full((C,), eps0, device=dev, dtype=torch.float32)
    # timing rule: inputs larger than the 126 MB L2, or an L2 flush between timed iterations
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8) if needs_l2_flush(name, C) else None
    stream = torch.cuda.current_stream().cuda_stream
    draw = [0]

    def step():
        opts = _cabi.HmcOpts(L, 1, 0, wl["gibbs"], 1.05, 0.95, args.seed, draw[0], chain_base)
        model.hmc_run_device(wl["q"], wl["tau"], eps_t, opts, accepted=wl["accepted"], n_accepted=wl["nacc"],
                             stats=wl["stats"], stream=stream)
        draw[0] += 1

    n_eq = 0 if args.no_equilibrate else int(w.get("equilibrate", 0))
    for _ in range(n_eq + warmup):
        step()
    ctx.sync_all()
    wl["stats"].zero_()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.sync_all()
    t_wall = time.perf_counter()
    for k in range(steps):
        if flush is not None:
            flush.fill_(k)                      # L2 flush between timed iterations (not timed)
        ev[k][0].record()
        step()
        ev[k][1].record()
    torch.cuda.synchronize()
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    if flush is None:
        total_ms = ev[0][0].elapsed_time(ev[-1][1])   # one bracket over all K steps
    else:
        total_ms = float(sum(kernel_ms))
    stats = wl["stats"]
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(stats)                  # the diagnostics reduction (4 doubles, NCCL)
    total_ms = ctx.max_over_ranks(total_ms)
    ctx.sync_all()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.summary() if sampler else None
    st = stats.cpu().numpy()
    value = world * C * L * steps / (total_ms * 1e-3)

    # ---- e2e: the same sweep through the host-buffer C-ABI call (pinned host memory) --------
    e2e = None
    if with_e2e:
        qh = torch.from_numpy(wl["q_host"].copy()).pin_memory().numpy()
        th = torch.full((C,), wl["tau0"], dtype=torch.float32).pin_memory().numpy()
        eh = torch.full((C,), eps0, dtype=torch.float32).pin_memory().numpy()
        ah = torch.zeros(C, dtype=torch.uint8).pin_memory().numpy()
        n_e2e = max(2, min(steps, 5))
        from ctypes import byref

        def host_step(k):
            opts = _cabi.HmcOpts(L, 1, 0, wl["gibbs"], 1.05, 0.95, args.seed, 1000 + k, chain_base)
            _cabi.check(_cabi.lib().binfb_hmc_run_host(
                model._h, _cabi.ptr(qh), _cabi.ptr(th), None, _cabi.ptr(eh), C, byref(opts), None, None,
                None, _cabi.ptr(ah), None, None, None, None, None, None))
        host_step(0)
        ctx.sync_all()
        t0 = time.perf_counter()
        for k in range(n_e2e):
            host_step(1 + k)
        el = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = dict(value=world * C * L * n_e2e / el, unit="leapfrog steps/s",
                   h2d_bytes_per_step=int(4 * C * D + 8 * C),
                   # back: the state, the accept flags, the precision when the sweep updates it (the step sizes
                   # only while they adapt)
                   d2h_bytes_per_step=int(4 * C * D + C + (4 * C if wl["gibbs"] else 0)),
                   steps=n_e2e, ms_per_step=el * 1e3 / n_e2e,
                   api="binfb_hmc_run_host: pinned host buffers in and out, synchronous; the state copies are "
                       "chunked on two copy streams and overlap the one trajectory launch")
    model.close()
    if ctx.rank != 0:
        return None
    ms_kernel = float(np.mean(kernel_ms))
    roofline = chromatin_roofline(ctx, wl, ms_kernel, recorded_traffic(name, C == w["chains"]))
    return dict(metric="HMC leapfrog steps/s (chains x steps)", value=value, unit="leapfrog steps/s",
                n_gpus=world, steps=steps, warmup=warmup, ms_per_step=total_ms / steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=dict(workload_config(name, world, C, eps0, not args.no_equilibrate),
                            **({"contact_function": "algebraic (experiment; the headline is logistic)"}
                               if wl.get("contact") == "algebraic" else {})),
                acceptance_rate=float(st[0] / st[1]) if st[1] else None,
                e2e=e2e,
                gpu_launches=steps, wall_ms=wall_ms, clocks=clocks,
                roofline=roofline)


def rex_leg(ctx, args, steps):
    """BASELINE.json configs[4]: a tempered ensemble of 8 inverse temperatures x (64 x n_gpus) columns of the
    1000-bead chromatin posterior, 512 replicas per GPU, one exchange attempt after every sweep.  Exchanges swap
    LABELS (temperature index, beta, step size): per attempt the ranks all-gather 16 bytes per chain over NCCL /
    NVLink, no state moves and nothing synchronises the host.  The ladder starts geometric in [0.05, 1] (where
    no exchange is ever accepted with 3000 degrees of freedom) and is re-spaced from the measured mean
    log-likelihoods during the untimed warm-up."""
    import torch
    from binf_b200 import _cabi
    from binf_b200.distributed import ChainShard, ReplicaExchangeDriver
    wl = make_hmc_workload(ctx, args, "rex")
    w, C, L, model = wl["w"], wl["C"], wl["L"], wl["model"]
    T = 8
    betas0 = [float(b) for b in np.geomspace(1.0, 0.05, T)]
    eps = torch.full((C,), w["eps"], device=ctx.dev, dtype=torch.float32)
    shard = ChainShard(model, wl["q"], wl["tau"], eps, L, gibbs_mode=wl["gibbs"], seed=args.seed,
                       chain_base=ctx.rank * C)
    drv = ReplicaExchangeDriver.for_shard(shard, ctx.rank, ctx.world, betas0, seed=args.seed + 1)
    eps.mul_(1.0 / torch.sqrt(shard.beta))               # step size ~ 1/sqrt(beta)
    # ---- untimed: equilibrate, measure log-likelihoods, re-space the ladder (twice) -----------------
    drv.run(20)
    drv.rex.reset_stats()
    drv.run(20)
    drv.adapt(target=0.35)
    for _ in range(2):
        drv.run(15)
        drv.rex.reset_stats()
        drv.run(25)
        drv.adapt(target=0.35)
    drv.run(15)
    drv.rex.reset_stats()
    ctx.sync_all()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
    steps = max(steps, 20)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    shard.stats.zero_()
    flush = torch.empty(256 << 20, device=ctx.dev, dtype=torch.uint8)   # 18 MB of state: flush L2 between iterations
    ctx.sync_all()
    for k in range(steps):
        flush.fill_(k)                          # (not timed)
        ev[k][0].record()
        shard.sweep()
        ev[k][1].record()
        drv.n_sweeps += 1
        drv.rex.swap(shard.last_chi2(), shard.tau, shard.eps, shard.n_data)
        ev[k][2].record()
    torch.cuda.synchronize()
    sweep_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    swap_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    total_ms = ctx.max_over_ranks(float(sum(e[0].elapsed_time(e[2]) for e in ev)))
    rates = drv.swap_rates()
    stats = shard.stats
    if ctx.world > 1:
        import torch.distributed as dist
        dist.all_reduce(stats)
    ctx.sync_all()
    clocks = sampler.summary() if sampler else None
    st = stats.cpu().numpy()
    model.close()
    if ctx.rank != 0:
        return None
    return dict(metric="HMC leapfrog steps/s (chains x steps), tempered ensemble with an exchange attempt per sweep",
                value=ctx.world * C * L * steps / (total_ms * 1e-3), unit="leapfrog steps/s", n_gpus=ctx.world,
                steps=steps, ms_per_step=total_ms / steps, sweep_ms=sweep_ms, swap_ms=swap_ms,
                swap_overhead_frac=swap_ms / sweep_ms,
                config=dict(workload=w["name"], replicas_per_gpu=C, temperatures=T, columns=drv.rex.n_columns,
                            rows_per_gpu=drv.rex.rows, ladder_start=betas0, ladder=drv.betas,
                            l2="L2 flushed (256 MiB fill) between timed steps",
                            exchange="label swap: all-gather of 16 B per chain per attempt (%d B per rank), "
                                     "no state moves" % (16 * C),
                            collective="NCCL all_gather_into_tensor" if ctx.world > 1 else
                                       "none (all temperatures on one device)"),
                swap_rates=rates, acceptance_rate=float(st[0] / st[1]) if st[1] else None,
                gpu_launches=steps * 4, clocks=clocks,
                roofline=chromatin_roofline(ctx, wl, sweep_ms, None))


def sink_leg(ctx, args, steps, warmup):
    """one step = one binfb_sink_push of the configs[2] state (4096 chains x 3000 dof): ring
    copy (every 2nd sweep) + float64 running moments + MAP tracking.  Roofline: HBM copy bandwidth."""
    import torch
    from binf_b200 import _cabi
    local, world, rank, dev = ctx.local, ctx.world, ctx.rank, ctx.dev
    C, D = args.chains or 4096, 3000
    sink = _cabi.Sink(C, D, capacity=8, burn_in=0, thin=2, track_map=True, device=local)
    g = torch.Generator(device=dev).manual_seed(args.seed)
    q = torch.randn(C, D, device=dev, generator=g)
    tau = torch.rand(C, device=dev, generator=g)
    logp = torch.randn(C, device=dev, generator=g, dtype=torch.float64)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(max(warmup, 3)):
        sink.push(q, tau, logp, stream)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # a push is ~100 us: launch it from a CUDA graph so that host-side launch jitter (e.g. the
    # nvidia-smi clock sampler holding the driver lock) is not mistaken for kernel time
    steps = max(steps, 50)
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for k in range(steps):
                logp.add_(0.01)                       # every sweep improves: MAP state rewritten each time
                sink.push(q, tau, logp, side.cuda_stream)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    with torch.cuda.stream(side):
        e0.record(side)
        for _ in range(reps):
            graph.replay()
        e1.record(side)
    torch.cuda.synchronize()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1) / (steps * reps))
    clocks = sampler.summary() if sampler else None
    # algorithmic bytes per element and push: read q 4, moments RMW 2 x (8 + 8), MAP state write 4,
    # ring write 4 on every 2nd sweep
    bytes_per_push = C * D * (4 + 32 + 4 + 2.0)
    peak = float(measured_peaks().get("hbm_gbs", 6543.7))
    achieved = bytes_per_push / (ms * 1e-3) / 1e9
    # e2e: the same push from pinned host memory (H2D inside the timed region)
    qh = q.cpu().pin_memory()
    th, lh = tau.cpu().pin_memory(), logp.cpu().pin_memory()
    t0 = time.perf_counter()
    n_e2e = 5
    for _ in range(n_e2e):
        q.copy_(qh, non_blocking=True), tau.copy_(th, non_blocking=True), logp.copy_(lh, non_blocking=True)
        sink.push(q, tau, logp, stream)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    sink.close()
    if rank != 0:
        return None
    return {
        "metric": "sample-sink state elements absorbed/s (chains x dim per sweep)", "value": world * C * D / (ms * 1e-3),
        "unit": "elements/s", "n_gpus": world, "steps": steps * reps, "warmup": max(warmup, 3) + steps,
        "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 state, f64 moments",
        "data": "synthetic",
        "config": {"workload": "sink_c%d_d%d_thin2_map" % (C, D), "chains_per_gpu": C, "dim": D,
                   "l2": "working set 295 MB of moments + 49 MB state per push > 126 MB L2"},
        "e2e": {"value": world * C * D / (e2e_ms * 1e-3), "unit": "elements/s",
                "h2d_bytes_per_step": C * D * 4 + C * 12, "d2h_bytes_per_step": 0, "steps": n_e2e,
                "api": "Sink.push after an H2D copy of the state from pinned host memory"},
        "gpu_launches": steps * reps, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic("sink"), "bytes_per_launch": bytes_per_push,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"},
        "cpu_baseline": None}


def brief(leg):
    """what an `extra` entry keeps of a leg: the numbers the headline line carries for its own workload"""
    if leg is None:
        return None
    keep = ("metric", "value", "unit", "steps", "ms_per_step", "config", "acceptance_rate", "e2e", "gpu_launches",
            "clocks", "roofline", "swap_rates", "sweep_ms", "swap_ms", "swap_overhead_frac", "dtype")
    return {k: leg[k] for k in keep if k in leg}


def run_ours(args):
    ctx = Ctx()
    line = None
    extras = {}
    if args.workload == "sink":
        line = sink_leg(ctx, args, args.steps, args.warmup)
    elif args.workload == "rex":
        line = rex_leg(ctx, args, args.steps)
    else:
        line = hmc_leg(ctx, args, args.workload, args.steps, args.warmup, chains=args.chains or None,
                       eps=args.eps or None, with_e2e=not args.no_e2e)
    # ---- the other configurations of BASELINE.json, short legs on the same box in the same run -----------
    if args.workload == "chromatin" and not args.no_extra and not args.chains and not args.roles \
            and args.contact == "logistic":
        k = max(3, min(args.steps, 6))
        import copy
        args_alg = copy.copy(args)
        args_alg.contact = "algebraic"   # SURVEY.md A.2's other contact function (2 MUFU + 21 FMA-pipe ops per pair)
        for name, fn in (("poly", lambda: hmc_leg(ctx, args, "poly", max(k, 10), 3)),
                         ("chromatin_algebraic", lambda: hmc_leg(ctx, args_alg, "chromatin", k, 3, eps=6.5e-3,
                                                                 with_e2e=False)),
                         ("generic", lambda: hmc_leg(ctx, args, "generic", max(k, 10), 3, with_e2e=False)),
                         ("chromatin5k", lambda: hmc_leg(ctx, args, "chromatin5k", 3, 3, with_e2e=False)),
                         ("sink", lambda: sink_leg(ctx, args, 50, 3)),
                         ("rex", lambda: rex_leg(ctx, args, 20))):
            try:
                extras[name] = brief(fn())
            except Exception as exc:  # a failing side leg must not take the headline with it
                extras[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            ctx.sync_all()
    if ctx.rank == 0:
        if args.workload in ("chromatin", "poly", "chromatin5k", "generic"):
            line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_seconds) \
                if ctx.world == 1 and not args.no_cpu and args.contact == "logistic" else None
        if extras:
            line["extra"] = extras
            line["gpu_launches"] = line["gpu_launches"] + sum(
                (e or {}).get("gpu_launches", 0) for e in extras.values())
        emit(json.dumps(line))
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="chromatin", choices=sorted(WORKLOADS) + ["sink"])  # rex: N >= 2
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--eps", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--roles", type=int, default=0, help="chromatin: force the warps per chain (experiments)")
    ap.add_argument("--ev-k", type=float, default=0.0, help="chromatin: excluded-volume strength (0 = off)")
    ap.add_argument("--chrom-sets", type=int, default=-1,
                    help="chromatin: 0 = all chain groups in one pass-major item sequence (experiments)")
    ap.add_argument("--contact", default="logistic", choices=["logistic", "algebraic"],
                    help="chromatin: contact function of the forward model (experiments; the headline is logistic)")
    ap.add_argument("--chrom-warps", type=int, default=0,
                    help="chromatin: cap on the chains per CTA (experiments)")
    ap.add_argument("--host-chunks", type=int, default=0,
                    help="chromatin e2e: pieces the state travels in through binfb_hmc_run_host (experiments)")
    ap.add_argument("--no-equilibrate", action="store_true", help="skip the untimed equilibration sweeps")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="default workload only: skip the short poly / chromatin5k / sink / rex legs")
    args = ap.parse_args()
    _capture_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload == "sink" and args.impl == "reference":
        emit(json.dumps({"impl": "reference", "unavailable": "the sink workload has no reference arm "
                          "(the reference keeps a Python list of deep copies of one chain)"}))
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
