#!/usr/bin/env python
"""Benchmark of the B200-native HMC hot path (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload chromatin|poly]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port)

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

FLOP_PER_PAIR = 31.0          # per unordered bead pair and force evaluation
FLOP_PER_DATUM = 14.0         # per chain, datum and force evaluation (K = 4)

WORKLOADS = {
    "chromatin": dict(name="chromatin_n1000_c4096_L20_gibbs", n_beads=1000, chains=4096, L=20,
                      eps=1.5e-3, alpha=2.0, d_c=2.5, k_bb=4.0, l0=1.0, noise=0.05),
    "poly": dict(name="poly_K4_N1000_c65536_L20", n_data=1000, chains=65536, L=20, eps=0.009,
                 tau=2.5),
    # BASELINE.json configs[3]: 5000 beads, chains sharded across the GPUs (4 chains per SM and GPU)
    "chromatin5k": dict(name="chromatin_n5000_c592_L20_gibbs", n_beads=5000, chains=592, L=20,
                        eps=5e-4, alpha=2.0, d_c=2.5, k_bb=4.0, l0=1.0, noise=0.05),
    # BASELINE.json configs[4]: one inverse temperature per rank (geometric in [0.05, 1]), 512 chains
    # per rank, a neighbour swap attempt (NCCL send/recv over NVLink) after every sweep
    "rex": dict(name="chromatin_n1000_rex_512_per_rank_L20", n_beads=1000, chains=512, L=20,
                eps=1.5e-3, alpha=2.0, d_c=2.5, k_bb=4.0, l0=1.0, noise=0.05),
}


# --------------------------------------------------------------------------------------------
# synthetic inputs; generated without touching oracle/ so the product arm
# never imports the checker
# --------------------------------------------------------------------------------------------
def chromatin_inputs(w, chains, seed):
    n = w["n_beads"]
    rng = np.random.RandomState(0)
    X = np.cumsum(rng.normal(size=(n, 3)) * w["l0"], axis=0)
    X -= X.mean(axis=0)
    i, j = np.triu_indices(n, 1)
    d = np.sqrt(np.sum((X[i] - X[j]) ** 2, axis=-1) + 1e-12)
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
    workload = "chromatin" if workload in ("rex", "chromatin5k") else workload
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
                config=dict(workload=w["name"], note="oracle port of the reference's numpy path; "
                            "the reference itself is Python 2 + CSB and cannot travel to the box"),
                cpu_baseline=best,
                e2e=dict(value=best["value"], unit="leapfrog steps/s", h2d_bytes_per_step=0,
                         d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t0)
    emit(json.dumps(line))


# --------------------------------------------------------------------------------------------
# the sample sink (HBM-bound), measured next to the sweep it follows
# --------------------------------------------------------------------------------------------
def run_sink(args):
    """One step = one binfb_sink_push of the configs[2] state (4096 chains x 3000 dof): ring copy
    (every 2nd sweep) + float64 running moments + MAP tracking.  Roofline: HBM copy bandwidth."""
    import torch
    from binf_b200 import _cabi
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    C, D = args.chains or 4096, 3000
    sink = _cabi.Sink(C, D, capacity=8, burn_in=0, thin=2, track_map=True, device=local)
    g = torch.Generator(device=dev).manual_seed(args.seed)
    q = torch.randn(C, D, device=dev, generator=g)
    tau = torch.rand(C, device=dev, generator=g)
    logp = torch.randn(C, device=dev, generator=g, dtype=torch.float64)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(max(args.warmup, 3)):
        sink.push(q, tau, logp, stream)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # a push is ~100 us: launch it from a CUDA graph so that host-side launch jitter (e.g. the
    # nvidia-smi clock sampler holding the driver lock) is not mistaken for kernel time
    steps = max(args.steps, 50)
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
    ms = e0.elapsed_time(e1) / (steps * reps)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.summary() if sampler else None
    # algorithmic bytes per element and push: read q 4, moments RMW 2 x (8 + 8), MAP state write 4,
    # ring write 4 on every 2nd sweep
    bytes_per_push = C * D * (4 + 32 + 4 + 2.0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6543.7))
    achieved = bytes_per_push / (ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["sink"]["dram_bytes_per_launch"]
    except Exception:
        pass
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
    if rank == 0:
        emit(json.dumps({
            "metric": "sample-sink state elements absorbed/s (chains x dim per sweep)", "value": world * C * D / (ms * 1e-3),
            "unit": "elements/s", "n_gpus": world, "steps": steps * reps, "warmup": max(args.warmup, 3) + steps,
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
                         "traffic": traffic, "bytes_per_launch": bytes_per_push,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"},
            "cpu_baseline": None}))


# the product arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from binf_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    C = args.chains or w["chains"]
    L = w["L"]
    eps0 = args.eps or w["eps"]
    chain_base = rank * C

    rex = None
    if args.workload in ("chromatin", "rex", "chromatin5k"):
        y, q_host = chromatin_inputs(w, C, rank)
        model = _cabi.Model.chromatin(w["n_beads"], y, w["alpha"], w["d_c"], w["k_bb"], w["l0"], 0.0,
                                      1.0, 1.0, device=local, roles=args.roles, ev_k=args.ev_k, ev_d=1.5)
        tau0, gibbs = 100.0, _cabi.GIBBS_TAU_FIRST
        units = float(model.n_data)            # pairs per force evaluation
        flop_per_launch = FLOP_PER_PAIR * units * (L + 1) * C
        bytes_per_launch = 2.0 * 4 * q_host.shape[1] * C + 4.0 * units
    else:
        xs, ys, q_host = poly_inputs(w, C, rank)
        model = _cabi.Model.polynomial(xs, ys, 4, np.zeros(4), 5 * np.ones(4), 1.0, 1.0, device=local)
        tau0, gibbs = w["tau"], _cabi.GIBBS_NONE
        units = float(w["n_data"])
        flop_per_launch = (FLOP_PER_DATUM * (L + 1) + 4) * units * C
        bytes_per_launch = 2.0 * 4 * 4 * C
    D = q_host.shape[1]

    q = torch.from_numpy(q_host).to(dev)
    tau = torch.full((C,), tau0, device=dev, dtype=torch.float32)
    eps = torch.full((C,), eps0, device=dev, dtype=torch.float32)
    accepted = torch.zeros(C, device=dev, dtype=torch.uint8)
    nacc = torch.zeros(C, device=dev, dtype=torch.int32)
    stats = torch.zeros(4, device=dev, dtype=torch.float64)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8) if args.workload == "poly" else None
    stream = torch.cuda.current_stream().cuda_stream
    draw = [0]

    beta = None
    if args.workload == "rex":
        from binf_b200.distributed import ReplicaExchange
        betas = [float(b) for b in np.geomspace(1.0, 0.05, world)] if world > 1 else [1.0]
        beta = torch.full((C,), betas[rank], device=dev, dtype=torch.float32)
        rex = ReplicaExchange(rank, world, betas[rank], seed=args.seed)
        chi2 = torch.zeros(C, device=dev, dtype=torch.float64)
        chain_base = rank * C

    launches = [0]

    def step():
        opts = _cabi.HmcOpts(L, 1, 0, gibbs, 1.05, 0.95, args.seed, draw[0], chain_base)
        model.hmc_run_device(q, tau, eps, opts, beta=beta, accepted=accepted, n_accepted=nacc,
                             stats=stats, stream=stream)
        draw[0] += 1
        launches[0] += 1
        if rex is not None and world > 1:
            model.logprob_grad_device(q, tau, chi2=chi2, stream=stream)
            t = tau.double()
            ll = -0.5 * t * chi2 + 0.5 * units * torch.log(t)
            launches[0] += 1 + (2 if rex.swap(q, tau, ll, betas) is not None else 0)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    stats.zero_()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    sync_all()
    t_wall = time.perf_counter()
    launches[0] = 0
    for k in range(args.steps):
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
    if world > 1:
        dist.all_reduce(stats)                  # the diagnostics reduction (4 doubles, NCCL)
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    sync_all()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.summary() if sampler else None
    st = stats.cpu().numpy()
    value = world * C * L * args.steps / (total_ms * 1e-3)

    # ---- e2e: the same sweep through the host-buffer C-ABI call (pinned host memory) --------
    e2e = None
    if not args.no_e2e and rex is None:
        qh = torch.from_numpy(q_host.copy()).pin_memory().numpy()
        th = torch.full((C,), tau0, dtype=torch.float32).pin_memory().numpy()
        eh = torch.full((C,), eps0, dtype=torch.float32).pin_memory().numpy()
        ah = np.empty(C, dtype=np.uint8)
        n_e2e = max(2, min(args.steps, 5))
        from ctypes import byref

        def host_step(k):
            opts = _cabi.HmcOpts(L, 1, 0, gibbs, 1.05, 0.95, args.seed, 1000 + k, chain_base)
            _cabi.check(_cabi.lib().binfb_hmc_run_host(
                model._h, _cabi.ptr(qh), _cabi.ptr(th), None, _cabi.ptr(eh), C, byref(opts), None, None,
                None, _cabi.ptr(ah), None, None, None, None, None, None))
        host_step(0)
        sync_all()
        t0 = time.perf_counter()
        for k in range(n_e2e):
            host_step(1 + k)
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = dict(value=world * C * L * n_e2e / el, unit="leapfrog steps/s",
                   h2d_bytes_per_step=int(4 * C * D + 8 * C),
                   # back: the state, the accept flags, the precision when the sweep updates it (the step sizes
                   # only while they adapt)
                   d2h_bytes_per_step=int(4 * C * D + C + (4 * C if gibbs else 0)),
                   steps=n_e2e, api="binfb_hmc_run_host (pinned host buffers in, synchronous)")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel ------------------------------------------------
    ms_kernel = float(np.mean(kernel_ms))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    mb = _cabi.microbench(local, 3000)
    achieved = flop_per_launch / (ms_kernel * 1e-3) / 1e12
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.workload in tj and C == w["chains"]:
            traffic = tj[args.workload]["dram_bytes_per_launch"]
    except Exception:
        pass
    sfu = None
    if args.workload in ("chromatin", "rex", "chromatin5k"):
        # the binding pipe of the pair kernel: 3 MUFU (rsqrt, ex2, rcp) per bead pair
        sfu_gops = 3.0 * units * (L + 1) * C / (ms_kernel * 1e-3) / 1e9
        sfu = dict(ops_per_pair=3, achieved_gops=sfu_gops, peak_gops=mb["mufu_gops"],
                   frac=sfu_gops / mb["mufu_gops"],
                   note="MUFU issues at 16 lanes/clk/SM: 24 SMSP-cycles per warp-pair vs 19 on the FMA "
                        "pipe, so the special-function unit bounds this kernel at 31/(2*24) = 64.6 % of "
                        "the FP32-FMA peak")
    roofline = dict(
        bound="fp32_fma", achieved=achieved, peak=mb["ffma_tflops"], unit="TFLOP/s",
        frac=achieved / mb["ffma_tflops"], traffic=traffic, sfu=sfu,
        peak_source="live FFMA issue microbenchmark in this run (binfb_microbench); "
                    "MEASURED_PEAKS.json has no non-tensor FP32 figure",
        peak_formula_tflops=148 * 128 * 2 * 1.965e9 / 1e12,
        ffma2_tflops=mb["ffma2_tflops"], mufu_gops=mb["mufu_gops"],
        flop_per_launch=flop_per_launch, kernel_ms=ms_kernel,
        hbm_sanity=dict(algorithmic_gb_per_launch=bytes_per_launch / 1e9,
                        achieved_gbs=bytes_per_launch / (ms_kernel * 1e-3) / 1e9,
                        peak_gbs=peaks.get("hbm_gbs"),
                        note="compulsory HBM bytes are O(chains x dim) per trajectory: not the bound"))
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.workload, args.cpu_seconds)
    line = dict(metric="HMC leapfrog steps/s (chains x steps)", value=value, unit="leapfrog steps/s",
                n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=total_ms / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=dict(workload=w["name"], chains_per_gpu=C, leapfrog_steps=L, timestep=eps0,
                            parallelism="chain-sharded x%d (no data-path collective)" % world,
                            l2="working set %.0f MB per step > 126 MB L2" % (3 * 4 * C * D / 1e6)
                            if flush is None else "L2 flushed (256 MiB fill) between timed steps",
                            gibbs="precision update fused in front of each trajectory"
                            if gibbs else "none",
                            replica_exchange=None if rex is None else dict(
                                betas=betas, swap_rate_rank0=rex.n_swapped / max(1, rex.n_attempted))),
                acceptance_rate=float(st[0] / st[1]) if st[1] else None,
                e2e=e2e, gpu_launches=launches[0], wall_ms=wall_ms, clocks=clocks, roofline=roofline,
                cpu_baseline=cpu)
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    _capture_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload == "sink":
        if args.impl == "reference":
            emit(json.dumps({"impl": "reference", "unavailable": "the sink workload has no reference arm "
                              "(the reference keeps a Python list of deep copies of one chain)"}))
        else:
            run_sink(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
