"""Throughput of the chromatin kernel across bead counts (plan: roles per chain, chains per CTA)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from binf_b200 import _cabi  # noqa: E402

dev = torch.device("cuda")
peak = _cabi.microbench(0)["ffma_tflops"]
for n in (200, 500, 700, 1000, 1400, 2000, 3000, 5000):
    rng = np.random.RandomState(n)
    X = np.cumsum(rng.normal(size=(n, 3)), axis=0)
    M = n * (n - 1) // 2
    y = rng.uniform(size=M).astype(np.float32) * 0.1
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0)
    _, plan = _cabi.chromatin_stream_layout(n, None if False else y)
    C = 148 * plan["chains_per_cta"] * 4
    q = torch.as_tensor((X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))).astype(np.float32), device=dev)
    tau = torch.full((C,), 50.0, device=dev)
    eps = torch.full((C,), 1e-3, device=dev)
    L = 10
    st = torch.cuda.current_stream().cuda_stream

    def step(d):
        m.hmc_run_device(q, tau, eps, _cabi.HmcOpts(L, 1, 0, 1, 1.05, 0.95, 1, d, 0), stream=st)
    for d in range(2):
        step(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for d in range(3):
        step(2 + d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    tf = 31.0 * M * (L + 1) * C / (ms * 1e-3) / 1e12
    print("n=%5d roles=%2d chains/CTA=%2d C=%5d  %8.2f ms  %5.1f TFLOP/s = %4.1f %% of FFMA peak"
          % (n, plan["roles"], plan["chains_per_cta"], C, ms, tf, 100 * tf / peak), flush=True)
    m.close()
