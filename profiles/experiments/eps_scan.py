"""acceptance of the benchmark workloads at equilibrium as a function of the step size (bench.py's eps choice):
equilibrate with the fused sweep, then measure the acceptance of 8 sweeps per candidate step size, each from
a copy of the equilibrated state"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from binf_b200 import _cabi


class A:  # the arguments make_hmc_workload looks at
    roles, ev_k, chrom_sets = 0, 0.0, -1


ctx = bench.Ctx()
which = sys.argv[1] if len(sys.argv) > 1 else "chromatin"
n_eq = int(sys.argv[2]) if len(sys.argv) > 2 else 150
cands = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.004, 0.006, 0.008, 0.010, 0.012]
wl = bench.make_hmc_workload(ctx, A, which, chains=int(os.environ.get("CHAINS", "0")) or None)
C, L, model = wl["C"], wl["L"], wl["model"]
eps = torch.full((C,), float(os.environ.get("EQ_EPS", cands[0])), device=ctx.dev)
stream = torch.cuda.current_stream().cuda_stream
stats = wl["stats"]
for k in range(n_eq):
    model.hmc_run_device(wl["q"], wl["tau"], eps, _cabi.HmcOpts(L, 1, 0, wl["gibbs"], 1.05, 0.95, 5, k, 0), stats=stats, stream=stream)
    if (k + 1) % 25 == 0:
        torch.cuda.synchronize()
        s = stats.cpu().numpy(); stats.zero_()
        print("equilibration sweep %d: acceptance %.3f, tau %.1f" % (k + 1, s[0] / s[1], float(wl["tau"].mean())), flush=True)
q_eq, tau_eq = wl["q"].clone(), wl["tau"].clone()
for e in cands:
    q, tau = q_eq.clone(), tau_eq.clone()
    eps.fill_(e)
    stats.zero_()
    for k in range(8):
        model.hmc_run_device(q, tau, eps, _cabi.HmcOpts(L, 1, 0, wl["gibbs"], 1.05, 0.95, 6, 1000 + k, 0), stats=stats, stream=stream)
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    print("%s eps %.4g: acceptance %.3f  mean min(1,exp(-dH)) %.3f" % (which, e, s[0] / s[1], s[3] / s[1]), flush=True)
