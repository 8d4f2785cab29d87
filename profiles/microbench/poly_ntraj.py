import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from binf_b200 import _cabi
xs = np.linspace(-2, 2, 1000); rng = np.random.RandomState(0)
ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2., -4., 1., 1.5]), 1/np.sqrt(2.5))
C = 65536
q0 = (np.ones((C, 4)) + 0.1*np.random.RandomState(1).normal(size=(C, 4))).astype(np.float32)
dev = torch.device('cuda')
m = _cabi.Model.polynomial(xs, ys, 4, np.zeros(4), 5*np.ones(4), 1.0, 1.0)
q = torch.from_numpy(q0).to(dev); tau = torch.full((C,), 2.5, device=dev); eps = torch.full((C,), 0.009, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for nt in (1, 2, 8, 32):
    def step(d):
        opts = _cabi.HmcOpts(20, nt, 0, 0, 1.05, 0.95, 1, d, 0)
        m.hmc_run_device(q, tau, eps, opts, stream=stream)
    for d in range(3): step(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for d in range(10): step(100+d*nt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print('n_traj=%2d: %.4f ms per launch, %.4f ms per trajectory, frac %.3f' % (nt, ms, ms/nt, (14*21+4)*1000*C/(ms/nt*1e-3)/72.5e12), flush=True)
