import sys, json, subprocess
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from binf_b200 import _cabi
xs = np.linspace(-2, 2, 1000); rng = np.random.RandomState(0)
ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2., -4., 1., 1.5]), 1/np.sqrt(2.5))
C = 65536
q0 = (np.ones((C, 4)) + 0.1*np.random.RandomState(1).normal(size=(C, 4))).astype(np.float32)
dev = torch.device('cuda')
for (J, G, blk) in [(2, 4, -1), (1, 2, -1), (4, 4, -1), (4, 8, -1), (2, 8, -1)]:
    m = _cabi.Model.polynomial(xs, ys, 4, np.zeros(4), 5*np.ones(4), 1.0, 1.0)
    m.set_option('poly.chains_per_thread', J); m.set_option('poly.group', G); m.set_option('poly.block', blk)
    q = torch.from_numpy(q0).to(dev); tau = torch.full((C,), 2.5, device=dev); eps = torch.full((C,), 0.009, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    def step(d):
        opts = _cabi.HmcOpts(20, 1, 0, 0, 1.05, 0.95, 1, d, 0)
        m.hmc_run_device(q, tau, eps, opts, stream=stream)
    for d in range(5): step(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for d in range(20): step(5+d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    print('J=%d G=%d block=%d: %.4f ms  frac %.3f' % (J, G, blk, ms, (14*21+4)*1000*C/(ms*1e-3)/72.5e12), flush=True)
