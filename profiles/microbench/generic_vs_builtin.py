"""The user-defined-model path (NVRTC, generic_kernel.cuh) against the built-in polynomial kernel on the same
cubic: 65,536 chains x 1000 data points, L = 20."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from binf_b200 import _cabi
CODE = """
__device__ float binfb_mock(const float *theta, const float *x, float *dmock) {
    const float t = x[0];
    dmock[0] = 1.0f; dmock[1] = t; dmock[2] = t * t; dmock[3] = t * t * t;
    return fmaf(fmaf(fmaf(theta[3], t, theta[2]), t, theta[1]), t, theta[0]);
}
"""
CODE_POWERS = """
__device__ float binfb_mock(const float *theta, const float *x, float *dmock) {
    // the abscissae carry x, x^2, x^3 (what the built-in model precomputes on the host)
    dmock[0] = 1.0f; dmock[1] = x[0]; dmock[2] = x[1]; dmock[3] = x[2];
    return fmaf(theta[3], x[2], fmaf(theta[2], x[1], fmaf(theta[1], x[0], theta[0])));
}
"""
xs = np.linspace(-2, 2, 1000); rng = np.random.RandomState(0)
ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2., -4., 1., 1.5]), 1/np.sqrt(2.5))
C = 65536
q0 = (np.ones((C, 4)) + 0.1*np.random.RandomState(1).normal(size=(C, 4))).astype(np.float32)
dev = torch.device('cuda')
def gen(code, x, flags=0, rows="smem", split=1):
    os.environ["BINFB_GENERIC_ROWS"] = rows   # experiment switch read at model creation (generic.cu)
    m = _cabi.Model.generic(code, 4, x, ys, np.zeros(4), 5*np.ones(4), 1.0, 1.0, flags=flags)
    m.set_option("generic.split", split)      # 1: begin / middle / end kernels per trajectory, 0: one fused launch
    return m
x1, x3, P, S = xs[:, None].copy(), np.stack([xs, xs**2, xs**3], 1), 0, _cabi.FLAG_GENERIC_SCALAR
models = {"builtin": _cabi.Model.polynomial(xs, ys, 4, np.zeros(4), 5*np.ones(4), 1.0, 1.0),
          "generic (default: pairs, one launch, smem)": gen(CODE, x1, P, split=-1),
          "generic, powers as abscissae (default)": gen(CODE_POWERS, x3, P, split=-1),
          "generic, one launch, smem, one chain": gen(CODE, x1, S, split=0),
          "generic, split, chain pairs": gen(CODE, x1, P),
          "generic, split, chain pairs, powers": gen(CODE_POWERS, x3, P),
          "generic, split, one chain per lane": gen(CODE, x1, S),
          "generic, one launch, const, chain pairs": gen(CODE, x1, P, rows="const", split=0),
          "generic, one launch, const, one chain": gen(CODE, x1, S, rows="const", split=0)}
stream = torch.cuda.current_stream().cuda_stream
for name, m in list(models.items()) + list(reversed(list(models.items()))):
    q = torch.from_numpy(q0).to(dev); tau = torch.full((C,), 2.5, device=dev); eps = torch.full((C,), 0.009, device=dev)
    def step(d):
        opts = _cabi.HmcOpts(20, 1, 0, 0, 1.05, 0.95, 1, d, 0)
        m.hmc_run_device(q, tau, eps, opts, stream=stream)
    for d in range(10): step(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for d in range(40): step(10 + d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    print('%-44s %.4f ms per trajectory of %d chains (%.2f G leapfrog steps/s)' % (name, ms, C, C * 20 / ms / 1e6), flush=True)
