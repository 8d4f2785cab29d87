"""ctypes binding of libbinf_b200.so -- the only route from Python to the CUDA kernels.

There is no CPU fallback: if the library is missing or cannot be loaded, importing a
compute entry point raises.  Signatures mirror include/binf_b200.h one to one.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbinf_b200.so")

OK, EINVAL, ECUDA, EUNSUPPORTED, ENOMEM = 0, -1, -2, -3, -4
MODEL_POLYNOMIAL, MODEL_CHROMATIN, MODEL_GENERIC = 1, 2, 3
FLAG_PRIOR_GRAD = 1
FLAG_GENERIC_SCALAR = 2
FLAG_CONTACT_ALGEBRAIC = 4
GIBBS_NONE, GIBBS_TAU_FIRST, GIBBS_TAU_LAST = 0, 1, 2
SINK_TRACK_MAP = 1
REX_MAX_TEMPS, REX_RECORD_BYTES = 64, 16


class BinfB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("binf_b200 error %d: %s" % (code, message))
        self.code = code


class HmcOpts(C.Structure):
    _fields_ = [("n_steps", C.c_int32), ("n_traj", C.c_int32), ("n_adapt", C.c_int32),
                ("gibbs_mode", C.c_int32), ("adapt_up", C.c_double), ("adapt_down", C.c_double),
                ("seed", C.c_uint64), ("draw", C.c_uint64), ("chain_base", C.c_uint64)]


_vp, _i, _d, _u64, _ll = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_longlong
_pi, _pd, _pll = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_longlong)

# name -> (restype, argtypes); kept in the order of include/binf_b200.h
SIGNATURES = {
    "binfb_version": (_i, []),
    "binfb_last_error": (C.c_char_p, []),
    "binfb_device_count": (_i, [_pi]),
    "binfb_device_props": (_i, [_i, _pi, _pi, _pi, _pi]),
    "binfb_model_create_polynomial": (_i, [_vp, _vp, _i, _i, _vp, _vp, _d, _d, C.c_uint, _i,
                                           C.POINTER(_vp)]),
    "binfb_model_create_chromatin": (_i, [_i, _vp, _d, _d, _d, _d, _d, _d, _d, C.c_uint, _i,
                                          C.POINTER(_vp)]),
    "binfb_model_create_generic": (_i, [C.c_char_p, _i, _i, _vp, _vp, _i, _vp, _vp, _d, _d, C.c_uint, _i,
                                        C.POINTER(_vp)]),
    "binfb_generic_compile_check": (_i, [C.c_char_p, _i, _i, C.c_char_p, _i]),
    "binfb_model_destroy": (_i, [_vp]),
    "binfb_model_info": (_i, [_vp, _pi, _pi, _pll, _pi]),
    "binfb_model_set_gamma_prior": (_i, [_vp, _d, _d]),
    "binfb_model_set_option": (_i, [_vp, C.c_char_p, _d]),
    "binfb_model_get_option": (_i, [_vp, C.c_char_p, C.POINTER(_d)]),
    "binfb_logprob_grad": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "binfb_logprob_grad_host": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "binfb_forward_host": (_i, [_vp, _vp, _i, _vp]),
    "binfb_hmc_run": (_i, [_vp, _vp, _vp, _vp, _vp, _i, C.POINTER(HmcOpts)] + [_vp] * 11),
    "binfb_hmc_run_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, C.POINTER(HmcOpts)] + [_vp] * 10),
    "binfb_gibbs_precision": (_i, [_vp, _vp, _vp, _vp, _i, _u64, _u64, _u64, _vp, _vp, _vp]),
    "binfb_gibbs_precision_host": (_i, [_vp, _vp, _vp, _vp, _i, _u64, _u64, _u64, _vp, _vp]),
    "binfb_swap_decide": (_i, [_vp, _vp, _d, _d, _i, _u64, _u64, _u64, _u64, _vp, _vp]),
    "binfb_swap_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "binfb_hmc_last_chi2": (_i, [_vp, _i, _vp, _vp]),
    "binfb_rex_pack": (_i, [_vp, _vp, _vp, _vp, _i, _d, _vp, _vp]),
    "binfb_rex_decide": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _u64, _u64, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "binfb_rex_select": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "binfb_sink_create": (_i, [_i, _i, _i, _i, _i, C.c_uint, _i, C.POINTER(_vp)]),
    "binfb_sink_destroy": (_i, [_vp]),
    "binfb_sink_info": (_i, [_vp, _pll, _pll, _pll]),
    "binfb_sink_push": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "binfb_sink_push_host": (_i, [_vp, _vp, _vp, _vp]),
    "binfb_sink_summary": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "binfb_sink_summary_host": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "binfb_sink_sums_host": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "binfb_sink_moments_host": (_i, [_vp, _vp, _vp]),
    "binfb_sink_read_host": (_i, [_vp, _ll, _ll, _vp, _vp]),
    "binfb_sink_map_host": (_i, [_vp, _vp, _vp, _vp]),
    "binfb_rwmc_run": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _u64, _u64, _u64] + [_vp] * 6),
    "binfb_rwmc_run_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _u64, _u64, _u64] + [_vp] * 5),
    "binfb_posterior_predictive_host": (_i, [_vp, _vp, _ll, _i, _vp, _vp, _i, _vp, _i]),
    "binfb_rng_fill_host": (_i, [_u64, _u64, _u64, _i, _i, _d, _vp, _vp, _vp, _i]),
    "binfb_chromatin_stream_layout": (_i, [_i, _vp, _i, _i, _vp, _ll, _pll, _pi]),
    "binfb_microbench": (_i, [_i, _i, _pd, _pd, _pd, _pd]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libbinf_b200.so not found at %s -- build it with `python -m binf_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code):
    if code != OK:
        raise BinfB200Error(code, lib().binfb_last_error().decode())


def ptr(a):
    """void* of a numpy array (None -> NULL), an int device pointer, or a torch tensor."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())  # torch tensor


def f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


_TORCH_NAMES = {"float32": "torch.float32", "float64": "torch.float64", "int32": "torch.int32",
                "uint8": "torch.uint8"}


def dev_arg(name, t, device, dtype, shape):
    """Validate one tensor argument of a device entry point and return it.  The kernels take raw pointers, so
    a float64, non-contiguous, wrong-device or wrongly shaped tensor would be silently reinterpreted (or read
    out of bounds); `None` passes through (optional outputs)."""
    if t is None:
        return None
    if not hasattr(t, "data_ptr"):
        raise ValueError("%s: expected a CUDA torch tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda or t.device.index != device:
        raise ValueError("%s: tensor lives on %s, the model on cuda:%d" % (name, t.device, device))
    if str(t.dtype) != _TORCH_NAMES[dtype]:
        raise ValueError("%s: dtype %s, expected %s" % (name, t.dtype, _TORCH_NAMES[dtype]))
    if not t.is_contiguous():
        raise ValueError("%s: tensor is not contiguous" % name)
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s: shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return t


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    n = C.c_int(0)
    check(lib().binfb_device_count(C.byref(n)))
    return n.value


def device_props(device=0):
    sm, smem, clk, cc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib().binfb_device_props(device, C.byref(sm), C.byref(smem), C.byref(clk), C.byref(cc)))
    return dict(sm_count=sm.value, smem_optin=smem.value, clock_khz=clk.value, cc=cc.value)


def microbench(device=0, iters=2000):
    a, b, c, d = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    check(lib().binfb_microbench(device, iters, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
    return dict(ffma_tflops=a.value, ffma2_tflops=b.value, mufu_gops=c.value, sm_clock_mhz=d.value)


def chromatin_stream_layout(n_beads, y_pairs, roles=0, smem_bytes=0):
    """Host-only: the contact stream exactly as the pair kernel consumes it, and the plan."""
    y = f32(y_pairs)
    n_floats = C.c_longlong()
    plan = (C.c_int * 8)()
    check(lib().binfb_chromatin_stream_layout(n_beads, ptr(y), roles, smem_bytes, None, 0,
                                              C.byref(n_floats), plan))
    out = np.empty(n_floats.value, dtype=np.float32)
    check(lib().binfb_chromatin_stream_layout(n_beads, ptr(y), roles, smem_bytes, ptr(out), out.size,
                                              None, None))
    keys = ("quads", "partner_steps", "row_blocks", "roles", "slots_per_row_block", "chains_per_cta",
            "stage_steps", "ring_depth")
    return out, dict(zip(keys, list(plan)))


def rng_fill(seed, draw, chain_base, n_chains, dim, gamma_shape=1.0, device=0):
    normals = np.empty((n_chains, dim), dtype=np.float32)
    uniforms = np.empty(n_chains, dtype=np.float32)
    gammas = np.empty(n_chains, dtype=np.float64)
    check(lib().binfb_rng_fill_host(seed, draw, chain_base, n_chains, dim, gamma_shape,
                                    ptr(normals), ptr(uniforms), ptr(gammas), device))
    return normals, uniforms, gammas


class Model(object):
    """Owning wrapper of a `binfb_model*`."""

    def __init__(self, handle):
        self._h = handle
        kind, dim, n_data, dev = C.c_int(), C.c_int(), C.c_longlong(), C.c_int()
        check(lib().binfb_model_info(self._h, C.byref(kind), C.byref(dim), C.byref(n_data),
                                     C.byref(dev)))
        self.kind, self.dim, self.n_data, self.device = kind.value, dim.value, n_data.value, dev.value

    @classmethod
    def polynomial(cls, xs, ys, n_coeff, prior_mean=None, prior_var=None, gamma_shape=1.0,
                   gamma_rate=1.0, flags=0, device=0):
        xs, ys = f64(xs), f64(ys)
        pm, pv = f64(prior_mean), f64(prior_var)
        h = C.c_void_p()
        check(lib().binfb_model_create_polynomial(ptr(xs), ptr(ys), len(xs), n_coeff, ptr(pm),
                                                  ptr(pv), gamma_shape, gamma_rate, flags, device,
                                                  C.byref(h)))
        return cls(h)

    @classmethod
    def chromatin(cls, n_beads, y_pairs, alpha, d_c, k_bb, l0, conf_s=0.0, gamma_shape=1.0,
                  gamma_rate=1.0, flags=0, device=0, roles=0, ev_k=0.0, ev_d=0.0, contact="logistic"):
        """roles: warps per chain (0 = heuristic; forcing it is a test hook, flags bits 8..12);
        ev_k, ev_d: excluded-volume prior -ev_k sum max(0, ev_d - d_ij)^4 (0 = off);
        contact: "logistic" 1/(1+exp(-z)) or "algebraic" 1/2 (1 + z/sqrt(1+z^2)), z = alpha (d_c - d)."""
        if contact not in ("logistic", "algebraic"):
            raise ValueError("contact: 'logistic' or 'algebraic', got %r" % (contact,))
        flags |= (roles & 0x1f) << 8
        if contact == "algebraic":
            flags |= FLAG_CONTACT_ALGEBRAIC
        y = f32(y_pairs)
        assert y.shape == (n_beads * (n_beads - 1) // 2,)
        h = C.c_void_p()
        check(lib().binfb_model_create_chromatin(n_beads, ptr(y), alpha, d_c, k_bb, l0, conf_s,
                                                 gamma_shape, gamma_rate, flags, device,
                                                 C.byref(h)))
        model = cls(h)
        if ev_k > 0.0:
            model.set_option("chrom.ev_k", ev_k)
            model.set_option("chrom.ev_d", ev_d)
        return model

    @classmethod
    def generic(cls, device_code, n_params, xs, ys, prior_mean=None, prior_var=None, gamma_shape=1.0,
                gamma_rate=1.0, flags=0, device=0):
        """User-defined per-datum forward model given as CUDA device code (compiled with NVRTC)."""
        xs, ys = f64(xs), f64(ys)
        xs = xs.reshape(len(ys), -1)
        pm, pv = f64(prior_mean), f64(prior_var)
        h = C.c_void_p()
        check(lib().binfb_model_create_generic(device_code.encode(), n_params, xs.shape[1], ptr(xs), ptr(ys),
                                               len(ys), ptr(pm), ptr(pv), gamma_shape, gamma_rate, flags,
                                               device, C.byref(h)))
        return cls(h)

    def close(self):
        if self._h is not None:
            lib().binfb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gamma_prior(self, shape, rate):
        check(lib().binfb_model_set_gamma_prior(self._h, shape, rate))

    def set_option(self, key, value):
        check(lib().binfb_model_set_option(self._h, key.encode(), float(value)))

    def get_option(self, key):
        v = C.c_double()
        check(lib().binfb_model_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    # ---- host-buffer entry points (numpy in, numpy out) ------------------------------------
    def logprob_grad(self, q, tau, beta=None, want_grad=True):
        q = f32(q).reshape(-1, self.dim)
        n = q.shape[0]
        tau = np.ascontiguousarray(np.broadcast_to(f32(tau), (n,)))
        beta = None if beta is None else np.ascontiguousarray(np.broadcast_to(f32(beta), (n,)))
        logp, chi2 = np.empty(n), np.empty(n)
        grad = np.empty_like(q) if want_grad else None
        check(lib().binfb_logprob_grad_host(self._h, ptr(q), ptr(tau), ptr(beta), n, ptr(logp),
                                            ptr(grad), ptr(chi2)))
        return logp, grad, chi2

    def forward(self, q):
        q = f32(q).reshape(-1, self.dim)
        mock = np.empty((q.shape[0], self.n_data), dtype=np.float32)
        check(lib().binfb_forward_host(self._h, ptr(q), q.shape[0], ptr(mock)))
        return mock

    def hmc_run(self, q, tau, eps, n_steps, n_traj=1, beta=None, p0=None, u=None,
                gamma_draws=None, n_adapt=0, adapt_up=1.05, adapt_down=0.95,
                gibbs_mode=GIBBS_NONE, seed=0, draw=0, chain_base=0, want_end=False):
        """Returns a dict; q/tau/eps are NOT modified in place (copies are returned)."""
        q = np.array(f32(q).reshape(-1, self.dim))
        n = q.shape[0]
        tau = np.array(np.broadcast_to(f32(tau), (n,)))
        eps = np.array(np.broadcast_to(f32(eps), (n,)))
        beta = None if beta is None else np.ascontiguousarray(np.broadcast_to(f32(beta), (n,)))
        p0 = None if p0 is None else f32(p0).reshape(n, self.dim)
        u = None if u is None else f32(u).reshape(n)
        gamma_draws = None if gamma_draws is None else f64(gamma_draws).reshape(n)
        opts = HmcOpts(n_steps, n_traj, n_adapt, gibbs_mode, adapt_up, adapt_down, seed, draw,
                       chain_base)
        accepted = np.empty(n, dtype=np.uint8)
        e0, e1 = np.empty(n), np.empty(n)
        q_end = np.empty_like(q) if want_end else None
        p_end = np.empty_like(q) if want_end else None
        nacc = np.empty(n, dtype=np.int32)
        stats = np.zeros(4)
        check(lib().binfb_hmc_run_host(self._h, ptr(q), ptr(tau), ptr(beta), ptr(eps), n,
                                       C.byref(opts), ptr(p0), ptr(u), ptr(gamma_draws),
                                       ptr(accepted), ptr(e0), ptr(e1), ptr(q_end), ptr(p_end),
                                       ptr(nacc), ptr(stats)))
        return dict(q=q, tau=tau, eps=eps, accepted=accepted.astype(bool), e_before=e0,
                    e_after=e1, q_end=q_end, p_end=p_end, n_accepted=nacc, stats=stats)

    def gibbs_precision(self, q, tau, beta=None, gamma_draws=None, seed=0, draw=0, chain_base=0):
        q = f32(q).reshape(-1, self.dim)
        n = q.shape[0]
        tau = np.array(np.broadcast_to(f32(tau), (n,)))
        beta = None if beta is None else np.ascontiguousarray(np.broadcast_to(f32(beta), (n,)))
        gamma_draws = None if gamma_draws is None else f64(gamma_draws).reshape(n)
        chi2 = np.empty(n)
        check(lib().binfb_gibbs_precision_host(self._h, ptr(q), ptr(tau), ptr(beta), n, seed, draw,
                                               chain_base, ptr(gamma_draws), ptr(chi2)))
        return tau, chi2

    def rwmc_run(self, q, tau, stepsize, n_moves=1, beta=None, change=None, u=None, seed=0, draw=0,
                 chain_base=0):
        """Random-walk Metropolis moves (RWMCSampler.sample); returns a dict, q is not modified."""
        q = np.array(f32(q).reshape(-1, self.dim))
        n = q.shape[0]
        tau = np.ascontiguousarray(np.broadcast_to(f32(tau), (n,)))
        stepsize = np.ascontiguousarray(np.broadcast_to(f32(stepsize), (n,)))
        beta = None if beta is None else np.ascontiguousarray(np.broadcast_to(f32(beta), (n,)))
        change = None if change is None else f32(change).reshape(n, self.dim)
        u = None if u is None else f32(u).reshape(n)
        accepted, nacc, logp = np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.int32), np.empty(n)
        check(lib().binfb_rwmc_run_host(self._h, ptr(q), ptr(tau), ptr(beta), ptr(stepsize), n, n_moves,
                                        seed, draw, chain_base, ptr(change), ptr(u), ptr(accepted),
                                        ptr(nacc), ptr(logp)))
        return dict(q=q, accepted=accepted.astype(bool), n_accepted=nacc, logp=logp)

    def _chains_of(self, q):
        """number of chains of a device state [C, dim] (validated)"""
        if not hasattr(q, "data_ptr") or q.dim() != 2 or q.shape[1] != self.dim:
            raise ValueError("q: expected a CUDA tensor of shape [n_chains, %d], got %s"
                             % (self.dim, tuple(getattr(q, "shape", ())) or type(q).__name__))
        n = int(q.shape[0])
        dev_arg("q", q, self.device, "float32", (n, self.dim))
        return n

    def rwmc_run_device(self, q, tau, stepsize, n_moves=1, beta=None, seed=0, draw=0, chain_base=0,
                        accepted=None, n_accepted=None, logp=None, stream=None):
        n, d = self._chains_of(q), self.device
        dev_arg("tau", tau, d, "float32", (n,)), dev_arg("stepsize", stepsize, d, "float32", (n,))
        dev_arg("beta", beta, d, "float32", (n,)), dev_arg("accepted", accepted, d, "uint8", (n,))
        dev_arg("n_accepted", n_accepted, d, "int32", (n,)), dev_arg("logp", logp, d, "float64", (n,))
        check(lib().binfb_rwmc_run(self._h, ptr(q), ptr(tau), ptr(beta), ptr(stepsize), q.shape[0], n_moves,
                                   seed, draw, chain_base, None, None, ptr(accepted), ptr(n_accepted),
                                   ptr(logp), ptr(stream)))

    # ---- device-pointer entry points (torch CUDA tensors; asynchronous on `stream`) ----------
    def hmc_run_device(self, q, tau, eps, opts, beta=None, p0=None, u=None, gamma_draws=None,
                       accepted=None, e_before=None, e_after=None, q_end=None, p_end=None,
                       n_accepted=None, stats=None, stream=None):
        n, d = self._chains_of(q), self.device
        dev_arg("tau", tau, d, "float32", (n,)), dev_arg("eps", eps, d, "float32", (n,))
        dev_arg("beta", beta, d, "float32", (n,)), dev_arg("p0", p0, d, "float32", (n, self.dim))
        dev_arg("u", u, d, "float32", (n,)), dev_arg("gamma_draws", gamma_draws, d, "float64", (n,))
        dev_arg("accepted", accepted, d, "uint8", (n,)), dev_arg("e_before", e_before, d, "float64", (n,))
        dev_arg("e_after", e_after, d, "float64", (n,)), dev_arg("q_end", q_end, d, "float32", (n, self.dim))
        dev_arg("p_end", p_end, d, "float32", (n, self.dim)), dev_arg("n_accepted", n_accepted, d, "int32", (n,))
        dev_arg("stats", stats, d, "float64", (4,))
        check(lib().binfb_hmc_run(self._h, ptr(q), ptr(tau), ptr(beta), ptr(eps), q.shape[0],
                                  C.byref(opts), ptr(p0), ptr(u), ptr(gamma_draws), ptr(accepted),
                                  ptr(e_before), ptr(e_after), ptr(q_end), ptr(p_end),
                                  ptr(n_accepted), ptr(stats), ptr(stream)))

    def logprob_grad_device(self, q, tau, beta=None, logp=None, grad=None, chi2=None, stream=None):
        n, d = self._chains_of(q), self.device
        dev_arg("tau", tau, d, "float32", (n,)), dev_arg("beta", beta, d, "float32", (n,))
        dev_arg("logp", logp, d, "float64", (n,)), dev_arg("grad", grad, d, "float32", (n, self.dim))
        dev_arg("chi2", chi2, d, "float64", (n,))
        check(lib().binfb_logprob_grad(self._h, ptr(q), ptr(tau), ptr(beta), q.shape[0], ptr(logp),
                                       ptr(grad), ptr(chi2), ptr(stream)))

    def hmc_last_chi2(self, chi2, stream=None):
        """chi2 [C] f64 <- chi^2 of every chain's current state, as left by the last hmc_run_device"""
        n = int(chi2.shape[0])
        dev_arg("chi2", chi2, self.device, "float64", (n,))
        check(lib().binfb_hmc_last_chi2(self._h, n, ptr(chi2), ptr(stream)))

    def gibbs_precision_device(self, q, tau, chi2, beta=None, gamma_draws=None, seed=0, draw=0,
                               chain_base=0, stream=None):
        n, d = self._chains_of(q), self.device
        dev_arg("tau", tau, d, "float32", (n,)), dev_arg("chi2", chi2, d, "float64", (n,))
        dev_arg("beta", beta, d, "float32", (n,)), dev_arg("gamma_draws", gamma_draws, d, "float64", (n,))
        check(lib().binfb_gibbs_precision(self._h, ptr(q), ptr(tau), ptr(beta), q.shape[0], seed,
                                          draw, chain_base, ptr(gamma_draws), ptr(chi2),
                                          ptr(stream)))


def generic_compile_check(device_code, n_params, x_dim=1):
    """(ok, log): run NVRTC on a generic model's device code without touching a GPU"""
    buf = C.create_string_buffer(1 << 16)
    rc = lib().binfb_generic_compile_check(device_code.encode(), n_params, x_dim, buf, len(buf))
    msg = buf.value.decode(errors="replace")
    if rc != OK and not msg:
        msg = lib().binfb_last_error().decode(errors="replace")
    return rc == OK, msg


def posterior_predictive(coeffs, precision, xs, ys, device=0):
    """predict (binf/example/misc.py:3-16) for many points at once"""
    coeffs = f32(coeffs)
    coeffs = coeffs.reshape(-1, coeffs.shape[-1])
    precision = f32(precision).reshape(-1)
    assert len(precision) == len(coeffs)
    xs, ys = np.atleast_1d(f64(xs)), np.atleast_1d(f64(ys))
    assert xs.shape == ys.shape
    out = np.empty(xs.size)
    check(lib().binfb_posterior_predictive_host(ptr(coeffs), ptr(precision), len(coeffs), coeffs.shape[1],
                                                ptr(np.ascontiguousarray(xs.ravel())),
                                                ptr(np.ascontiguousarray(ys.ravel())), xs.size, ptr(out), device))
    return out.reshape(xs.shape)


class Sink(object):
    """Owning wrapper of a `binfb_sink*` (sample ring, running moments, MAP candidates)."""

    def __init__(self, n_chains, dim, capacity=0, burn_in=0, thin=1, track_map=False, device=0):
        h = C.c_void_p()
        check(lib().binfb_sink_create(n_chains, dim, capacity, burn_in, thin,
                                      SINK_TRACK_MAP if track_map else 0, device, C.byref(h)))
        self._h = h
        self.device = device
        self.n_chains, self.dim, self.capacity, self.burn_in, self.thin = n_chains, dim, capacity, burn_in, thin
        self.track_map = track_map

    def close(self):
        if self._h is not None:
            lib().binfb_sink_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
        check(lib().binfb_sink_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(n_pushed=a.value, n_moment=b.value, n_kept=c.value)

    def push(self, q, aux=None, logp=None, stream=None):
        """q/aux/logp: torch CUDA tensors (device path, asynchronous) or numpy arrays (host path)."""
        if hasattr(q, "data_ptr"):
            dev_arg("q", q, self.device, "float32", (self.n_chains, self.dim))
            dev_arg("aux", aux, self.device, "float32", (self.n_chains,))
            dev_arg("logp", logp, self.device, "float64", (self.n_chains,))
            check(lib().binfb_sink_push(self._h, ptr(q), ptr(aux), ptr(logp), ptr(stream)))
        else:
            q = f32(q).reshape(self.n_chains, self.dim)
            aux = None if aux is None else np.ascontiguousarray(np.broadcast_to(f32(aux), (self.n_chains,)))
            logp = None if logp is None else np.ascontiguousarray(np.broadcast_to(f64(logp), (self.n_chains,)))
            check(lib().binfb_sink_push_host(self._h, ptr(q), ptr(aux), ptr(logp)))

    def summary(self):
        out = [np.empty(self.dim) for _ in range(4)]
        check(lib().binfb_sink_summary_host(self._h, *[ptr(o) for o in out]))
        return dict(mean=out[0], var=out[1], rhat=out[2], ess_per_chain=out[3])

    def sums(self):
        """raw per-dimension sums of the summary (see binf_b200.distributed.merge_sink_sums)"""
        out = [np.empty(self.dim) for _ in range(4)]
        check(lib().binfb_sink_sums_host(self._h, *[ptr(o) for o in out]))
        return dict(pivot=out[0], s1=out[1], s2=out[2], s3=out[3], n_chains=self.n_chains,
                    n=self.info()["n_moment"])

    def summary_device(self, mean=None, var=None, rhat=None, ess=None, stream=None):
        check(lib().binfb_sink_summary(self._h, ptr(mean), ptr(var), ptr(rhat), ptr(ess), ptr(stream)))

    def moments(self):
        mean, var = np.empty((self.n_chains, self.dim)), np.empty((self.n_chains, self.dim))
        check(lib().binfb_sink_moments_host(self._h, ptr(mean), ptr(var)))
        return mean, var

    def read(self, first=None, count=None):
        """kept samples still in the ring (default: all of them): q [count, C, dim], aux [count, C]"""
        n_kept = self.info()["n_kept"]
        oldest = max(0, n_kept - self.capacity)
        first = oldest if first is None else first
        count = n_kept - first if count is None else count
        q = np.empty((count, self.n_chains, self.dim), dtype=np.float32)
        aux = np.empty((count, self.n_chains), dtype=np.float32)
        check(lib().binfb_sink_read_host(self._h, first, count, ptr(q), ptr(aux)))
        return q, aux

    def map_estimate(self):
        logp = np.empty(self.n_chains)
        q = np.empty((self.n_chains, self.dim), dtype=np.float32)
        aux = np.empty(self.n_chains, dtype=np.float32)
        check(lib().binfb_sink_map_host(self._h, ptr(logp), ptr(q), ptr(aux)))
        return logp, q, aux
