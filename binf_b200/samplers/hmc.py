"""Batched Hamiltonian Monte Carlo on the device behind the reference's HMCSampler interface
(reference: binf/samplers/hmc.py:15-191).

`HMCSampler(pdf, state, timestep, nsteps, ...)` keeps the reference's constructor, attributes and
methods.  What changes underneath: `sample()` lowers `pdf` once (binf_b200.lowering) and runs the
whole transition -- momentum draw, L leapfrog steps with L+1 fused force evaluations, Metropolis
test, step-size adaption -- as ONE launch of the fused CUDA kernel through the C ABI
(binfb_hmc_run[_host]), for every chain of the batch at once.

State conventions
  * `state` of shape (D,) (numpy): one chain, `sample()` returns a (D,) float64 array like the
    reference.
  * `state` of shape (C, D) (numpy): C independent chains, `sample()` returns (C, D) float64.
  * `state` a CUDA torch tensor (C, D) float32: chains stay resident in HBM, `sample()` returns a
    tensor; nothing crosses PCIe.
The per-chain precision is read from the pdf's bound `precision` parameter at every call, scalar
or of shape (C,).
"""
from collections import namedtuple

import numpy as np

from binf_b200 import _cabi
from binf_b200.lowering import lower, _is_tensor

HMCSampleStats = namedtuple("HMCSampleStats", "accepted stepsize")


class HMCSampler(object):
    def __init__(self, pdf, state, timestep, nsteps, timestep_adaption_limit=0,
                 adaption_uprate=1.05, adaption_downrate=0.95, variable_name=None, seed=None,
                 chain_base=0):
        self._pdf = None
        self._lowered = None
        self.pdf = pdf
        self._state = None
        self._eps = None
        self._eps_dev = None
        self.nsteps = nsteps
        self.timestep_adaption_limit = timestep_adaption_limit
        self.adaption_uprate = adaption_uprate
        self.adaption_downrate = adaption_downrate
        self._variable_name = variable_name
        self._timestep0 = timestep
        self.state = state
        self._last_move_accepted = 0
        self.n_accepted = 0
        self.counter = 0
        # Philox key: drawn from numpy's global RNG (the reference's only source of randomness,
        # hmc.py:146,151) unless given, so that np.random.seed(...) still makes runs repeatable
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.chain_base = int(chain_base)
        self._draw = 0
        self.last_energies = None

    # -- pdf / state ------------------------------------------------------------------------------
    @property
    def pdf(self):
        return self._pdf

    @pdf.setter
    def pdf(self, value):
        self._pdf = value
        self._lowered = None  # lowered lazily at the next sample()

    @property
    def state(self):
        return self._state

    @state.setter
    def state(self, value):
        self._state = value
        n = self._n_chains()
        if n is not None and (self._eps is None or len(self._eps) != n):
            eps = np.asarray(self._timestep0 if self._eps is None else self.timestep, dtype=np.float64)
            self._eps = np.array(np.broadcast_to(eps, (n,)), dtype=np.float64)
            self._eps_dev = None

    def _n_chains(self):
        s = self._state
        if s is None or np.ndim(s) == 0 and not _is_tensor(s):
            return None
        return 1 if len(s.shape) == 1 else int(s.shape[0])

    @property
    def timestep(self):
        if self._eps is None:
            return self._timestep0
        if self._eps_dev is not None:
            self._eps = self._eps_dev.detach().cpu().numpy().astype(np.float64)
        return float(self._eps[0]) if len(self._eps) == 1 else self._eps.copy()

    @timestep.setter
    def timestep(self, value):
        self._timestep0 = value
        if self._eps is not None:
            self._eps = np.array(np.broadcast_to(np.asarray(value, dtype=np.float64), self._eps.shape))
            self._eps_dev = None

    @property
    def acceptance_rate(self):
        if self.counter > 0:
            n = self.n_accepted
            mean = float(n.double().mean()) if _is_tensor(n) else float(np.mean(n))
            return mean / float(self.counter)
        return 0.0

    @property
    def variable_name(self):
        return "HMC" if self._variable_name is None else self._variable_name

    @property
    def last_move_accepted(self):
        return self._last_move_accepted

    @property
    def last_draw_stats(self):
        return {self.variable_name: HMCSampleStats(self.last_move_accepted, self.timestep)}

    # -- lowering -------------------------------------------------------------------------------
    def _lower(self):
        if self._lowered is None:
            dim = int(self._state.shape[-1])
            low = lower(self._pdf, n_coeff=dim)
            if low is None:
                raise NotImplementedError(
                    "HMCSampler: %r is not made of model classes the CUDA kernels implement "
                    "(polynomial / chromatin forward model + GaussianErrorModel + Gamma/Gaussian/"
                    "backbone priors); there is no CPU fallback" % (self._pdf,))
            if self._variable_name is not None and low.variable != self._variable_name:
                raise ValueError("sampler variable %r does not match the pdf's sampled variable %r"
                                 % (self._variable_name, low.variable))
            if low.dim != dim:
                raise ValueError("state has dimension %d, the model %d" % (dim, low.dim))
            self._lowered = low
        self._lowered.refresh()
        return self._lowered

    # -- the reference's two methods -----------------------------------------------------------
    def _leapfrog(self, q, p, timestep, nsteps):
        """Integrate Hamilton's equations for `nsteps` leapfrog steps (hmc.py:92-125); returns
        the end point (q, p).  One fused launch with injected momenta."""
        low = self._lower()
        q2 = np.asarray(q, dtype=np.float64).reshape(-1, low.dim)
        n = len(q2)
        r = low.model.hmc_run(q2, low.tau(n), np.broadcast_to(np.asarray(timestep, dtype=np.float64), (n,)),
                              nsteps, beta=low.beta(n), p0=np.asarray(p).reshape(n, low.dim),
                              u=np.full(n, 0.5), want_end=True)
        qe, pe = r["q_end"].astype(np.float64), r["p_end"].astype(np.float64)
        return (qe[0], pe[0]) if np.ndim(q) == 1 else (qe, pe)

    def _copy_state(self, state):
        return state.clone() if _is_tensor(state) else np.array(state, copy=True)

    def sample(self, p0=None, u=None, n_traj=1):
        """One HMC transition per chain (hmc.py:136-164).  `p0` / `u` inject the momenta and the
        Metropolis uniforms (parity tests); `n_traj` > 1 fuses several transitions into one launch."""
        low = self._lower()
        n_adapt = int(np.clip(self.timestep_adaption_limit - 1 - self.counter, 0, n_traj))  # hmc.py:153-157
        if _is_tensor(self._state):
            new_state = self._sample_device(low, n_traj, n_adapt, p0, u)
        else:
            new_state = self._sample_host(low, n_traj, n_adapt, p0, u)
        self.counter += n_traj
        self._draw += n_traj
        return new_state

    def _sample_host(self, low, n_traj, n_adapt, p0, u):
        single = np.ndim(self._state) == 1
        q = np.asarray(self._state, dtype=np.float64).reshape(-1, low.dim)
        n = len(q)
        r = low.model.hmc_run(q, low.tau(n), self._eps, self.nsteps, n_traj=n_traj, beta=low.beta(n),
                              p0=p0, u=u, n_adapt=n_adapt, adapt_up=self.adaption_uprate,
                              adapt_down=self.adaption_downrate, seed=self.seed, draw=self._draw,
                              chain_base=self.chain_base)
        self._eps = r["eps"].astype(np.float64)
        acc = r["accepted"]
        self.last_energies = (r["e_before"], r["e_after"])
        self.n_accepted = self.n_accepted + (int(r["n_accepted"][0]) if single else r["n_accepted"].astype(np.int64))
        self._last_move_accepted = bool(acc[0]) if single else acc
        new = r["q"].astype(np.float64)
        self._state = new[0] if single else new
        return self._copy_state(self._state)

    def _sample_device(self, low, n_traj, n_adapt, p0, u):
        import torch
        q = self._state
        if q.dim() != 2:
            raise ValueError("HMCSampler: a device-resident state must have shape [n_chains, dim], got %s"
                             % (tuple(q.shape),))
        n = q.shape[0]
        dev = q.device
        if p0 is not None and not _is_tensor(p0):
            p0 = torch.as_tensor(np.ascontiguousarray(p0, dtype=np.float32).reshape(n, low.dim), device=dev)
        if u is not None and not _is_tensor(u):
            u = torch.as_tensor(np.ascontiguousarray(u, dtype=np.float32).reshape(n), device=dev)
        if self._eps_dev is None:
            self._eps_dev = torch.as_tensor(self._eps, dtype=torch.float32, device=dev)
            self._acc_dev = torch.zeros(n, dtype=torch.uint8, device=dev)
            self._nacc_dev = torch.zeros(n, dtype=torch.int32, device=dev)
            self._e0_dev = torch.zeros(n, dtype=torch.float64, device=dev)
            self._e1_dev = torch.zeros(n, dtype=torch.float64, device=dev)
        tau = low.tau(n)
        tau_dev = tau if _is_tensor(tau) else torch.as_tensor(tau, dtype=torch.float32, device=dev)
        beta = low.beta(n)
        beta_dev = None if beta is None else torch.as_tensor(beta, dtype=torch.float32, device=dev)
        opts = _cabi.HmcOpts(self.nsteps, n_traj, n_adapt, _cabi.GIBBS_NONE, self.adaption_uprate,
                             self.adaption_downrate, self.seed, self._draw, self.chain_base)
        low.model.hmc_run_device(q, tau_dev, self._eps_dev, opts, beta=beta_dev, p0=p0, u=u,
                                 accepted=self._acc_dev, e_before=self._e0_dev, e_after=self._e1_dev,
                                 n_accepted=self._nacc_dev, stream=torch.cuda.current_stream().cuda_stream)
        self.n_accepted = self.n_accepted + self._nacc_dev.to(torch.int64)
        self._last_move_accepted = self._acc_dev.bool()
        self.last_energies = (self._e0_dev, self._e1_dev)
        return self._copy_state(q)

    def _adapt_timestep(self):
        """Reference hook (hmc.py:183-191).  The adaption itself runs inside the kernel; this
        method applies the same rule on the host for callers that invoke it directly."""
        acc = np.asarray(self._last_move_accepted)
        self._eps = np.where(acc, self._eps * self.adaption_uprate, self._eps * self.adaption_downrate)
        self._eps_dev = None
