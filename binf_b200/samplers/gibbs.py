"""Gibbs sampler (reference: binf/samplers/gibbs.py:11-191).

Same contract as the reference: one conditional pdf per state variable, sub-samplers visited in
sorted-name order, the conditional pdfs' bound parameters refreshed from the state before every
sub-sample.  When the sub-samplers are this package's HMCSampler + GammaSampler over a lowered
posterior, `sample()` runs the whole sweep -- conjugate precision update and HMC trajectory -- as
ONE launch of the fused kernel (BINFB_GIBBS_TAU_FIRST / _LAST according to the sorted order)."""
from collections import OrderedDict

import numpy as np

from binf_b200.samplers import State, AbstractMC


class GibbsSampler(object):
    def __init__(self, pdf, state, subsamplers, fuse="auto"):
        self._state = state
        self._pdf = pdf
        self._subsamplers = subsamplers
        self._conditional_pdfs = {}
        self._fuse = fuse
        self._setup_conditional_pdfs()
        self._update_subsampler_states()

    def _setup_conditional_pdfs(self):
        """condition the joint pdf on everything but `var`, for every state variable
        (gibbs.py:40-52)"""
        variables = self.state.variables
        for var in variables:
            fixed = {x: variables[x] for x in variables if x != var}
            fixed.update({x: self._pdf[x].value for x in self._pdf.parameters
                          if x in self._pdf._original_variables})
            cond = self._pdf.conditional_factory(**fixed)
            self._conditional_pdfs[var] = cond
            self.subsamplers[var].pdf = cond

    def _update_conditional_pdf_params(self):
        """push the current state into every conditional pdf (gibbs.py:54-62)"""
        variables = self._state.variables
        for pdf in self._conditional_pdfs.values():
            for param in pdf.parameters:
                if param in variables:
                    pdf[param].set(variables[param])

    def _checkstate(self, state):
        if type(state) is not dict:
            raise TypeError(state)

    @property
    def pdf(self):
        return self._pdf

    @pdf.setter
    def pdf(self, value):
        self._pdf = value
        self._setup_conditional_pdfs()

    @property
    def state(self):
        return self._state

    @property
    def subsamplers(self):
        return self._subsamplers

    def update_samplers(self, **samplers):
        self._subsamplers.update(**samplers)

    def _update_subsampler_states(self):
        variables = self.state.variables
        for var in variables:
            sub = self.subsamplers[var]
            sub.state = State(variables[var]) if isinstance(sub, AbstractMC) else variables[var]

    def _update_state(self, **variables):
        for var, value in variables.items():
            if type(value) is State:
                value = value.position
            self._state.update_variables(**{var: value})

    # -- sweep --------------------------------------------------------------------------------
    def _fused_plan(self):
        """(hmc sampler, gamma sampler, gibbs_mode) when the sweep can run as one fused launch."""
        if self._fuse is False:
            return None
        from binf_b200 import _cabi
        from binf_b200.samplers.hmc import HMCSampler
        from binf_b200.example.samplers import GammaSampler
        names = sorted(self._pdf.variables)
        if len(names) != 2 or "precision" not in names:
            return None
        other = [v for v in names if v != "precision"][0]
        hmc, gam = self.subsamplers.get(other), self.subsamplers.get("precision")
        if type(hmc) is not HMCSampler or type(gam) is not GammaSampler:
            return None
        if hmc._variable_name not in (None, other):
            return None
        mode = _cabi.GIBBS_TAU_FIRST if names[0] == "precision" else _cabi.GIBBS_TAU_LAST
        return hmc, gam, other, mode

    def sample(self):
        """One sweep over the variables in sorted-name order (gibbs.py:136-151)."""
        self._update_subsampler_states()
        plan = self._fused_plan()
        if plan is not None:
            return self._sample_fused(*plan)
        for var in sorted(self._pdf.variables):
            self._update_conditional_pdf_params()
            new = self.subsamplers[var].sample()
            self._update_state(**{var: new})
        return self._state

    def _sample_fused(self, hmc, gam, other, mode):
        from binf_b200.lowering import _is_tensor
        self._update_conditional_pdf_params()
        low = hmc._lower()
        # the Gamma prior the precision sub-sampler sees (its own conditional pdf: quirk Q2 applies)
        prior = gam._get_prior()
        low.model.set_gamma_prior(prior.shape, prior.rate)
        q = self._state.variables[other]
        n_adapt = int(np.clip(hmc.timestep_adaption_limit - 1 - hmc.counter, 0, 1))
        if _is_tensor(q):
            return self._sample_fused_device(hmc, gam, other, mode, low, q, n_adapt)
        single = np.ndim(q) == 1
        q2 = np.asarray(q, dtype=np.float64).reshape(-1, low.dim)
        n = len(q2)
        tau = np.broadcast_to(np.asarray(self._state.variables["precision"], dtype=np.float64), (n,))
        r = low.model.hmc_run(q2, tau, hmc._eps, hmc.nsteps, beta=low.beta(n), n_adapt=n_adapt,
                              adapt_up=hmc.adaption_uprate, adapt_down=hmc.adaption_downrate,
                              gibbs_mode=mode, seed=hmc.seed, draw=hmc._draw, chain_base=hmc.chain_base)
        hmc._eps = r["eps"].astype(np.float64)
        hmc.counter += 1
        hmc._draw += 1
        hmc.n_accepted = hmc.n_accepted + (int(r["n_accepted"][0]) if single else r["n_accepted"].astype(np.int64))
        hmc._last_move_accepted = bool(r["accepted"][0]) if single else r["accepted"]
        new_q = r["q"].astype(np.float64)
        new_tau = r["tau"].astype(np.float64)
        hmc._state = new_q[0] if single else new_q
        gam.state = float(new_tau[0]) if single else new_tau
        self._update_state(**{other: hmc._copy_state(hmc._state), "precision": gam.state})
        low.refresh()
        return self._state

    def _sample_fused_device(self, hmc, gam, other, mode, low, q, n_adapt):
        """chains resident in HBM: q [C, D] and precision [C] are CUDA tensors updated in place"""
        import torch
        from binf_b200 import _cabi
        n, dev = q.shape[0], q.device
        tau = self._state.variables["precision"]
        if not hasattr(tau, "data_ptr"):
            tau = torch.as_tensor(np.broadcast_to(np.asarray(tau, dtype=np.float32), (n,)).copy(), device=dev)
        tau = tau.to(torch.float32).contiguous()
        hmc.state = q
        if hmc._eps_dev is None:
            hmc._eps_dev = torch.as_tensor(hmc._eps, dtype=torch.float32, device=dev)
            hmc._acc_dev = torch.zeros(n, dtype=torch.uint8, device=dev)
            hmc._nacc_dev = torch.zeros(n, dtype=torch.int32, device=dev)
            hmc._e0_dev = torch.zeros(n, dtype=torch.float64, device=dev)
            hmc._e1_dev = torch.zeros(n, dtype=torch.float64, device=dev)
        beta = low.beta(n)
        beta_dev = None if beta is None else torch.as_tensor(beta, dtype=torch.float32, device=dev)
        opts = _cabi.HmcOpts(hmc.nsteps, 1, n_adapt, mode, hmc.adaption_uprate, hmc.adaption_downrate,
                             hmc.seed, hmc._draw, hmc.chain_base)
        low.model.hmc_run_device(q, tau, hmc._eps_dev, opts, beta=beta_dev, accepted=hmc._acc_dev,
                                 e_before=hmc._e0_dev, e_after=hmc._e1_dev, n_accepted=hmc._nacc_dev,
                                 stream=torch.cuda.current_stream().cuda_stream)
        hmc.counter += 1
        hmc._draw += 1
        hmc.n_accepted = hmc.n_accepted + hmc._nacc_dev.to(torch.int64)
        hmc._last_move_accepted = hmc._acc_dev.bool()
        gam.state = tau
        self._update_state(**{other: q, "precision": tau})
        low.refresh()
        return self._state

    @property
    def last_draw_stats(self):
        return {k: v.last_draw_stats[k] for k, v in self.subsamplers.items()
                if getattr(v, "last_draw_stats", None) is not None}

    @property
    def sampling_stats(self):
        out = OrderedDict()
        for s in self.subsamplers.values():
            if "sampling_stats" in dir(s):
                out.update(s.sampling_stats)
        return out
