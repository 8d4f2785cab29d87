"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden.py

The reference has no golden vectors for the HMC path, so the vectors are
outputs of the reference itself, imported in memory from /root/reference by
oracle/ref_import.py.  Randomness is injected: `numpy.random.normal/uniform/gamma` are
replaced while the reference's samplers run so that momenta, uniforms and Gamma variates
are known inputs (the reference draws them at samplers/hmc.py:146,151 and
example/samplers.py:47).  Each case is also evaluated with the numpy port
(oracle/binf_port.py, oracle/chromatin_port.py) and the script aborts if they disagree,
so a committed fixture certifies port == reference on that case.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402
import binf_port as port  # noqa: E402
import chromatin_port as chrom  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
polyval = np.polynomial.polynomial.polyval


class InjectedRandom(object):
    """Context manager feeding queued values to numpy.random.normal/uniform/gamma."""

    def __init__(self, normals=(), uniforms=(), gammas=()):
        self.normals, self.uniforms, self.gammas = list(normals), list(uniforms), list(gammas)

    def __enter__(self):
        self._saved = (np.random.normal, np.random.uniform, np.random.gamma)

        def normal(loc=0.0, scale=1.0, size=None):
            v = self.normals.pop(0)
            assert np.shape(v) == (tuple(size) if size is not None else ()), (np.shape(v), size)
            return np.array(v, dtype=np.float64)

        def uniform(low=0.0, high=1.0, size=None):
            return float(self.uniforms.pop(0))

        def gamma(shape, scale=1.0, size=None):
            self.gamma_shapes.append(float(shape))
            return float(self.gammas.pop(0))

        self.gamma_shapes = []
        np.random.normal, np.random.uniform, np.random.gamma = normal, uniform, gamma
        return self

    def __exit__(self, *exc):
        np.random.normal, np.random.uniform, np.random.gamma = self._saved
        return False


def close(a, b, tol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.max(np.abs(a))) if a.size else 1.0)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol * scale, (a, b)


def polynomial_case(binf, name, n_data, n_chains, nsteps, timestep, seed, start="ones"):
    from binf.example.misc import make_posterior
    from binf.samplers.hmc import HMCSampler

    rng = np.random.RandomState(seed)
    xs = np.linspace(-2, 2, n_data)
    true_c = np.array([2.0, -4.0, 1.0, 1.5])
    ys = rng.normal(polyval(xs, true_c), 1.0 / np.sqrt(2.5))
    tau = 2.5
    post = make_posterior(xs, ys, polyval)
    cond = post.conditional_factory(precision=tau)
    gp = [p for p in cond.priors.values() if "precision" in p._original_variables][0]
    cprior = [p for p in cond.priors.values() if "coefficients" in p.variables][0]

    if start == "ones":
        q0 = np.ones((n_chains, 4)) + 0.1 * rng.normal(size=(n_chains, 4))
        q0[0] = 1.0
    else:  # equilibrium start: exact Gaussian conditional
        V = np.vstack([xs ** i for i in range(4)])
        A = tau * V.dot(V.T) + np.diag(np.ones(4) / 5.0)
        q0 = rng.multivariate_normal(np.linalg.solve(A, tau * V.dot(ys)), np.linalg.inv(A),
                                     size=n_chains)
    p0 = rng.normal(size=(n_chains, 4))
    u = rng.uniform(size=n_chains)

    out = dict(xs=xs, ys=ys, tau=tau, q0=q0, p0=p0, u=u, nsteps=nsteps, timestep=timestep,
               prior_means=cprior["means"].value, prior_variances=cprior["variances"].value,
               gamma_shape=gp.shape, gamma_rate=gp.rate,
               full_gamma_rate=[p for p in post.priors.values()
                                if "precision" in p.variables][0].rate)
    logp, grad, full_logp = [], [], []
    qe, pe, eb, ea, acc, qn = [], [], [], [], [], []
    for c in range(n_chains):
        logp.append(cond.log_prob(coefficients=q0[c].copy()))
        grad.append(cond.gradient(coefficients=q0[c].copy()))
        full_logp.append(post.log_prob(coefficients=q0[c].copy(), precision=tau))
        # the reference sampler, with its own leapfrog (hmc.py:92-125)
        s = HMCSampler(cond, q0[c].copy(), timestep, nsteps, variable_name="coefficients")
        ql, pl = s._leapfrog(q0[c].copy(), p0[c].copy(), timestep, nsteps)
        qe.append(ql), pe.append(pl)
        eb.append(-cond.log_prob(coefficients=q0[c].copy()) + 0.5 * np.sum(p0[c] ** 2))
        ea.append(-cond.log_prob(coefficients=ql.copy()) + 0.5 * np.sum(pl ** 2))
        with InjectedRandom(normals=[p0[c]], uniforms=[u[c]]):
            new = s.sample()
        acc.append(bool(s.last_move_accepted)), qn.append(new)
    out.update(log_prob=logp, gradient=grad, full_log_prob=full_logp, q_end=qe, p_end=pe,
               e_before=eb, e_after=ea, accepted=acc, q_new=qn)

    # ---- the port must agree with the reference on every number ----
    pp = port.PolynomialPosterior(xs, ys, out["prior_means"], out["prior_variances"],
                                  gp.shape, gp.rate)
    close(pp.log_prob(q0, tau), logp), close(pp.gradient(q0, tau), grad)
    r = port.hmc_sample(lambda q: pp.log_prob(q, tau), lambda q: pp.gradient(q, tau),
                        q0, timestep, nsteps, p0, u)
    close(r["q_end"], qe), close(r["p_end"], pe), close(r["e_before"], eb)
    close(r["e_after"], ea), close(r["q"], qn)
    assert list(r["accepted"]) == acc
    np.savez(os.path.join(GOLDEN, name + ".npz"), **{k: np.asarray(v) for k, v in out.items()})
    print(name, "acceptance", np.mean(acc), "max |dH|", np.max(np.abs(np.array(ea) - eb)))


def polynomial_gibbs_case(binf, name, n_sweeps, seed):
    """Reference GibbsSampler(HMCSampler + GammaSampler) with injected randomness, incl.
    step-size adaption (hmc.py:153-157) and the sweep order (gibbs.py:146-149)."""
    from binf.example.misc import make_posterior
    from binf.example.samplers import GammaSampler
    from binf.samplers import BinfState
    from binf.samplers.gibbs import GibbsSampler
    from binf.samplers.hmc import HMCSampler

    rng = np.random.RandomState(seed)
    n_data = 20
    xs = np.linspace(-2, 2, n_data)
    ys = rng.normal(polyval(xs, [2.0, -4.0, 1.0, 1.5]), 1.0 / np.sqrt(2.5))
    post = make_posterior(xs, ys, polyval)
    c0, tau0 = np.ones(4), 1.0
    timestep, nsteps, limit = 0.02, 20, 6
    hmc = HMCSampler(post.conditional_factory(precision=tau0), c0.copy(), timestep, nsteps,
                     timestep_adaption_limit=limit, variable_name="coefficients")
    gam = GammaSampler(post.conditional_factory(coefficients=c0), tau0)
    gibbs = GibbsSampler(post, BinfState(dict(coefficients=c0.copy(), precision=tau0)),
                         {"coefficients": hmc, "precision": gam})
    p0 = rng.normal(size=(n_sweeps, 4))
    u = rng.uniform(size=n_sweeps)
    gdraw = rng.gamma(0.5 * n_data, size=n_sweeps)
    cs, taus, accs, steps = [], [], [], []
    with InjectedRandom(normals=list(p0), uniforms=list(u), gammas=list(gdraw)) as inj:
        for _ in range(n_sweeps):
            st = gibbs.sample()
            cs.append(np.array(st.variables["coefficients"]))
            taus.append(float(st.variables["precision"]))
            accs.append(bool(hmc.last_move_accepted)), steps.append(hmc.timestep)
        shapes = inj.gamma_shapes
    gp = [p for p in gam.pdf.priors.values() if "precision" in p.variables][0]
    cp = [p for p in hmc.pdf.priors.values() if "coefficients" in p.variables][0]
    hgp = [p for p in hmc.pdf.priors.values() if "precision" in p._original_variables][0]

    # ---- port replay ----
    pp = port.PolynomialPosterior(xs, ys, cp["means"].value, cp["variances"].value,
                                  hgp.shape, hgp.rate)
    c, tau, dt = c0.copy(), tau0, timestep
    for k in range(n_sweeps):
        r = port.hmc_sample(lambda q: pp.log_prob(q, tau), lambda q: pp.gradient(q, tau),
                            c, dt, nsteps, p0[k], u[k])
        dt = float(port.adapt_timestep(dt, r["accepted"], k + 1, limit))
        c = r["q"]
        shape, rate = port.gamma_precision_params(pp.chi2(c), n_data, gp.shape, gp.rate)
        assert abs(shape - shapes[k]) < 1e-12
        tau = gdraw[k] / rate
        close(c, cs[k]), close(tau, taus[k]), close(dt, steps[k])
        assert bool(r["accepted"]) == accs[k]
    np.savez(os.path.join(GOLDEN, name + ".npz"), xs=xs, ys=ys, c0=c0, tau0=tau0,
             timestep=timestep, nsteps=nsteps, limit=limit, p0=p0, u=u, gamma_draws=gdraw,
             coefficients=np.array(cs), precision=np.array(taus), accepted=np.array(accs),
             timesteps=np.array(steps), gamma_shapes=np.array(shapes),
             prior_means=cp["means"].value, prior_variances=cp["variances"].value,
             hmc_gamma_shape=hgp.shape, hmc_gamma_rate=hgp.rate,
             gam_gamma_shape=gp.shape, gam_gamma_rate=gp.rate)
    print(name, "acceptance", np.mean(accs), "tau", taus[-1], "dt", steps[-1])


def chromatin_case(binf, name, n_beads, n_chains, nsteps, timestep, seed, tau=50.0, ev_k=0.0, ev_d=0.0,
                   contact="logistic"):
    """The reference's Posterior / Likelihood (dense J.dot(g)) / HMCSampler driving the
    build-defined chromatin model at small n."""
    from binf.samplers.hmc import HMCSampler

    alpha, d_c, k_bb, l0 = 2.0, 2.5, 4.0, 1.0
    X, y = chrom.synthetic_chromatin(n_beads, alpha, d_c, l0, 0.05, seed, contact=contact)
    model = chrom.ChromatinModel(n_beads, y, alpha, d_c, k_bb, l0, conf_s=0.0,
                                 gamma_shape=1.0, gamma_rate=1.0, ev_k=ev_k, ev_d=ev_d, contact=contact)
    post = chrom.reference_posterior(binf, model)
    cond = post.conditional_factory(precision=tau)
    rng = np.random.RandomState(seed + 1)
    q0 = X.reshape(-1)[None, :] + 0.1 * rng.normal(size=(n_chains, 3 * n_beads))
    p0 = rng.normal(size=q0.shape)
    u = rng.uniform(size=n_chains)
    logp, grad, qe, pe, eb, ea, acc, qn = [], [], [], [], [], [], [], []
    for c in range(n_chains):
        logp.append(cond.log_prob(structure=q0[c].copy()))
        grad.append(cond.gradient(structure=q0[c].copy()))
        s = HMCSampler(cond, q0[c].copy(), timestep, nsteps, variable_name="structure")
        ql, pl = s._leapfrog(q0[c].copy(), p0[c].copy(), timestep, nsteps)
        qe.append(ql), pe.append(pl)
        eb.append(-cond.log_prob(structure=q0[c].copy()) + 0.5 * np.sum(p0[c] ** 2))
        ea.append(-cond.log_prob(structure=ql.copy()) + 0.5 * np.sum(pl ** 2))
        with InjectedRandom(normals=[p0[c]], uniforms=[u[c]]):
            new = s.sample()
        acc.append(bool(s.last_move_accepted)), qn.append(new)
    gp = [p for p in cond.priors.values() if "precision" in p._original_variables][0]

    # ---- matrix-free port vs the reference's dense-Jacobian path ----
    m2 = chrom.ChromatinModel(n_beads, y, alpha, d_c, k_bb, l0, 0.0, gp.shape, gp.rate, ev_k, ev_d, contact)
    for c in range(n_chains):
        close(m2.log_prob(q0[c], tau), logp[c]), close(m2.gradient(q0[c], tau), grad[c], 1e-9)
        r = port.hmc_sample(lambda q: m2.log_prob(q, tau), lambda q: m2.gradient(q, tau),
                            q0[c], timestep, nsteps, p0[c], u[c])
        close(r["q_end"], qe[c], 1e-9), close(r["p_end"], pe[c], 1e-9)
        close(r["e_before"], eb[c]), close(r["e_after"], ea[c], 1e-9)
        assert bool(r["accepted"]) == acc[c]
    # central finite differences of -log_prob
    h, g_fd = 1e-6, np.zeros(3 * n_beads)
    for k in range(3 * n_beads):
        e = np.zeros(3 * n_beads)
        e[k] = h
        g_fd[k] = -(m2.log_prob(q0[0] + e, tau) - m2.log_prob(q0[0] - e, tau)) / (2 * h)
    close(g_fd, grad[0], 1e-6)
    np.savez(os.path.join(GOLDEN, name + ".npz"), n_beads=n_beads, y=y, alpha=alpha, d_c=d_c,
             k_bb=k_bb, l0=l0, tau=tau, gamma_shape=gp.shape, gamma_rate=gp.rate, ev_k=ev_k, ev_d=ev_d,
             q0=q0, p0=p0, u=u, nsteps=nsteps, timestep=timestep, log_prob=np.array(logp),
             gradient=np.array(grad), q_end=np.array(qe), p_end=np.array(pe),
             e_before=np.array(eb), e_after=np.array(ea), accepted=np.array(acc),
             q_new=np.array(qn), **({} if contact == "logistic" else {"contact": contact}))
    print(name, "acceptance", np.mean(acc), "max |dH|", np.max(np.abs(np.array(ea) - eb)))


def _acceptance_worker(args):
    n_beads, n_chains, seed, tau, timestep, nsteps, lo, hi = args
    (alpha, d_c, k_bb, l0), y, q0, p0, u = chrom.acceptance_inputs(n_beads, n_chains, seed)
    model = chrom.ChromatinModel(n_beads, y, alpha, d_c, k_bb, l0)
    out = []
    for c in range(lo, hi):
        r = port.hmc_sample(lambda q: model.log_prob(q, tau), lambda q: model.gradient(q, tau), q0[c], timestep,
                            nsteps, p0[c], u[c])
        out.append((r["e_after"] - r["e_before"], bool(r["accepted"])))
    return out


def chromatin_acceptance_case(binf, name, n_beads, n_chains, nsteps, timestep, seed, tau=60.0):
    """L = 20 acceptance at a WORKING step size over >= 10^4 chain-trajectories.  The
    decision being matched is binf/samplers/hmc.py:151.  The first chains go through the reference's own
    HMCSampler (dense J.dot(g)); all of them through the port (asserted equal on the shared chains), whose
    energy differences and decisions are stored.  Inputs are seeded and regenerated by the test."""
    import multiprocessing as mp
    from binf.samplers.hmc import HMCSampler
    (alpha, d_c, k_bb, l0), y, q0, p0, u = chrom.acceptance_inputs(n_beads, n_chains, seed)
    model = chrom.ChromatinModel(n_beads, y, alpha, d_c, k_bb, l0)
    cond = chrom.reference_posterior(binf, model).conditional_factory(precision=tau)
    gp = [p for p in cond.priors.values() if "precision" in p._original_variables][0]
    assert (gp.shape, gp.rate) == (1.0, 1.0)
    n_ref = 6
    ref_acc, ref_dh = [], []
    for c in range(n_ref):
        s = HMCSampler(cond, q0[c].copy(), timestep, nsteps, variable_name="structure")
        ql, pl = s._leapfrog(q0[c].copy(), p0[c].copy(), timestep, nsteps)
        ref_dh.append(-cond.log_prob(structure=ql.copy()) + 0.5 * np.sum(pl ** 2)
                      + cond.log_prob(structure=q0[c].copy()) - 0.5 * np.sum(p0[c] ** 2))
        with InjectedRandom(normals=[p0[c]], uniforms=[u[c]]):
            s.sample()
        ref_acc.append(bool(s.last_move_accepted))
    workers = max(1, len(os.sched_getaffinity(0)))
    bounds = np.linspace(0, n_chains, workers * 4 + 1).astype(int)
    jobs = [(n_beads, n_chains, seed, tau, timestep, nsteps, int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:])]
    with mp.get_context("fork").Pool(workers) as pool:
        res = [x for part in pool.map(_acceptance_worker, jobs) for x in part]
    dh = np.array([r[0] for r in res])
    acc = np.array([r[1] for r in res])
    close(dh[:n_ref], ref_dh, 1e-7)
    assert list(acc[:n_ref]) == ref_acc
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), n_beads=n_beads, n_chains=n_chains, seed=seed, tau=tau,
                        nsteps=nsteps, timestep=timestep, y_checksum=float(np.sum(y.astype(np.float64))),
                        dh=dh, accepted=acc, u=u)
    print(name, "acceptance", acc.mean(), "mean min(1, exp(-dH))", np.mean(np.exp(np.minimum(0.0, -dh))))


def user_model_case(binf, name, n_data, n_chains, nsteps, timestep, seed):
    """A USER-DEFINED forward model written against the reference's own extension point
    (AbstractForwardModel._evaluate / _evaluate_jacobi_matrix, binf/model/forwardmodels.py:30-38):
    mock_n = a exp(-k x_n) + c, variables (a, k, c) registered as `coefficients` so that the
    reference's GaussianPrior / GammaPrior / GaussianErrorModel / Likelihood / Posterior / HMCSampler
    run it unchanged.  The fixture pins what the NVRTC-lowered device model must reproduce."""
    from binf import ArrayParameter
    from binf.model.forwardmodels import AbstractForwardModel
    from binf.example.likelihood import GaussianErrorModel
    from binf.example.priors import GammaPrior, GaussianPrior
    from binf.pdf.likelihoods import Likelihood
    from binf.pdf.posteriors import Posterior
    from binf.samplers.hmc import HMCSampler

    class Decay(AbstractForwardModel):
        def __init__(self, xses):
            super(Decay, self).__init__("decay")
            self.xses = xses
            self._register_variable("coefficients", differentiable=True)
            self.update_var_param_types(coefficients=ArrayParameter)
            self._set_original_variables()

        def _evaluate(self, coefficients):
            a, k, c = coefficients
            return a * np.exp(-k * self.xses) + c

        def _evaluate_jacobi_matrix(self, coefficients):
            a, k, c = coefficients
            e = np.exp(-k * self.xses)
            return np.vstack([e, -a * self.xses * e, np.ones_like(self.xses)])

        def clone(self):
            copy = self.__class__(self.xses)
            self._set_parameters(copy)
            return copy

    rng = np.random.RandomState(seed)
    xs = np.linspace(0.0, 4.0, n_data)
    true = np.array([3.0, 1.2, 0.5])
    tau = 25.0
    ys = rng.normal(true[0] * np.exp(-true[1] * xs) + true[2], 1.0 / np.sqrt(tau))
    lik = Likelihood("points", Decay(xs), GaussianErrorModel(ys))
    means, variances = np.array([2.0, 1.0, 0.0]), np.array([4.0, 1.0, 2.0])
    priors = (GammaPrior(1.0, 0.2), GaussianPrior(means=means, variances=variances))
    post = Posterior({lik.name: lik}, {p.name: p for p in priors})
    cond = post.conditional_factory(precision=tau)
    gp = [p for p in cond.priors.values() if "precision" in p._original_variables][0]
    q0 = true[None] + 0.05 * rng.normal(size=(n_chains, 3))
    p0 = rng.normal(size=q0.shape)
    u = rng.uniform(size=n_chains)
    logp, grad, qe, pe, eb, ea, acc, qn, mock = [], [], [], [], [], [], [], [], []
    for c in range(n_chains):
        logp.append(cond.log_prob(coefficients=q0[c].copy()))
        grad.append(cond.gradient(coefficients=q0[c].copy()))
        mock.append(lik.forward_model(coefficients=q0[c].copy()))
        s = HMCSampler(cond, q0[c].copy(), timestep, nsteps, variable_name="coefficients")
        ql, pl = s._leapfrog(q0[c].copy(), p0[c].copy(), timestep, nsteps)
        qe.append(ql), pe.append(pl)
        eb.append(-cond.log_prob(coefficients=q0[c].copy()) + 0.5 * np.sum(p0[c] ** 2))
        ea.append(-cond.log_prob(coefficients=ql.copy()) + 0.5 * np.sum(pl ** 2))
        with InjectedRandom(normals=[p0[c]], uniforms=[u[c]]):
            new = s.sample()
        acc.append(bool(s.last_move_accepted)), qn.append(new)
    # the port with the same functor must agree
    um = port.UserModelPosterior(xs, ys, lambda th, x: th[..., 0:1] * np.exp(-th[..., 1:2] * x) + th[..., 2:3],
                                 lambda th, x: np.stack([np.exp(-th[..., 1:2] * x),
                                                         -th[..., 0:1] * x * np.exp(-th[..., 1:2] * x),
                                                         np.ones_like(th[..., 0:1] * x)], axis=-2),
                                 means, variances, gp.shape, gp.rate)
    close(um.log_prob(q0, tau), logp), close(um.gradient(q0, tau), grad)
    r = port.hmc_sample(lambda q: um.log_prob(q, tau), lambda q: um.gradient(q, tau), q0, timestep, nsteps, p0, u)
    close(r["q_end"], qe), close(r["p_end"], pe), close(r["e_before"], eb), close(r["e_after"], ea)
    assert list(r["accepted"]) == acc
    np.savez(os.path.join(GOLDEN, name + ".npz"), xs=xs, ys=ys, tau=tau, q0=q0, p0=p0, u=u, nsteps=nsteps,
             timestep=timestep, prior_means=means, prior_variances=variances, gamma_shape=gp.shape,
             gamma_rate=gp.rate, log_prob=np.array(logp), gradient=np.array(grad), mock=np.array(mock),
             q_end=np.array(qe), p_end=np.array(pe), e_before=np.array(eb), e_after=np.array(ea),
             accepted=np.array(acc), q_new=np.array(qn))
    print(name, "acceptance", np.mean(acc), "max |dH|", np.max(np.abs(np.array(ea) - eb)))


def rwmc_predict_case(binf, name, n_data, n_chains, n_moves, stepsize, seed):
    """RWMCSampler.sample (binf/example/samplers.py:78-92) with the proposal displacement and the
    uniform injected, chain by chain and move by move; and predict (binf/example/misc.py:3-16) over
    the visited states."""
    from binf.example.misc import make_posterior, predict
    from binf.example.samplers import RWMCSampler
    from binf.samplers import BinfState

    rng = np.random.RandomState(seed)
    xs = np.linspace(-2, 2, n_data)
    ys = rng.normal(polyval(xs, np.array([2.0, -4.0, 1.0, 1.5])), 1.0 / np.sqrt(2.5))
    tau = 2.5
    post = make_posterior(xs, ys, polyval)
    cond = post.conditional_factory(precision=tau)
    gp = [p for p in cond.priors.values() if "precision" in p._original_variables][0]
    cprior = [p for p in cond.priors.values() if "coefficients" in p.variables][0]
    pp = port.PolynomialPosterior(xs, ys, cprior["means"].value, cprior["variances"].value, gp.shape, gp.rate)
    q0 = np.array([2.0, -4.0, 1.0, 1.5]) + 0.3 * rng.normal(size=(n_chains, 4))
    change = rng.uniform(-stepsize, stepsize, size=(n_moves, n_chains, 4))
    u = rng.uniform(size=(n_moves, n_chains))
    states = np.empty((n_moves + 1, n_chains, 4))
    states[0] = q0
    accepted = np.zeros((n_moves, n_chains), dtype=bool)
    saved = (np.random.uniform, np.random.random)
    try:
        for c in range(n_chains):
            smp = RWMCSampler(cond, q0[c].copy(), stepsize)
            pstate = q0[c].copy()
            for k in range(n_moves):
                np.random.uniform = lambda low=0.0, high=1.0, size=None, v=change[k, c]: np.array(v)
                np.random.random = lambda v=u[k, c]: float(v)
                before = smp._n_accepted_moves
                states[k + 1, c] = smp.sample()
                accepted[k, c] = smp._n_accepted_moves > before
                r = port.rwmc_sample(lambda x: pp.log_prob(x, tau), pstate, change[k, c], u[k, c])
                assert r["accepted"] == accepted[k, c]
                close(r["state"], states[k + 1, c])
                pstate = r["state"]
    finally:
        np.random.uniform, np.random.random = saved
    # predictive density over the visited states (each with its own precision draw)
    taus = rng.gamma(6.0, 0.5, size=(n_moves + 1) * n_chains)
    flat = states.reshape(-1, 4)
    samples = [BinfState(dict(coefficients=flat[i], precision=float(taus[i]))) for i in range(len(flat))]
    gx, gy = np.meshgrid(np.linspace(-2.5, 2.5, 7), np.linspace(-12.0, 12.0, 9))
    pred = np.array([predict(float(a), float(b), samples, polyval) for a, b in zip(gx.ravel(), gy.ravel())])
    close(pred, [port.predict(a, b, flat, taus) for a, b in zip(gx.ravel(), gy.ravel())], tol=1e-12)
    np.savez(os.path.join(GOLDEN, name + ".npz"), xs=xs, ys=ys, tau=tau, stepsize=stepsize, q0=q0,
             change=change, u=u, states=states, accepted=accepted, gamma_shape=gp.shape,
             gamma_rate=gp.rate, prior_means=cprior["means"].value, prior_variances=cprior["variances"].value,
             pred_x=gx.ravel(), pred_y=gy.ravel(), pred_tau=taus, pred=pred)
    print("wrote", name, "acceptance", accepted.mean())


def algebraic_cases(binf):
    """the algebraic contact function 1/2 (1 + z / sqrt(1 + z^2)), alone and with excluded volume"""
    chromatin_case(binf, "chromatin_alg_n26", n_beads=26, n_chains=6, nsteps=8, timestep=0.005, seed=13,
                   contact="algebraic")
    chromatin_case(binf, "chromatin_alg_ev_n22", n_beads=22, n_chains=5, nsteps=6, timestep=0.004, seed=14,
                   ev_k=5.0, ev_d=1.6, contact="algebraic")


def main():
    binf = ref_import.install()
    os.makedirs(GOLDEN, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "rwmc":     # only the fixtures added with SURVEY 8f rank 4
        rwmc_predict_case(binf, "poly_rwmc_n20", n_data=20, n_chains=24, n_moves=6, stepsize=0.1, seed=9)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "ev":       # the fixture added with the excluded-volume prior
        chromatin_case(binf, "chromatin_ev_n28", n_beads=28, n_chains=8, nsteps=6, timestep=0.004, seed=11,
                       ev_k=5.0, ev_d=1.6)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "alg":      # the fixtures added with the algebraic contact function
        algebraic_cases(binf)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "accept":   # acceptance at a working step size
        chromatin_acceptance_case(binf, "chromatin_accept_n64", n_beads=64, n_chains=10240, nsteps=20,
                                  timestep=0.032, seed=12)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "user":     # the fixture added with SURVEY 8f rank 2
        user_model_case(binf, "user_decay_n200", n_data=200, n_chains=24, nsteps=10, timestep=0.012, seed=10)
        return
    # config 1 shape (example_script.py:17-26), seeded
    polynomial_case(binf, "poly_n20", n_data=20, n_chains=16, nsteps=20, timestep=0.02, seed=0)
    # config 2 shape: N = 1000 data points, L = 20
    polynomial_case(binf, "poly_n1000", n_data=1000, n_chains=16, nsteps=20, timestep=0.004,
                    seed=1)
    polynomial_case(binf, "poly_n1000_L5", n_data=1000, n_chains=8, nsteps=5, timestep=0.004,
                    seed=2)
    polynomial_case(binf, "poly_n1000_mode", n_data=1000, n_chains=48, nsteps=20,
                    timestep=0.011, seed=6, start="mode")
    polynomial_case(binf, "poly_n77_mode", n_data=77, n_chains=32, nsteps=7,
                    timestep=0.04, seed=7, start="mode")
    polynomial_gibbs_case(binf, "poly_gibbs_n20", n_sweeps=12, seed=3)
    chromatin_case(binf, "chromatin_n24", n_beads=24, n_chains=6, nsteps=5, timestep=0.005,
                   seed=4)
    chromatin_case(binf, "chromatin_n37_L20", n_beads=37, n_chains=4, nsteps=20,
                   timestep=0.004, seed=5)
    chromatin_case(binf, "chromatin_n30_big_step", n_beads=30, n_chains=12, nsteps=10,
                   timestep=0.05, seed=8)
    chromatin_case(binf, "chromatin_ev_n28", n_beads=28, n_chains=8, nsteps=6, timestep=0.004, seed=11,
                   ev_k=5.0, ev_d=1.6)
    chromatin_acceptance_case(binf, "chromatin_accept_n64", n_beads=64, n_chains=10240, nsteps=20,
                              timestep=0.032, seed=12)
    algebraic_cases(binf)
    rwmc_predict_case(binf, "poly_rwmc_n20", n_data=20, n_chains=24, n_moves=6, stepsize=0.1, seed=9)
    user_model_case(binf, "user_decay_n200", n_data=200, n_chains=24, nsteps=10, timestep=0.012, seed=10)


if __name__ == "__main__":
    main()
