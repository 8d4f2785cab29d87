"""GPU parity: chromatin pair kernel vs the reference-generated golden vectors (small n, the
reference's dense J.dot(g) path) and vs the matrix-free oracle at n = 1000, plus
size-independent properties (Newton's third law, translation invariance, reversibility)."""
import numpy as np
import pytest

from conftest import load_golden
import binf_port as port
import chromatin_port as chrom

pytestmark = pytest.mark.gpu
CASES = ["chromatin_n24", "chromatin_n37_L20", "chromatin_n30_big_step", "chromatin_ev_n28",
         "chromatin_alg_n26", "chromatin_alg_ev_n22"]


def make_model(g):
    from binf_b200 import _cabi
    return _cabi.Model.chromatin(int(g["n_beads"]), g["y"], float(g["alpha"]), float(g["d_c"]),
                                 float(g["k_bb"]), float(g["l0"]), 0.0, float(g["gamma_shape"]),
                                 float(g["gamma_rate"]), ev_k=float(g.get("ev_k", 0.0)),
                                 ev_d=float(g.get("ev_d", 0.0)), contact=str(g.get("contact", "logistic")))


def inf_norm(a):
    return np.max(np.abs(a), axis=-1, keepdims=True)


@pytest.mark.parametrize("name", CASES)
def test_logprob_and_gradient_vs_reference(gpu, name):
    g = load_golden(name)
    m = make_model(g)
    logp, grad, chi2 = m.logprob_grad(g["q0"], float(g["tau"]))
    np.testing.assert_allclose(logp, g["log_prob"], rtol=1e-5)
    assert np.all(np.abs(grad - g["gradient"]) <= 1e-4 * inf_norm(g["gradient"]))


@pytest.mark.parametrize("name", CASES)
def test_trajectory_vs_reference(gpu, name):
    g = load_golden(name)
    m = make_model(g)
    L = int(g["nsteps"])
    r = m.hmc_run(g["q0"], float(g["tau"]), float(g["timestep"]), L, p0=g["p0"], u=g["u"], want_end=True)
    tol = 1e-4 if L <= 5 else 1e-3
    assert np.all(np.abs(r["q_end"] - g["q_end"]) <= tol * inf_norm(g["q_end"]))
    assert np.all(np.abs(r["p_end"] - g["p_end"]) <= tol * np.maximum(inf_norm(g["p_end"]), 1.0))
    np.testing.assert_allclose(r["e_before"], g["e_before"], rtol=1e-5)
    dh_ref = g["e_after"] - g["e_before"]
    assert np.all(np.abs((r["e_after"] - r["e_before"]) - dh_ref) <= 5e-3 + 1e-3 * np.abs(dh_ref))
    decided = np.abs(np.log(g["u"]) + dh_ref) > 0.02
    assert np.array_equal(r["accepted"][decided], g["accepted"][decided])


@pytest.mark.parametrize("contact", ["logistic", "algebraic"])
@pytest.mark.parametrize("n", [2, 5, 8, 64, 130, 257])
def test_ragged_sizes_vs_oracle(gpu, n, contact):
    """quad padding, odd/even quad counts, partial row blocks; both contact functions"""
    from binf_b200 import _cabi
    X, y = chrom.synthetic_chromatin(n, seed=n, contact=contact)
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 3.0, 1.0, conf_s=20.0, gamma_shape=2.0, gamma_rate=0.5, contact=contact)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 3.0, 1.0, conf_s=20.0, gamma_shape=2.0, gamma_rate=0.5, contact=contact)
    rng = np.random.RandomState(n)
    C = 11
    q = X.reshape(-1)[None] + 0.2 * rng.normal(size=(C, 3 * n))
    tau = rng.uniform(20, 80, size=C)
    beta = rng.uniform(0.2, 1.0, size=C)
    logp, grad, chi2 = m.logprob_grad(q, tau, beta)
    for c in range(C):
        assert logp[c] == pytest.approx(o.log_prob(q[c], tau[c], beta[c]), rel=1e-5)
        ref = o.gradient(q[c], tau[c], beta[c])
        assert np.all(np.abs(grad[c] - ref) <= 1e-4 * np.max(np.abs(ref)))
        assert chi2[c] == pytest.approx(o.chi2(q[c]), rel=1e-5)


@pytest.mark.parametrize("n,roles", [(2000, 0), (2000, 2), (700, 2), (1000, 4), (500, 2), (1000, 0)])
def test_multi_warp_chains_vs_oracle(gpu, n, roles):
    """2 and 4 warps per chain (partner-step ranges split between the warps of a chain); (1000, 4) and
    (500, 2) need the LOCKSTEP kernels (role ranges of exactly 32 slots), and (1000, 0) with 9 chains
    takes the small-batch route (the alternative 4-role lockstep plan and its own contact stream)"""
    from binf_b200 import _cabi
    X, y = chrom.synthetic_chromatin(n, seed=n)
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0, roles=roles)
    rng = np.random.RandomState(n)
    C = 9
    q = X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))
    logp, grad, chi2 = m.logprob_grad(q, 150.0)
    for c in (0, 8):
        assert logp[c] == pytest.approx(o.log_prob(q[c], 150.0), rel=1e-5)
        ref = o.gradient(q[c], 150.0)
        assert np.all(np.abs(grad[c] - ref) <= 1e-4 * np.max(np.abs(ref)))
    f = grad.reshape(C, n, 3).sum(axis=1)
    assert np.all(np.abs(f) <= 1e-3 * np.max(np.abs(grad)))
    p0, u = rng.normal(size=(C, 3 * n)), np.full(C, 0.5)
    r = m.hmc_run(q, 150.0, 0.002, 2, p0=p0, u=u, want_end=True)
    ref = port.hmc_sample(lambda x: o.log_prob(x, 150.0), lambda x: o.gradient(x, 150.0), q[3], 0.002,
                          2, p0[3], 0.5)
    assert np.max(np.abs(r["q_end"][3] - ref["q_end"])) <= 1e-4 * np.max(np.abs(ref["q_end"]))
    assert (r["e_after"][3] - r["e_before"][3]) == pytest.approx(ref["e_after"] - ref["e_before"], abs=3e-2)


@pytest.mark.parametrize("roles", [0, 1])
def test_n1000_vs_matrix_free_oracle(gpu, roles):
    """Config-3 size: 1000 beads, 499,500 pairs per force evaluation."""
    from binf_b200 import _cabi
    n = 1000
    X, y = chrom.synthetic_chromatin(n, seed=0)
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0, roles=roles)
    rng = np.random.RandomState(1)
    C = 20  # not a multiple of the chains-per-CTA
    q = X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))
    tau = np.full(C, 400.0)
    logp, grad, chi2 = m.logprob_grad(q, tau)
    for c in (0, 7, 19):
        assert logp[c] == pytest.approx(o.log_prob(q[c], 400.0), rel=1e-5)
        ref = o.gradient(q[c], 400.0)
        assert np.all(np.abs(grad[c] - ref) <= 1e-4 * np.max(np.abs(ref)))
    # Newton's third law: likelihood + backbone forces are pairwise => they sum to zero
    f = grad.reshape(C, n, 3).sum(axis=1)
    assert np.all(np.abs(f) <= 1e-3 * np.max(np.abs(grad)))
    # translation invariance of log_prob
    logp2, _, _ = m.logprob_grad(q + np.tile([3.0, -2.0, 1.0], n)[None], tau, want_grad=False)
    np.testing.assert_allclose(logp2, logp, rtol=2e-6)
    # short trajectory against the oracle integrator
    p0, u = rng.normal(size=(2, 3 * n)), np.array([0.5, 0.5])
    r = m.hmc_run(q[:2], 400.0, 0.002, 3, p0=p0, u=u, want_end=True)
    for c in range(2):
        ref = port.hmc_sample(lambda x: o.log_prob(x, 400.0), lambda x: o.gradient(x, 400.0), q[c],
                              0.002, 3, p0[c], u[c])
        assert np.max(np.abs(r["q_end"][c] - ref["q_end"])) <= 1e-4 * np.max(np.abs(ref["q_end"]))
        assert np.max(np.abs(r["p_end"][c] - ref["p_end"])) <= 1e-4 * np.max(np.abs(ref["p_end"]))
        assert r["e_before"][c] == pytest.approx(ref["e_before"], rel=1e-6)
        assert (r["e_after"][c] - r["e_before"][c]) == pytest.approx(ref["e_after"] - ref["e_before"], abs=2e-2)


@pytest.mark.parametrize("roles,contact", [(0, "logistic"), (8, "logistic"), (0, "algebraic"), (8, "algebraic")])
def test_n5000_many_warps_per_chain(gpu, roles, contact):
    """Config-4 size: 5000 beads (12,497,500 pairs per force evaluation), one chain per CTA, 16 warps
    per chain (the plan's choice: three 32 KiB ring stages) or 8.  Oracle: float64 forces on a subset of beads (chunked), chi^2 over all pairs."""
    from binf_b200 import _cabi
    n, alpha, d_c, k_bb, l0, tau = 5000, 2.0, 2.5, 4.0, 1.0, 120.0
    X, y = chrom.synthetic_chromatin(n, alpha, d_c, l0, 0.05, seed=7, contact=contact)
    m = _cabi.Model.chromatin(n, y, alpha, d_c, k_bb, l0, roles=roles, contact=contact)
    rng = np.random.RandomState(3)
    C = 3
    q = X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))
    logp, grad, chi2 = m.logprob_grad(q, tau)
    # ---- float64 oracle, chunked over rows --------------------------------------------------
    iu = np.triu_indices(n, 1)
    Y = np.zeros((n, n), dtype=np.float32)
    Y[iu] = y
    Y = Y + Y.T
    Xc = q[1].reshape(n, 3)
    chi2_ref, rows = 0.0, np.r_[0:40, 2480:2520, 4960:5000]
    f_ref = np.zeros((len(rows), 3))
    for lo in range(0, n, 500):
        blk = np.arange(lo, min(n, lo + 500))
        diff = Xc[blk][:, None, :] - Xc[None, :, :]
        d = np.sqrt(np.sum(diff * diff, axis=-1) + 1e-12)
        mm, dmdd = chrom.contact_function(d, alpha, d_c, contact)
        res = mm - Y[blk]
        res[np.arange(len(blk)), blk] = 0.0
        chi2_ref += 0.5 * np.sum(res * res)
        sel = np.isin(blk, rows)
        if sel.any():
            w = res[sel] * dmdd[sel] / d[sel]
            f_ref[np.isin(rows, blk)] = tau * np.einsum("ij,ija->ia", w, diff[sel])
    assert chi2[1] == pytest.approx(chi2_ref, rel=1e-5)
    b = Xc[1:] - Xc[:-1]
    db = np.sqrt(np.sum(b * b, axis=-1) + 1e-12)
    cb = (k_bb * (db - l0) / db)[:, None] * b
    gp = np.zeros((n, 3))
    gp[1:] += cb
    gp[:-1] -= cb
    g = grad[1].reshape(n, 3)
    assert np.max(np.abs(g[rows] - (f_ref + gp[rows]))) <= 1e-4 * np.max(np.abs(g))
    e_lik = -0.5 * tau * chi2_ref + 0.5 * len(y) * np.log(tau)
    e_pri = -0.5 * k_bb * np.sum((db - l0) ** 2) + (1.0 - 1.0) * np.log(tau) - tau * 1.0
    assert logp[1] == pytest.approx(e_lik + e_pri, rel=1e-5)
    # pairwise forces sum to zero; log_prob is translation invariant
    assert np.all(np.abs(grad.reshape(C, n, 3).sum(axis=1)) <= 1e-3 * np.max(np.abs(grad)))
    logp2, _, _ = m.logprob_grad(q + np.tile([1.0, 2.0, -3.0], n)[None], tau, want_grad=False)
    np.testing.assert_allclose(logp2, logp, rtol=2e-6)
    # a short reversible trajectory
    p0 = rng.normal(size=q.shape)
    u = np.full(C, 1e-30)
    fwd = m.hmc_run(q, tau, 0.001, 3, p0=p0, u=u, want_end=True)
    back = m.hmc_run(fwd["q_end"], tau, 0.001, 3, p0=-fwd["p_end"], u=u, want_end=True)
    assert np.max(np.abs(back["q_end"] - q)) < 2e-4 * np.max(np.abs(q))
    assert np.max(np.abs(fwd["e_after"] - fwd["e_before"])) < 1.0


def test_reversibility_and_energy_conservation(gpu):
    from binf_b200 import _cabi
    n = 257
    X, y = chrom.synthetic_chromatin(n, seed=3)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0)
    rng = np.random.RandomState(4)
    C = 37
    q0 = X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))
    p0 = rng.normal(size=q0.shape)
    u = np.full(C, 1e-30)
    fwd = m.hmc_run(q0, 100.0, 0.003, 12, p0=p0, u=u, want_end=True)
    assert fwd["accepted"].all()
    assert np.max(np.abs(fwd["e_after"] - fwd["e_before"])) < 0.5
    back = m.hmc_run(fwd["q_end"], 100.0, 0.003, 12, p0=-fwd["p_end"], u=u, want_end=True)
    assert np.max(np.abs(back["q_end"] - q0)) < 2e-4 * np.max(np.abs(q0))
    assert np.max(np.abs(back["p_end"] + p0)) < 5e-3


def test_fused_gibbs_sweeps_and_multi_trajectory(gpu):
    """tau-first Gibbs coupling ('precision' < 'structure' in the reference's sorted sweep) with an
    injected Gamma variate, then statistical sanity of many fused sweeps with Philox draws."""
    from binf_b200 import _cabi
    n = 64
    X, y = chrom.synthetic_chromatin(n, seed=9)
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0, gamma_shape=1.0, gamma_rate=1.0)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0, gamma_shape=1.0, gamma_rate=1.0)
    rng = np.random.RandomState(2)
    C = 24
    q0 = X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))
    gd = rng.gamma(0.5 * o.n_pairs, size=C)
    p0, u = rng.normal(size=q0.shape), rng.uniform(size=C)
    r = m.hmc_run(q0, 1.0, 0.004, 4, p0=p0, u=u, gamma_draws=gd, gibbs_mode=_cabi.GIBBS_TAU_FIRST, want_end=True)
    for c in range(C):
        shape, rate = port.gamma_precision_params(o.chi2(q0[c]), o.n_pairs, 1.0, 1.0)
        tau_c = gd[c] / rate
        assert r["tau"][c] == pytest.approx(tau_c, rel=2e-5)
        ref = port.hmc_sample(lambda x: o.log_prob(x, tau_c), lambda x: o.gradient(x, tau_c), q0[c],
                              0.004, 4, p0[c], u[c])
        assert np.max(np.abs(r["q_end"][c] - ref["q_end"])) <= 1e-4 * np.max(np.abs(ref["q_end"]))
        assert r["e_before"][c] == pytest.approx(ref["e_before"], rel=1e-5)
    # 40 fused sweeps: precision concentrates near 1/noise^2 = 400, acceptance is healthy
    r = m.hmc_run(q0, 100.0, 0.002, 10, n_traj=40, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=5)
    assert 150 < np.median(r["tau"]) < 600
    assert r["n_accepted"].mean() / 40 > 0.5
    assert r["stats"][1] == C * 40 and r["stats"][0] == r["n_accepted"].sum()
    # step-size adaption: accepted => x1.05, rejected => x0.95 (hmc.py:188-191)
    r = m.hmc_run(q0, 100.0, 0.002, 5, n_traj=3, n_adapt=3, seed=6)
    expect = 0.002 * 1.05 ** r["n_accepted"] * 0.95 ** (3 - r["n_accepted"])
    np.testing.assert_allclose(r["eps"], expect, rtol=1e-5)


@pytest.mark.parametrize("n,roles,C", [(61, 0, 5), (300, 0, 3), (1000, 0, 3), (1000, 2, 2), (1000, 0, 1300)])
def test_excluded_volume_prior_vs_oracle(gpu, n, roles, C):
    """the quartic excluded-volume repulsion fused into the pair loop: log_prob, gradient, a short
    trajectory, for every kernel shape (1 / 2 / 4 lockstep roles, small-batch and full-batch plans)"""
    from binf_b200 import _cabi
    X, y = chrom.synthetic_chromatin(n, seed=n + 1)
    ev_k, ev_d, tau = 4.0, 1.8, 60.0
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0, ev_k=ev_k, ev_d=ev_d)
    plain = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0, roles=roles, ev_k=ev_k, ev_d=ev_d)
    rng = np.random.RandomState(n)
    q = (X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))).astype(np.float32).astype(np.float64)
    beta = np.linspace(1.0, 0.5, C)
    logp, grad, chi2 = m.logprob_grad(q, tau, beta=beta)
    for c in sorted({0, C - 1, C // 2}):
        assert o.prior_log_prob(q[c]) < plain.prior_log_prob(q[c]) - 1.0     # the term is active
        assert logp[c] == pytest.approx(o.log_prob(q[c], tau, beta[c]), rel=1e-5)
        ref = o.gradient(q[c], tau, beta[c])
        assert np.all(np.abs(grad[c] - ref) <= 1e-4 * np.max(np.abs(ref)))
    p0, u = rng.normal(size=(C, 3 * n)), np.full(C, 0.5)
    r = m.hmc_run(q, tau, 0.002, 3, beta=beta, p0=p0, u=u, want_end=True)
    c = C // 2
    ref = port.hmc_sample(lambda x: o.log_prob(x, tau, beta[c]), lambda x: o.gradient(x, tau, beta[c]), q[c],
                          0.002, 3, p0[c], 0.5)
    assert np.max(np.abs(r["q_end"][c] - ref["q_end"])) <= 1e-4 * np.max(np.abs(ref["q_end"]))
    assert r["e_before"][c] == pytest.approx(ref["e_before"], rel=1e-5)
    assert (r["e_after"][c] - r["e_before"][c]) == pytest.approx(ref["e_after"] - ref["e_before"], abs=3e-2)


def test_excluded_volume_with_fused_gibbs_sweep(gpu):
    """precision-first fused sweep with the excluded-volume term == precision update, then trajectory
    (the kernel draws tau before the first sweep from chi^2 of the incoming state)"""
    from binf_b200 import _cabi
    n, C = 120, 6
    X, y = chrom.synthetic_chromatin(n, seed=4)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0, ev_k=3.0, ev_d=1.7)
    rng = np.random.RandomState(2)
    q = X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))
    p0, u, gd = rng.normal(size=q.shape), rng.uniform(size=C), rng.gamma(3000.0, size=C)
    fused = m.hmc_run(q, 55.0, 0.002, 4, p0=p0, u=u, gamma_draws=gd, gibbs_mode=_cabi.GIBBS_TAU_FIRST, want_end=True)
    tau1, _ = m.gibbs_precision(q, 55.0, gamma_draws=gd)
    split = m.hmc_run(q, tau1, 0.002, 4, p0=p0, u=u, want_end=True)
    np.testing.assert_allclose(fused["tau"], tau1, rtol=1e-6)
    np.testing.assert_allclose(fused["q_end"], split["q_end"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(fused["e_after"], split["e_after"], rtol=1e-9)
    # several fused sweeps on Philox streams: tau stays where the data put it, chains keep moving
    r = m.hmc_run(q, 55.0, 0.002, 4, n_traj=5, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=3)
    assert np.all(r["tau"] > 5.0) and np.all(r["tau"] < 5000.0) and r["n_accepted"].min() >= 3


def test_excluded_volume_through_the_python_api(gpu):
    from binf_b200.chromatin import make_chromatin_posterior
    n = 48
    X, y = chrom.synthetic_chromatin(n, seed=9)
    post = make_chromatin_posterior(n, y, ev_k=2.0, ev_d=1.5)
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0, ev_k=2.0, ev_d=1.5)
    q = X.reshape(-1) + 0.05 * np.random.RandomState(1).normal(size=3 * n)
    cond = post.conditional_factory(precision=40.0)
    assert cond.log_prob(structure=q) == pytest.approx(o.log_prob(q, 40.0), rel=1e-5)
    ref = o.gradient(q, 40.0)
    assert np.max(np.abs(cond.gradient(structure=q) - ref)) <= 1e-4 * np.max(np.abs(ref))


@pytest.mark.parametrize("alpha,d_c", [(0.35, 6.0), (1.0, 2.0), (6.5, 1.2), (12.0, 3.0)])
def test_other_exponent_slopes_vs_oracle(gpu, alpha, d_c):
    """the pair loop works on positions scaled by alpha*log2(e) (pair_block.cuh, SCALED): every constant
    that carries a length (softening, backbone, conformational prior, excluded volume, drift) has to follow"""
    from binf_b200 import _cabi
    n, C, L = 70, 7, 4
    rng = np.random.RandomState(3)
    X = np.cumsum(rng.normal(size=(n, 3)), axis=0)
    iu = np.triu_indices(n, 1)
    d = np.sqrt(((X[iu[0]] - X[iu[1]]) ** 2).sum(-1))
    y = (1.0 / (1.0 + np.exp(np.minimum(alpha * (d - d_c), 60.0))) + 0.05 * rng.normal(size=d.shape)).astype(np.float32)
    kw = dict(conf_s=15.0, gamma_shape=2.0, gamma_rate=0.5, ev_k=1.5, ev_d=1.4)
    o = chrom.ChromatinModel(n, y, alpha, d_c, 3.0, 1.1, **kw)
    m = _cabi.Model.chromatin(n, y, alpha, d_c, 3.0, 1.1, **kw)
    q = X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))
    tau, beta = rng.uniform(20, 60, size=C), rng.uniform(0.3, 1.0, size=C)
    logp, grad, chi2 = m.logprob_grad(q, tau, beta)
    for c in range(C):
        assert logp[c] == pytest.approx(o.log_prob(q[c], tau[c], beta[c]), rel=1e-5)
        ref = o.gradient(q[c], tau[c], beta[c])
        assert np.all(np.abs(grad[c] - ref) <= 1e-4 * np.max(np.abs(ref)))
        assert chi2[c] == pytest.approx(o.chi2(q[c]), rel=1e-5)
    p0, u = rng.normal(size=q.shape), rng.uniform(size=C)
    eps = 2e-3
    r = m.hmc_run(q, tau, eps, L, beta=beta, p0=p0, u=u, want_end=True)
    for c in range(C):
        ref = port.hmc_sample(lambda x: o.log_prob(x, tau[c], beta[c]), lambda x: o.gradient(x, tau[c], beta[c]),
                              q[c], eps, L, p0[c], u[c])
        assert np.max(np.abs(r["q_end"][c] - ref["q_end"])) <= 1e-4 * np.max(np.abs(ref["q_end"]))
        assert np.max(np.abs(r["p_end"][c] - ref["p_end"])) <= 1e-4 * max(1.0, np.max(np.abs(ref["p_end"])))
        assert r["e_before"][c] == pytest.approx(ref["e_before"], rel=1e-5)
        assert abs((r["e_after"][c] - r["e_before"][c]) - (ref["e_after"] - ref["e_before"])) <= 5e-3


def test_exponent_slope_limits_are_checked(gpu):
    from binf_b200 import _cabi
    y = np.zeros(10 * 9 // 2, dtype=np.float32)
    for alpha, d_c in [(0.0, 2.5), (-2.0, 2.5), (50.0, 2.5), (float("nan"), 2.5)]:
        with pytest.raises((_cabi.BinfB200Error, ValueError)):
            _cabi.Model.chromatin(10, y, alpha, d_c, 3.0, 1.0)


def test_n6000_two_stage_ring_fallback(gpu):
    """beyond ~5576 beads a chain leaves no room for the third stage of the 16-role contact ring: the plan
    falls back to two stages.  Size-independent properties (the full oracle is checked at n = 5000): chi^2
    against float64 on a row subset is replaced by Newton's third law, translation invariance, agreement of
    energy-only and energy+gradient passes, and a reversible trajectory"""
    from binf_b200 import _cabi
    n, alpha, d_c, k_bb, l0, tau = 6000, 2.0, 2.5, 4.0, 1.0, 120.0
    X, y = chrom.synthetic_chromatin(n, alpha, d_c, l0, 0.05, seed=11)
    m = _cabi.Model.chromatin(n, y, alpha, d_c, k_bb, l0)
    rng = np.random.RandomState(4)
    C = 2
    q = X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))
    logp, grad, chi2 = m.logprob_grad(q, tau)
    # chi^2 of chain 0 in float64, chunked over rows
    Xc = q[0].reshape(n, 3)
    iu = np.triu_indices(n, 1)
    Y = np.zeros((n, n), dtype=np.float32)
    Y[iu] = y
    chi2_ref = 0.0
    for lo in range(0, n, 500):
        blk = np.arange(lo, min(n, lo + 500))
        d = np.sqrt(np.sum((Xc[blk][:, None, :] - Xc[None, :, :]) ** 2, axis=-1) + 1e-12)
        with np.errstate(over="ignore"):
            res = 1.0 / (1.0 + np.exp(alpha * (d - d_c))) - Y[blk]
        res[np.arange(n)[None, :] <= blk[:, None]] = 0.0     # pairs i < j only
        chi2_ref += np.sum(res * res)
    assert chi2[0] == pytest.approx(chi2_ref, rel=1e-5)
    assert np.all(np.abs(grad.reshape(C, n, 3).sum(axis=1)) <= 1e-3 * np.max(np.abs(grad)))
    logp2, _, _ = m.logprob_grad(q + np.tile([1.0, 2.0, -3.0], n)[None], tau, want_grad=False)
    np.testing.assert_allclose(logp2, logp, rtol=2e-6)
    p0 = rng.normal(size=q.shape)
    u = np.full(C, 1e-30)
    fwd = m.hmc_run(q, tau, 0.001, 2, p0=p0, u=u, want_end=True)
    back = m.hmc_run(fwd["q_end"], tau, 0.001, 2, p0=-fwd["p_end"], u=u, want_end=True)
    assert np.max(np.abs(back["q_end"] - q)) < 2e-4 * np.max(np.abs(q))


def test_pipelined_host_call_equals_the_serial_one(gpu):
    """binfb_hmc_run_host overlaps the state copies with the ONE trajectory launch (the copy-in stream opens the
    kernel's gate chunk by chunk, the copy-out stream waits on per-chunk completion counters).  Same launch,
    same arithmetic: the result must be bit-identical to copy-in / run / copy-out, including Gibbs updates,
    rejected chains (their states come back unchanged) and a ragged last chunk."""
    from binf_b200 import _cabi
    n, C = 100, 4099                     # 4.9 MB of state: above the pipelining threshold; C not a multiple of W
    X, y = chrom.synthetic_chromatin(n, seed=7)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0)
    rng = np.random.RandomState(3)
    q0 = (X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))).astype(np.float32)
    eps = np.where(np.arange(C) % 3 == 0, 0.05, 0.004).astype(np.float32)      # a third of the chains rejects
    out = {}
    for mode in (1, 0):
        m.set_option("host.pipeline", mode)
        out[mode] = m.hmc_run(q0, 60.0, eps, 6, n_traj=2, n_adapt=2, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=5, draw=3)
    for key in ("q", "tau", "eps", "accepted", "e_before", "e_after", "n_accepted"):
        np.testing.assert_array_equal(out[1][key], out[0][key], err_msg=key)
    np.testing.assert_allclose(out[1]["stats"], out[0]["stats"], rtol=1e-12)   # (float64 atomics: order varies)
    acc = out[1]["n_accepted"]
    assert 0 < (acc == 0).sum() and (acc > 0).sum() > C // 2
    np.testing.assert_array_equal(out[1]["q"][acc == 0], q0[acc == 0])
    # injected uniforms travel on the kernel's stream in the pipelined path too
    m.set_option("host.pipeline", 1)
    u = rng.uniform(size=C).astype(np.float32)
    r1 = m.hmc_run(q0, 60.0, 0.004, 5, u=u, seed=1)
    m.set_option("host.pipeline", 0)
    r0 = m.hmc_run(q0, 60.0, 0.004, 5, u=u, seed=1)
    np.testing.assert_array_equal(r1["q"], r0["q"])
    assert r1["accepted"].mean() > 0.5
    # the number of pieces the state travels in is a knob ("host.chunks", 2 .. 64): same bits for every setting
    m.set_option("host.pipeline", 1)
    for chunks in (2, 7, 64):
        m.set_option("host.chunks", chunks)
        assert m.get_option("host.chunks") == chunks
        r = m.hmc_run(q0, 60.0, eps, 6, n_traj=2, n_adapt=2, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=5, draw=3)
        for key in ("q", "tau", "eps", "accepted", "n_accepted"):
            np.testing.assert_array_equal(r[key], out[0][key], err_msg="%s with %d chunks" % (key, chunks))
    m.set_option("host.chunks", 1000)
    assert m.get_option("host.chunks") == 64


def test_benchmarked_launch_shape_vs_oracle(gpu):
    """the instantiation the benchmark runs -- chrom_kernel<2,2,0,4,0>, 8 chains per CTA on
    all 148 SMs, no excluded volume, 1000 beads -- checked against the float64 oracle AT THAT LAUNCH SHAPE:
    log_prob / gradient, and a full L = 20 trajectory with a bound on the error of dH, on chains that sit in
    different CTAs and at different positions inside a CTA."""
    from binf_b200 import _cabi
    n, C, tau, eps, L = 1000, 1200, 400.0, 0.0015, 20
    X, y = chrom.synthetic_chromatin(n, seed=0)
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0)
    _, plan = _cabi.chromatin_stream_layout(n, y)
    assert plan["roles"] == 2 and plan["chains_per_cta"] == 8 and plan["stage_steps"] == 4 and plan["ring_depth"] == 4
    assert (C + 7) // 8 >= gpu["sm_count"]            # every SM gets a full group of 8 chains: the full-batch plan
    rng = np.random.RandomState(11)
    q = (X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))).astype(np.float32).astype(np.float64)
    check = (0, 5, 603, C - 1)                        # first / middle / last group, different slots of a group
    logp, grad, chi2 = m.logprob_grad(q, tau)
    for c in check:
        assert logp[c] == pytest.approx(o.log_prob(q[c], tau), rel=1e-5)
        ref = o.gradient(q[c], tau)
        assert np.all(np.abs(grad[c] - ref) <= 1e-4 * np.max(np.abs(ref)))
    p0 = rng.normal(size=q.shape)
    u = rng.uniform(size=C)
    r = m.hmc_run(q, tau, eps, L, p0=p0, u=u, want_end=True)
    for c in check:
        ref = port.hmc_sample(lambda x: o.log_prob(x, tau), lambda x: o.gradient(x, tau), q[c], eps, L, p0[c], u[c])
        assert np.max(np.abs(r["q_end"][c] - ref["q_end"])) <= 1e-4 * np.max(np.abs(ref["q_end"]))
        assert np.max(np.abs(r["p_end"][c] - ref["p_end"])) <= 2e-3 * np.max(np.abs(ref["p_end"]))
        assert r["e_before"][c] == pytest.approx(ref["e_before"], rel=1e-6)
        dh_ref = ref["e_after"] - ref["e_before"]
        assert abs((r["e_after"][c] - r["e_before"][c]) - dh_ref) <= 5e-2, (c, dh_ref)
        if abs(np.log(u[c]) + dh_ref) > 0.1:
            assert bool(r["accepted"][c]) == bool(ref["accepted"])
    # every chain of the batch behaves like the checked ones (no slot of any CTA is special): the starting points
    # are out of equilibrium, so dH is systematically negative, but its spread over the batch is small
    dh = r["e_after"] - r["e_before"]
    assert np.all(np.isfinite(dh)) and np.max(np.abs(dh - np.median(dh))) < 1.5


def test_acceptance_rate_parity_at_a_working_step_size(gpu):
    """SURVEY.md A.4 item 4 / hmc.py:151: 10,240 chains x one L = 20 trajectory of the 64-bead posterior at a
    step size that rejects one proposal in six.  Fixture: energy differences and decisions of the float64 port
    (pinned to the reference's own HMCSampler on its first chains) for seeded inputs the test regenerates.
    (1) same momenta and uniforms: acceptance rates within 1 % absolute -- in fact decisions identical wherever
    the uniform is not within 0.02 of the threshold; (2) the kernel's own Philox momenta and uniforms, 4
    trajectories per chain from the same starting points: rate within 1 % of the port's."""
    from binf_b200 import _cabi
    g = load_golden("chromatin_accept_n64")
    n, C, L = int(g["n_beads"]), int(g["n_chains"]), int(g["nsteps"])
    tau, eps = float(g["tau"]), float(g["timestep"])
    (alpha, d_c, k_bb, l0), y, q0, p0, u = chrom.acceptance_inputs(n, C, int(g["seed"]))
    assert float(np.sum(y.astype(np.float64))) == float(g["y_checksum"])       # same inputs as the fixture
    np.testing.assert_array_equal(u, g["u"])
    m = _cabi.Model.chromatin(n, y, alpha, d_c, k_bb, l0)
    r = m.hmc_run(q0, tau, eps, L, p0=p0, u=u)
    rate_ref, rate = g["accepted"].mean(), r["accepted"].mean()
    assert 0.7 < rate_ref < 0.9
    assert abs(rate - rate_ref) < 0.01
    dh = r["e_after"] - r["e_before"]
    assert np.max(np.abs(dh - g["dh"])) < 2e-2                               # fp32 forces vs float64, 20 steps
    decided = np.abs(np.log(u) + g["dh"]) > 0.02
    assert decided.mean() > 0.9
    np.testing.assert_array_equal(r["accepted"][decided], g["accepted"][decided])
    # independent randomness: 4 x 10,240 chain-trajectories, each from the fixture's starting points
    acc = np.concatenate([m.hmc_run(q0, tau, eps, L, seed=100 + k, draw=k)["accepted"] for k in range(4)])
    se = np.sqrt(rate_ref * (1 - rate_ref) * (1.0 / acc.size + 1.0 / C))
    assert abs(acc.mean() - rate_ref) < 0.01 + 2 * se


def test_full_size_properties_of_the_benchmarked_configuration(gpu):
    """BASELINE.json configs[2] at its FULL size (1000 beads, 4096 chains, L = 20, fused precision update),
    where the float64 oracle would take hours: size-independent properties instead.
    (1) chain sharding: the random streams are keyed by the global chain id, so a contiguous
        slice of the batch run on its own with chain_base = first index -- what another rank of a chain-sharded
        job does -- ends BIT-identically where the slice keeps the launch plan and the groups' row-block rotation
        (same chains per CTA, first index a multiple of chains-per-CTA x row blocks = 64), and within fp32
        summation-order differences where a small slice runs in the small-batch plan (other role split);
    (2) reversibility of the leapfrog map (hmc.py:116-123) on every chain of the batch;
    (3) Newton's third law: the pair forces of the likelihood and the bond forces sum to zero per chain."""
    from binf_b200 import _cabi
    n, C, L, eps, tau = 1000, 4096, 20, 0.009, 100.0
    X, y = chrom.synthetic_chromatin(n, seed=0)
    m = _cabi.Model.chromatin(n, y, 2.0, 2.5, 4.0, 1.0, gamma_shape=1.0, gamma_rate=1.0)
    rng = np.random.RandomState(21)
    q0 = (X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))).astype(np.float32)
    kw = dict(gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=17, draw=4)
    full = m.hmc_run(q0, tau, eps, L, n_traj=2, **kw)
    assert np.all(np.isfinite(full["q"])) and np.all(full["tau"] > 0)
    assert 0.2 < full["accepted"].mean() <= 1.0
    for lo, hi in ((0, 1100), (2048, 4096)):
        part = m.hmc_run(q0[lo:hi], tau, eps, L, n_traj=2, chain_base=lo, **kw)
        for key in ("q", "tau", "accepted", "n_accepted", "e_before", "e_after"):
            np.testing.assert_array_equal(part[key], full[key][lo:hi], err_msg="%s of chains %d..%d" % (key, lo, hi))
    one = m.hmc_run(q0, tau, eps, L, **kw)
    lo, hi = 1000, 1100                      # 100 chains: the small-batch plan (4 warps per chain, lockstep)
    part = m.hmc_run(q0[lo:hi], tau, eps, L, chain_base=lo, **kw)
    np.testing.assert_allclose(part["tau"], one["tau"][lo:hi], rtol=1e-5)
    np.testing.assert_allclose(part["e_before"], one["e_before"][lo:hi], rtol=1e-6)
    assert np.max(np.abs((part["e_after"] - part["e_before"]) - (one["e_after"] - one["e_before"])[lo:hi])) < 5e-2
    assert (part["accepted"] == one["accepted"][lo:hi]).mean() > 0.9
    # (2) forward and back with flipped momenta, every chain
    p0 = rng.normal(size=q0.shape).astype(np.float32)
    u = np.full(C, 1e-30)
    fwd = m.hmc_run(q0, tau, 0.002, L, p0=p0, u=u, want_end=True)
    assert fwd["accepted"].all()
    back = m.hmc_run(fwd["q_end"], tau, 0.002, L, p0=-fwd["p_end"], u=u, want_end=True)
    scale = np.max(np.abs(q0))
    assert np.max(np.abs(back["q_end"] - q0)) < 3e-4 * scale
    assert np.max(np.abs(back["p_end"] + p0)) < 2e-2
    # (3) no net force on any chain (the Gaussian confinement is off in this model)
    logp, grad, _ = m.logprob_grad(q0[:512], tau)
    net = grad.reshape(512, n, 3).sum(axis=1)
    assert np.max(np.abs(net)) < 2e-4 * np.max(np.abs(grad)) * np.sqrt(n)


@pytest.mark.parametrize("n,roles,C,ev_k", [(61, 0, 5, 0.0), (300, 0, 3, 0.0), (500, 2, 4, 3.0), (1000, 0, 3, 0.0),
                                            (1000, 2, 2, 0.0), (1000, 0, 1300, 2.0), (2000, 0, 2, 0.0)])
def test_algebraic_contact_function_vs_oracle(gpu, n, roles, C, ev_k):
    """mock = 1/2 (1 + z / sqrt(1 + z^2)), z = alpha (d_c - d) (BINFB_FLAG_CONTACT_ALGEBRAIC) in every
    kernel shape -- one warp per chain, the small-batch lockstep plan, forced roles, the full-batch plan with 8
    chains per CTA -- with and without excluded volume and tempering: forward model, log_prob, gradient and a
    short trajectory against the float64 oracle (pinned to the reference's Posterior / HMCSampler through the
    fixtures chromatin_alg_n26 / chromatin_alg_ev_n22)."""
    from binf_b200 import _cabi
    alpha, d_c = 1.7, 2.2
    X, y = chrom.synthetic_chromatin(n, alpha, d_c, seed=n + 1, contact="algebraic")
    o = chrom.ChromatinModel(n, y, alpha, d_c, 4.0, 1.0, ev_k=ev_k, ev_d=1.5, contact="algebraic")
    m = _cabi.Model.chromatin(n, y, alpha, d_c, 4.0, 1.0, roles=roles, ev_k=ev_k, ev_d=1.5, contact="algebraic")
    assert m.get_option("chrom.algebraic") == 1
    rng = np.random.RandomState(n)
    q = (X.reshape(-1)[None] + 0.1 * rng.normal(size=(C, 3 * n))).astype(np.float32).astype(np.float64)
    tau, beta = 80.0, np.linspace(0.3, 1.0, C)
    check = sorted(set([0, C // 2, C - 1]))
    if n <= 300:
        mock = m.forward(q[:1])[0]
        np.testing.assert_allclose(mock, o.forward(q[0]), atol=2e-6)
    logp, grad, chi2 = m.logprob_grad(q, tau, beta=beta)
    for c in check:
        assert logp[c] == pytest.approx(o.log_prob(q[c], tau, beta[c]), rel=1e-5)
        assert chi2[c] == pytest.approx(o.chi2(q[c]), rel=1e-5)
        ref = o.gradient(q[c], tau, beta[c])
        assert np.all(np.abs(grad[c] - ref) <= 1e-4 * np.max(np.abs(ref)))
    p0, u = rng.normal(size=q.shape), rng.uniform(size=C)
    r = m.hmc_run(q, tau, 0.003, 4, beta=beta, p0=p0, u=u, want_end=True)
    for c in check:
        ref = port.hmc_sample(lambda x: o.log_prob(x, tau, beta[c]), lambda x: o.gradient(x, tau, beta[c]), q[c],
                              0.003, 4, p0[c], u[c])
        assert np.max(np.abs(r["q_end"][c] - ref["q_end"])) <= 1e-4 * np.max(np.abs(ref["q_end"]))
        assert np.max(np.abs(r["p_end"][c] - ref["p_end"])) <= 1e-3 * max(1.0, np.max(np.abs(ref["p_end"])))
        assert r["e_before"][c] == pytest.approx(ref["e_before"], rel=1e-5)
        assert abs((r["e_after"][c] - r["e_before"][c]) - (ref["e_after"] - ref["e_before"])) <= 2e-2


def test_algebraic_contact_function_through_the_python_api(gpu):
    """make_chromatin_posterior(contact="algebraic") -> Posterior.log_prob / gradient, the forward model, a Gibbs
    sweep: the lowered model is the algebraic kernel (values against the oracle), clones keep the choice"""
    from binf_b200.chromatin import make_chromatin_posterior, ContactForwardModel
    n = 40
    X, y = chrom.synthetic_chromatin(n, seed=5, contact="algebraic")
    o = chrom.ChromatinModel(n, y, 2.0, 2.5, 4.0, 1.0, contact="algebraic")
    post = make_chromatin_posterior(n, y, contact="algebraic")
    q = X.reshape(-1) + 0.05 * np.random.RandomState(1).normal(size=3 * n)
    cond = post.conditional_factory(precision=30.0)
    assert cond.log_prob(structure=q) == pytest.approx(o.log_prob(q, 30.0), rel=1e-5)
    ref = o.gradient(q, 30.0)
    assert np.max(np.abs(cond.gradient(structure=q) - ref)) <= 1e-4 * np.max(np.abs(ref))
    fwm = ContactForwardModel(n, 2.0, 2.5, contact="algebraic")
    assert fwm.clone().contact == "algebraic"
    np.testing.assert_allclose(fwm(structure=q), o.forward(q), atol=2e-6)
    logistic = make_chromatin_posterior(n, y).conditional_factory(precision=30.0)
    assert abs(logistic.log_prob(structure=q) - o.log_prob(q, 30.0)) > 1e-3   # a different model
    with pytest.raises(ValueError):
        ContactForwardModel(n, 2.0, 2.5, contact="tanh")
