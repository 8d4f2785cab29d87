"""GPU parity: fused polynomial HMC kernel vs the golden vectors of the reference and the
oracle port.  Tolerances are the fp32-vs-fp64 bounds of SURVEY.md A.4(4): log_prob rel 1e-5,
gradient 1e-4 of |g|_inf, short trajectories rel 1e-4 (L <= 7) / 2e-3 (L = 20)."""
import numpy as np
import pytest

from conftest import load_golden
import binf_port as port

pytestmark = pytest.mark.gpu
CASES = ["poly_n20", "poly_n1000", "poly_n1000_L5", "poly_n1000_mode", "poly_n77_mode"]


def make_model(g, flags=0):
    from binf_b200 import _cabi
    return _cabi.Model.polynomial(g["xs"], g["ys"], 4, g["prior_means"], g["prior_variances"],
                                  float(g["gamma_shape"]), float(g["gamma_rate"]), flags=flags)


def inf_norm(a):
    return np.max(np.abs(a), axis=-1, keepdims=True)


@pytest.mark.parametrize("name", CASES)
def test_logprob_and_gradient(gpu, name):
    g = load_golden(name)
    m = make_model(g)
    logp, grad, chi2 = m.logprob_grad(g["q0"], float(g["tau"]))
    np.testing.assert_allclose(logp, g["log_prob"], rtol=1e-5)
    assert np.all(np.abs(grad - g["gradient"]) <= 1e-4 * inf_norm(g["gradient"]))
    pp = port.PolynomialPosterior(g["xs"], g["ys"], g["prior_means"], g["prior_variances"], 1.0, 1.0)
    np.testing.assert_allclose(chi2, pp.chi2(g["q0"]), rtol=1e-5)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("group", [None, 1, 4, 32])
def test_trajectory_energies_accept(gpu, name, group):
    g = load_golden(name)
    m = make_model(g)
    if group is not None:
        m.set_option("poly.group", group)
    L = int(g["nsteps"])
    r = m.hmc_run(g["q0"], float(g["tau"]), float(g["timestep"]), L, p0=g["p0"], u=g["u"],
                  want_end=True)
    tol = 1e-4 if L <= 7 else 2e-3
    assert np.all(np.abs(r["q_end"] - g["q_end"]) <= tol * inf_norm(g["q_end"]))
    assert np.all(np.abs(r["p_end"] - g["p_end"]) <= tol * np.maximum(inf_norm(g["p_end"]), 1.0))
    np.testing.assert_allclose(r["e_before"], g["e_before"], rtol=1e-5)
    dh_ref = g["e_after"] - g["e_before"]
    dh = r["e_after"] - r["e_before"]
    assert np.all(np.abs(dh - dh_ref) <= 2e-2 + 1e-3 * np.abs(dh_ref))
    # the accept decision must agree wherever it is not a coin flip at fp32 resolution
    margin = np.abs(np.log(g["u"]) + dh_ref)
    decided = margin > 0.05
    assert decided.sum() >= len(decided) // 2
    assert np.array_equal(r["accepted"][decided], g["accepted"][decided])
    new = np.where(r["accepted"][:, None], r["q_end"], g["q0"])
    np.testing.assert_allclose(r["q"], new, rtol=1e-6)


def test_all_template_shapes_agree(gpu):
    """every (lanes-per-chain, chains-per-thread) instantiation computes the same thing"""
    g = load_golden("poly_n1000_mode")
    base = None
    for grp in (1, 2, 4, 8, 16, 32):
        for j in (1, 2, 4):
            if j == 4 and grp not in (4, 8):
                continue
            m = make_model(g)
            for ur in ((1, 0) if (j == 2 and grp <= 8) else (1,)):   # uniform-row mapping on / off
                m.set_option("poly.group", grp), m.set_option("poly.chains_per_thread", j)
                m.set_option("poly.uniform_rows", ur)
                r = m.hmc_run(g["q0"], 2.5, float(g["timestep"]), 20, p0=g["p0"], u=g["u"], want_end=True)
                if base is None:
                    base = r
                assert np.all(np.abs(r["q_end"] - base["q_end"]) <= 2e-3 * inf_norm(base["q_end"]))
                np.testing.assert_allclose(r["e_before"], base["e_before"], rtol=1e-6)


@pytest.mark.parametrize("name", ["poly_n1000_L5", "poly_n77_mode", "poly_n20"])
@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_uniform_row_mapping(gpu, name, warps):
    """the uniform-row mapping (rows in constant memory, W warps per chain set, cross-warp reduction in
    shared memory) against the golden vectors: log-prob / gradient, trajectories, several trajectories
    per launch with rejections (the state array is the rejection fallback across warps), ragged chain
    counts, row counts that do not divide by W, and two models taking turns on the constant bank"""
    g = load_golden(name)
    m = make_model(g)
    m.set_option("poly.chains_per_thread", 2), m.set_option("poly.group", warps)
    logp, grad, chi2 = m.logprob_grad(g["q0"], float(g["tau"]))
    np.testing.assert_allclose(logp, g["log_prob"], rtol=1e-5)
    assert np.all(np.abs(grad - g["gradient"]) <= 1e-4 * inf_norm(g["gradient"]))
    L = int(g["nsteps"])
    r = m.hmc_run(g["q0"], float(g["tau"]), float(g["timestep"]), L, p0=g["p0"], u=g["u"], want_end=True)
    tol = 1e-4 if L <= 7 else 2e-3
    assert np.all(np.abs(r["q_end"] - g["q_end"]) <= tol * inf_norm(g["q_end"]))
    np.testing.assert_allclose(r["e_before"], g["e_before"], rtol=1e-5)
    margin = np.abs(np.log(g["u"]) + g["e_after"] - g["e_before"])
    assert np.array_equal(r["accepted"][margin > 0.05], g["accepted"][margin > 0.05])
    # a second model with other data in between, then the same runs with the regular mapping: a ragged
    # number of chains (not a multiple of 64), 3 trajectories per launch with a step size that rejects a lot
    other = make_model(dict(g, ys=g["ys"][::-1].copy()))
    other.set_option("poly.chains_per_thread", 2), other.set_option("poly.group", warps)
    rng = np.random.RandomState(5)
    C = 203
    q0 = g["q0"][0] + 0.05 * rng.normal(size=(C, 4))
    ref = make_model(g)
    ref.set_option("poly.chains_per_thread", 2), ref.set_option("poly.group", warps)
    ref.set_option("poly.uniform_rows", 0)
    eps = 4.0 * float(g["timestep"])
    outs = []
    for model in (m, ref):
        other.logprob_grad(q0, 1.0)
        outs.append(model.hmc_run(q0, float(g["tau"]), eps, 6, n_traj=3, seed=11, want_end=True))
    a, b = outs
    same = a["n_accepted"] == b["n_accepted"]
    assert same.mean() > 0.95 and 0.02 < (a["n_accepted"] < 3).mean()
    assert np.all(np.abs(a["q"][same] - b["q"][same]) <= 2e-3 * inf_norm(b["q"][same]))
    np.testing.assert_allclose(a["e_before"][same], b["e_before"][same], rtol=1e-4)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 6, 8])
def test_other_degrees(gpu, k):
    from binf_b200 import _cabi
    rng = np.random.RandomState(k)
    xs = np.linspace(-1.5, 1.5, 333)
    ys = rng.normal(size=333)
    mean, var = np.zeros(k), 5 * np.ones(k)
    pp = port.PolynomialPosterior(xs, ys, mean, var, 1.0, 1.0)
    m = _cabi.Model.polynomial(xs, ys, k, mean, var, 1.0, 1.0)
    q0 = rng.normal(size=(37, k)) * 0.3
    logp, grad, _ = m.logprob_grad(q0, 1.7)
    np.testing.assert_allclose(logp, pp.log_prob(q0, 1.7), rtol=1e-5)
    ref = pp.gradient(q0, 1.7)
    assert np.all(np.abs(grad - ref) <= 1e-4 * inf_norm(ref))
    p0, u = rng.normal(size=q0.shape), rng.uniform(size=37)
    r = m.hmc_run(q0, 1.7, 0.01, 6, p0=p0, u=u, want_end=True)
    o = port.hmc_sample(lambda q: pp.log_prob(q, 1.7), lambda q: pp.gradient(q, 1.7), q0, 0.01, 6, p0, u)
    assert np.all(np.abs(r["q_end"] - o["q_end"]) <= 1e-4 * np.maximum(inf_norm(o["q_end"]), 1.0))
    np.testing.assert_allclose(r["e_after"], o["e_after"], rtol=1e-5, atol=1e-3)


def test_prior_gradient_flag_fixes_quirk_q1(gpu):
    from binf_b200 import _cabi
    g = load_golden("poly_n20")
    m = make_model(g, flags=_cabi.FLAG_PRIOR_GRAD)
    _, grad, _ = m.logprob_grad(g["q0"], float(g["tau"]))
    want = g["gradient"] + (g["q0"] - g["prior_means"]) / g["prior_variances"]
    assert np.all(np.abs(grad - want) <= 1e-4 * inf_norm(want))


def test_gibbs_sweep_matches_reference_sampler(gpu):
    """GibbsSampler(HMCSampler + GammaSampler) of the reference with injected randomness, replayed
    on the device one sweep at a time (sorted order: coefficients, then precision)."""
    from binf_b200 import _cabi
    g = load_golden("poly_gibbs_n20")
    m = _cabi.Model.polynomial(g["xs"], g["ys"], 4, g["prior_means"], g["prior_variances"],
                               float(g["hmc_gamma_shape"]), float(g["hmc_gamma_rate"]))
    c, tau, eps = g["c0"][None, :], float(g["tau0"]), float(g["timestep"])
    limit = int(g["limit"])
    for k in range(len(g["u"])):
        n_adapt = 1 if (k + 1) < limit else 0
        r = m.hmc_run(c, tau, eps, int(g["nsteps"]), p0=g["p0"][k][None], u=g["u"][k:k + 1],
                      gamma_draws=g["gamma_draws"][k:k + 1], n_adapt=n_adapt,
                      gibbs_mode=_cabi.GIBBS_TAU_LAST)
        c, tau, eps = r["q"], float(r["tau"][0]), float(r["eps"][0])
        assert bool(r["accepted"][0]) == bool(g["accepted"][k])
        np.testing.assert_allclose(c[0], g["coefficients"][k], rtol=2e-3, atol=2e-3)
        assert tau == pytest.approx(float(g["precision"][k]), rel=5e-3)
        assert eps == pytest.approx(float(g["timesteps"][k]), rel=1e-5)


def test_separate_precision_update(gpu):
    g = load_golden("poly_n1000_mode")
    m = make_model(g)
    pp = port.PolynomialPosterior(g["xs"], g["ys"], g["prior_means"], g["prior_variances"], 1.0, 1.0)
    gd = np.random.RandomState(0).gamma(500.0, size=len(g["q0"]))
    tau, chi2 = m.gibbs_precision(g["q0"], 2.5, gamma_draws=gd)
    shape, rate = port.gamma_precision_params(pp.chi2(g["q0"]), 1000, 1.0, 1.0)
    np.testing.assert_allclose(tau, gd / rate, rtol=1e-5)
    assert shape == 500.0


def test_posterior_moments_65536_chains(gpu):
    """Config-2 shape at full size: 65,536 chains x 1,000 data x L = 20.  The conditional
    posterior is Gaussian with mean A^-1 tau V y and covariance A^-1; with
    the reference's quirk Q1 the force omits the prior but the Metropolis test restores it."""
    from binf_b200 import _cabi
    g = load_golden("poly_n1000")
    xs, ys, tau = g["xs"], g["ys"], 2.5
    V = np.vstack([xs ** i for i in range(4)])
    A = tau * V.dot(V.T) + np.diag(1.0 / g["prior_variances"])
    mean, cov = np.linalg.solve(A, tau * V.dot(ys)), np.linalg.inv(A)
    C = 65536
    rng = np.random.RandomState(5)
    q = rng.multivariate_normal(mean, cov, size=C)
    m = make_model(g)
    r = m.hmc_run(q, tau, 0.009, 20, n_traj=30, seed=11, draw=0)
    acc = r["n_accepted"].mean() / 30
    assert 0.6 < acc < 0.99
    assert r["stats"][1] == C * 30 and r["stats"][0] == r["n_accepted"].sum()
    qs = r["q"].astype(np.float64)
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(qs.mean(0) - mean) < 6 * sd / np.sqrt(C))
    assert np.all(np.abs(qs.std(0) / sd - 1) < 0.02)
    # different seed => different draws; same seed => identical (Philox streams are keyed)
    r2 = m.hmc_run(q, tau, 0.009, 20, n_traj=2, seed=11, draw=0)
    r3 = m.hmc_run(q, tau, 0.009, 20, n_traj=2, seed=11, draw=0)
    r4 = m.hmc_run(q, tau, 0.009, 20, n_traj=2, seed=12, draw=0)
    assert np.array_equal(r2["q"], r3["q"]) and not np.array_equal(r2["q"], r4["q"])


def test_rng_streams(gpu):
    from binf_b200 import _cabi
    n, u, gm = _cabi.rng_fill(seed=3, draw=7, chain_base=100, n_chains=4096, dim=64, gamma_shape=250.0)
    assert abs(n.mean()) < 4 / np.sqrt(n.size) and abs(n.std() - 1) < 0.01
    assert abs(np.mean(n ** 4) - 3.0) < 0.1
    assert 0 < u.min() and u.max() <= 1 and abs(u.mean() - 0.5) < 0.02
    assert abs(gm.mean() / 250.0 - 1) < 0.005 and abs(gm.var() / 250.0 - 1) < 0.1
    n2, _, _ = _cabi.rng_fill(seed=3, draw=7, chain_base=101, n_chains=4095, dim=64)
    assert np.array_equal(n[1:], n2)   # keyed by global chain id => sharding-invariant


def test_harmonic_oscillator_closed_form(gpu):
    """K = 1: E(c) = 1/2 tau sum (c - y_n)^2 is a 1-D harmonic oscillator with k = tau N around the
    data mean, for which one leapfrog step is a known linear map."""
    from binf_b200 import _cabi
    rng = np.random.RandomState(3)
    N, tau, eps, L = 64, 1.7, 0.013, 9
    xs, ys = np.linspace(-1, 1, N), rng.normal(size=N)
    m = _cabi.Model.polynomial(xs, ys, 1, np.zeros(1), np.full(1, 1e30), 1.0, 1.0)
    C = 257
    q0, p0 = rng.normal(size=(C, 1)), rng.normal(size=(C, 1))
    r = m.hmc_run(q0, tau, eps, L, p0=p0, u=np.full(C, 1e-30), want_end=True)
    k, x0 = tau * N, ys.mean()
    # kick-drift-kick step as a matrix acting on (q - x0, p)
    half = np.array([[1.0, 0.0], [-0.5 * eps * k, 1.0]])
    drift = np.array([[1.0, eps], [0.0, 1.0]])
    step = np.linalg.matrix_power(half.dot(drift).dot(half), L)
    z = step.dot(np.vstack([q0[:, 0] - x0, p0[:, 0]]))
    np.testing.assert_allclose(r["q_end"][:, 0], z[0] + x0, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(r["p_end"][:, 0], z[1], rtol=2e-5, atol=2e-4)
    # energy error O((eps omega)^2) relative, and exact reversibility to fp32 round-off
    assert np.max(np.abs(r["e_after"] - r["e_before"]) / np.abs(r["e_before"])) < (eps ** 2 * k)
    back = m.hmc_run(r["q_end"], tau, eps, L, p0=-r["p_end"], u=np.full(C, 1e-30), want_end=True)
    np.testing.assert_allclose(back["q_end"], q0, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(back["p_end"], -p0, rtol=1e-4, atol=1e-4)


def test_acceptance_rate_parity_with_reference_sampler(gpu):
    """L = 20, N = 1000: acceptance rate of 32,768 device chain-trajectories (Philox momenta) vs
    12,288 chain-trajectories of the CPU port of the reference sampler, same equilibrium ensemble:
    difference < 1 % absolute."""
    g = load_golden("poly_n1000_mode")
    xs, ys, tau = g["xs"], g["ys"], 2.5
    V = np.vstack([xs ** i for i in range(4)])
    A = tau * V.dot(V.T) + np.diag(1.0 / g["prior_variances"])
    mean, cov = np.linalg.solve(A, tau * V.dot(ys)), np.linalg.inv(A)
    rng = np.random.RandomState(9)
    pp = port.PolynomialPosterior(xs, ys, g["prior_means"], g["prior_variances"], 1.0, 1.0)
    eps, L = 0.0105, 20
    m = make_model(g)
    C = 32768
    r = m.hmc_run(rng.multivariate_normal(mean, cov, size=C), tau, eps, L, seed=123)
    acc_gpu = r["accepted"].mean()
    qc = rng.multivariate_normal(mean, cov, size=12288)
    rc = port.hmc_sample(lambda q: pp.log_prob(q, tau), lambda q: pp.gradient(q, tau), qc, eps, L,
                         rng.normal(size=qc.shape), rng.uniform(size=len(qc)))
    acc_cpu = rc["accepted"].mean()
    assert 0.5 < acc_cpu < 0.98
    assert abs(acc_gpu - acc_cpu) < 0.01 + 3 * np.sqrt(0.25 / 12288)
    # mean Metropolis probability from the device diagnostics agrees with the accept fraction
    assert abs(r["stats"][3] / r["stats"][1] - acc_gpu) < 0.01


def test_error_behaviour_and_edge_shapes(gpu):
    from binf_b200 import _cabi
    g = load_golden("poly_n20")
    m = make_model(g)
    with pytest.raises(_cabi.BinfB200Error) as e:      # injected momenta need n_traj == 1
        m.hmc_run(g["q0"], 2.5, 0.02, 5, n_traj=2, p0=g["p0"])
    assert e.value.code == _cabi.EINVAL
    with pytest.raises(_cabi.BinfB200Error):
        m.hmc_run(g["q0"], 2.5, 0.02, 0)               # L >= 1
    # a single chain, a single datum, chain counts that are not multiples of anything
    one = m.hmc_run(g["q0"][:1], 2.5, 0.02, 20, p0=g["p0"][:1], u=g["u"][:1], want_end=True)
    assert np.all(np.abs(one["q_end"][0] - g["q_end"][0]) <= 2e-3 * np.abs(g["q_end"][0]).max())
    tiny = _cabi.Model.polynomial(np.array([0.5]), np.array([1.0]), 4, np.zeros(4), np.ones(4), 1.0, 1.0)
    lp, grad, chi2 = tiny.logprob_grad(np.ones((3, 4)), 2.0)
    mock = 1 + 0.5 + 0.25 + 0.125
    np.testing.assert_allclose(chi2, (mock - 1.0) ** 2, rtol=1e-6)
    np.testing.assert_allclose(grad[0], 2.0 * (mock - 1.0) * 0.5 ** np.arange(4), rtol=1e-5)
    for C in (1, 31, 33, 1023, 4097):
        rng = np.random.RandomState(C)
        q = g["q0"][rng.randint(0, len(g["q0"]), size=C)]
        lp, _, _ = m.logprob_grad(q, 2.5, want_grad=False)
        pp = port.PolynomialPosterior(g["xs"], g["ys"], g["prior_means"], g["prior_variances"], 1.0, 1.0)
        np.testing.assert_allclose(lp, pp.log_prob(q, 2.5), rtol=1e-5)
    # NaN energies reject (hmc.py:151: `u < exp(nan)` is False) and leave the state untouched
    bad = g["q0"][:4].copy()
    bad[1, 2] = np.nan
    r = m.hmc_run(bad, 2.5, 0.02, 5, seed=3)
    assert not r["accepted"][1] and np.isnan(r["q"][1, 2]) and np.isfinite(r["q"][0]).all()
    # data sets larger than shared memory are streamed in chunks (N = 20,000 -> 320 KB of rows)
    rng = np.random.RandomState(0)
    xs = np.linspace(-1.2, 1.2, 20000)
    ys = rng.normal(size=20000)
    big = _cabi.Model.polynomial(xs, ys, 4, np.zeros(4), 5 * np.ones(4), 1.0, 1.0)
    pb = port.PolynomialPosterior(xs, ys, np.zeros(4), 5 * np.ones(4), 1.0, 1.0)
    q = rng.normal(size=(70, 4)) * 0.2
    lp, grad, _ = big.logprob_grad(q, 0.7)
    np.testing.assert_allclose(lp, pb.log_prob(q, 0.7), rtol=1e-5)
    ref = pb.gradient(q, 0.7)
    assert np.all(np.abs(grad - ref) <= 1e-4 * inf_norm(ref))
    p0, u = rng.normal(size=q.shape), rng.uniform(size=70)
    r = big.hmc_run(q, 0.7, 0.002, 4, p0=p0, u=u, want_end=True)
    o = port.hmc_sample(lambda c: pb.log_prob(c, 0.7), lambda c: pb.gradient(c, 0.7), q, 0.002, 4, p0, u)
    assert np.all(np.abs(r["q_end"] - o["q_end"]) <= 1e-4 * np.maximum(inf_norm(o["q_end"]), 1.0))


def test_uniform_row_mapping_many_chains(gpu):
    """more chains than one wave of chain sets (several iterations per thread, one warp per set): the
    uniform-row and the regular mapping agree chain by chain"""
    g = load_golden("poly_n1000_L5")
    C = 600001
    rng = np.random.RandomState(2)
    q0 = g["q0"][0] + 0.05 * rng.normal(size=(C, 4))
    outs = []
    for ur in (1, 0):
        m = make_model(g)
        m.set_option("poly.uniform_rows", ur)
        outs.append(m.hmc_run(q0, float(g["tau"]), float(g["timestep"]), 3, seed=3, want_end=True))
    a, b = outs
    assert np.all(np.abs(a["q_end"] - b["q_end"]) <= 1e-4 * inf_norm(b["q_end"]))
    np.testing.assert_allclose(a["e_before"], b["e_before"], rtol=1e-5)
    assert (a["accepted"] == b["accepted"]).mean() > 0.999
