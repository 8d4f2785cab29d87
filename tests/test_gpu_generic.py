"""User-defined per-datum forward models lowered through NVRTC (binfb_model_create_generic) against
the reference running the same model through its own extension point
(tests/golden/user_decay_n200.npz: AbstractForwardModel._evaluate/_evaluate_jacobi_matrix ->
Likelihood -> Posterior -> HMCSampler), and against the built-in polynomial kernel."""
import numpy as np
import pytest

import binf_port as port
from conftest import load_golden

pytestmark = pytest.mark.gpu

DECAY_CODE = """
__device__ float binfb_mock(const float *theta, const float *x, float *dmock) {
    const float e = expf(-theta[1] * x[0]);
    dmock[0] = e; dmock[1] = -theta[0] * x[0] * e; dmock[2] = 1.0f;
    return theta[0] * e + theta[2];
}
"""
POLY_CODE = """
__device__ float binfb_mock(const float *theta, const float *x, float *dmock) {
    float v = theta[GEN_K - 1], pw = 1.0f;
    for (int k = GEN_K - 2; k >= 0; --k) v = fmaf(v, x[0], theta[k]);
    for (int k = 0; k < GEN_K; ++k) { dmock[k] = pw; pw *= x[0]; }
    return v;
}
"""


def test_decay_model_vs_reference(gpu):
    from binf_b200 import _cabi
    g = load_golden("user_decay_n200")
    m = _cabi.Model.generic(DECAY_CODE, 3, g["xs"], g["ys"], g["prior_means"], g["prior_variances"],
                            float(g["gamma_shape"]), float(g["gamma_rate"]))
    assert m.kind == _cabi.MODEL_GENERIC and m.dim == 3 and m.n_data == 200
    tau = float(g["tau"])
    logp, grad, _ = m.logprob_grad(g["q0"], tau)
    np.testing.assert_allclose(logp, g["log_prob"], rtol=1e-5)
    assert np.max(np.abs(grad - g["gradient"])) <= 1e-4 * np.max(np.abs(g["gradient"]))
    np.testing.assert_allclose(m.forward(g["q0"]), g["mock"], rtol=1e-5, atol=1e-6)
    r = m.hmc_run(g["q0"], tau, float(g["timestep"]), int(g["nsteps"]), p0=g["p0"], u=g["u"], want_end=True)
    assert np.max(np.abs(r["q_end"] - g["q_end"])) <= 1e-4 * np.max(np.abs(g["q_end"]))
    assert np.max(np.abs(r["p_end"] - g["p_end"])) <= 1e-3 * np.max(np.abs(g["p_end"]))
    np.testing.assert_allclose(r["e_before"], g["e_before"], rtol=1e-5)
    dh_ref = g["e_after"] - g["e_before"]
    clear = np.abs(np.log(g["u"]) + dh_ref) > 0.05
    np.testing.assert_array_equal(r["accepted"][clear], g["accepted"][clear])
    assert not r["accepted"].all() or g["accepted"].all()


def test_user_model_through_the_reference_api(gpu):
    """the drop-in claim: Likelihood / Posterior / HMCSampler / GibbsSampler over a DeviceForwardModel"""
    from binf_b200.example.likelihood import GaussianErrorModel
    from binf_b200.example.priors import GammaPrior, GaussianPrior
    from binf_b200.example.samplers import GammaSampler
    from binf_b200.model.forwardmodels import DeviceForwardModel
    from binf_b200.pdf.likelihoods import Likelihood
    from binf_b200.pdf.posteriors import Posterior
    from binf_b200.samplers import BinfState
    from binf_b200.samplers.gibbs import GibbsSampler
    from binf_b200.samplers.hmc import HMCSampler
    g = load_golden("user_decay_n200")
    fwm = DeviceForwardModel("decay", g["xs"], "coefficients", 3, DECAY_CODE)
    lik = Likelihood("points", fwm, GaussianErrorModel(g["ys"]))
    priors = (GammaPrior(1.0, 0.2), GaussianPrior(means=g["prior_means"], variances=g["prior_variances"]))
    post = Posterior({lik.name: lik}, {p.name: p for p in priors})
    tau = float(g["tau"])
    cond = post.conditional_factory(precision=tau)
    assert cond.log_prob(coefficients=g["q0"][0]) == pytest.approx(g["log_prob"][0], rel=1e-5)
    np.testing.assert_allclose(cond.gradient(coefficients=g["q0"][3]), g["gradient"][3], rtol=2e-4, atol=1e-2)
    np.testing.assert_allclose(fwm(coefficients=g["q0"][1]), g["mock"][1], rtol=1e-5, atol=1e-6)
    s = HMCSampler(cond, g["q0"][0].copy(), float(g["timestep"]), int(g["nsteps"]), variable_name="coefficients")
    new = s.sample(p0=g["p0"][0:1], u=g["u"][0:1])
    np.testing.assert_allclose(new, g["q_new"][0], rtol=1e-4, atol=1e-5)
    # Gibbs over (coefficients, precision) with 512 chains: posterior mean of the decay rate near the truth
    C = 512
    rng = np.random.RandomState(0)
    start = BinfState(dict(coefficients=np.array([3.0, 1.2, 0.5]) + 0.02 * rng.normal(size=(C, 3)),
                           precision=np.full(C, 20.0)))
    hmc = HMCSampler(post.conditional_factory(precision=start.variables["precision"]),
                     start.variables["coefficients"], 0.004, 10, variable_name="coefficients", seed=5)
    gam = GammaSampler(post.conditional_factory(coefficients=start.variables["coefficients"]),
                       start.variables["precision"], seed=6)
    gibbs = GibbsSampler(post, start, {"coefficients": hmc, "precision": gam})
    acc = np.zeros(3)
    for i in range(60):
        st = gibbs.sample()
        if i >= 20:
            acc += st.variables["coefficients"].mean(axis=0)
    mean = acc / 40
    assert abs(mean[1] - 1.2) < 0.1 and abs(mean[0] - 3.0) < 0.2
    assert 5.0 < np.mean(st.variables["precision"]) < 80.0 and hmc.acceptance_rate > 0.5


def test_generic_polynomial_matches_builtin_kernel(gpu):
    """the polynomial written as a user functor reproduces the hand-written polynomial kernel"""
    from binf_b200 import _cabi
    g = load_golden("poly_n1000")
    args = (g["prior_means"], g["prior_variances"], float(g["gamma_shape"]), float(g["gamma_rate"]))
    builtin = _cabi.Model.polynomial(g["xs"], g["ys"], 4, *args)
    generic = _cabi.Model.generic(POLY_CODE, 4, g["xs"], g["ys"], *args)
    tau = float(g["tau"])
    lb, gb, cb = builtin.logprob_grad(g["q0"], tau)
    lg, gg, cg = generic.logprob_grad(g["q0"], tau)
    np.testing.assert_allclose(lg, lb, rtol=2e-6)
    np.testing.assert_allclose(lg, g["log_prob"], rtol=1e-5)
    assert np.max(np.abs(gg - gb)) <= 2e-5 * np.max(np.abs(gb))
    rb = builtin.hmc_run(g["q0"], tau, float(g["timestep"]), int(g["nsteps"]), p0=g["p0"], u=g["u"], want_end=True)
    rg = generic.hmc_run(g["q0"], tau, float(g["timestep"]), int(g["nsteps"]), p0=g["p0"], u=g["u"], want_end=True)
    assert np.max(np.abs(rg["q_end"] - rb["q_end"])) <= 1e-4 * np.max(np.abs(rb["q_end"]))
    assert np.max(np.abs(rg["q_end"] - g["q_end"])) <= 2e-3 * np.max(np.abs(g["q_end"]))
    # Philox path: same seeds -> same momenta; fused Gibbs sweeps stay close to the built-in kernel's
    a = builtin.hmc_run(g["q0"], tau, 0.003, 8, n_traj=3, gibbs_mode=_cabi.GIBBS_TAU_LAST, seed=9)
    b = generic.hmc_run(g["q0"], tau, 0.003, 8, n_traj=3, gibbs_mode=_cabi.GIBBS_TAU_LAST, seed=9)
    same = a["n_accepted"] == b["n_accepted"]
    assert same.mean() > 0.9
    np.testing.assert_allclose(b["tau"][same], a["tau"][same], rtol=1e-3)
    # RWMC and the precision update run on the generic model too (they only need log_prob)
    r = generic.rwmc_run(g["q0"], tau, 0.01, n_moves=5, seed=3)
    assert r["logp"].shape == (len(g["q0"]),) and np.all(np.isfinite(r["logp"]))


BRANCHY_CODE = """
__device__ float binfb_mock(const float *theta, const float *x, float *dmock) {
    // a kink: the code branches on a value, which a pair of chains cannot do together
    const float z = theta[0] * x[0] + theta[1];
    if (z > 0.0f) { dmock[0] = x[0]; dmock[1] = 1.0f; return z; }
    dmock[0] = 0.1f * x[0]; dmock[1] = 0.1f; return 0.1f * z;
}
"""


def test_chain_pairs_packed_build_equals_the_scalar_build(gpu):
    """Where the rows fit the constant bank the user's code is compiled over pairs of chains (packed FP32); each
    component performs the scalar build's operations, so the two builds agree to the last bit wherever the
    code has no a*b+c the scalar compiler may contract (the decay model has: tolerance there); odd chain
    counts leave half a pair empty; code that cannot be compiled over pairs falls back to one chain per lane."""
    from binf_b200 import _cabi
    g = load_golden("poly_n1000")
    args = (g["prior_means"], g["prior_variances"], float(g["gamma_shape"]), float(g["gamma_rate"]))
    packed = _cabi.Model.generic(POLY_CODE, 4, g["xs"], g["ys"], *args)
    scalar = _cabi.Model.generic(POLY_CODE, 4, g["xs"], g["ys"], *args, flags=_cabi.FLAG_GENERIC_SCALAR)
    assert packed.get_option("generic.packed") == 1 and scalar.get_option("generic.packed") == 0
    assert packed.get_option("generic.uniform_rows") == 1
    tau = float(g["tau"])
    q0 = np.concatenate([g["q0"], g["q0"][:1] + 0.01])[:len(g["q0"]) | 1]  # odd chain count
    lp, gp, cp = packed.logprob_grad(q0, tau)
    ls, gs, cs = scalar.logprob_grad(q0, tau)
    np.testing.assert_array_equal(lp, ls), np.testing.assert_array_equal(gp, gs), np.testing.assert_array_equal(cp, cs)
    a = packed.hmc_run(q0, tau, 0.012, 9, n_traj=4, n_adapt=2, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=11, want_end=True)
    b = scalar.hmc_run(q0, tau, 0.012, 9, n_traj=4, n_adapt=2, gibbs_mode=_cabi.GIBBS_TAU_FIRST, seed=11, want_end=True)
    for k in ("q", "tau", "eps", "q_end", "p_end", "accepted", "n_accepted", "e_before", "e_after"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert a["n_accepted"].sum() > 0
    # a step size that rejects part of the proposals: the rejected half of a pair keeps its state
    for eps in np.arange(0.010, 0.020, 0.001):
        a = packed.hmc_run(q0, tau, eps, 9, n_traj=3, seed=12, want_end=True)
        if 0 < a["n_accepted"].sum() < 3 * len(q0):
            break
    else:
        raise AssertionError("no step size with partial acceptance")
    b = scalar.hmc_run(q0, tau, eps, 9, n_traj=3, seed=12, want_end=True)
    for k in ("q", "q_end", "p_end", "accepted", "n_accepted", "e_before", "e_after"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    np.testing.assert_array_equal(packed.forward(q0[:3]), scalar.forward(q0[:3]))
    # the decay model: expf and products the scalar compiler may fuse differently
    gd = load_golden("user_decay_n200")
    dargs = (gd["prior_means"], gd["prior_variances"], float(gd["gamma_shape"]), float(gd["gamma_rate"]))
    dp = _cabi.Model.generic(DECAY_CODE, 3, gd["xs"], gd["ys"], *dargs)
    ds = _cabi.Model.generic(DECAY_CODE, 3, gd["xs"], gd["ys"], *dargs, flags=_cabi.FLAG_GENERIC_SCALAR)
    assert dp.get_option("generic.packed") == 1
    l1, g1, _ = dp.logprob_grad(gd["q0"], float(gd["tau"]))
    l2, g2, _ = ds.logprob_grad(gd["q0"], float(gd["tau"]))
    np.testing.assert_allclose(l1, l2, rtol=1e-6)
    assert np.max(np.abs(g1 - g2)) <= 2e-5 * np.max(np.abs(g2))
    # a branch on a value: built one chain per lane, same entry points
    rng = np.random.RandomState(3)
    xs = rng.uniform(-1, 1, size=150)
    ys = np.where(1.5 * xs + 0.2 > 0, 1.5 * xs + 0.2, 0.1 * (1.5 * xs + 0.2)) + 0.05 * rng.normal(size=150)
    m = _cabi.Model.generic(BRANCHY_CODE, 2, xs[:, None].copy(), ys, np.zeros(2), 4.0 * np.ones(2), 1.0, 1.0)
    assert m.get_option("generic.packed") == 0
    th = np.array([1.4, 0.25]) + 0.01 * rng.normal(size=(9, 2))
    z = th[:, :1] * xs[None, :] + th[:, 1:]
    mock = np.where(z > 0, z, 0.1 * z)
    logp, grad, chi2 = m.logprob_grad(th, 50.0)
    np.testing.assert_allclose(chi2, ((mock - ys) ** 2).sum(-1), rtol=1e-5)
    np.testing.assert_allclose(m.forward(th), mock, rtol=1e-5, atol=1e-6)


def test_split_trajectory_equals_the_fused_launch(gpu):
    """Launches with enough work run a trajectory as three kernels (begin / middle / end); chain by chain they
    perform the fused kernel's operations: same states, energies, accept decisions, step sizes, precisions."""
    from binf_b200 import _cabi
    g = load_golden("poly_n1000")
    args = (g["prior_means"], g["prior_variances"], float(g["gamma_shape"]), float(g["gamma_rate"]))
    rng = np.random.RandomState(5)
    q0 = np.array([2.0, -4.0, 1.0, 1.5]) + 0.02 * rng.normal(size=(333, 4))
    for flags in (0, _cabi.FLAG_GENERIC_SCALAR):
        for code, xs in ((POLY_CODE, g["xs"]), ):
            out = []
            for split in (0, 1):
                m = _cabi.Model.generic(code, 4, xs, g["ys"], *args, flags=flags)
                m.set_option("generic.split", split)
                assert m.get_option("generic.split") == split
                for mode in (_cabi.GIBBS_TAU_FIRST, _cabi.GIBBS_TAU_LAST, _cabi.GIBBS_NONE):
                    out.append(m.hmc_run(q0, 2.5, 0.013, 7, n_traj=4, n_adapt=3, gibbs_mode=mode, seed=21,
                                         want_end=True))
                out.append(m.hmc_run(q0[:5], 2.5, 0.004, 1, n_traj=1, seed=2, want_end=True))   # L = 1
            h = len(out) // 2
            for a, b in zip(out[:h], out[h:]):
                for k in ("q", "tau", "eps", "q_end", "p_end", "accepted", "n_accepted", "e_before", "e_after"):
                    np.testing.assert_array_equal(a[k], b[k], err_msg=k)
                np.testing.assert_allclose(a["stats"], b["stats"], rtol=1e-12)
            assert 0 < out[0]["n_accepted"].sum() < 4 * len(q0)
    # the lanes-per-chain mapping over global memory (rows beyond the constant bank) splits the same way
    n = 3500
    xs = np.sort(rng.uniform(-2, 2, size=n))
    ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2.0, -4.0, 1.0, 1.5]), 1 / np.sqrt(2.5))
    res = []
    for split in (0, 1):
        m = _cabi.Model.generic(POLY_CODE, 4, xs[:, None].copy(), ys, np.zeros(4), 5.0 * np.ones(4), 1.0, 0.2)
        assert m.get_option("generic.uniform_rows") == 0
        m.set_option("generic.split", split)
        res.append(m.hmc_run(q0[:40], 2.5, 0.006, 5, n_traj=2, gibbs_mode=_cabi.GIBBS_TAU_LAST, seed=4, want_end=True))
    for k in ("q", "tau", "q_end", "p_end", "accepted", "e_after"):
        np.testing.assert_array_equal(res[0][k], res[1][k], err_msg=k)


def test_compile_error_is_reported(gpu):
    from binf_b200 import _cabi
    with pytest.raises(_cabi.BinfB200Error) as e:
        _cabi.Model.generic("__device__ float binfb_mock(const float *t, const float *x, float *d) { return zz; }",
                            2, np.zeros(4), np.zeros(4))
    assert "zz" in str(e.value) and "user_model.cu" in str(e.value)


@pytest.mark.parametrize("n_data", [100, 1000, 3001, 5000])
def test_generic_polynomial_row_mappings_vs_oracle(gpu, n_data):
    """short data sets take the uniform-row mapping (rows in the module's constant bank: one warp per set
    below 128 rows, four above), long ones (> 48 KiB of rows) the lanes-per-chain mapping over global
    memory; ragged chain and row counts; several trajectories per launch with rejections"""
    from binf_b200 import _cabi
    rng = np.random.RandomState(n_data)
    xs = np.sort(rng.uniform(-2, 2, size=n_data))
    ys = rng.normal(np.polynomial.polynomial.polyval(xs, [2.0, -4.0, 1.0, 1.5]), 1 / np.sqrt(2.5))
    pm, pv = np.zeros(4), 5.0 * np.ones(4)
    m = _cabi.Model.generic(POLY_CODE, 4, xs[:, None].copy(), ys, pm, pv, 1.0, 0.2)
    pp = port.PolynomialPosterior(xs, ys, pm, pv, 1.0, 0.2)
    C = 77
    q0 = np.array([2.0, -4.0, 1.0, 1.5]) + 0.02 * rng.normal(size=(C, 4))
    logp, grad, chi2 = m.logprob_grad(q0, 2.5)
    np.testing.assert_allclose(logp, pp.log_prob(q0, 2.5), rtol=1e-5)
    ref = pp.gradient(q0, 2.5)
    assert np.all(np.abs(grad - ref) <= 1e-4 * np.max(np.abs(ref), axis=-1, keepdims=True))
    np.testing.assert_allclose(chi2, pp.chi2(q0), rtol=1e-5)
    eps, L = 0.2 / n_data ** 0.5, 6
    p0, u = rng.normal(size=q0.shape), rng.uniform(size=C)
    r = m.hmc_run(q0, 2.5, eps, L, p0=p0, u=u, want_end=True)
    o = port.hmc_sample(lambda c: pp.log_prob(c, 2.5), lambda c: pp.gradient(c, 2.5), q0, eps, L, p0, u)
    assert np.all(np.abs(r["q_end"] - o["q_end"]) <= 1e-4 * np.max(np.abs(o["q_end"])))
    np.testing.assert_allclose(r["e_after"], o["e_after"], rtol=1e-5)
    # Philox momenta, three trajectories per launch, a step size that rejects often: the same chains on the
    # built-in polynomial kernel (same streams, same arithmetic up to summation order)
    b = _cabi.Model.polynomial(xs, ys, 4, pm, pv, 1.0, 0.2)
    ra = m.hmc_run(q0, 2.5, 4 * eps, L, n_traj=3, seed=5, want_end=True)
    rb = b.hmc_run(q0, 2.5, 4 * eps, L, n_traj=3, seed=5, want_end=True)
    same = ra["n_accepted"] == rb["n_accepted"]
    assert same.mean() > 0.9 and (ra["n_accepted"] < 3).any()
    assert np.all(np.abs(ra["q"][same] - rb["q"][same]) <= 2e-3 * np.max(np.abs(rb["q"][same])))
