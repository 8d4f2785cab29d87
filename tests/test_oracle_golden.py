"""The oracle port against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py), and against the live reference when /root/reference exists."""
import numpy as np
import pytest

from conftest import load_golden
import binf_port as port
import chromatin_port as chrom
import ref_import

POLY_CASES = ["poly_n20", "poly_n1000", "poly_n1000_L5", "poly_n1000_mode", "poly_n77_mode"]
CHROM_CASES = ["chromatin_n24", "chromatin_n37_L20", "chromatin_n30_big_step", "chromatin_ev_n28",
               "chromatin_alg_n26", "chromatin_alg_ev_n22"]


def _poly(g):
    return port.PolynomialPosterior(g["xs"], g["ys"], g["prior_means"], g["prior_variances"],
                                    float(g["gamma_shape"]), float(g["gamma_rate"]))


@pytest.mark.parametrize("name", POLY_CASES)
def test_polynomial_port_matches_reference_vectors(name):
    g = load_golden(name)
    pp, tau = _poly(g), float(g["tau"])
    np.testing.assert_allclose(pp.log_prob(g["q0"], tau), g["log_prob"], rtol=1e-12)
    np.testing.assert_allclose(pp.gradient(g["q0"], tau), g["gradient"], rtol=1e-10, atol=1e-9)
    r = port.hmc_sample(lambda q: pp.log_prob(q, tau), lambda q: pp.gradient(q, tau), g["q0"],
                        float(g["timestep"]), int(g["nsteps"]), g["p0"], g["u"])
    np.testing.assert_allclose(r["q_end"], g["q_end"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r["p_end"], g["p_end"], rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(r["e_before"], g["e_before"], rtol=1e-12)
    np.testing.assert_allclose(r["e_after"], g["e_after"], rtol=1e-10)
    assert np.array_equal(r["accepted"], g["accepted"])
    np.testing.assert_allclose(r["q"], g["q_new"], rtol=1e-9, atol=1e-10)


def test_quirks_are_in_the_vectors():
    g = load_golden("poly_n20")
    # Q2: conditional pdfs carry rate == shape (1.0) while the full posterior has 0.2
    assert float(g["gamma_rate"]) == 1.0 and float(g["full_gamma_rate"]) == 0.2
    np.testing.assert_allclose(g["full_log_prob"] - g["log_prob"], float(g["tau"]) * 0.8, rtol=1e-9)
    # Q1: the gradient omits the Gaussian prior term (c - mu)/v
    pp = _poly(g)
    fixed = port.PolynomialPosterior(g["xs"], g["ys"], g["prior_means"], g["prior_variances"],
                                     1.0, 1.0, prior_grad=True)
    diff = fixed.gradient(g["q0"], 2.5) - pp.gradient(g["q0"], 2.5)
    np.testing.assert_allclose(diff, (g["q0"] - g["prior_means"]) / g["prior_variances"], rtol=1e-9)
    # SURVEY.md 8(c) anchors for chain 0 = ones(4)
    np.testing.assert_allclose(g["log_prob"][0], -600.6869482862852, rtol=1e-13)
    np.testing.assert_allclose(g["gradient"][0], [-68.00394064, 284.268345, -102.47753112,
                                                  714.87851861], rtol=1e-8)


def test_gibbs_sweep_port_matches_reference_vectors():
    g = load_golden("poly_gibbs_n20")
    pp = port.PolynomialPosterior(g["xs"], g["ys"], g["prior_means"], g["prior_variances"],
                                  float(g["hmc_gamma_shape"]), float(g["hmc_gamma_rate"]))
    c, tau, dt = g["c0"].copy(), float(g["tau0"]), float(g["timestep"])
    n_data, limit = len(g["xs"]), int(g["limit"])
    for k in range(len(g["u"])):
        r = port.hmc_sample(lambda q: pp.log_prob(q, tau), lambda q: pp.gradient(q, tau), c, dt,
                            int(g["nsteps"]), g["p0"][k], g["u"][k])
        dt = float(port.adapt_timestep(dt, r["accepted"], k + 1, limit))
        c = r["q"]
        shape, rate = port.gamma_precision_params(pp.chi2(c), n_data, float(g["gam_gamma_shape"]),
                                                  float(g["gam_gamma_rate"]))
        assert shape == pytest.approx(float(g["gamma_shapes"][k]))
        tau = g["gamma_draws"][k] / rate
        np.testing.assert_allclose(c, g["coefficients"][k], rtol=1e-9)
        assert tau == pytest.approx(float(g["precision"][k]), rel=1e-10)
        assert dt == pytest.approx(float(g["timesteps"][k]), rel=1e-12)
        assert bool(r["accepted"]) == bool(g["accepted"][k])
    # Q4: limit = 6 => exactly 5 adaptions
    assert len(set(np.round(g["timesteps"][4:], 12))) == 1


@pytest.mark.parametrize("name", CHROM_CASES)
def test_chromatin_port_matches_reference_vectors(name):
    g = load_golden(name)
    m = chrom.ChromatinModel(int(g["n_beads"]), g["y"], float(g["alpha"]), float(g["d_c"]),
                             float(g["k_bb"]), float(g["l0"]), 0.0, float(g["gamma_shape"]),
                             float(g["gamma_rate"]), float(g.get("ev_k", 0.0)), float(g.get("ev_d", 0.0)),
                             contact=str(g.get("contact", "logistic")))
    tau = float(g["tau"])
    for c in range(g["q0"].shape[0]):
        assert m.log_prob(g["q0"][c], tau) == pytest.approx(float(g["log_prob"][c]), rel=1e-12)
        np.testing.assert_allclose(m.gradient(g["q0"][c], tau), g["gradient"][c], rtol=1e-8, atol=1e-8)
        r = port.hmc_sample(lambda q: m.log_prob(q, tau), lambda q: m.gradient(q, tau), g["q0"][c],
                            float(g["timestep"]), int(g["nsteps"]), g["p0"][c], g["u"][c])
        np.testing.assert_allclose(r["q_end"], g["q_end"][c], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(r["p_end"], g["p_end"][c], rtol=1e-7, atol=1e-8)
        assert bool(r["accepted"]) == bool(g["accepted"][c])


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present on this box")
def test_live_reference_reproduces_golden_vectors():
    """Build container only: the fixtures are what the unmodified reference computes today."""
    binf = ref_import.install()
    from binf.example.misc import make_posterior
    g = load_golden("poly_n20")
    polyval = np.polynomial.polynomial.polyval
    cond = make_posterior(g["xs"], g["ys"], polyval).conditional_factory(precision=float(g["tau"]))
    for c in range(4):
        assert cond.log_prob(coefficients=g["q0"][c].copy()) == pytest.approx(g["log_prob"][c], rel=1e-13)
        np.testing.assert_allclose(cond.gradient(coefficients=g["q0"][c].copy()), g["gradient"][c], rtol=1e-12)
    gc = load_golden("chromatin_n24")
    m = chrom.ChromatinModel(int(gc["n_beads"]), gc["y"], float(gc["alpha"]), float(gc["d_c"]),
                             float(gc["k_bb"]), float(gc["l0"]), 0.0, 1.0, 1.0)
    condc = chrom.reference_posterior(binf, m).conditional_factory(precision=float(gc["tau"]))
    np.testing.assert_allclose(condc.gradient(structure=gc["q0"][0].copy()), gc["gradient"][0], rtol=1e-10)
    ga = load_golden("chromatin_alg_ev_n22")     # the algebraic contact function + excluded volume
    assert str(ga["contact"]) == "algebraic"
    m = chrom.ChromatinModel(int(ga["n_beads"]), ga["y"], float(ga["alpha"]), float(ga["d_c"]), float(ga["k_bb"]),
                             float(ga["l0"]), 0.0, 1.0, 1.0, float(ga["ev_k"]), float(ga["ev_d"]), contact="algebraic")
    conda = chrom.reference_posterior(binf, m).conditional_factory(precision=float(ga["tau"]))
    np.testing.assert_allclose(conda.gradient(structure=ga["q0"][1].copy()), ga["gradient"][1], rtol=1e-10)
    assert conda.log_prob(structure=ga["q0"][1].copy()) == pytest.approx(float(ga["log_prob"][1]), rel=1e-13)


def test_rwmc_and_predict_port_vs_reference_outputs():
    """oracle restatement of RWMCSampler.sample / predict vs the fixture generated by running the
    reference itself (oracle/make_golden.py rwmc)"""
    g = load_golden("poly_rwmc_n20")
    pp = port.PolynomialPosterior(g["xs"], g["ys"], g["prior_means"], g["prior_variances"],
                                  float(g["gamma_shape"]), float(g["gamma_rate"]))
    tau = float(g["tau"])
    n_moves, n_chains = g["u"].shape
    for c in range(n_chains):
        state = g["q0"][c]
        for k in range(n_moves):
            r = port.rwmc_sample(lambda x: pp.log_prob(x, tau), state, g["change"][k, c], g["u"][k, c])
            assert r["accepted"] == bool(g["accepted"][k, c])
            np.testing.assert_allclose(r["state"], g["states"][k + 1, c], rtol=0, atol=1e-12)
            state = r["state"]
    flat = g["states"].reshape(-1, 4)
    pred = [port.predict(a, b, flat, g["pred_tau"]) for a, b in zip(g["pred_x"], g["pred_y"])]
    np.testing.assert_allclose(pred, g["pred"], rtol=1e-12)


DECAY_F = staticmethod(lambda th, x: th[..., 0:1] * np.exp(-th[..., 1:2] * x) + th[..., 2:3])
DECAY_J = staticmethod(lambda th, x: np.stack([np.exp(-th[..., 1:2] * x), -th[..., 0:1] * x * np.exp(-th[..., 1:2] * x),
                                               np.ones_like(th[..., 0:1] * x)], axis=-2))


def test_user_model_port_vs_reference_outputs():
    """a user-defined AbstractForwardModel run by the reference itself (oracle/make_golden.py user)"""
    g = load_golden("user_decay_n200")
    um = port.UserModelPosterior(g["xs"], g["ys"], DECAY_F.__func__, DECAY_J.__func__, g["prior_means"],
                                 g["prior_variances"], float(g["gamma_shape"]), float(g["gamma_rate"]))
    tau = float(g["tau"])
    np.testing.assert_allclose(um.log_prob(g["q0"], tau), g["log_prob"], rtol=1e-12)
    np.testing.assert_allclose(um.gradient(g["q0"], tau), g["gradient"], rtol=1e-10, atol=1e-9)
    r = port.hmc_sample(lambda q: um.log_prob(q, tau), lambda q: um.gradient(q, tau), g["q0"],
                        float(g["timestep"]), int(g["nsteps"]), g["p0"], g["u"])
    np.testing.assert_allclose(r["q_end"], g["q_end"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r["e_after"], g["e_after"], rtol=1e-10)
    assert np.array_equal(r["accepted"], g["accepted"]) and not g["accepted"].all()


def test_port_reproduces_a_sample_of_the_acceptance_fixture():
    """chromatin_accept_n64: energy differences and decisions of 10,240 seeded L = 20 trajectories (written by
    oracle/make_golden.py, whose first chains went through the reference's own HMCSampler); spot-check chains
    spread over the fixture against the port, and the fixture's own consistency (hmc.py:151)"""
    g = load_golden("chromatin_accept_n64")
    n, C = int(g["n_beads"]), int(g["n_chains"])
    (alpha, d_c, k_bb, l0), y, q0, p0, u = chrom.acceptance_inputs(n, C, int(g["seed"]))
    assert float(np.sum(y.astype(np.float64))) == float(g["y_checksum"])
    model = chrom.ChromatinModel(n, y, alpha, d_c, k_bb, l0)
    tau, eps, L = float(g["tau"]), float(g["timestep"]), int(g["nsteps"])
    for c in (0, 1, 4097, C - 1):
        r = port.hmc_sample(lambda q: model.log_prob(q, tau), lambda q: model.gradient(q, tau), q0[c], eps, L, p0[c], u[c])
        assert abs((r["e_after"] - r["e_before"]) - g["dh"][c]) < 1e-9
        assert bool(r["accepted"]) == bool(g["accepted"][c])
    np.testing.assert_array_equal(g["accepted"], u < np.exp(np.clip(-g["dh"], -308, 709)))
    assert C >= 10000 and 0.7 < g["accepted"].mean() < 0.9


@pytest.mark.parametrize("contact", ["logistic", "algebraic"])
def test_chromatin_port_gradient_is_the_gradient_of_its_energy(contact):
    """SURVEY.md A.4 item 3 on the oracle itself, for both contact functions, with excluded volume, confinement and
    tempering switched on: central finite differences of -log_prob against gradient()."""
    n = 12
    X, y = chrom.synthetic_chromatin(n, 1.6, 2.1, seed=3, contact=contact)
    m = chrom.ChromatinModel(n, y, 1.6, 2.1, 3.0, 1.0, conf_s=4.0, ev_k=2.0, ev_d=1.4, contact=contact)
    rng = np.random.RandomState(0)
    q = X.reshape(-1) + 0.2 * rng.normal(size=3 * n)
    tau, beta, h = 35.0, 0.7, 1e-6
    g = m.gradient(q, tau, beta)
    fd = np.array([-(m.log_prob(q + h * e, tau, beta) - m.log_prob(q - h * e, tau, beta)) / (2 * h)
                   for e in np.eye(3 * n)])
    np.testing.assert_allclose(g, fd, rtol=1e-6, atol=1e-6)
    # the dense Jacobian the reference's Likelihood contracts (likelihoods.py:152-155) gives the same force
    jg = m.jacobian_dense(q).dot(tau * (m.forward(q) - m.y)) * beta
    np.testing.assert_allclose(jg, m.likelihood_gradient(q, tau, beta), rtol=1e-10, atol=1e-12)
    d = np.linspace(0.05, 9.0, 50)
    f0, df = chrom.contact_function(d, 1.6, 2.1, contact)
    fp, _ = chrom.contact_function(d + 1e-6, 1.6, 2.1, contact)
    fm, _ = chrom.contact_function(d - 1e-6, 1.6, 2.1, contact)
    np.testing.assert_allclose(df, (fp - fm) / 2e-6, rtol=1e-6, atol=1e-9)
    assert np.all((f0 > 0) & (f0 < 1)) and np.all(df < 0)          # a contact probability, falling with distance
