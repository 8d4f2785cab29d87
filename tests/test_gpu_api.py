"""GPU: the reference-facing Python classes (Posterior, Likelihood, HMCSampler, GibbsSampler,
GammaSampler) against the golden vectors the unmodified reference produced."""
import numpy as np
import pytest

from conftest import load_golden
import binf_port as port

pytestmark = pytest.mark.gpu
polyval = np.polynomial.polynomial.polyval


def poly_posterior(g):
    from binf_b200.example.misc import make_posterior
    return make_posterior(g["xs"], g["ys"], polyval)


def test_posterior_log_prob_and_gradient(gpu):
    g = load_golden("poly_n20")
    post = poly_posterior(g)
    cond = post.conditional_factory(precision=float(g["tau"]))
    for c in range(4):
        assert cond.log_prob(coefficients=g["q0"][c].copy()) == pytest.approx(g["log_prob"][c], rel=1e-5)
        grad = cond.gradient(coefficients=g["q0"][c].copy())
        assert grad.shape == (4,) and np.all(np.abs(grad - g["gradient"][c]) <= 1e-4 * np.abs(g["gradient"][c]).max())
        # full posterior: both variables passed, the un-cloned Gamma prior has rate 0.2 (quirk Q2)
        assert post.log_prob(coefficients=g["q0"][c].copy(), precision=float(g["tau"])) == pytest.approx(
            g["full_log_prob"][c], rel=1e-5)
    # batched evaluation and per-chain precision
    lp = cond.log_prob(coefficients=g["q0"])
    np.testing.assert_allclose(lp, g["log_prob"], rtol=1e-5)
    # the SURVEY.md anchors
    assert cond.log_prob(coefficients=np.ones(4)) == pytest.approx(-600.6869482862852, rel=1e-5)
    lik = post.likelihoods["points"]
    pp = port.PolynomialPosterior(g["xs"], g["ys"], np.zeros(4), 5 * np.ones(4), 1.0, 1.0)
    assert lik.log_prob(coefficients=np.ones(4), precision=1.0) == pytest.approx(
        float(pp.likelihood_log_prob(np.ones(4), 1.0)), rel=1e-5)
    mock = lik.forward_model(coefficients=g["q0"][1])
    np.testing.assert_allclose(mock, polyval(g["xs"], g["q0"][1]), rtol=2e-6, atol=1e-5)


@pytest.mark.parametrize("name", ["poly_n20", "poly_n1000_mode"])
def test_hmc_sampler_matches_reference_sampler(gpu, name):
    from binf_b200.samplers.hmc import HMCSampler
    g = load_golden(name)
    cond = poly_posterior(g).conditional_factory(precision=float(g["tau"]))
    L, dt = int(g["nsteps"]), float(g["timestep"])
    # one chain at a time, numpy (D,) in and out like the reference
    for c in range(3):
        s = HMCSampler(cond, g["q0"][c].copy(), dt, L, variable_name="coefficients")
        ql, pl = s._leapfrog(g["q0"][c].copy(), g["p0"][c].copy(), dt, L)
        assert np.all(np.abs(ql - g["q_end"][c]) <= 2e-3 * np.abs(g["q_end"][c]).max())
        assert np.all(np.abs(pl - g["p_end"][c]) <= 2e-3 * max(1.0, np.abs(g["p_end"][c]).max()))
        new = s.sample(p0=g["p0"][c][None], u=g["u"][c:c + 1])
        assert new.shape == (4,) and new.dtype == np.float64
        if abs(np.log(g["u"][c]) + g["e_after"][c] - g["e_before"][c]) > 0.05:
            assert bool(s.last_move_accepted) == bool(g["accepted"][c])
            np.testing.assert_allclose(new, g["q_new"][c], rtol=2e-3, atol=2e-3)
        assert s.counter == 1 and s.n_accepted == int(s.last_move_accepted)
    # the whole batch in one launch
    s = HMCSampler(cond, g["q0"].copy(), dt, L, variable_name="coefficients")
    new = s.sample(p0=g["p0"], u=g["u"])
    decided = np.abs(np.log(g["u"]) + g["e_after"] - g["e_before"]) > 0.05
    assert np.array_equal(np.asarray(s.last_move_accepted)[decided], g["accepted"][decided])
    assert new.shape == g["q0"].shape
    assert s.acceptance_rate == pytest.approx(np.mean(s.last_move_accepted))
    assert set(s.last_draw_stats) == {"coefficients"}


def test_gibbs_sampler_generic_path_matches_reference(gpu):
    """sub-samplers called one after the other (fuse=False) with the reference's injected draws"""
    from binf_b200.example.samplers import GammaSampler
    from binf_b200.samplers import BinfState
    from binf_b200.samplers.gibbs import GibbsSampler
    from binf_b200.samplers.hmc import HMCSampler
    g = load_golden("poly_gibbs_n20")
    post = poly_posterior(g)
    c0, tau0 = g["c0"].copy(), float(g["tau0"])
    hmc = HMCSampler(post.conditional_factory(precision=tau0), c0.copy(), float(g["timestep"]),
                     int(g["nsteps"]), timestep_adaption_limit=int(g["limit"]), variable_name="coefficients")
    gam = GammaSampler(post.conditional_factory(coefficients=c0), tau0)
    gibbs = GibbsSampler(post, BinfState(dict(coefficients=c0.copy(), precision=tau0)),
                         {"coefficients": hmc, "precision": gam}, fuse=False)
    hmc_sample, gam_sample = hmc.sample, gam.sample
    for k in range(len(g["u"])):
        hmc.sample = lambda k=k: hmc_sample(p0=g["p0"][k][None], u=g["u"][k:k + 1])
        gam.sample = lambda k=k: gam_sample(gamma_draws=g["gamma_draws"][k:k + 1])
        st = gibbs.sample()
        assert bool(hmc.last_move_accepted) == bool(g["accepted"][k])
        np.testing.assert_allclose(st.variables["coefficients"], g["coefficients"][k], rtol=3e-3, atol=3e-3)
        assert st.variables["precision"] == pytest.approx(float(g["precision"][k]), rel=5e-3)
        assert hmc.timestep == pytest.approx(float(g["timesteps"][k]), rel=1e-5)
    assert gibbs.last_draw_stats["coefficients"].stepsize == pytest.approx(float(g["timesteps"][-1]), rel=1e-5)


def test_gibbs_sampler_fused_sweeps_posterior_moments(gpu):
    """2,048 chains x 150 fused sweeps: the precision marginal of the reference's Gibbs sampler
    (mean 2.54, sd 0.90 from 4,000 reference sweeps) and the acceptance rate."""
    from binf_b200.example.samplers import make_sampler
    from binf_b200.samplers import BinfState
    g = load_golden("poly_n20")
    post = poly_posterior(g)
    C = 2048
    start = BinfState(dict(coefficients=np.tile([2.0, -4.0, 1.0, 1.5], (C, 1)), precision=np.full(C, 2.5)))
    gibbs = make_sampler(post, 0.02, start, nsteps=20, seed=7)
    assert gibbs._fused_plan() is not None
    taus = []
    for k in range(150):
        st = gibbs.sample()
        if k >= 50:
            taus.append(st.variables["precision"].copy())
    taus = np.concatenate(taus)
    assert taus.mean() == pytest.approx(2.54, abs=0.12)
    assert taus.std() == pytest.approx(0.90, abs=0.12)
    assert 0.7 < gibbs.subsamplers["coefficients"].acceptance_rate < 0.95   # reference: 0.86
    # independent check of the same marginal with the CPU port of the reference sampler
    rng = np.random.RandomState(0)
    pp = port.PolynomialPosterior(g["xs"], g["ys"], np.zeros(4), 5 * np.ones(4), 1.0, 1.0)
    c, tau, ref = np.array([2.0, -4.0, 1.0, 1.5]), 2.5, []
    for k in range(1500):
        r = port.hmc_sample(lambda q: pp.log_prob(q, tau), lambda q: pp.gradient(q, tau), c, 0.02, 20,
                            rng.normal(size=4), rng.uniform())
        c = r["q"]
        tau = float(port.gamma_precision_sample(pp.chi2(c), 20, 1.0, 1.0, rng))
        ref.append(tau)
    assert taus.mean() == pytest.approx(np.mean(ref[300:]), abs=0.2)


def test_chromatin_api_matches_reference(gpu):
    from binf_b200.chromatin import make_chromatin_posterior
    from binf_b200.samplers.hmc import HMCSampler
    g = load_golden("chromatin_n30_big_step")
    n = int(g["n_beads"])
    post = make_chromatin_posterior(n, g["y"], float(g["alpha"]), float(g["d_c"]), float(g["k_bb"]),
                                    float(g["l0"]))
    assert post.variables == {"structure", "precision"} and post.differentiable_variables == {"structure"}
    cond = post.conditional_factory(precision=float(g["tau"]))
    for c in range(3):
        assert cond.log_prob(structure=g["q0"][c].copy()) == pytest.approx(g["log_prob"][c], rel=1e-5)
        grad = cond.gradient(structure=g["q0"][c].copy())
        assert np.all(np.abs(grad - g["gradient"][c]) <= 1e-4 * np.abs(g["gradient"][c]).max())
    s = HMCSampler(cond, g["q0"].copy(), float(g["timestep"]), int(g["nsteps"]), variable_name="structure")
    s.sample(p0=g["p0"], u=g["u"])
    decided = np.abs(np.log(g["u"]) + g["e_after"] - g["e_before"]) > 0.02
    assert np.array_equal(np.asarray(s.last_move_accepted)[decided], g["accepted"][decided])
    np.testing.assert_allclose(s.last_energies[0], g["e_before"], rtol=1e-5)
    mock = post.likelihoods["points"].forward_model(structure=g["q0"][0])
    import chromatin_port as chrom
    ref = chrom.ChromatinModel(n, g["y"], float(g["alpha"]), float(g["d_c"]), 4.0, 1.0).forward(g["q0"][0])
    np.testing.assert_allclose(mock, ref, rtol=1e-4, atol=2e-6)


def test_device_resident_state(gpu):
    """torch CUDA tensors in, tensors out: chains never leave HBM"""
    import torch
    from binf_b200.samplers.hmc import HMCSampler
    g = load_golden("poly_n1000_mode")
    cond = poly_posterior(g).conditional_factory(precision=float(g["tau"]))
    q = torch.as_tensor(g["q0"], dtype=torch.float32, device="cuda")
    p0 = torch.as_tensor(g["p0"], dtype=torch.float32, device="cuda")
    u = torch.as_tensor(g["u"], dtype=torch.float32, device="cuda")
    s = HMCSampler(cond, q, float(g["timestep"]), int(g["nsteps"]), variable_name="coefficients")
    new = s.sample(p0=p0, u=u)
    torch.cuda.synchronize()
    assert new.is_cuda and new.shape == q.shape
    decided = np.abs(np.log(g["u"]) + g["e_after"] - g["e_before"]) > 0.05
    assert np.array_equal(s.last_move_accepted.cpu().numpy()[decided], g["accepted"][decided])
    np.testing.assert_allclose(s.last_energies[0].cpu().numpy(), g["e_before"], rtol=1e-5)
    for _ in range(5):
        s.sample()
    assert s.counter == 6 and 0.0 < s.acceptance_rate <= 1.0


def test_fused_gibbs_on_device_resident_state_matches_host_path(gpu):
    """the same fused sweeps with numpy state (host buffers) and with CUDA tensors (chains never
    leave HBM): identical Philox streams => identical chains"""
    import torch
    import chromatin_port as chrom
    from binf_b200.chromatin import make_chromatin_posterior
    from binf_b200.example.samplers import GammaSampler
    from binf_b200.samplers import BinfState
    from binf_b200.samplers.gibbs import GibbsSampler
    from binf_b200.samplers.hmc import HMCSampler
    n, C = 64, 48
    X, y = chrom.synthetic_chromatin(n, seed=11)
    rng = np.random.RandomState(5)
    q0 = (X.reshape(-1)[None] + 0.05 * rng.normal(size=(C, 3 * n))).astype(np.float32)

    def build(structure, precision):
        post = make_chromatin_posterior(n, y)
        hmc = HMCSampler(post.conditional_factory(precision=precision), structure, 0.003, 8,
                         timestep_adaption_limit=4, variable_name="structure", seed=21)
        gam = GammaSampler(post.conditional_factory(structure=structure), precision, seed=22)
        return GibbsSampler(post, BinfState(dict(structure=structure, precision=precision)),
                            {"structure": hmc, "precision": gam})

    host = build(q0.astype(np.float64), np.full(C, 100.0))
    dev = build(torch.as_tensor(q0, device="cuda"), torch.full((C,), 100.0, device="cuda"))
    assert host._fused_plan() is not None and dev._fused_plan() is not None
    for _ in range(5):
        sh, sd = host.sample(), dev.sample()
    torch.cuda.synchronize()
    qd = sd.variables["structure"].cpu().numpy()
    np.testing.assert_allclose(qd, sh.variables["structure"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(sd.variables["precision"].cpu().numpy(), sh.variables["precision"], rtol=1e-6)
    hh, hd = host.subsamplers["structure"], dev.subsamplers["structure"]
    assert hh.acceptance_rate == pytest.approx(hd.acceptance_rate)
    np.testing.assert_allclose(hd.timestep, hh.timestep, rtol=1e-6)
    assert 100.0 < np.median(sh.variables["precision"]) < 1000.0     # noise sd 0.05 => tau ~ 400


@pytest.mark.parametrize("argv", [["--sweeps", "900"], ["--sweeps", "300", "--chains", "64", "--hmc", "--sink"]])
def test_reference_example_script_runs_on_the_device(gpu, argv, capsys):
    """BASELINE configs[0]: the reference's example_script.py flow (make_posterior, make_sampler, the
    sample loop, thinning, get_MAP, predict) unchanged, through install_as_binf()"""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("example_script", os.path.join(ROOT, "examples", "example_script.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    coeffs, precisions, dens = mod.main(argv)
    assert np.all(np.abs(coeffs.mean(axis=0) - [2.0, -4.0, 1.0, 1.5]) < [1.2, 1.2, 0.6, 0.6])
    assert 0.5 < precisions.mean() < 8.0
    assert np.all(dens >= 0) and dens.max() > 0.05
    out = capsys.readouterr().out
    assert "posterior mean" in out and (int(argv[1]) <= 500 or "acceptance rate" in out)


@pytest.mark.parametrize("argv", [["--beads", "96", "--chains", "128", "--sweeps", "40"],
                                  ["--beads", "64", "--chains", "64", "--sweeps", "24", "--excluded-volume", "1.0"],
                                  ["--beads", "80", "--chains", "96", "--sweeps", "30", "--contact", "algebraic"]])
def test_chromatin_inference_example(gpu, argv):
    """examples/chromatin_inference.py: posterior behind the reference API -> lowered model -> fused Gibbs
    sweeps -> on-device sink -> cross-rank summary (single rank here)"""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("chromatin_inference",
                                                  os.path.join(ROOT, "examples", "chromatin_inference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.main(argv)
    assert out["chains"] == int(argv[3])
    assert 0.5 < out["acceptance"] <= 1.0
    assert np.isfinite(out["precision"]) and out["precision"] > 1.0
    assert out["contact_drmsd"] < 1.0          # chains start 0.3 from the truth and the data hold them there
