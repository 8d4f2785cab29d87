"""Sample sink (binfb_sink_*) against the list-and-slice bookkeeping of the reference's driver
(oracle/sink_port.py: example_script.py:32-42, binf/example/misc.py:18-22)."""
import numpy as np
import pytest

import sink_port

pytestmark = pytest.mark.gpu


def _run(C, D, n_sweeps, burn_in, thin, capacity, seed, nan_logp=False):
    from binf_b200 import _cabi
    rng = np.random.RandomState(seed)
    sink = _cabi.Sink(C, D, capacity, burn_in, thin, track_map=True)
    ref = sink_port.ListSink(burn_in, thin, capacity)
    drift = rng.normal(size=(C, D))
    for t in range(n_sweeps):
        q = (30.0 + drift + rng.normal(size=(C, D))).astype(np.float32)
        aux = rng.gamma(3.0, size=C).astype(np.float32)
        logp = rng.normal(size=C) * 10 - 1000.0
        if nan_logp and t % 3 == 1:
            logp[::5] = np.nan
        sink.push(q, aux, logp)
        ref.append(q, aux, logp)
    return sink, ref


@pytest.mark.parametrize("C,D,n,burn,thin,cap", [(64, 12, 40, 7, 3, 100), (33, 7, 25, 0, 1, 8),
                                                 (5, 3000, 12, 2, 2, 3), (1, 4, 30, 10, 4, 2)])
def test_ring_moments_map_vs_lists(gpu, C, D, n, burn, thin, cap):
    sink, ref = _run(C, D, n, burn, thin, cap, seed=C + D)
    info = sink.info()
    q_ref, a_ref, _ = ref.thinned()
    assert info["n_pushed"] == n and info["n_moment"] == n - burn
    assert info["n_kept"] == len(ref.samples[burn::thin])
    # ring = the last `cap` kept samples, bit-exact copies
    q, aux = sink.read()
    assert q.shape[0] == len(q_ref)
    np.testing.assert_array_equal(q, np.array(q_ref, dtype=np.float32))
    np.testing.assert_array_equal(aux, np.array(a_ref, dtype=np.float32))
    # running moments (float64 Welford vs two-pass numpy): round-off only
    mean, var = sink.moments()
    m_ref, v_ref = ref.moments()
    np.testing.assert_allclose(mean, m_ref, rtol=1e-13, atol=0)
    np.testing.assert_allclose(var, v_ref, rtol=1e-10, atol=0)
    # MAP per chain over the kept samples: identical selection
    logp, qm, am = sink.map_estimate()
    qm_ref, am_ref, l_ref = ref.get_MAP()
    np.testing.assert_array_equal(logp, l_ref)
    np.testing.assert_array_equal(qm, qm_ref.astype(np.float32))
    np.testing.assert_array_equal(am, am_ref.astype(np.float32))


def test_summary_rhat_ess(gpu):
    sink, ref = _run(512, 20, 60, 10, 1, 0, seed=3)
    s, r = sink.summary(), ref.summary()
    for k in ("mean", "var", "rhat", "ess_per_chain"):
        np.testing.assert_allclose(s[k], r[k], rtol=1e-9, err_msg=k)
    # chains here have different means (drift ~ N(0,1)) on top of unit noise: R-hat well above 1
    assert np.all(s["rhat"] > 1.2)


def test_summary_well_mixed_chains(gpu):
    """independent N(mu, 1) draws: R-hat -> 1 and ESS per chain -> n (CLT bounds over 4096 chains)"""
    from binf_b200 import _cabi
    C, D, n = 4096, 8, 200
    rng = np.random.RandomState(0)
    sink = _cabi.Sink(C, D)
    for t in range(n):
        sink.push((5.0 + rng.normal(size=(C, D))).astype(np.float32))
    s = sink.summary()
    assert np.all(np.abs(s["mean"] - 5.0) < 5.0 / np.sqrt(C * n))
    assert np.all(np.abs(s["var"] - 1.0) < 0.01)
    assert np.all(np.abs(s["rhat"] - 1.0) < 2e-3)
    assert np.all(np.abs(s["ess_per_chain"] / n - 1.0) < 0.1)


def test_nan_logp_never_wins_and_ties_keep_first(gpu):
    sink, ref = _run(40, 8, 20, 0, 1, 4, seed=11, nan_logp=True)
    logp, qm, _ = sink.map_estimate()
    qm_ref, _, l_ref = ref.get_MAP()
    np.testing.assert_array_equal(logp, l_ref)
    np.testing.assert_array_equal(qm, qm_ref.astype(np.float32))
    from binf_b200 import _cabi
    s2 = _cabi.Sink(3, 4, 2, track_map=True)
    a, b = np.ones((3, 4), np.float32), 2 * np.ones((3, 4), np.float32)
    s2.push(a, None, np.zeros(3))
    s2.push(b, None, np.zeros(3))     # equal log-prob: the first sample stays (numpy.argmax)
    _, q, _ = s2.map_estimate()
    np.testing.assert_array_equal(q, a)


def test_device_tensors_and_errors(gpu):
    import torch
    from binf_b200 import _cabi
    from binf_b200.samplers.sink import SampleSink
    C, D = 128, 3000
    sink = SampleSink(C, D, capacity=2, thin=2)
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randn(C, D, device="cuda", generator=g) for _ in range(5)]
    tau = torch.rand(C, device="cuda", generator=g)
    for x in xs:
        sink.append(x, aux=tau)
    torch.cuda.synchronize()
    assert len(sink) == 3 and sink.n_sweeps == 5
    q, aux = sink.samples()
    np.testing.assert_array_equal(q, torch.stack([xs[2], xs[4]]).cpu().numpy())
    np.testing.assert_array_equal(aux[0], tau.cpu().numpy())
    mean, var = sink.moments()
    np.testing.assert_allclose(mean, torch.stack(xs).double().mean(0).cpu().numpy(), rtol=0, atol=1e-14)
    with pytest.raises(_cabi.BinfB200Error):
        sink.samples(first=0, count=1)          # overwritten: no longer in the ring
    with pytest.raises(_cabi.BinfB200Error):
        sink.get_MAP()                          # created without track_map
    with pytest.raises(_cabi.BinfB200Error):
        _cabi.Sink(4, 4).summary()              # fewer than 2 sweeps


def test_sums_merge_across_sinks_equals_one_sink(gpu):
    """two sinks holding disjoint chain ranges (as two ranks would) merge to the summary of one sink
    over all chains"""
    from binf_b200 import _cabi
    from binf_b200.distributed import merge_sink_sums
    rng = np.random.RandomState(8)
    C, D, n = 96, 12, 30
    a, b, whole = _cabi.Sink(40, D), _cabi.Sink(C - 40, D), _cabi.Sink(C, D)
    drift = rng.normal(size=(C, D))
    for t in range(n):
        x = (10.0 + drift + rng.normal(size=(C, D))).astype(np.float32)
        a.push(x[:40]), b.push(x[40:]), whole.push(x)
    merged, ref = merge_sink_sums([a.sums(), b.sums()]), whole.summary()
    for k in ("mean", "var", "rhat", "ess_per_chain"):
        np.testing.assert_allclose(merged[k], ref[k], rtol=1e-9, err_msg=k)
