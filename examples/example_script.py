"""The reference's example_script.py (polynomial fit, BASELINE.json configs[0]) on the B200 path.

The body below is the reference's script line for line -- same imports (resolved to binf_b200 by
install_as_binf), same construction calls, same `samples.append(deepcopy(gips.sample()))` loop, same
thinning slice, `get_MAP` and `predict` -- minus the matplotlib plots, plus:
  --chains C    run C independent chains at once instead of one (state arrays get a leading axis)
  --hmc         sample the coefficients with the fused HMC kernel instead of RWMC
  --sink        keep the samples on the device with SampleSink instead of a Python list

    python examples/example_script.py --sweeps 3000
    python examples/example_script.py --sweeps 3000 --chains 4096 --hmc --sink
"""
import argparse
import os
import sys
from copy import deepcopy

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binf_b200  # noqa: E402

binf_b200.install_as_binf()

from binf.samplers import BinfState  # noqa: E402
from binf.example.samplers import make_sampler  # noqa: E402
from binf.example.misc import get_MAP, make_posterior, predict  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweeps", type=int, default=30000)
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--hmc", action="store_true")
    ap.add_argument("--sink", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    np.random.seed(args.seed)

    n_data_points = 20
    real_coeffs = np.array([2.0, -4.0, 1.0, 1.5])
    real_precision = 2.5
    polynomial = np.polynomial.polynomial.polyval
    xses = np.linspace(-2, 2, n_data_points)
    ys = np.random.normal(loc=polynomial(xses, real_coeffs), scale=1.0 / np.sqrt(real_precision))

    if args.chains:
        start = BinfState(dict(coefficients=np.ones((args.chains, 4)), precision=np.ones(args.chains)))
    else:
        start = BinfState(dict(coefficients=np.ones(4), precision=1.0))
    posterior = make_posterior(xses, ys, polynomial)
    gips = make_sampler(posterior, 0.02, start, nsteps=20) if args.hmc else make_sampler(posterior, 0.1, start)

    burn_in, thin = (2 * args.sweeps) // 3, 20
    samples, sink = [], None
    if args.sink:
        from binf.samplers.sink import SampleSink
        sink = SampleSink(max(args.chains, 1), 4, capacity=(args.sweeps - burn_in + thin - 1) // thin,
                          burn_in=burn_in, thin=thin)
    for i in range(args.sweeps):
        state = gips.sample()
        if sink is not None:
            sink.append(state, variable="coefficients", aux="precision")
        else:
            samples.append(deepcopy(state))
        if i % 500 == 0 and i > 0:
            print("#### Gibbs sampling step {} ####".format(i))
            print("acceptance rate: {}".format(gips.last_draw_stats["coefficients"]))

    if sink is not None:
        coeffs, precisions = sink.samples()                      # [n_kept, C, 4], [n_kept, C]
        summary = sink.summary()
        print("posterior mean of the coefficients:", summary["mean"], " R-hat:", summary["rhat"])
        flat_c, flat_t = coeffs.reshape(-1, 4), precisions.reshape(-1)
    else:
        samples_thin = samples[burn_in::thin]
        log_probs = np.array([np.mean(posterior.log_prob(**x.variables)) for x in samples_thin])
        MAP_coeffs, MAP_precision = get_MAP(samples_thin, log_probs)
        print("MAP coefficients:", np.asarray(MAP_coeffs).reshape(-1, 4).mean(axis=0), "precision:",
              np.mean(MAP_precision))
        flat_c = np.array([x.variables["coefficients"] for x in samples_thin]).reshape(-1, 4)
        flat_t = np.array([x.variables["precision"] for x in samples_thin]).reshape(-1)
    print("posterior mean:", flat_c.mean(axis=0), "true:", real_coeffs)
    print("precision mean:", flat_t.mean(), "true:", real_precision)
    # the prediction tube of the reference's last plot, as numbers: predictive density on a grid
    gx, gy = np.meshgrid(np.linspace(-2, 2, 5), np.linspace(-10, 10, 9))
    dens = predict(gx, gy, (flat_c, flat_t))
    print("predictive density, column sums (should be ~ 1 / dy):", dens.sum(axis=0) * (gy[1, 0] - gy[0, 0]))
    return flat_c, flat_t, dens


if __name__ == "__main__":
    main()
