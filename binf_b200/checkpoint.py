"""Checkpoint / resume of a sampling run (the reference has none -- its driver
keeps deep copies in a Python list and numpy's global RNG state, example_script.py:32-34).

A run on the B200 path is fully described by a handful of arrays: the state variables (q [C, D],
precision [C]), the per-chain step sizes, the counters, and the Philox coordinates (seed, draw index,
chain base) -- the random streams are counter-based, so a resumed run continues bit-for-bit where the
saved one stopped.  `save` writes them to one .npz; `load` puts them back into freshly constructed
samplers of the same kinds.

    save("run.npz", gibbs)                      # or an HMCSampler / RWMCSampler / GammaSampler
    ...
    gibbs = make_sampler(posterior, 0.02, start_state, nsteps=20)   # same construction as the original run
    load("run.npz", gibbs)
"""
import numpy as np


def _np(x):
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _like(template, value):
    """value (numpy) converted to the kind of `template` (CUDA tensor / numpy array / python scalar)"""
    if hasattr(template, "detach"):
        import torch
        return torch.as_tensor(value, dtype=template.dtype, device=template.device)
    if np.ndim(template) == 0 and not isinstance(template, np.ndarray):
        return type(template)(np.asarray(value).reshape(()).item()) if np.ndim(value) == 0 else np.array(value)
    return np.array(value)


def _counter_like(state, value):
    """acceptance counters in the kind the sampler's own update produces: a python int for one chain, an int64
    CUDA tensor next to a device-resident state (hmc.py / example/samplers.py add `_nacc_dev` to it), else numpy"""
    if not np.ndim(value):
        return int(value)
    if hasattr(state, "detach"):
        import torch
        return torch.as_tensor(np.asarray(value), dtype=torch.int64, device=state.device)
    return np.array(value)


def _sampler_state(s):
    from binf_b200.example.samplers import GammaSampler, RWMCSampler
    from binf_b200.samplers.hmc import HMCSampler
    if isinstance(s, HMCSampler):
        return dict(kind="hmc", state=_np(s.state), eps=_np(s.timestep), counter=s.counter, draw=s._draw,
                    seed=s.seed, chain_base=s.chain_base, n_accepted=_np(s.n_accepted))
    if isinstance(s, RWMCSampler):
        return dict(kind="rwmc", state=_np(s.state), stepsize=_np(s.stepsize), n_moves=s._n_moves, draw=s._draw,
                    seed=s.seed, chain_base=s.chain_base, n_accepted=_np(s._n_accepted_moves))
    if isinstance(s, GammaSampler):
        return dict(kind="gamma", state=_np(s.state), draw=s._draw, seed=s.seed, chain_base=s.chain_base)
    raise TypeError("cannot checkpoint a %s" % type(s).__name__)


def _restore_sampler(s, d):
    kind = str(d["kind"])
    if kind == "hmc":
        s.state = _like(s.state, d["state"])
        s.timestep = d["eps"] if np.ndim(d["eps"]) else float(d["eps"])
        s.counter, s._draw = int(d["counter"]), int(d["draw"])
        s.seed, s.chain_base = int(d["seed"]), int(d["chain_base"])
        s.n_accepted = _counter_like(s.state, d["n_accepted"])
    elif kind == "rwmc":
        s.state = _like(s.state, d["state"])
        s.stepsize = d["stepsize"] if np.ndim(d["stepsize"]) else float(d["stepsize"])
        s._n_moves, s._draw = int(d["n_moves"]), int(d["draw"])
        s.seed, s.chain_base = int(d["seed"]), int(d["chain_base"])
        s._n_accepted_moves = _counter_like(s.state, d["n_accepted"])
    elif kind == "gamma":
        s.state = _like(s.state, d["state"])
        s._draw, s.seed, s.chain_base = int(d["draw"]), int(d["seed"]), int(d["chain_base"])
    else:
        raise ValueError("unknown sampler kind in checkpoint: %r" % kind)


def save(path, sampler):
    """sampler: a GibbsSampler (its state and every sub-sampler are saved) or a single sampler"""
    from binf_b200.samplers.gibbs import GibbsSampler
    flat = {}
    if isinstance(sampler, GibbsSampler):
        flat["__gibbs__"] = np.array(sorted(sampler.subsamplers))
        for name, value in sampler.state.variables.items():
            flat["var/" + name] = _np(value)
        for name, sub in sampler.subsamplers.items():
            for k, v in _sampler_state(sub).items():
                flat["sub/%s/%s" % (name, k)] = np.asarray(v)
    else:
        for k, v in _sampler_state(sampler).items():
            flat["one/" + k] = np.asarray(v)
    np.savez(path, **flat)


def load(path, sampler):
    """restore into `sampler`, constructed like the one that was saved; returns it"""
    from binf_b200.samplers.gibbs import GibbsSampler
    z = np.load(path, allow_pickle=False)
    if isinstance(sampler, GibbsSampler):
        if "__gibbs__" not in z.files:
            raise ValueError("checkpoint does not hold a GibbsSampler")
        if sorted(sampler.subsamplers) != list(z["__gibbs__"]):
            raise ValueError("checkpoint sub-samplers %s do not match %s"
                             % (list(z["__gibbs__"]), sorted(sampler.subsamplers)))
        current = sampler.state.variables
        for name in current:
            sampler.state.update_variables(**{name: _like(current[name], z["var/" + name])})
        for name, sub in sampler.subsamplers.items():
            prefix = "sub/%s/" % name
            _restore_sampler(sub, {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)})
        sampler._update_subsampler_states()
    else:
        _restore_sampler(sampler, {k[4:]: z[k] for k in z.files if k.startswith("one/")})
    return sampler
