"""ORACLE (test infrastructure, not product code): host restatement of the label-swap replica exchange
(binfb_rex_pack / binfb_rex_decide / binfb_rex_select, csrc/misc.cu) in numpy.

The reference has no replica-exchange code (it is alluded to in docstrings only:
binf/samplers/hmc.py:171-177, binf/samplers/gibbs.py:117,143); the scheme is build-defined:
replica a at inverse temperature beta_a holds a state with untempered log-likelihood l_a, its partner b
likewise; the exchange is accepted iff u < exp(-(beta_a - beta_b)(l_a - l_b)).  Parity of the device
kernels against this file is bit-exact on the decisions (the Philox4x32-10 stream is restated below).

Used by tests/ only: the gloo (CPU) tests run the protocol of binf_b200.distributed with `HostOps` in
place of the C-ABI kernels, the GPU tests compare the kernels with it."""
import numpy as np

RECORD = np.dtype([("ll", "<f8"), ("tidx", "<i4"), ("eps", "<f4")])   # 16 bytes, the wire format
assert RECORD.itemsize == 16
RNG_SWAP = 4
M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(seed, chain, draw_lo, kind_elem):
    """csrc/common.cuh:philox4x32_10, vectorised over `chain` (uint64 array); returns the x word"""
    chain = np.asarray(chain, dtype=np.uint64)
    c = [chain & M32, chain >> np.uint64(32), np.full(chain.shape, draw_lo, dtype=np.uint64),
         np.full(chain.shape, kind_elem, dtype=np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & M32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & M32]
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c[0]


def unit_open0(x):
    """csrc/common.cuh:u32_to_unit_open0 (float32 arithmetic)"""
    return ((x >> np.uint64(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)


def swap_uniform(seed, attempt, lo, column):
    key = (int(seed) ^ ((int(lo) + 1) * 0xD6E8FEB86659FD93)) & 0xFFFFFFFFFFFFFFFF
    kind = ((RNG_SWAP << 24) | ((int(attempt) >> 32) & 0xFFFFFF)) & 0xFFFFFFFF
    return unit_open0(philox4x32_10(key, column, int(attempt) & 0xFFFFFFFF, kind)).astype(np.float64)


def pack(chi2, tau, eps, tidx, n_data):
    """records of the local chains: log L = -tau chi^2 / 2 + n_data log(tau) / 2"""
    rec = np.empty(len(chi2), dtype=RECORD)
    t = np.asarray(tau, dtype=np.float32).astype(np.float64)
    rec["ll"] = -0.5 * t * np.asarray(chi2, dtype=np.float64) + 0.5 * float(n_data) * np.log(t)
    rec["tidx"] = tidx
    rec["eps"] = eps
    return rec


def decide(records_all, world, rank, n_chains, n_columns, betas, seed, attempt, ll_shift=0.0):
    """One attempt seen from `rank`.  records_all: RECORD array [world * n_chains] (rank-major).
    Returns (tidx, beta, eps, accept, pair_counts [T-1, 2], temp_stats [T, 3]) of the local chains."""
    betas = np.asarray(betas, dtype=np.float64)
    T = len(betas)
    allr = np.asarray(records_all).reshape(world, n_chains)
    me = allr[rank]
    rows = n_chains // n_columns
    tidx = me["tidx"].copy()
    beta = betas[np.clip(tidx, 0, T - 1)].astype(np.float32)
    eps = me["eps"].copy()
    accept = np.zeros(n_chains, dtype=np.uint8)
    pair_counts = np.zeros((max(T - 1, 0), 2), dtype=np.uint64)
    temp_stats = np.zeros((T, 3))
    for i in range(n_chains):
        k, col = int(me["tidx"][i]), i % n_columns
        if 0 <= k < T and me["ll"][i] == me["ll"][i]:
            d = me["ll"][i] - ll_shift
            temp_stats[k] += (1.0, d, d * d)
        kp = k + 1 if ((k + (attempt & 1)) & 1) == 0 else k - 1
        if not (0 <= k < T and 0 <= kp < T):
            continue
        other = None
        for r in range(world):
            for row in range(rows):
                o = allr[r, row * n_columns + col]
                if o["tidx"] == kp:
                    other = o
                    break
            if other is not None:
                break
        if other is None:
            continue
        lo = min(k, kp)
        delta = (betas[k] - betas[kp]) * (me["ll"][i] - other["ll"])
        u = swap_uniform(seed, attempt, lo, np.array([col]))[0]
        acc = bool(delta == delta) and u < np.exp(min(709.0, max(-308.0, -delta)))
        if k == lo:
            pair_counts[lo, 0] += 1
            pair_counts[lo, 1] += int(acc)
        if acc:
            tidx[i], beta[i], eps[i], accept[i] = kp, np.float32(betas[kp]), other["eps"], 1
    return tidx, beta, eps, accept, pair_counts, temp_stats


def select(q, aux, tidx, k_sel, n_columns):
    q = np.asarray(q)
    out_q = np.zeros((n_columns, q.shape[1]), dtype=q.dtype)
    out_aux = np.zeros(n_columns, dtype=np.float32)
    for c in range(q.shape[0]):
        if tidx[c] == k_sel:
            out_q[c % n_columns] = q[c]
            if aux is not None:
                out_aux[c % n_columns] = aux[c]
    return out_q, out_aux


class HostOps(object):
    """drop-in for binf_b200.distributed.DeviceOps on CPU tensors (gloo tests of the protocol)"""

    def pack(self, chi2, tau, eps, tidx, n_data, records):
        import torch
        rec = pack(chi2.numpy(), tau.numpy(), eps.numpy(), tidx.numpy(), n_data)
        records.copy_(torch.from_numpy(rec.view(np.uint8).copy()))

    def decide(self, records_all, world, rank, n_chains, n_columns, betas, seed, attempt, ll_shift, tidx, beta, eps,
               accept, pair_counts, temp_stats):
        import torch
        allr = records_all.numpy().view(RECORD)
        t, b, e, a, pc, ts = decide(allr, world, rank, n_chains, n_columns, betas, seed, attempt, ll_shift)
        tidx.copy_(torch.from_numpy(t)), beta.copy_(torch.from_numpy(b)), eps.copy_(torch.from_numpy(e))
        if accept is not None:
            accept.copy_(torch.from_numpy(a))
        if pair_counts is not None and pc.size:
            pair_counts += torch.from_numpy(pc.astype(np.int64)).reshape(pair_counts.shape)
        if temp_stats is not None:
            temp_stats += torch.from_numpy(ts).reshape(temp_stats.shape)

    def select(self, q, aux, tidx, k_sel, n_columns, out_q, out_aux):
        import torch
        oq, oa = select(q.numpy(), None if aux is None else aux.numpy(), tidx.numpy(), k_sel, n_columns)
        out_q.copy_(torch.from_numpy(oq))
        if out_aux is not None:
            out_aux.copy_(torch.from_numpy(oa))
