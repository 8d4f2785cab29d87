"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, and the
host-only layout pass of the chromatin contact stream covers every bead pair exactly once."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from binf_b200 import _cabi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "binf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(binfb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    handle = _cabi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), "missing export: " + name
    assert set(declared) == set(_cabi.SIGNATURES), set(declared) ^ set(_cabi.SIGNATURES)
    assert handle.binfb_version() == 100


def test_error_reporting_without_gpu():
    handle = _cabi.lib()
    assert handle.binfb_model_destroy(None) == 0
    rc = handle.binfb_model_info(None, None, None, None, None)
    assert rc == _cabi.EINVAL and b"null model" in handle.binfb_last_error()


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 9, 24, 37, 130, 257, 1000])
def test_contact_stream_covers_every_pair_once(n):
    m = n * (n - 1) // 2
    y = (np.arange(m, dtype=np.float64) + 1.0).astype(np.float32)  # unique non-zero tags
    stream, q, t = _cabi.chromatin_stream_layout(n, y)
    n_pad = (n + 3) // 4 * 4
    assert q == n_pad // 4
    ks, nrb = q // 2, (q + 31) // 32
    assert t == nrb * (ks + 1)
    s = stream.reshape(-1, 4, 32, 4)[:t]            # [step][row r][lane][col c]
    assert stream.size % (4 * 128 * 4) == 0          # padded to whole bulk-copy stages
    assert not stream.reshape(-1, 4, 32, 4)[t:].any()
    iu = np.triu_indices(n, 1)
    tag = np.zeros((n, n), dtype=np.float32)
    tag[iu] = y
    seen = np.zeros((n, n), dtype=np.int64)
    # replay the kernel's schedule: lane l of row block rb owns quad a, partner quad (a+k)%q
    for rb in range(nrb):
        for k in range(ks + 1):
            step = s[rb * (ks + 1) + k]
            for lane in range(32):
                a = rb * 32 + lane
                if a >= q:
                    assert not step[:, lane, :].any()
                    continue
                b = (a + k) % q
                for r in range(4):
                    for c in range(4):
                        v = step[r, lane, c]
                        if v == 0:
                            continue
                        i, j = 4 * a + r, 4 * b + c
                        assert i < n and j < n and i != j
                        lo, hi = min(i, j), max(i, j)
                        assert tag[lo, hi] == v
                        seen[lo, hi] += 1
    assert (seen[iu] == 1).all() and seen.sum() == m
