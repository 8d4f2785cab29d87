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


def test_argument_validation_needs_no_gpu():
    import ctypes as C
    h = _cabi.lib()
    out = C.c_void_p()
    xs = np.zeros(4)
    # more coefficients than the kernels are instantiated for -> EUNSUPPORTED, before any CUDA call
    rc = h.binfb_model_create_polynomial(_cabi.ptr(xs), _cabi.ptr(xs), 4, 9, None, None, 1.0, 1.0, 0, 0, C.byref(out))
    assert rc == _cabi.EUNSUPPORTED and b"n_coeff" in h.binfb_last_error()
    rc = h.binfb_model_create_polynomial(None, _cabi.ptr(xs), 4, 4, None, None, 1.0, 1.0, 0, 0, C.byref(out))
    assert rc == _cabi.EINVAL
    rc = h.binfb_model_create_chromatin(1, _cabi.ptr(xs.astype(np.float32)), 2.0, 2.5, 1.0, 1.0, 0.0, 1.0, 1.0, 0, 0, C.byref(out))
    assert rc == _cabi.EINVAL and b"n_beads" in h.binfb_last_error()
    assert h.binfb_hmc_run(None, None, None, None, None, 1, None, *([None] * 11)) == _cabi.EINVAL
    assert h.binfb_chromatin_stream_layout(1000, None, 3, 0, None, 0, None, None) == _cabi.EINVAL   # roles not a power of 2
    # a structure that cannot fit one chain in shared memory is refused by the plan (chains per CTA = 0)
    plan = (C.c_int * 8)()
    nf = C.c_longlong()
    _cabi.check(h.binfb_chromatin_stream_layout(12000, None, 0, 0, None, 0, C.byref(nf), plan))
    assert plan[5] == 0


def test_launch_plan_heuristic():
    y = np.zeros(1, dtype=np.float32)
    for n, roles, chains in [(64, 1, 16), (500, 1, 16), (1000, 2, 8), (2000, 4, 4), (5000, 16, 1)]:
        h = _cabi.lib()
        import ctypes as C
        plan = (C.c_int * 8)()
        nf = C.c_longlong()
        _cabi.check(h.binfb_chromatin_stream_layout(n, None, 0, 0, None, 0, C.byref(nf), plan))
        assert (plan[3], plan[5]) == (roles, chains), (n, list(plan))
        # roles never overlap on a partner quad: Lr - ring drift > 31
        if plan[3] > 1:
            assert plan[4] - (plan[7] * plan[6] // plan[3] - 1) > 31


@pytest.mark.parametrize("n,roles", [(2, 1), (3, 1), (4, 1), (5, 1), (8, 1), (9, 1), (24, 1),
                                     (37, 1), (130, 1), (257, 1), (1000, 1), (1000, 2), (700, 2),
                                     (1001, 0), (1400, 4)])
def test_contact_stream_covers_every_pair_once(n, roles):
    m = n * (n - 1) // 2
    y = (np.arange(m, dtype=np.float64) % 16777213 + 1.0).astype(np.float32)  # non-zero tags
    stream, plan = _cabi.chromatin_stream_layout(n, y, roles)
    q, ks, nrb = plan["quads"], plan["partner_steps"], plan["row_blocks"]
    R, lr = plan["roles"], plan["slots_per_row_block"]
    assert q == (n + 3) // 4 and ks == q // 2 and nrb == (q + 31) // 32
    ss = plan["stage_steps"]
    spr = ss // R
    assert lr == -(-(-(-(ks + 1) // R)) // spr) * spr and (roles == 0 or R == roles)
    assert stream.size % (ss * 128 * 4) == 0          # whole bulk-copy stages
    s = stream.reshape(-1, 4, 32, 4)                   # [step][row r][lane][col c]
    iu = np.triu_indices(n, 1)
    tag = np.zeros((n, n), dtype=np.float32)
    tag[iu] = y
    seen = np.zeros((n, n), dtype=np.int64)
    used = np.zeros(s.shape[0], dtype=bool)
    # replay the kernel's schedule: role r of row block rb, slot sl works on partner offset
    # k = r*Lr + sl; lane l owns quad a = 32 rb + l, partner quad (a + k) % q
    for rb in range(nrb):
        for sl in range(lr):
            for role in range(R):
                k = role * lr + sl
                t = (rb * lr + sl) * R + role
                used[t] = True
                step = s[t]
                if k > ks:
                    assert not step.any()
                    continue
                lanes = np.arange(32)
                a = rb * 32 + lanes
                ok = a < q
                assert not step[:, ~ok, :].any()
                b = (a + k) % q
                for r in range(4):
                    for c in range(4):
                        v = step[r, :, c]
                        nz = (v != 0) & ok
                        i, j = 4 * a[nz] + r, 4 * b[nz] + c
                        assert (i < n).all() and (j < n).all() and (i != j).all()
                        lo, hi = np.minimum(i, j), np.maximum(i, j)
                        assert (tag[lo, hi] == v[nz]).all()
                        np.add.at(seen, (lo, hi), 1)
    assert not s[~used].any()
    assert (seen[iu] == 1).all() and seen.sum() == m


def test_sink_argument_validation_needs_no_gpu():
    import ctypes as C
    h = _cabi.lib()
    out = C.c_void_p()
    for args in [(0, 4, 1, 0, 1), (4, 0, 1, 0, 1), (4, 4, -1, 0, 1), (4, 4, 1, -1, 1), (4, 4, 1, 0, 0)]:
        assert h.binfb_sink_create(*args, 0, 0, C.byref(out)) == _cabi.EINVAL
    assert h.binfb_sink_push(None, None, None, None, None) == _cabi.EINVAL
    assert b"null sink" in h.binfb_last_error()
    assert h.binfb_sink_destroy(None) == 0


def test_list_sink_oracle_matches_reference_slicing():
    """the oracle's bookkeeping is literally the reference driver's list + slice + argmax"""
    import sink_port
    rng = np.random.RandomState(0)
    ref = sink_port.ListSink(burn_in=5, thin=3)
    qs, ls = [], []
    for t in range(20):
        q, l = rng.normal(size=(2, 3)), rng.normal(size=2)
        qs.append(q), ls.append(l)
        ref.append(q, None, l)
    thin = qs[5::3]                                       # example_script.py:41
    np.testing.assert_array_equal(np.array(ref.thinned()[0]), np.array(thin))
    lp = np.array(ls[5::3])
    for c in range(2):                                    # misc.py:18-22 per chain
        np.testing.assert_array_equal(ref.get_MAP()[0][c], thin[int(np.argmax(lp[:, c]))][c])


DECAY_CODE = """
__device__ float binfb_mock(const float *theta, const float *x, float *dmock) {
    const float e = expf(-theta[1] * x[0]);
    dmock[0] = e; dmock[1] = -theta[0] * x[0] * e; dmock[2] = 1.0f;
    return theta[0] * e + theta[2];
}
"""


def test_generic_model_device_code_compiles_without_gpu():
    """NVRTC runs on the host: user code + the generic fused kernels compile for sm_100a here"""
    ok, log = _cabi.generic_compile_check(DECAY_CODE, 3, 1)
    assert ok, log
    ok, log = _cabi.generic_compile_check("__device__ float binfb_mock(const float *t, const float *x, float *d)"
                                          " { return nope(t[0]); }", 2, 1)
    assert not ok and "user_model.cu(1)" in log and "nope" in log      # the log points into the user's code
    h = _cabi.lib()
    assert h.binfb_generic_compile_check(None, 3, 1, None, 0) == _cabi.EINVAL
    assert h.binfb_generic_compile_check(DECAY_CODE.encode(), 17, 1, None, 0) == _cabi.EINVAL


def test_device_forward_model_descriptor():
    from binf_b200.model.forwardmodels import DeviceForwardModel
    xs = np.linspace(0, 1, 8)
    m = DeviceForwardModel("decay", xs, "rates", 3, DECAY_CODE)
    assert "rates" in m.variables and m.check_device_code() == ""
    twin = m.clone()
    assert twin.device_code == m.device_code and twin.n_params == 3
    with pytest.raises(ValueError):
        DeviceForwardModel("bad", xs, "rates", 3, "int x;")
    with pytest.raises(ValueError):
        DeviceForwardModel("bad", xs, "rates", 3, "__device__ float binfb_mock(const float*a,const float*b,float*c){return q;}"
                           ).check_device_code()


def test_lockstep_plans_are_reported_and_unsafe_role_counts_refused():
    import ctypes as C
    h = _cabi.lib()
    nf, plan = C.c_longlong(), (C.c_int * 8)()
    # n = 1000, 4 roles: role ranges of 32 slots -> only safe in lockstep; the layout pass still plans it
    _cabi.check(h.binfb_chromatin_stream_layout(1000, None, 4, 0, None, 0, C.byref(nf), plan))
    assert list(plan)[3] == 4 and list(plan)[4] == 32 and list(plan)[5] == 4
    # n = 1000, 8 roles: ranges of 16 slots can collide even in lockstep -> no chains per CTA
    _cabi.check(h.binfb_chromatin_stream_layout(1000, None, 8, 0, None, 0, C.byref(nf), plan))
    assert list(plan)[5] == 0


@pytest.mark.parametrize("n,roles", [(1000, 2), (1000, 4), (500, 2), (700, 2), (2000, 4), (5000, 8), (5000, 16), (1400, 4)])
def test_roles_never_touch_the_same_partner_quad(n, roles):
    """Host replay of the race-freedom argument of the pair sweep (compute-sanitizer's racecheck is not
    available on the GPU pool).  Within a row block, role r at slot s read-modify-writes the partner
    quads {(a + r*Lr + s) mod Q : a in the row block}.  Free-running roles may be up to
    NS*SPR - 1 slots apart, LOCKSTEP roles are always at the same slot: no two roles may ever address
    the same quad within that window."""
    import ctypes as C
    h = _cabi.lib()
    nf, plan = C.c_longlong(), (C.c_int * 8)()
    _cabi.check(h.binfb_chromatin_stream_layout(n, None, roles, 0, None, 0, C.byref(nf), plan))
    q, ks, nrb, R, lr, W, ss, ns = list(plan)
    assert R == roles and W >= 1
    spr = ss // R
    free_ok = lr - (ns * spr - 1) >= 34
    drift = ns * spr - 1 if free_ok else 0             # lockstep plans: the roles move together
    assert free_ok or lr >= 32
    for rb in (0, nrb - 1):
        a = rb * 32 + np.arange(32)
        a = a[a < q]
        for s in range(0, lr, max(1, lr // 7)):
            touched = {}
            for role in range(R):
                for ds in range(-drift, drift + 1):     # every slot the other roles can be at meanwhile
                    sl = s + (ds if role else 0)
                    if not 0 <= sl < lr:
                        continue
                    k = role * lr + sl
                    if k > ks:
                        continue
                    for b in (a + k) % q:
                        owner = touched.setdefault(int(b), role)
                        assert owner == role, "roles %d and %d meet on quad %d (n=%d)" % (owner, role, b, n)
