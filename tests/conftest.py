import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the shared library is a build artefact (git-ignored): compile it when the checkout is fresh
    # (nvcc cross-compiles sm_100a without a GPU; incremental builds are no-ops)
    from binf_b200 import _cabi, build
    if not os.path.exists(_cabi.LIB_PATH):
        build.build()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def _gpu_available():
    try:
        from binf_b200 import _cabi
        return _cabi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """Skip (never silently fall back) when no CUDA device is visible."""
    if not _gpu_available():
        pytest.skip("no CUDA device visible")
    from binf_b200 import _cabi
    props = _cabi.device_props(0)
    assert props["cc"] >= 100, "binf_b200 targets sm_100a only, found sm_%d" % props["cc"]
    return props
