"""SURVEY.md section 7, step 2's gate: the reference's OWN unit tests (binf/tests/pdf/__init__.py,
binf/tests/pdf/likelihoods.py, binf/tests/samplers/gibbs.py -- 18 cases) run unchanged against the mirror
package, the way a user of the reference would switch over:

    import binf_b200; binf_b200.install_as_binf()        # `binf...` and the `csb...` names the tests import

The test sources are read from the reference tree and compiled in memory (nothing is copied); the run happens
in a subprocess so that the aliased module names do not leak into the other tests (oracle/ref_import.py
registers the REAL reference under the same names).  Skipped where the reference tree does not exist (the GPU
box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BINF_REFERENCE_ROOT", "/root/reference")

RUNNER = r'''
import os, sys, types, unittest
sys.path.insert(0, %(root)r)
import binf_b200
binf_b200.install_as_binf(alias_csb=True)
import csb
assert getattr(csb, "__binf_b200_shim__", False), "a real CSB shadows the shim"
ref_tests = os.path.join(%(ref)r, "binf", "tests")


def load(modname, relpath, is_pkg=False):
    mod = types.ModuleType(modname)
    path = os.path.join(ref_tests, relpath)
    mod.__file__ = path
    if is_pkg:
        mod.__path__ = []
    sys.modules[modname] = mod
    with open(path) as fh:
        exec(compile(fh.read(), path, "exec"), mod.__dict__)
    return mod


pkg = types.ModuleType("binf.tests"); pkg.__path__ = []; sys.modules["binf.tests"] = pkg
mods = [load("binf.tests.pdf", os.path.join("pdf", "__init__.py"), True),
        load("binf.tests.pdf.likelihoods", os.path.join("pdf", "likelihoods.py"))]
sp = types.ModuleType("binf.tests.samplers"); sp.__path__ = []; sys.modules["binf.tests.samplers"] = sp
mods.append(load("binf.tests.samplers.gibbs", os.path.join("samplers", "gibbs.py")))
suite = unittest.TestSuite()
for m in mods:
    suite.addTests(unittest.defaultTestLoader.loadTestsFromModule(m))
res = unittest.TextTestRunner(verbosity=1, stream=sys.stderr).run(suite)
print("RAN %%d FAILURES %%d ERRORS %%d" %% (res.testsRun, len(res.failures), len(res.errors)))
sys.exit(0 if res.wasSuccessful() else 1)
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "binf", "tests", "pdf", "__init__.py")),
                    reason="reference tree not present")
def test_reference_unit_tests_pass_against_the_mirror():
    r = subprocess.run([sys.executable, "-c", RUNNER % dict(root=ROOT, ref=REF)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "RAN 18 FAILURES 0 ERRORS 0" in r.stdout, r.stdout


def test_csb_alias_hands_out_the_mirrors_parameter_classes():
    """reference-style imports resolve after install_as_binf(); without the alias a foreign, duck-typed
    parameter object is adopted (wrapped) instead of being rejected"""
    code = r'''
import sys
sys.path.insert(0, %(root)r)
import binf_b200
binf_b200.install_as_binf()
from csb.statistics.pdf.parameterized import Parameter, AbstractParameter, ParameterizedDensity
from csb.statistics.samplers import State
from csb.statistics.samplers.mc.singlechain import AbstractSingleChainMC
from csb.numeric import exp, log, log_sum_exp
from csb.core import OrderedDict
import binf, binf.pdf, binf.samplers.hmc
from binf_b200 import params
assert Parameter is params.Parameter and AbstractParameter is params.AbstractParameter
assert binf.ArrayParameter is binf_b200.ArrayParameter
assert exp(1e4) == exp(709.0) and log(0.0) == log(1e-308)
from binf.pdf import AbstractBinfPDF


class Foreign(object):               # what a parameter of a real CSB install looks like from outside
    def __init__(self, value, name):
        self._v, self.name = value, name
    value = property(lambda self: self._v)
    def set(self, v):
        self._v = v


class Pdf(AbstractBinfPDF):
    def __init__(self):
        super(Pdf, self).__init__("p")
        self._register("A")
        self["A"] = Foreign(2.0, "A")
        self._register_variable("x")
    def _evaluate_log_prob(self, x):
        return -0.5 * self["A"].value * x * x


p = Pdf()
assert p.log_prob(x=3.0) == -9.0
follower = Parameter(0.0, "A")
follower.bind_to(p["A"])
p["A"].set(4.0)
assert follower.value == 4.0 and p.log_prob(x=1.0) == -2.0
try:
    p["A"] = 3.0
    raise SystemExit("a bare float was accepted as a parameter object")
except TypeError:
    pass
print("ALIAS OK")
'''
    r = subprocess.run([sys.executable, "-c", code % dict(root=ROOT)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALIAS OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
