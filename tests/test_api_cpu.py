"""API compatibility of the Python layer with the reference: the scenarios of the reference's own
unit tests (binf/tests/pdf/__init__.py:44-105, binf/tests/pdf/likelihoods.py:79-119,
binf/tests/samplers/gibbs.py:41-133) with their known answers, run against binf_b200's classes.
They exercise the generic plumbing with user-defined (mock) components -- no GPU involved."""
import numpy as np
import pytest

import binf_b200
from binf_b200 import ArrayParameter
from binf_b200.params import Parameter
from binf_b200.pdf import AbstractBinfPDF
from binf_b200.pdf.likelihoods import Likelihood
from binf_b200.model.errormodels import AbstractErrorModel
from binf_b200.model.forwardmodels import AbstractForwardModel
from binf_b200.samplers import BinfState
from binf_b200.samplers.gibbs import GibbsSampler
from binf_b200.samplers.hmc import HMCSampler


class QuadraticPDF(AbstractBinfPDF):
    """log p = -1/2 A (x^2 + y^2), A = 2"""

    def __init__(self, name="QuadraticPDF"):
        super().__init__(name=name)
        self._register("ParamA")
        self["ParamA"] = Parameter(2.0, "ParamA")
        self._register_variable("x")
        self._register_variable("y")
        self.update_var_param_types(x=Parameter, y=Parameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, x, y):
        return -0.5 * self["ParamA"].value * (x ** 2 + y ** 2)

    def clone(self):
        copy = self.__class__()
        copy.set_fixed_variables_from_pdf(self)
        return copy


# ---- AbstractBinfPDF --------------------------------------------------------------------------
def test_fix_variables():
    pdf = QuadraticPDF()
    pdf.fix_variables(y=5.0)
    assert pdf.variables == {"x"} and pdf["y"].value == 5.0
    with pytest.raises(ValueError):
        pdf.fix_variables(z=2.0)


def test_conditional_factory_twice():
    cond = QuadraticPDF().conditional_factory(x=5.0)
    assert "x" in cond.parameters and cond["x"].value == 5.0 and cond.variables == {"y"}
    assert cond.log_prob(y=2.0) == -29.0
    cond2 = cond.conditional_factory(y=2.0)
    assert cond2["y"].value == 2.0 and len(cond2.variables) == 0 and cond2.log_prob() == -29.0


def test_set_fixed_variables_from_pdf():
    a, b = QuadraticPDF(), QuadraticPDF()
    a.fix_variables(y=2.0)
    b.set_fixed_variables_from_pdf(a)
    assert "y" in b.parameters and b["y"].value == 2.0


def test_log_prob_and_types_and_completion():
    pdf = QuadraticPDF()
    assert pdf.log_prob(x=3, y=2) == -13.0

    class Other(Parameter):
        pass
    pdf.update_var_param_types(x=Other)
    assert pdf.var_param_types["x"] is Other
    pdf = QuadraticPDF()
    pdf.fix_variables(x=7.0)
    variables = {"y": 2.34}
    pdf._complete_variables(variables)
    assert variables == {"y": 2.34, "x": 7.0}


def test_error_conventions():
    pdf = QuadraticPDF()
    with pytest.raises(ValueError):
        pdf._register_variable("x")            # duplicate (binf/__init__.py:54-57)
    with pytest.raises(ValueError):
        pdf._delete_variable("nope")           # unknown (binf/__init__.py:75)
    with pytest.raises(AttributeError):
        pdf["nope"]                            # ParameterNotFoundError(AttributeError)
    with pytest.raises(NotImplementedError):
        pdf.gradient(x=1.0, y=2.0)             # missing gradient (binf/__init__.py:158)
    with pytest.raises(ValueError):
        pdf(x=1.0)                             # wrong arity (binf/__init__.py:112-116)


# ---- Likelihood ---------------------------------------------------------------------------------
class SquareErrorModel(AbstractErrorModel):
    def __init__(self):
        super().__init__("StupidErrorModel")
        self._register("ParamB")
        self["ParamB"] = Parameter(4.0, "ParamB")
        self._register_variable("mock_data", differentiable=True)
        self._register_variable("a")
        self.update_var_param_types(mock_data=ArrayParameter, a=Parameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, mock_data, a):
        return a * np.sum(mock_data ** 2)

    def _evaluate_gradient(self, mock_data, a):
        return a * 2.0 * mock_data

    def clone(self):
        copy = self.__class__()
        copy.set_fixed_variables_from_pdf(self)
        return copy


class FixedForwardModel(AbstractForwardModel):
    def __init__(self, parameters=()):
        super().__init__("testfwm", parameters)
        self._register_variable("X")
        self._register_variable("b")
        self.update_var_param_types(X=ArrayParameter, b=Parameter)
        self._set_original_variables()

    def _evaluate(self, X, b):
        return b * np.array([1.0, 2.0, 3.0])

    def _evaluate_jacobi_matrix(self, X, b):
        return b * np.array([[2.0, 1.0, 1.0], [1.0, 2.0, 2.0]])

    def clone(self):
        pass


def _likelihood():
    return Likelihood("testL", FixedForwardModel(parameters=[Parameter(2.0, "ParamA")]), SquareErrorModel())


def test_likelihood_parameter_binding_both_ways():
    L = _likelihood()
    assert L["ParamA"].value == 2.0 and L["ParamB"].value == 4.0
    L["ParamA"].set(3.0)
    assert L.forward_model["ParamA"].value == 3.0
    L["ParamB"].set(7.0)
    assert L.error_model["ParamB"].value == 7.0


def test_likelihood_split_variables():
    fwm, em = _likelihood()._split_variables({"X": np.array([1.0, 2.0]), "a": 5.0, "b": 2.0})
    assert set(fwm) == {"X", "b"} and set(em) == {"a"}


def test_likelihood_known_answers():
    L = _likelihood()
    assert L.log_prob(X=np.array([1.2, 4.2, 54.5]), a=2.0, b=3.0) == 252.0
    a, b = 2.0, 3.0
    assert np.all(L.gradient(X=np.array([1.2, 4.2]), a=a, b=b) == np.array([14 * a * b ** 2, 22 * a * b ** 2]))


# ---- GibbsSampler ---------------------------------------------------------------------------------
class DoublingSampler(object):
    def __init__(self, variable_name):
        self.pdf, self.state, self.variable_name = None, 5.0, variable_name

    @property
    def last_draw_stats(self):
        return {self.variable_name: {"testlastdrawstats{}".format(self.state): self.state}}

    @property
    def sampling_stats(self):
        return {"testsamplingstats{}".format(self.state): self.state}

    def sample(self):
        other = "y" if "y" in self.pdf.parameters else "x"
        return self.state * 2.0 * self.pdf[other].value


def _gibbs():
    return GibbsSampler(QuadraticPDF(), BinfState({"x": 2.0, "y": 3.0}),
                        {"x": DoublingSampler("x"), "y": DoublingSampler("y")})


def test_gibbs_conditional_pdfs():
    g = _gibbs()
    assert set(g._conditional_pdfs) == {"x", "y"}
    assert g._conditional_pdfs["x"]["y"].value == 3.0 and g._conditional_pdfs["y"]["x"].value == 2.0
    assert g.subsamplers["x"].pdf["y"].value == 3.0 and g.subsamplers["y"].pdf["x"].value == 2.0
    assert len(g._conditional_pdfs["x"].variables) == 1
    g.state.update_variables(x=5.0)
    g._update_conditional_pdf_params()
    assert g._conditional_pdfs["y"]["x"].value == 5.0


def test_gibbs_update_samplers_and_states():
    g = _gibbs()
    s = DoublingSampler("x")
    s.pdf = QuadraticPDF()
    s.pdf["ParamA"].set(23.0)
    g.update_samplers(x=s)
    assert g.subsamplers["x"].pdf["ParamA"].value == 23.0
    # the only place the reference constructs an HMCSampler (tests/samplers/gibbs.py:83-94)
    g = GibbsSampler(QuadraticPDF(), BinfState({"x": 2.0, "y": 3.0}),
                     {"x": DoublingSampler("x"), "y": HMCSampler(QuadraticPDF(), 1.0, 0.1, 12)})
    g.state.update_variables(x=5.0, y=2.3)
    g._update_subsampler_states()
    assert g.subsamplers["x"].state == 5.0 and g.subsamplers["y"].state == 2.3
    g._update_state(x=34.0)
    assert g.state.variables["x"] == 34.0


def test_gibbs_sweep_order_and_stats():
    g = _gibbs()
    g._update_state(x=0.5)
    s = g.sample()
    assert s.variables == g.state.variables
    assert s.variables["x"] == 3.0 and s.variables["y"] == 18.0     # x first, y sees the new x
    g = _gibbs()
    stats = g.last_draw_stats
    assert set(stats) == {"x", "y"} and "testlastdrawstats2.0" in stats["x"] and "testlastdrawstats3.0" in stats["y"]
    assert set(g.sampling_stats) == {"testsamplingstats2.0", "testsamplingstats3.0"}


def test_hmc_sampler_attributes_match_reference():
    s = HMCSampler(QuadraticPDF(), np.zeros(3), 0.1, 12, timestep_adaption_limit=5, variable_name="x", seed=1)
    assert (s.timestep, s.nsteps, s.timestep_adaption_limit) == (0.1, 12, 5)
    assert (s.adaption_uprate, s.adaption_downrate, s.n_accepted, s.counter) == (1.05, 0.95, 0, 0)
    assert s.acceptance_rate == 0.0 and s.variable_name == "x" and s.last_move_accepted == 0
    assert HMCSampler(QuadraticPDF(), np.zeros(3), 0.1, 12).variable_name == "HMC"
    assert s.last_draw_stats == {"x": (0, 0.1)}
    with pytest.raises(NotImplementedError):        # user pdfs are not lowered: no CPU fallback
        s.sample()


def test_parameters_follow_their_base():
    base, dep = Parameter(1.0, "p"), Parameter(5.0, "p")
    dep.bind_to(base)
    assert dep.value == 1.0
    base.set(3.5)
    assert dep.value == 3.5
    with pytest.raises(ValueError):
        dep.set(2.0)
    arr = ArrayParameter([1, 2, 3], "a")
    assert arr.value.shape == (3,)
    assert Parameter(np.array([1.0, 2.0]), "tau").value.shape == (2,)   # per-chain precision


def test_example_wiring_and_quirks():
    from binf_b200.example.misc import make_posterior
    xs = np.linspace(-2, 2, 20)
    post = make_posterior(xs, np.zeros(20), np.polynomial.polynomial.polyval)
    assert post.variables == {"coefficients", "precision"}
    assert post.differentiable_variables == {"coefficients"}
    cond = post.conditional_factory(precision=2.5)
    assert cond.variables == {"coefficients"} and cond["precision"].value == 2.5
    rate = lambda p: [x for x in p.priors.values() if hasattr(x, "rate")][0].rate
    assert rate(post) == 0.2 and rate(cond) == 1.0          # quirk Q2 (example/priors.py:29)
    cond["precision"].set(4.0)                              # reaches the error model lazily
    assert cond.likelihoods["points"].error_model["precision"].value == 4.0
    import sys
    binf = binf_b200.install_as_binf()
    try:
        from binf.pdf.posteriors import Posterior as P2
        assert P2 is type(post) and binf.ArrayParameter is ArrayParameter
    finally:
        for k in [k for k in sys.modules if k == "binf" or k.startswith("binf.")]:
            del sys.modules[k]
