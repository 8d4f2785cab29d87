"""Posterior = sum of likelihood and prior components (reference: binf/pdf/posteriors.py:15-211)."""
import numpy as np

from binf_b200.pdf import AbstractBinfPDF


class Posterior(AbstractBinfPDF):
    def __init__(self, likelihoods, priors, name="the one and only posterior"):
        super(Posterior, self).__init__(name)
        self._likelihoods = likelihoods
        self._priors = priors
        self.beta = None  # inverse temperature of the likelihood terms (replica exchange); None = 1
        self._setup_parameters()
        self._components = dict(**self.priors)
        self._components.update(**self.likelihoods)
        self._register_component_variables(*self._get_component_variables())
        self._set_original_variables()

    @property
    def likelihoods(self):
        return self._likelihoods

    @property
    def priors(self):
        return self._priors

    # -- wiring ---------------------------------------------------------------------------------
    def _setup_parameters(self):
        """one parameter per distinct component parameter name; all components follow it
        (posteriors.py:44-55)"""
        for group in (self.likelihoods, self.priors):
            for comp in group.values():
                for p in comp.parameters:
                    if p not in self.parameters:
                        self._register(p)
                        self[p] = type(comp[p])(comp[p].value, comp[p].name)
                    comp[p].bind_to(self[p])

    def _get_component_variables(self):
        names, fixed, diff, types = [], [], [], []
        for comp in self._components.values():
            for v in comp.variables:
                names.append(v)
                types.append(comp.var_param_types[v])
                if v in comp.differentiable_variables:
                    diff.append(v)
            fixed.extend(p for p in comp.parameters if p in comp._original_variables)
        return names, set(fixed), set(diff), types

    def _register_component_variables(self, names, fixed_vars, diff_vars, var_param_types):
        for var in set(names):
            self._register_variable(str(var), differentiable=var in diff_vars)
        self._original_variables.update(fixed_vars)
        self.update_var_param_types(**dict(zip(names, var_param_types)))

    def _get_component_variables_list(self):
        return {c: c.variables for c in self._components.values()}

    # -- evaluation -----------------------------------------------------------------------------
    def _lowered(self, variables):
        from binf_b200.lowering import lower
        v = variables.get("coefficients")
        return lower(self, n_coeff=None if v is None else int(np.shape(v)[-1]))

    def _evaluate_components(self, **model_parameters):
        out = []
        for comp, comp_vars in self._get_component_variables_list().items():
            lp = comp.log_prob(**{v: model_parameters[v] for v in comp_vars})
            if self.beta is not None and comp in self.likelihoods.values():
                lp = self.beta * lp
            out.append(lp)
        return out

    def _evaluate_log_prob(self, **model_parameters):
        low = self._lowered(model_parameters)
        if low is not None:
            return low.log_prob(model_parameters[low.variable], model_parameters)
        return np.sum(self._evaluate_components(**model_parameters), axis=0)

    def _evaluate_gradient(self, **variables):
        low = self._lowered(variables)
        if low is not None:
            return low.gradient(variables[low.variable], variables)
        size = sum(len(variables[v]) if hasattr(variables[v], "__len__") else 1
                   for v in variables if v in self.differentiable_variables)
        res = np.zeros(size)
        for comp in self._components.values():
            # only components that declare a differentiable variable contribute
            # (posteriors.py:182-185; this is what drops the Gaussian prior, quirk Q1)
            if len(comp.variables) > 0 and len(comp.differentiable_variables) > 0:
                g = comp.gradient(**{x: variables[x] for x in variables if x in comp.variables})
                if self.beta is not None and comp in self.likelihoods.values():
                    g = self.beta * g
                res = res + g
        return res

    def clone(self):
        copy = self.__class__({k: v.clone() for k, v in self.likelihoods.items()},
                              {k: v.clone() for k, v in self.priors.items()}, self.name)
        copy.beta = self.beta
        copy.set_fixed_variables_from_pdf(self)
        return copy

    def conditional_factory(self, **fixed_vars):
        copy = self.__class__({k: v.conditional_factory(**fixed_vars) for k, v in self.likelihoods.items()},
                              {k: v.conditional_factory(**fixed_vars) for k, v in self.priors.items()},
                              self.name)
        copy.beta = self.beta
        return copy
