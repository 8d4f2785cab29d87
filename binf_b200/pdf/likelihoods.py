"""Likelihood = forward model + error model (reference: binf/pdf/likelihoods.py:12-175)."""
import numpy as np

from binf_b200.pdf import AbstractBinfPDF


class Likelihood(AbstractBinfPDF):
    def __init__(self, name, forward_model, error_model):
        super(Likelihood, self).__init__(name)
        self._forward_model = forward_model
        self._error_model = error_model
        self._inherit_variables()
        self._setup_parameters()
        self._set_original_variables()

    # -- wiring --------------------------------------------------------------------------------
    def _inherit_from(self, component, skip=()):
        """take over a component's variables; the ones it already holds as fixed parameters only
        count as "original" variables (likelihoods.py:42-77)"""
        for v in component._original_variables:
            if v in skip:
                continue
            if v in component.parameters:
                self._original_variables.add(v)
            else:
                self._register_variable(v, differentiable=v in component.differentiable_variables)
            self.update_var_param_types(**{v: component.var_param_types[v]})

    def _inherit_fwm_variables(self):
        self._inherit_from(self._forward_model)

    def _inherit_em_variables(self):
        self._inherit_from(self._error_model, skip=("mock_data",))

    def _inherit_variables(self):
        self._inherit_fwm_variables()
        self._inherit_em_variables()

    def _setup_parameters(self):
        """own copy of every component parameter; the component's parameter follows it
        (likelihoods.py:79-88)"""
        for component in (self._forward_model, self._error_model):
            for p in component.get_params():
                self._register(p.name)
                self[p.name] = type(p)(p.value, p.name)
                p.bind_to(self[p.name])

    def _setup_fixed_variable_parameters(self):
        for model in (self._forward_model, self._error_model):
            for p in model.get_params():
                model[p.name] = model.var_param_types[p.name](self[p.name].value, p.name)
                model[p.name].bind_to(self[p.name])

    @property
    def forward_model(self):
        return self._forward_model

    @property
    def error_model(self):
        return self._error_model

    def _split_variables(self, variables):
        fwm = {v: variables[v] for v in variables if v in self.forward_model.variables}
        em = {v: variables[v] for v in variables if v in self.error_model.variables}
        return fwm, em

    # -- evaluation ------------------------------------------------------------------------------
    def _lowered(self, variables):
        from binf_b200.lowering import lower
        try:
            low = lower(self, n_coeff=self._n_free(variables))
        except Exception:
            raise
        return low

    def _n_free(self, variables):
        v = variables.get("coefficients")
        return None if v is None else int(np.shape(v)[-1])

    def _evaluate_log_prob(self, **variables):
        low = self._lowered(variables)
        if low is not None:  # fused on the device: forward model + error model + reduction
            return low.log_prob(variables[low.variable], variables)
        fwm_variables, em_variables = self._split_variables(variables)
        mock_data = self.forward_model(**fwm_variables)
        return self.error_model.log_prob(mock_data=mock_data, **em_variables)

    def _evaluate_gradient(self, **variables):
        low = self._lowered(variables)
        if low is not None:  # J(theta).dot(dE/dmock) without ever forming J
            return low.gradient(variables[low.variable], variables)
        fwm_variables, em_variables = self._split_variables(variables)
        mock_data = self.forward_model(**fwm_variables)
        jac = self.forward_model.jacobi_matrix(**fwm_variables)
        em_grad = self.error_model.gradient(mock_data=mock_data, **em_variables)
        return jac.dot(em_grad)

    def clone(self):
        return self.__class__(self.name, self.forward_model.clone(), self.error_model.clone())

    def conditional_factory(self, **fixed_vars):
        fwm = self.forward_model.clone()
        fwm.fix_variables(**fwm._get_variables_intersection(fixed_vars))
        em = self.error_model.conditional_factory(
            **self.error_model._get_variables_intersection(fixed_vars))
        return self.__class__(self.name, fwm, em)
