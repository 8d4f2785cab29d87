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
    def _device(self, variables):
        """the fused device evaluation of this likelihood, or None for user-defined models"""
        from binf_b200.lowering import lower
        coeffs = variables.get("coefficients")
        return lower(self, n_coeff=None if coeffs is None else int(np.shape(coeffs)[-1]))

    def _host_pieces(self, variables):
        """generic plumbing for user-defined models: mock data and the error model's arguments"""
        fwm_args, em_args = self._split_variables(variables)
        return self.forward_model(**fwm_args), fwm_args, em_args

    def _evaluate_log_prob(self, **variables):
        dev = self._device(variables)
        if dev is not None:  # forward model + error model + reduction in one kernel
            return dev.log_prob(variables[dev.variable], variables)
        mock, _, em_args = self._host_pieces(variables)
        return self.error_model.log_prob(mock_data=mock, **em_args)

    def _evaluate_gradient(self, **variables):
        dev = self._device(variables)
        if dev is not None:  # J(theta) . dE/dmock without ever forming J (likelihoods.py:148-155)
            return dev.gradient(variables[dev.variable], variables)
        mock, fwm_args, em_args = self._host_pieces(variables)
        jacobian = self.forward_model.jacobi_matrix(**fwm_args)
        return jacobian.dot(self.error_model.gradient(mock_data=mock, **em_args))

    def clone(self):
        return type(self)(self.name, self.forward_model.clone(), self.error_model.clone())

    def conditional_factory(self, **fixed_vars):
        """freeze variables in copies of both components and wire a new likelihood around them
        (likelihoods.py:165-174)"""
        model = self.forward_model.clone()
        model.fix_variables(**model._get_variables_intersection(fixed_vars))
        errors = self.error_model
        errors = errors.conditional_factory(**errors._get_variables_intersection(fixed_vars))
        return type(self)(self.name, model, errors)
