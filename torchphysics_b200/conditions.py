"""Conditions with the reference's interface (``problem/conditions/condition.py``):
``Condition`` :30-77, ``SingleModuleCondition`` :215-309, ``MeanCondition`` :312-373,
``DeepRitzCondition`` :376-431, ``PINNCondition`` :434-495, ``SquaredError`` :11-27.

``forward(device, iteration)`` is the hot loop body: sample -> track coordinate leaves ->
data functions -> model -> user residual (bound by argument NAME) -> error -> reduce.  The
model call and every differential operator inside the residual resolve to the sm_100a jet
kernels (``models.py`` / ``operators.py``); what is left to torch is the user's pointwise
residual arithmetic and the scalar reduction.
"""
import torch

from . import jets
from .models import Parameter
from .samplers import StaticSampler
from .spaces import Points
from .userfn import UserFunction


class SquaredError(torch.nn.Module):
    """sum of squares over every axis but the first (reference ``condition.py:11-27``)."""

    def forward(self, x):
        return torch.sum(torch.square(x), dim=list(range(1, len(x.shape))))


class Condition(torch.nn.Module):
    def __init__(self, name=None, weight=1.0, track_gradients=True):
        super().__init__()
        self.name = name
        self.weight = weight
        self.track_gradients = track_gradients

    def forward(self, device="cpu", iteration=None):
        raise NotImplementedError

    def _setup_data_functions(self, data_functions, sampler):
        data_functions = dict(data_functions)
        for name in data_functions:
            data_functions[name] = UserFunction(data_functions[name])
        if isinstance(sampler, StaticSampler):
            # evaluated once on the stored points (reference condition.py:64-74)
            for name in data_functions:
                points = sampler.sample_points()
                data_functions[name] = UserFunction(data_functions[name](points))
        return data_functions

    def _move_static_data(self, device):
        pass


def _empty_parameter():
    return Points.empty()


class SingleModuleCondition(Condition):
    def __init__(self, module, sampler, residual_fn, error_fn, reduce_fn=torch.mean,
                 name="singlemodulecondition", track_gradients=True, data_functions={},
                 parameter=None, weight=1.0):
        super().__init__(name=name, weight=weight, track_gradients=track_gradients)
        self.module = module
        self.parameter = _empty_parameter() if parameter is None else parameter
        if isinstance(self.parameter, Parameter):
            self.register_parameter(name + "_params", self.parameter.as_tensor)
        self.sampler = sampler
        self.residual_fn = UserFunction(residual_fn)
        self.error_fn = error_fn
        self.reduce_fn = reduce_fn
        self.data_functions = self._setup_data_functions(data_functions, sampler)
        if self.sampler.is_adaptive:
            self.last_unreduced_loss = None

    def forward(self, device="cpu", iteration=None):
        if self.sampler.is_adaptive:
            x = self.sampler.sample_points(unreduced_loss=self.last_unreduced_loss, device=device)
            self.last_unreduced_loss = None
        else:
            x = self.sampler.sample_points(device=device)
        x_coordinates, x = x.track_coord_gradients()
        data = {}
        for name in self.data_functions:
            data[name] = self.data_functions[name](x_coordinates)
        with jets.scope(id(self)):
            y = self.module(x)
            unreduced_loss = self.error_fn(self.residual_fn(
                {**y.coordinates, **x_coordinates, **self.parameter.coordinates, **data}))
        if self.sampler.is_adaptive:
            self.last_unreduced_loss = unreduced_loss
        return self.reduce_fn(unreduced_loss)

    def _move_static_data(self, device):
        if self.sampler.is_static:
            for name in self.data_functions:
                fun = self.data_functions[name].fun
                if isinstance(fun, torch.Tensor):
                    self.data_functions[name].fun = fun.to(device)


class MeanCondition(SingleModuleCondition):
    def __init__(self, module, sampler, residual_fn, track_gradients=True, data_functions={},
                 parameter=None, name="meancondition", weight=1.0):
        super().__init__(module, sampler, residual_fn, error_fn=torch.nn.Identity(), reduce_fn=torch.mean,
                         name=name, track_gradients=track_gradients, data_functions=data_functions,
                         parameter=parameter, weight=weight)


class DeepRitzCondition(MeanCondition):
    def __init__(self, module, sampler, integrand_fn, track_gradients=True, data_functions={},
                 parameter=None, name="deepritzcondition", weight=1.0):
        super().__init__(module, sampler, integrand_fn, track_gradients=track_gradients,
                         data_functions=data_functions, parameter=parameter, name=name, weight=weight)


class PINNCondition(SingleModuleCondition):
    def __init__(self, module, sampler, residual_fn, track_gradients=True, data_functions={},
                 parameter=None, name="pinncondition", weight=1.0):
        super().__init__(module, sampler, residual_fn, error_fn=SquaredError(), reduce_fn=torch.mean,
                         name=name, track_gradients=track_gradients, data_functions=data_functions,
                         parameter=parameter, weight=weight)
