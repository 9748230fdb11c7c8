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


class _SquaredErrorMean(torch.autograd.Function):
    """mean over rows of the summed squares as ONE reduction kernel (``tpx_squared_error_mean``: warp-shuffle reduction, float64
    accumulation) with a one-kernel adjoint, instead of torch's square / sum / mean launches and their autograd nodes."""

    @staticmethod
    def forward(ctx, r):
        import ctypes
        from . import _lib
        r = r.contiguous()
        acc = torch.zeros(1, dtype=torch.float64, device=r.device)
        with _lib.on_device(r.device):
            _lib.check(_lib.lib().tpx_squared_error_mean(_lib.ptr(r), ctypes.c_int64(r.numel()), ctypes.c_int64(r.shape[0]),
                                                         _lib.ptr(acc), _lib.stream_ptr()))
        ctx.save_for_backward(r)
        return acc.to(torch.float32).reshape(())

    @staticmethod
    @torch.autograd.function.once_differentiable      # raw-pointer adjoint: a double backward raises instead of returning zeros
    def backward(ctx, g):
        import ctypes
        from . import _lib
        (r,) = ctx.saved_tensors
        g = g.to(device=r.device, dtype=torch.float32).contiguous()
        gr = torch.empty_like(r)
        with _lib.on_device(r.device):
            _lib.check(_lib.lib().tpx_squared_error_mean_backward(_lib.ptr(r), ctypes.c_int64(r.numel()), ctypes.c_int64(r.shape[0]),
                                                                  _lib.ptr(g), _lib.ptr(gr), _lib.stream_ptr()))
        return gr


class _MicroBatchedLoss(torch.autograd.Function):
    """loss of a condition with a mean reduction as the weighted sum over row slices of the batch, each slice taken through
    forward (saving its checkpoints) and reverse before the next one starts; the parameter gradients are accumulated here and
    handed to autograd in ``backward`` (scaled by the incoming adjoint)."""

    @staticmethod
    def forward(ctx, cond, x, chunks, *params):
        n = len(x)
        bounds = [round(i * n / chunks) for i in range(chunks + 1)]
        total, grads = None, [None] * len(params)
        with torch.enable_grad():
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                if hi == lo:
                    continue
                part = cond._loss_on(x[lo:hi]) * ((hi - lo) / n)
                gs = torch.autograd.grad(part, params, allow_unused=True)
                for i, g in enumerate(gs):
                    if g is not None:
                        grads[i] = g if grads[i] is None else grads[i] + g
                total = part.detach() if total is None else total + part.detach()
        ctx.grads = grads
        return total

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        return (None, None, None) + tuple(None if gr is None else gr * g for gr in ctx.grads)


def _fusable_residual(r):
    # (a 1-D residual is not a batch of rows for SquaredError: torch.sum(., dim=[]) reduces everything)
    return (isinstance(r, torch.Tensor) and r.is_cuda and r.dtype == torch.float32 and r.dim() >= 2 and r.numel() > 0
            and r.requires_grad)


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
        chunks = self._micro_batches(x)
        if chunks > 1:
            params = [p for p in self.parameters() if p.requires_grad]
            return _MicroBatchedLoss.apply(self, x, chunks, *params)
        return self._loss_on(x)

    def _micro_batches(self, x):
        """how many slices the batch is evaluated in.  A batch whose activation-jet checkpoints do not fit the HBM budget
        would otherwise run forward -> (recompute forward in chunks -> reverse): one forward sweep too many.  With a mean
        reduction the loss is the weighted sum of the slices' losses, so each slice runs forward (saving) -> reverse on its
        own and the gradients add up (``_MicroBatchedLoss``): two sweeps per point instead of three."""
        if self.sampler.is_adaptive or self.reduce_fn is not torch.mean or not torch.is_grad_enabled():
            return 1
        t = x.as_tensor
        eng_of = getattr(self.module, "_native_engine", None)
        if not t.is_cuda or t.dim() != 2 or eng_of is None:
            return 1
        spec = getattr(self.module, "_jet_hints", {}).get(id(self))
        eng = eng_of()
        if spec is None or eng is False:
            return 1
        # decided once per (batch size, jet, device): asking the driver for the free memory costs ~0.3 ms of host time per call
        # (cudaMemGetInfo), which a pipelined step should not pay every iteration
        cache = self.__dict__.setdefault("_micro_batch_cache", {})
        key = (t.shape[0], spec.key(), str(t.device))
        if key not in cache:
            from .engine import save_budget_bytes, within_save_budget
            need = eng.saved_bytes(spec, t.shape[0])
            if need == 0 or within_save_budget(need, t.device):
                cache[key] = 1
            else:
                budget = save_budget_bytes(t.device)
                cache[key] = min(int(-(-need // max(budget // 2, 1))), 64)
        return cache[key]

    def _loss_on(self, x):
        x_coordinates, x = x.track_coord_gradients()
        data = {}
        for name in self.data_functions:
            data[name] = self.data_functions[name](x_coordinates)
        with jets.scope(id(self)):
            y = self.module(x)
            residual = self.residual_fn({**y.coordinates, **x_coordinates, **self.parameter.coordinates, **data})
            # the PINN loss (SquaredError + mean, reference condition.py:11-27, 488-489) as one native reduction kernel;
            # adaptive samplers need the per-point loss and keep the generic path
            if (type(self.error_fn) is SquaredError and self.reduce_fn is torch.mean and not self.sampler.is_adaptive
                    and _fusable_residual(residual)):
                return _SquaredErrorMean.apply(residual)
            unreduced_loss = self.error_fn(residual)
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


class DeepONetSingleModuleCondition(Condition):
    """reference ``deeponet_condition.py:9-108``: trunk points are sampled once and repeated for
    every branch function, the branch functions are discretised on the branch grid, the
    residual sees model output, trunk coordinates, parameters, data functions and (when it
    names them) the branch functions evaluated at the trunk points."""

    def __init__(self, deeponet_model, branch_function_sampler, trunk_points_sampler, residual_fn, error_fn,
                 data_sampler=None, reduce_fn=torch.mean, name="singlemodulecondition", track_gradients=True,
                 data_functions={}, parameter=None, weight=1.0):
        super().__init__(name=name, weight=weight, track_gradients=track_gradients)
        self.net = deeponet_model
        self.branch_function_sampler = branch_function_sampler
        self.data_sampler = data_sampler
        self.parameter = _empty_parameter() if parameter is None else parameter
        if isinstance(self.parameter, Parameter):
            self.register_parameter(name + "_params", self.parameter.as_tensor)
        self.trunk_points_sampler = trunk_points_sampler
        self.residual_fn = UserFunction(residual_fn)
        self.error_fn = error_fn
        self.reduce_fn = reduce_fn
        self.data_functions = self._setup_data_functions(data_functions, self.trunk_points_sampler)
        fn_out = self.branch_function_sampler.function_set.function_space.output_space.variables
        self.eval_function_set_for_residual = (len(fn_out & set(self.residual_fn.args)) > 0
                                               or self.data_sampler is not None)
        if self.trunk_points_sampler.is_adaptive:
            self.last_unreduced_loss = None

    def forward(self, device="cpu", iteration=None):
        if self.trunk_points_sampler.is_adaptive:
            trunk_points = self.trunk_points_sampler.sample_points(unreduced_loss=self.last_unreduced_loss,
                                                                   device=device)
            self.last_unreduced_loss = None
        else:
            trunk_points = self.trunk_points_sampler.sample_points(device=device)
        # the F copies are a broadcast view: the trunk kernel evaluates the N points once
        base = trunk_points.as_tensor
        n_fn = self.branch_function_sampler.n_functions
        trunk_points = Points(base.unsqueeze(0).expand(n_fn, *base.shape), trunk_points.space)
        trunk_coordinates, trunk_points = trunk_points.track_coord_gradients()

        branch_functions = self.branch_function_sampler.sample_functions(device)
        if callable(branch_functions):
            branch_function_input = branch_functions(self.net.branch.grid)
        else:
            branch_function_input = branch_functions

        data = None
        if self.eval_function_set_for_residual:
            if self.data_sampler is not None:
                data = self.data_sampler.sample_functions(device)
                if callable(data):
                    data = data(trunk_points)
            elif callable(branch_functions):
                data = branch_functions(trunk_points)
            else:
                data = branch_functions

        with jets.scope(id(self)):
            model_out = self.net(trunk_points, branch_function_input, device=device)
            data_fn = {name: self.data_functions[name](trunk_coordinates) for name in self.data_functions}
            input_dict = {**model_out.coordinates, **trunk_coordinates, **self.parameter.coordinates, **data_fn}
            if data is not None:
                input_dict = {**input_dict, **data.coordinates}
            residual = self.residual_fn(input_dict)
            if (type(self.error_fn) is SquaredError and self.reduce_fn is torch.mean
                    and not self.trunk_points_sampler.is_adaptive and _fusable_residual(residual)):
                return _SquaredErrorMean.apply(residual)      # mean over functions of the summed squares
            unreduced_loss = self.error_fn(residual)
        if self.trunk_points_sampler.is_adaptive:
            self.last_unreduced_loss = unreduced_loss
        return self.reduce_fn(unreduced_loss)


class PIDeepONetCondition(DeepONetSingleModuleCondition):
    """mean over functions of the summed squared residual (reference ``deeponet_condition.py:110-151``)."""

    def __init__(self, deeponet_model, branch_function_sampler, trunk_points_sampler, residual_fn,
                 data_sampler=None, name="pi_deeponet_condition", track_gradients=True, data_functions={},
                 parameter=None, weight=1.0):
        super().__init__(deeponet_model, branch_function_sampler, trunk_points_sampler, residual_fn=residual_fn,
                         data_sampler=data_sampler, error_fn=SquaredError(), reduce_fn=torch.mean, name=name,
                         track_gradients=track_gradients, data_functions=data_functions, parameter=parameter,
                         weight=weight)
