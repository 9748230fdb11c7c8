"""Name-based argument binding for user callables.

``residual_fn(u, x, f)``, ``data_functions``, ``filter_fn`` and domain parameters are
plain Python callables whose *parameter names* select what they receive (variables of the
input/output spaces, learnable parameters, data functions).  Same semantics as the
reference's ``utils/user_fun.py`` (``UserFunction`` :13-245, ``DomainUserFunction``
:248-324), pinned by the reference's ``tests/tests_utils/test_user_function.py``.
"""
import copy
import inspect

import torch

from .spaces import Points


class UserFunction:
    def __init__(self, fun, defaults={}, args={}):
        if isinstance(fun, (UserFunction, DomainUserFunction)):
            self.fun, self.defaults, self.args = fun.fun, fun.defaults, fun.args
            return
        self.fun, self.defaults, self.args = fun, defaults, args
        if callable(fun) and self.defaults == {} and self.args == {}:
            spec = inspect.getfullargspec(fun)
            if spec.varargs is not None or spec.varkw is not None:
                raise ValueError("Variable arguments are not supported in UserFunctions. "
                                 "Please use keyword arguments.")
            self.args = spec.args + spec.kwonlyargs
            self.defaults = {}
            if spec.defaults is not None:
                n = len(spec.defaults)
                self.defaults = {self.args[-i]: spec.defaults[-i] for i in range(n, 0, -1)}

    # -- introspection -----------------------------------------------------------------
    @property
    def necessary_args(self):
        return [a for a in self.args if a not in self.defaults]

    @property
    def optional_args(self):
        return [a for a in self.args if a in self.defaults]

    def __name__(self):
        return self.fun.__name__

    # -- evaluation --------------------------------------------------------------------
    def _bind(self, given):
        for key in self.necessary_args:
            assert key in given, f"The argument '{key}' is necessary in {self.__name__()} but not given."
        bound = {k: given[k] for k in self.args if k in given}
        bound.update({k: self.defaults[k] for k in self.args if k not in given})
        return bound

    def __call__(self, args={}, vectorize=False):
        if isinstance(args, Points):
            args = args.coordinates
        bound = self._bind(args)
        if vectorize:
            return self.apply_to_batch(bound)
        return self.evaluate_function(**bound)

    def evaluate_function(self, **inp):
        return self.fun(**inp) if callable(self.fun) else self.fun

    def apply_to_batch(self, inp):
        batch = max(len(inp[k]) for k in inp)
        out = []
        for i in range(batch):
            row = {k: (inp[k][i] if len(inp[k]) == batch else inp[k]) for k in inp}
            res = self.fun(**row)
            if res is not None:
                out.append(res)
        return out

    def partially_evaluate(self, **args):
        if not callable(self.fun):
            return self.fun
        if all(a in args for a in self.necessary_args):
            bound = {k: args[k] for k in self.args if k in args}
            bound.update({k: self.defaults[k] for k in self.args if k not in args})
            return self.fun(**bound)
        clone = copy.deepcopy(self)
        clone.set_default(**args)
        return clone

    def set_default(self, **args):
        self.defaults.update({k: v for k, v in args.items() if k in self.args})

    def remove_default(self, *args, **kwargs):
        for key in list(args) + list(kwargs.keys()):
            self.defaults.pop(key)

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, copy.deepcopy(v, memo))
        return new


class DomainUserFunction(UserFunction):
    """Domain parameters: constants become float tensors, callables get a trailing axis
    (reference ``user_fun.py:270-324``)."""

    def __call__(self, args={}, device="cpu"):
        if isinstance(args, Points):
            args = args.coordinates
        if len(args) != 0:
            device = args[list(args.keys())[0]].device
        bound = self._bind(args)
        return self.evaluate_function(device=device, **bound)

    def evaluate_function(self, device="cpu", **inp):
        if callable(self.fun):
            val = self.fun(**inp)
            if not isinstance(val, torch.Tensor):
                val = torch.tensor(val, device=device)
            return val[:, None]
        if isinstance(self.fun, torch.Tensor):
            self.fun = self.fun.to(device)
            return self.fun
        return torch.tensor(self.fun, device=device).float()

    @property
    def is_constant(self):
        return not callable(self.fun)
