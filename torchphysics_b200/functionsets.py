"""Function sets and function samplers: the branch-side data format of the physics-informed
DeepONet path (reference ``problem/domains/functionsets/functionset.py:6-207``,
``custom_functionset.py:10-61``, ``functionset_operations.py:82-158``,
``problem/samplers/function_sampler.py:7-104``).

These are the callers' data formats on the branch side of SURVEY 8(a) rows a12/a13: which
functions enter a training step and how they are discretised on the sensor grid / evaluated
at the trunk points.  The arithmetic is the user's closed-form function on torch tensors;
parameter sampling goes through the native Philox sampler kernels (``samplers.py``).
"""
import torch

from .spaces import Points
from .userfn import UserFunction


class FunctionSet:
    """A family of functions of one ``FunctionSpace`` (reference ``functionset.py:6-207``)."""

    def __init__(self, function_space, function_set_size):
        self.function_space = function_space
        self.function_set_size = function_set_size

    @property
    def is_discretized(self):
        return False

    def is_discretization_of(self, function_set):
        return False

    def create_functions(self, device="cpu"):
        raise NotImplementedError

    def get_function(self, idx):
        raise NotImplementedError

    def append(self, other):
        """union of two sets of the same output space (reference ``functionset.py:139-158``)."""
        assert self.function_space.output_space == other.function_space.output_space, \
            "Both FunctionSets need the same output space!"
        if isinstance(other, FunctionSetCollection):
            return FunctionSetCollection(self.function_space, [self] + list(other.function_sets))
        return FunctionSetCollection(self.function_space, [self, other])

    def _transform_locations(self, locations):
        """one copy of the locations per selected function (reference ``functionset.py:192-204``)."""
        if len(locations.shape) == 1:
            locations = locations.unsqueeze(0)
        sub = locations[..., list(self.function_space.input_space.keys())].as_tensor
        if sub.shape[0] == 1:
            return torch.repeat_interleave(sub, len(self.current_idx), dim=0)
        return sub[:len(self.current_idx)]


class CustomFunctionSet(FunctionSet):
    """``custom_fn(input variables, sampled parameters)`` (reference ``custom_functionset.py:10-61``)."""

    def __init__(self, function_space, parameter_sampler, custom_fn):
        super().__init__(function_space, parameter_sampler.n_points)
        if not isinstance(custom_fn, UserFunction):
            custom_fn = UserFunction(custom_fn)
        self.custom_fn = custom_fn
        self.parameter_sampler = parameter_sampler

    def create_functions(self, device="cpu"):
        self.param_samples = self.parameter_sampler.sample_points(device=device)

    def get_function(self, idx):
        if isinstance(idx, int):
            idx = [idx]
        self.current_idx = idx
        return self._evaluate_fn_at_locations

    def _evaluate_fn_at_locations(self, locations):
        loc = self._transform_locations(locations)
        idx = self.current_idx
        if isinstance(idx, torch.Tensor):
            idx = idx.to(self.param_samples.as_tensor.device)
        par = self.param_samples.as_tensor[idx].unsqueeze(1).expand(-1, loc.shape[1], -1)
        fn_input = torch.cat([loc, par], dim=-1)
        out = self.custom_fn(Points(fn_input, self.function_space.input_space * self.param_samples.space))
        return Points(out, self.function_space.output_space)


class FunctionSetCollection(FunctionSet):
    """Several sets addressed by one running index (reference ``functionset_operations.py:82-158``)."""

    def __init__(self, function_space, function_sets):
        super().__init__(function_space, sum(fs.function_set_size for fs in function_sets))
        self.function_sets = list(function_sets)
        self.current_idx = []
        self.device = "cpu"

    @property
    def is_discretized(self):
        return any(fs.is_discretized for fs in self.function_sets)

    def append(self, other):
        more = other.function_sets if isinstance(other, FunctionSetCollection) else [other]
        return FunctionSetCollection(self.function_space, self.function_sets + list(more))

    def create_functions(self, device="cpu"):
        for fs in self.function_sets:
            fs.create_functions(device)
        self.device = device

    def get_function(self, idx):
        if isinstance(idx, int):
            idx = [idx]
        self.current_idx = torch.as_tensor(idx, dtype=torch.long).to(self.device)
        return self._evaluate_collection

    def _evaluate_collection(self, location):
        pieces, order, start = [], [], 0
        for fs in self.function_sets:
            hit = torch.where((self.current_idx >= start) & (self.current_idx < start + fs.function_set_size))[0]
            if len(hit) > 0:
                pieces.append(fs.get_function(self.current_idx[hit] - start)(location).as_tensor)
                order.append(hit)
            start += fs.function_set_size
        out = torch.cat(pieces, dim=0)
        # same gather as the reference (``functionset_operations.py:157``)
        return Points(out[torch.cat(order)], self.function_space.output_space)


class FunctionSampler:
    """Draws ``n_functions`` functions of a set per call (reference ``function_sampler.py:7-58``)."""

    def __init__(self, n_functions, function_set, function_creation_interval=0):
        self.n_functions = n_functions
        self.function_set = function_set
        self.function_creation_interval = function_creation_interval
        assert self.n_functions <= self.function_set.function_set_size, \
            "The number of functions to be sampled must be smaller than the function set size."
        self.iteration_counter = self.function_creation_interval
        self.current_indices = torch.zeros(n_functions, dtype=torch.int64)

    def _check_recreate_functions(self, device="cpu"):
        if self.iteration_counter >= self.function_creation_interval:
            self.function_set.create_functions(device=device)
            self.iteration_counter = -1
        self.iteration_counter += 1

    def sample_functions(self, device="cpu"):
        raise NotImplementedError


class FunctionSamplerRandomUniform(FunctionSampler):
    """random subset, ``torch.randperm`` on the host generator (reference ``function_sampler.py:61-67``).  While a CUDA graph is
    being captured (``Trainer(cuda_graph=True)``) the subset is drawn on the device instead -- ranks of uniform keys from torch's
    capture-aware CUDA generator -- so that every replay picks new functions without a host round trip."""

    def sample_functions(self, device="cpu"):
        self._check_recreate_functions(device=device)
        size = self.function_set.function_set_size
        if torch.device(device).type == "cuda" and torch.cuda.is_current_stream_capturing():
            self.current_indices = torch.rand(size, device=device).argsort()[:self.n_functions]
        else:
            self.current_indices = torch.randperm(size)[:self.n_functions]
        return self.function_set.get_function(self.current_indices)


class FunctionSamplerOrdered(FunctionSampler):
    """consecutive blocks, wrapping around (reference ``function_sampler.py:70-87``)."""

    def __init__(self, n_functions, function_set, function_creation_interval=0):
        super().__init__(n_functions, function_set, function_creation_interval)
        self.current_indices = torch.arange(self.n_functions, dtype=torch.int64)
        self.new_indieces = torch.zeros_like(self.current_indices, dtype=torch.int64)

    def sample_functions(self, device="cpu"):
        self._check_recreate_functions(device=device)
        self.current_indices = self.new_indieces.clone()
        out = self.function_set.get_function(self.current_indices)
        self.new_indieces = (self.current_indices + self.n_functions) % self.function_set.function_set_size
        return out


class FunctionSamplerCoupled(FunctionSampler):
    """re-uses the indices another sampler drew (reference ``function_sampler.py:90-104``)."""

    def __init__(self, function_set, coupled_sampler):
        super().__init__(coupled_sampler.n_functions, function_set, coupled_sampler.function_creation_interval)
        self.coupled_sampler = coupled_sampler

    def sample_functions(self, device="cpu"):
        self._check_recreate_functions(device=device)
        self.current_indices = self.coupled_sampler.current_indices
        return self.function_set.get_function(self.current_indices)
