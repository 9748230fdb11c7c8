"""Point samplers with the reference's interface (``problem/samplers``):
``RandomUniformSampler``, ``GridSampler``, ``LHSSampler`` and the sampler algebra
(``*`` product, ``+`` concatenation, ``append``, ``make_static``).

Host orchestration only -- the points come from the libtpx sampler kernels through
``domains.py``.  Reference: ``sampler_base.py:13-383``, ``random_samplers.py:12-90,163-238``,
``grid_samplers.py:13-105``.
"""
import ctypes
import math
import warnings

import torch

from . import _lib, domains
from .spaces import Points
from .userfn import UserFunction


class PointSampler:
    def __init__(self, n_points=None, density=None, filter_fn=None):
        self.n_points = n_points
        self.density = density
        self.length = None
        self.filter_fn = UserFunction(filter_fn) if filter_fn else None

    @classmethod
    def empty(cls, **kwargs):
        return EmptySampler().make_static()

    def set_length(self, length):
        self.length = length

    def __iter__(self):
        return self

    def __next__(self):
        return self.sample_points()

    def __len__(self):
        if self.length is not None:
            return self.length
        if self.n_points is not None:
            return self.n_points
        raise ValueError("The expected number of samples is not known yet. Set the length by using "
                         ".set_length, if this property is needed")

    def make_static(self, resample_interval=math.inf):
        return StaticSampler(self, resample_interval)

    @property
    def is_static(self):
        return isinstance(self, StaticSampler)

    @property
    def is_adaptive(self):
        return False

    def sample_points(self, params=Points.empty(), device="cpu"):
        if self.filter_fn:
            return self._sample_points_with_filter(params, device)
        return self._sample_points(params, device)

    def _sample_points(self, params=Points.empty(), device="cpu"):
        raise NotImplementedError

    def _sample_points_with_filter(self, params=Points.empty(), device="cpu"):
        raise NotImplementedError

    def __mul__(self, other):
        assert isinstance(other, PointSampler)
        return ProductSampler(self, other)

    def __add__(self, other):
        assert isinstance(other, PointSampler)
        return ConcatSampler(self, other)

    def append(self, other):
        assert isinstance(other, PointSampler)
        return AppendSampler(self, other)

    # -- helpers shared by the concrete samplers --------------------------------------
    @staticmethod
    def _repeat_params(params, n):
        return Points(torch.repeat_interleave(params.as_tensor, n, dim=0), params.space)

    def _sample_params_independent(self, sample_function, params, device):
        points = sample_function(n=self.n_points, d=self.density, device=device)
        count = len(points)
        self.set_length(count)
        repeated_params = self._repeat_params(params, count)
        return points.repeat(max(1, len(params))).join(repeated_params)

    def _apply_filter(self, sample_points):
        keep = self.filter_fn(sample_points)
        return sample_points[torch.where(keep)[0], ]

    def _cut_tensor_to_length_n(self, points):
        return points[: self.n_points, ]

    def _check_iteration_number(self, iterations, found):
        if iterations == 10:
            warnings.warn(f"Sampling points with filter did run 10 iterations and until now only found "
                          f"{found} from {self.n_points} points. This may take some time.")
        elif iterations >= 20 and found <= 1:
            raise RuntimeError("Run 20 iterations and could not find a single valid point for the "
                               "filter condition.")

    def _n_points_with_filter(self, sample_function, params, device):
        """random_samplers.py:60-90: oversample by the observed acceptance until n_points pass."""
        out = None
        for i in range(max(1, len(params))):
            ith = params[i, ] if len(params) > 0 else Points.empty()
            kept, found, total, iterations = None, 0, 0, 0
            n_try = self.n_points
            while found < self.n_points:
                new = sample_function(n_try, None, Points.empty(), device)
                new = new.join(self._repeat_params(ith, len(new)))
                total += n_try
                new = self._apply_filter(new)
                found = max(1, found + len(new))
                n_try = int(1.1 * (self.n_points - found) * total / found) + 10
                if n_try > 1.0e9:
                    raise RuntimeError("To many points need to be sampled to fulfill filter condition!")
                kept = new if kept is None else kept | new
                iterations += 1
                self._check_iteration_number(iterations, found)
            kept = self._cut_tensor_to_length_n(kept)
            out = kept if out is None else out | kept
        return out


class ProductSampler(PointSampler):
    def __init__(self, sampler_a, sampler_b):
        self.sampler_a, self.sampler_b = sampler_a, sampler_b
        super().__init__()

    def __len__(self):
        return self.length if self.length else len(self.sampler_a) * len(self.sampler_b)

    def sample_points(self, params=Points.empty(), device="cpu"):
        b_points = self.sampler_b.sample_points(params, device=device)
        a_points = self.sampler_a.sample_points(b_points, device=device)
        self.set_length(len(a_points))
        return a_points


class ConcatSampler(PointSampler):
    def __init__(self, sampler_a, sampler_b):
        self.sampler_a, self.sampler_b = sampler_a, sampler_b
        super().__init__()

    def __len__(self):
        return self.length if self.length else len(self.sampler_a) + len(self.sampler_b)

    def sample_points(self, params=Points.empty(), device="cpu"):
        a = self.sampler_a.sample_points(params, device=device)
        b = self.sampler_b.sample_points(params, device=device)
        self.set_length(len(a) + len(b))
        return a | b


class AppendSampler(PointSampler):
    def __init__(self, sampler_a, sampler_b):
        self.sampler_a, self.sampler_b = sampler_a, sampler_b
        super().__init__()

    def __len__(self):
        return self.length if self.length else len(self.sampler_a)

    def sample_points(self, params=Points.empty(), device="cpu"):
        a = self.sampler_a.sample_points(params, device=device)
        b = self.sampler_b.sample_points(params, device=device)
        self.set_length(len(a))
        return a.join(b)


class StaticSampler(PointSampler):
    """Samples once (or every ``resample_interval`` calls) and then returns the stored points
    (reference ``sampler_base.py:333-383``)."""

    def __init__(self, sampler, resample_interval=math.inf):
        super().__init__(sampler.n_points, sampler.density, None)
        self.filter_fn = sampler.filter_fn
        self.length = None
        self.sampler = sampler
        self.created_points = None
        self.resample_interval = resample_interval
        self.counter = 0

    def __len__(self):
        return self.length if self.length else len(self.sampler)

    def __next__(self):
        return self.created_points if self.created_points else self.sample_points()

    def sample_points(self, params=Points.empty(), device="cpu", **kwargs):
        self.counter += 1
        if self.created_points and self.counter < self.resample_interval:
            self.created_points = self.created_points.to(device)
            return self.created_points
        self.counter = 0
        self.created_points = self.sampler.sample_points(params, device=device, **kwargs)
        return self.created_points

    def make_static(self, resample_interval=math.inf):
        self.resample_interval = resample_interval
        return self


class EmptySampler(PointSampler):
    def __init__(self):
        super().__init__(n_points=0)

    def sample_points(self, params=Points.empty(), device="cpu", **kwargs):
        return Points.empty()


class DataSampler(PointSampler):
    """Returns the given points (reference ``data_samplers.py:11``)."""

    def __init__(self, points):
        if isinstance(points, Points):
            self.points = points
        elif isinstance(points, dict):
            self.points = Points.from_coordinates(points)
        else:
            raise TypeError("points should be one of Points or dict.")
        super().__init__(n_points=len(self.points))

    def sample_points(self, params=Points.empty(), device="cpu"):
        self.points = self.points.to(device)
        if len(params) <= 1:
            return self.points.join(self._repeat_params(params, len(self.points))) if len(params) else self.points
        rep = self.points.repeat(len(params))
        return rep.join(self._repeat_params(params, len(self.points)))


class RandomUniformSampler(PointSampler):
    """Uniform random points of a domain (reference ``random_samplers.py:12-90``)."""

    def __init__(self, domain, n_points=None, density=None, filter_fn=None):
        super().__init__(n_points=n_points, density=density, filter_fn=filter_fn)
        self.domain = domain

    def _sample_points(self, params=Points.empty(), device="cpu"):
        if self.n_points:
            pts = self.domain.sample_random_uniform(self.n_points, params=params, device=device)
            return pts.join(self._repeat_params(params, len(self)))
        return self._sample_params_independent(self.domain.sample_random_uniform, params, device)

    def _sample_points_with_filter(self, params=Points.empty(), device="cpu"):
        if self.n_points:
            return self._n_points_with_filter(self.domain.sample_random_uniform, params, device)
        return self._apply_filter(self._sample_points(params, device))


class GridSampler(PointSampler):
    """Equidistant points of a domain (reference ``grid_samplers.py:13-105``)."""

    def __init__(self, domain, n_points=None, density=None, filter_fn=None):
        super().__init__(n_points=n_points, density=density, filter_fn=filter_fn)
        self.domain = domain

    def _sample_points(self, params=Points.empty(), device="cpu"):
        return self._sample_params_independent(self.domain.sample_grid, params, device)

    def _sample_points_with_filter(self, params=Points.empty(), device="cpu"):
        if not self.n_points:
            return self._apply_filter(self._sample_points(params, device))
        out = None
        for i in range(max(1, len(params))):
            ith = params[i, ] if len(params) > 0 else Points.empty()
            new = self._filtered_grid(ith, self.n_points, device)
            if len(new) != self.n_points:
                if len(new) == 0:
                    warnings.warn("First iteration did not find any valid grid points, for the given filter. "
                                  "Will try again with n = 10 * self.n_points. Or else use only random points!")
                    scaled = int(10 * self.n_points)
                else:
                    scaled = int(self.n_points ** 2 / len(new))
                new = self._filtered_grid(ith, scaled, device)
                if len(new) != self.n_points:
                    rnd = RandomUniformSampler(self.domain, n_points=self.n_points)
                    rnd.filter_fn = self.filter_fn
                    new = new | rnd.sample_points(ith, device=device)
            new = self._cut_tensor_to_length_n(new)
            out = new if out is None else out | new
        return out

    def _filtered_grid(self, ith, n, device):
        pts = self.domain.sample_grid(n, params=Points.empty(), device=device)
        return self._apply_filter(pts.join(self._repeat_params(ith, len(pts))))


class LHSSampler(PointSampler):
    """Latin hypercube sampling in the bounding box, points outside the domain replaced by
    uniform random ones (reference ``random_samplers.py:163-238``)."""

    def __init__(self, domain, n_points):
        super().__init__(n_points=n_points)
        self.domain = domain

    def _sample_points(self, params=Points.empty(), device="cpu"):
        out = None
        for i in range(max(1, len(params))):
            ith = params[i, ] if len(params) > 0 else Points.empty()
            new = self._lhs(device).join(self._repeat_params(ith, self.n_points))
            inside = self.domain._contains(new)
            new = new[torch.where(inside)[0], ]
            if len(new) != self.n_points:
                rnd = RandomUniformSampler(self.domain, n_points=self.n_points - len(new))
                new = new | rnd.sample_points(ith, device=device)
            out = new if out is None else out | new
        return out

    def _lhs(self, device):
        dev = domains._compute_device(device)
        n, d = self.n_points, self.domain.dim
        lib = _lib.lib()
        with _lib.on_device(dev):
            box = torch.tensor(self.domain._box(), dtype=torch.float32, device=dev)
            keys = torch.empty((d, n), dtype=torch.int64, device=dev)
            key = domains._rng.next_key()
            _lib.check(lib.tpx_sample_keys(ctypes.c_uint64(key), ctypes.c_uint64(0), ctypes.c_int64(d * n),
                                           _lib.ptr(keys), _lib.stream_ptr()))
            perm = torch.argsort(keys, dim=1).contiguous()       # iid keys -> uniform permutation per axis
            out = torch.empty((n, d), dtype=torch.float32, device=dev)
            _lib.check(lib.tpx_sample_lhs(ctypes.c_int32(d), _lib.ptr(box), _lib.ptr(perm), ctypes.c_uint64(key),
                                          ctypes.c_uint64(0), ctypes.c_int64(n), _lib.ptr(out), ctypes.c_int32(d),
                                          ctypes.c_int32(0), _lib.stream_ptr()))
        return Points(out.to(device), self.domain.space)


class HostStreamSampler(PointSampler):
    """Collocation batches that live in pinned HOST memory, streamed to the device one batch ahead.

    ``batches`` is a sequence of (n, dim) float32 tensors (pinned here if they are not); ``sample_points``
    returns batch i on the device and starts the host->device copy of batch i+1 on a side stream, so the
    copy overlaps the kernels of step i (two device buffers; the caller's per-step read of the loss is what
    orders a buffer's reuse after the step that consumed it).  The reference has no counterpart: its samplers
    create points on ``device`` directly (``sampler_base.py:130``); this is the drop-in for points produced
    on the host."""

    def __init__(self, batches, space):
        batches = [b if b.is_pinned() else b.contiguous().pin_memory() for b in batches]
        assert len(batches) > 0 and all(b.shape == batches[0].shape and b.dtype == torch.float32 for b in batches)
        super().__init__(n_points=batches[0].shape[0])
        self.batches, self.space = batches, space
        self._i = 0
        self._dev = None

    def __len__(self):
        return self.n_points

    def _start_copy(self, slot, index):
        with torch.cuda.stream(self._stream):
            self._bufs[slot].copy_(self.batches[index % len(self.batches)], non_blocking=True)
            self._ready[slot].record(self._stream)

    def sample_points(self, params=Points.empty(), device="cpu"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("HostStreamSampler delivers to a CUDA device")
        if self._dev != dev:
            self._dev = dev
            with _lib.on_device(dev):
                self._stream = torch.cuda.Stream()
                self._bufs = [torch.empty(self.batches[0].shape, dtype=torch.float32, device=dev) for _ in range(2)]
                self._ready = [torch.cuda.Event(), torch.cuda.Event()]
                self._start_copy(self._i % 2, self._i)
        slot = self._i % 2
        torch.cuda.current_stream(dev).wait_event(self._ready[slot])
        # the other buffer was last read by step i-1, which the current stream has already finished enqueuing:
        # the copy stream waits for it before overwriting
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        self._start_copy(1 - slot, self._i + 1)
        self._i += 1
        return Points(self._bufs[slot], self.space)


class AdaptiveSampler(PointSampler):
    """Sampler that needs the per-point loss of the previous iteration (reference ``sampler_base.py:396-440``)."""

    @property
    def is_adaptive(self):
        return True

    def sample_points(self, unreduced_loss=None, params=Points.empty(), device="cpu"):
        raise NotImplementedError


class _RejectionAdaptive(AdaptiveSampler):
    def __init__(self, domain, n_points=None, density=None, filter_fn=None):
        super().__init__(n_points=n_points, density=density, filter_fn=filter_fn)
        self.domain = domain
        self.random_sampler = RandomUniformSampler(domain, n_points=n_points, density=density, filter_fn=filter_fn)
        self.last_points = None

    def _level(self, unreduced_loss):
        raise NotImplementedError

    def sample_points(self, unreduced_loss=None, params=Points.empty(), device="cpu"):
        new_points = self.random_sampler.sample_points(params, device=device)
        if self.last_points is None or unreduced_loss is None:
            self.last_points = new_points
        else:
            loss = unreduced_loss.detach()
            max_l, min_l = torch.max(loss), torch.min(loss)
            replace = loss < min_l + (max_l - min_l) * self._level(loss)
            self.last_points._t[replace, :] = new_points._t[replace, :]
        return self.last_points


class AdaptiveThresholdRejectionSampler(_RejectionAdaptive):
    """keeps points whose loss is above min + ratio (max - min), resamples the rest uniformly
    (reference ``random_samplers.py:241-292``)."""

    def __init__(self, domain, resample_ratio, n_points=None, density=None, filter_fn=None):
        super().__init__(domain, n_points=n_points, density=density, filter_fn=filter_fn)
        self.resample_ratio = resample_ratio

    def _level(self, loss):
        return self.resample_ratio


class AdaptiveRandomRejectionSampler(_RejectionAdaptive):
    """keeps a point with probability (loss - min) / (max - min) (reference ``random_samplers.py:295-345``)."""

    def _level(self, loss):
        return torch.rand_like(loss)


# function sets / function samplers live in the same namespaces as in the reference
from .functionsets import FunctionSampler, FunctionSamplerRandomUniform, FunctionSamplerOrdered, FunctionSamplerCoupled  # noqa: E402,F401
