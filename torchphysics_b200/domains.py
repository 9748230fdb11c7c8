"""Domains with the reference's interface (``problem/domains``): ``Interval``,
``Parallelogram``, ``Circle`` and the boolean / product operations ``Cut`` (-), ``Union``
(+), ``Intersection`` (&), ``Product`` (*), for CONSTANT domain parameters.

Every point that is produced or tested here is produced or tested by a libtpx kernel
(``csrc/sampler.cu``): Philox4x32-10 sampling, boolean membership programs and rejection
with warp-ballot stable compaction.  The classes below only translate a domain expression
into the C descriptors (``tpx_shape`` / ``tpx_region``) and keep the bookkeeping the
reference keeps (volumes, bounding boxes, point counts).  ``device`` says where the
result is delivered; the computation itself always runs on the current CUDA device.

Reference: ``domain.py:8-292``, ``domain1D/interval.py:7-93``,
``domain2D/parallelogram.py:7-156``, ``domain2D/circle.py:8-108``,
``domainoperations/{cut,union,intersection,product,sampler_helper}.py``.
"""
import ctypes
import math
import warnings

import numpy as np
import torch

from . import _lib
from .spaces import Points

_F32 = np.float32


# --------------------------------------------------------------------------------------
# random stream
# --------------------------------------------------------------------------------------
class _Rng:
    """One 64-bit Philox key per sampling call, drawn from torch's default CPU generator:
    ``torch.manual_seed`` therefore reproduces (and re-seeding replays) every sampler, and the
    points of a call do not depend on how many candidates earlier calls consumed."""

    def __init__(self):
        self.stream = 0
        self.device_key = None          # int64 tensor holding the 64-bit key on the GPU (CUDA-graph capture), or None

    def use_device_key(self, device):
        """From now on the samplers without rejection read their key from GPU memory and advance it with a one-thread
        kernel per call (``tpx_key_advance`` / ``tpx_sample_uniform_dkey``): a sampling call becomes capturable in a CUDA
        graph, and every replay of the graph draws fresh points.  The starting key comes from torch's generator."""
        if device is None:
            self.device_key = None
            return
        k = self.next_key()
        self.device_key = torch.tensor([k - (1 << 64) if k >= (1 << 63) else k], dtype=torch.int64, device=device)

    @staticmethod
    def _mix(x):
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return x ^ (x >> 31)

    def next_key(self):
        draw = int(torch.randint(0, 2 ** 62, (1,)).item())
        return self._mix(draw ^ self._mix(self.stream + 0x5851F42D))


_rng = _Rng()


def use_device_key(device):
    """see ``_Rng.use_device_key`` (``None`` switches back to host-drawn keys)"""
    _rng.use_device_key(device)


def set_rank_stream(rank):
    """Disjoint sample streams for data-parallel ranks."""
    _rng.stream = int(rank)


def _compute_device(device):
    dev = torch.device(device)
    if dev.type == "cuda":
        return dev
    if not torch.cuda.is_available():
        raise _lib.TpxLibraryError("sampling runs on the GPU (libtpx); no CUDA device is visible")
    return torch.device("cuda", torch.cuda.current_device())


def _const(value, what):
    if callable(value):
        raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED,
                            f"{what} given as a function of other variables: only constant domain "
                            "parameters are on the native path")
    if isinstance(value, torch.Tensor):
        value = value.detach().cpu().numpy()
    return np.asarray(value, dtype=_F32).reshape(-1)


def _region_struct(shapes, ops):
    if len(shapes) > _lib.TPX_MAX_SHAPES or len(ops) > _lib.TPX_MAX_OPS:
        raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "domain expression too large for one region program")
    rg = _lib.Region()
    rg.n_shapes, rg.n_ops = len(shapes), len(ops)
    for i, s in enumerate(shapes):
        rg.shapes[i] = s
    for i, o in enumerate(ops):
        rg.ops[i] = o
    return rg


class Domain:
    """Base class (reference ``domain.py:8-292``)."""

    def __init__(self, space, dim=None):
        self.space = space
        self.dim = space.dim if dim is None else dim
        self._user_volume = None
        self.necessary_variables = set()
        self._accept_rate = 1.0

    # -- algebra -------------------------------------------------------------------
    def __add__(self, other):
        if self.space != other.space:
            raise ValueError("united domains should lie in the same space.")
        return UnionDomain(self, other)

    __or__ = __add__

    def __sub__(self, other):
        if self.space != other.space:
            raise ValueError("complemented domains should lie in the same space.")
        return CutDomain(self, other)

    def __and__(self, other):
        if self.space != other.space:
            raise ValueError("Intersected domains should lie in the same space.")
        return IntersectionDomain(self, other)

    def __mul__(self, other):
        return ProductDomain(self, other)

    def __call__(self, **data):
        return self                               # constant domains do not depend on data

    # -- volume ----------------------------------------------------------------------
    def set_volume(self, volume):
        self._user_volume = float(_const(volume, "volume")[0])

    def _get_volume(self):
        raise NotImplementedError

    def _volume_value(self):
        return self._user_volume if self._user_volume is not None else float(self._get_volume())

    def volume(self, params=Points.empty(), device="cpu"):
        rows = max(1, len(params))
        return torch.full((rows, 1), self._volume_value(), dtype=torch.float32, device=device)

    def len_of_params(self, params):
        return max(1, len(params))

    def compute_n_from_density(self, d, params=Points.empty()):
        if len(params) > 1:
            raise ValueError("Sampling with a density is only possible for one given pair of parameters.")
        return int(math.ceil(_F32(d) * _F32(self._volume_value())))

    def _count(self, n, d, params):
        if n is None or (d and not n):
            n = self.compute_n_from_density(d, params)
        return int(n) * self.len_of_params(params)

    # -- native description ------------------------------------------------------------
    def _program(self, col):
        """(shapes, ops): postfix membership program with this domain's first column = col."""
        raise NotImplementedError

    def _generator(self, col):
        """(shape, shapes, ops) if uniform points are 'sample one shape, accept by a region'
        (ops == [] means accept all); None if the domain needs composition on the host."""
        return None

    def _box(self):
        raise NotImplementedError

    def bounding_box(self, params=Points.empty(), device="cpu"):
        return torch.tensor(self._box(), dtype=torch.float32, device=device)

    # -- membership --------------------------------------------------------------------
    def __contains__(self, points):
        return self._contains(points)

    def _contains(self, points, params=Points.empty()):
        pts = points[:, list(self.space.keys())].as_tensor if points.space != self.space else points.as_tensor
        return self._contains_rows(pts.detach(), 0, pts.device)

    def _contains_rows(self, rows, col, out_device):
        dev = _compute_device(rows.device)
        rows = rows.to(device=dev, dtype=torch.float32).contiguous()
        n, stride = rows.shape
        shapes, ops = self._program(col)
        rg = _region_struct(shapes, ops)
        mask = torch.empty(n, dtype=torch.uint8, device=dev)
        with _lib.on_device(dev):
            _lib.check(_lib.lib().tpx_region_contains(ctypes.byref(rg), _lib.ptr(rows), ctypes.c_int64(n),
                                                      ctypes.c_int32(stride), _lib.ptr(mask), _lib.stream_ptr()))
        return mask.bool().reshape(-1, 1).to(out_device)

    # -- sampling ----------------------------------------------------------------------
    def sample_random_uniform(self, n=None, d=None, params=Points.empty(), device="cpu"):
        if not n and d is not None and self._variable_count_with_density():
            return self._sample_random_with_d(d, params, device)
        rows = self._count(n, d, params)
        dev = _compute_device(device)
        out = torch.empty((rows, self.dim_cols()), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            self._fill_random(out, 0, rows)
        return Points(out.to(device), self.space)

    def sample_grid(self, n=None, d=None, params=Points.empty(), device="cpu"):
        if not n and d is not None and self._variable_count_with_density():
            return self._sample_grid_with_d(d, params, device)
        rows = self._count(n, d, params)
        dev = _compute_device(device)
        with _lib.on_device(dev):
            out = self._grid(rows, dev, exact=not d or bool(n))
        return Points(out.to(device), self.space)

    def dim_cols(self):
        return self.space.dim

    def _variable_count_with_density(self):
        return False

    def _fill_random(self, out, col, n):
        """n uniform points of this domain into columns [col, col+dim) of ``out`` (rows x stride)."""
        gen = self._generator(col)
        if gen is None:
            raise NotImplementedError
        shape, shapes, ops = gen
        lib = _lib.lib()
        stride = out.shape[1]
        dkey = _rng.device_key
        if dkey is not None and dkey.device == out.device:
            if ops:
                raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "rejection sampling reads its counters on the host and cannot "
                                    "run with a device-resident key (CUDA-graph capture of the step)")
            _lib.check(lib.tpx_key_advance(_lib.ptr(dkey), _lib.stream_ptr()))
            _lib.check(lib.tpx_sample_uniform_dkey(ctypes.byref(shape), _lib.ptr(dkey), ctypes.c_uint64(0), ctypes.c_int64(n),
                                                   _lib.ptr(out), ctypes.c_int32(stride), _lib.stream_ptr()))
            return
        key = _rng.next_key()
        if not ops:
            _lib.check(lib.tpx_sample_uniform(ctypes.byref(shape), ctypes.c_uint64(key), ctypes.c_uint64(0),
                                              ctypes.c_int64(n), _lib.ptr(out), ctypes.c_int32(stride),
                                              _lib.stream_ptr()))
            return
        if n == 0:
            return
        rg = _region_struct(shapes, ops)
        counters = torch.zeros(4, dtype=torch.int64, device=out.device)
        offset, filled, rounds = 0, 0, 0
        rate = min(max(self._accept_rate, 1e-4), 1.0)
        while filled < n:
            m = int((n - filled) / rate * 1.05) + 2048
            if m > 2_000_000_000:
                raise RuntimeError("Too many candidates needed to fill the domain by rejection.")
            nbytes = ctypes.c_size_t(0)
            _lib.check(lib.tpx_sample_reject_workspace_bytes(ctypes.c_int64(m), ctypes.byref(nbytes)))
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=out.device)
            _lib.check(lib.tpx_sample_uniform_reject(
                ctypes.byref(shape), ctypes.byref(rg), ctypes.c_uint64(key), ctypes.c_uint64(offset),
                ctypes.c_int64(m), ctypes.c_int64(n), _lib.ptr(out), ctypes.c_int32(stride),
                _lib.ptr(counters), _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_ptr()))
            offset += m
            c = counters.tolist()                 # the one host read of a rejection call
            filled, offered, accepted = c[0], c[1], c[3]
            rounds += 1
            if accepted == 0:
                if rounds >= 20:
                    raise RuntimeError("Rejection sampling found no valid point in 20 rounds.")
                rate = max(rate / 5.0, 1e-4)
            else:
                rate = accepted / offered
        self._accept_rate = rate

    def _grid(self, n, dev, exact=True):
        raise NotImplementedError


# --------------------------------------------------------------------------------------
# basic shapes
# --------------------------------------------------------------------------------------
class _ShapeDomain(Domain):
    kind = 0

    def _params(self):
        raise NotImplementedError

    def _shape(self, col):
        s = _lib.Shape()
        s.kind, s.col = self.kind, col
        for i, v in enumerate(self._params()):
            s.p[i] = float(v)
        return s

    def _program(self, col):
        return [self._shape(col)], [_lib.TPX_OP_SHAPE | (0 << 8)]

    def _generator(self, col):
        return self._shape(col), [], []

    def _grid_dims(self, n):
        return 0, 0

    def _grid(self, n, dev, exact=True):
        out = torch.empty((n, self.dim_cols()), dtype=torch.float32, device=dev)
        self._fill_grid(out, 0, n, exact)
        return out

    def _fill_grid(self, out, col, n, exact=True):
        n1, n2 = self._grid_dims(n)
        m = n1 * n2 if n1 else n
        _lib.check(_lib.lib().tpx_sample_grid(ctypes.byref(self._shape(col)), ctypes.c_int64(m), ctypes.c_int32(n1),
                                              ctypes.c_int32(n2), _lib.ptr(out), ctypes.c_int32(out.shape[1]),
                                              _lib.stream_ptr()))
        return m


class Interval(_ShapeDomain):
    """[lower_bound, upper_bound] (reference ``domain1D/interval.py:7-93``)."""
    kind = _lib.TPX_SHAPE_INTERVAL

    def __init__(self, space, lower_bound, upper_bound):
        assert space.dim == 1
        super().__init__(space, dim=1)
        self.lower_bound = float(_const(lower_bound, "lower_bound")[0])
        self.upper_bound = float(_const(upper_bound, "upper_bound")[0])

    def _params(self):
        return [self.lower_bound, self.upper_bound]

    def _get_volume(self):
        return _F32(self.upper_bound) - _F32(self.lower_bound)

    def _box(self):
        return [self.lower_bound, self.upper_bound]


class Parallelogram(_ShapeDomain):
    """origin + s (corner_1 - origin) + t (corner_2 - origin), 0 <= s,t <= 1
    (reference ``domain2D/parallelogram.py:7-156``)."""
    kind = _lib.TPX_SHAPE_PARALLELOGRAM

    def __init__(self, space, origin, corner_1, corner_2):
        assert space.dim == 2
        super().__init__(space, dim=2)
        self.origin = _const(origin, "origin")
        self.corner_1 = _const(corner_1, "corner_1")
        self.corner_2 = _const(corner_2, "corner_2")
        assert len(self.origin) == len(self.corner_1) == len(self.corner_2) == 2

    def _params(self):
        return [*self.origin, *self.corner_1, *self.corner_2]

    def _dirs(self):
        return self.corner_1 - self.origin, self.corner_2 - self.origin

    def _get_volume(self):
        d1, d2 = self._dirs()
        return d1[0] * d2[1] - d1[1] * d2[0]

    def _box(self):
        c3 = self.corner_1 + self.corner_2 - self.origin
        corners = np.stack([self.origin, self.corner_1, self.corner_2, c3])
        return [float(corners[:, 0].min()), float(corners[:, 0].max()),
                float(corners[:, 1].min()), float(corners[:, 1].max())]

    def _grid_dims(self, n):
        d1, d2 = self._dirs()
        l1 = np.sqrt(d1[0] * d1[0] + d1[1] * d1[1], dtype=_F32)
        l2 = np.sqrt(d2[0] * d2[0] + d2[1] * d2[1], dtype=_F32)
        n1 = int(np.sqrt(_F32(n) * l1 / l2, dtype=_F32))
        n2 = int(np.sqrt(_F32(n) * l2 / l1, dtype=_F32))
        return n1, n2

    def _grid(self, n, dev, exact=True):
        n1, n2 = self._grid_dims(n)
        m = n1 * n2
        rows = max(m, n) if exact else m
        out = torch.empty((rows, 2), dtype=torch.float32, device=dev)
        if m:
            self._fill_grid(out, 0, n)
        if exact and m < n:                        # top up with random points (parallelogram.py:147-152)
            self._fill_random(out[m:], 0, n - m)
        return out[:rows]


class Circle(_ShapeDomain):
    """disc of given center and radius (reference ``domain2D/circle.py:8-108``)."""
    kind = _lib.TPX_SHAPE_CIRCLE

    def __init__(self, space, center, radius):
        assert space.dim == 2
        super().__init__(space, dim=2)
        self.center = _const(center, "center")
        self.radius = float(_const(radius, "radius")[0])
        assert len(self.center) == 2

    def _params(self):
        return [*self.center, self.radius]

    def _get_volume(self):
        return _F32(np.pi) * _F32(self.radius) ** 2

    def _box(self):
        c, r = self.center, _F32(self.radius)
        return [float(c[0] - r), float(c[0] + r), float(c[1] - r), float(c[1] + r)]


# --------------------------------------------------------------------------------------
# boundaries of the basic shapes
# --------------------------------------------------------------------------------------
class BoundaryDomain(_ShapeDomain):
    """Boundary of a basic shape (reference ``domain.py:295-348``): a domain of one dimension
    less in the same space, sampled / tested / given normals by the libtpx boundary shapes."""

    def __init__(self, domain):
        assert isinstance(domain, Domain)
        super().__init__(space=domain.space, dim=domain.dim - 1)
        self.domain = domain
        self.necessary_variables = domain.necessary_variables

    def __call__(self, **data):
        return self

    def _params(self):
        return self.domain._params()

    def _box(self):
        return self.domain._box()

    def normal(self, points, params=Points.empty(), device="cpu"):
        """outward unit normals at ``points`` (which should lie on the boundary), shape (len(points), dim)."""
        if not isinstance(points, Points):
            points = Points(points, self.space)
        rows = points[:, list(self.space.keys())].as_tensor if points.space != self.space else points.as_tensor
        out_device = rows.device
        dev = _compute_device(out_device)
        rows = rows.detach().to(device=dev, dtype=torch.float32).contiguous()
        n, stride = rows.shape
        out = torch.empty((n, self.space.dim), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(_lib.lib().tpx_boundary_normal(ctypes.byref(self._shape(0)), _lib.ptr(rows), ctypes.c_int64(n),
                                                      ctypes.c_int32(stride), _lib.ptr(out), _lib.stream_ptr()))
        return out.to(out_device)


class IntervalBoundary(BoundaryDomain):
    """both end points (reference ``interval.py:96-149``)."""
    kind = _lib.TPX_SHAPE_INTERVAL_BOUNDARY

    def __init__(self, domain):
        assert isinstance(domain, Interval)
        super().__init__(domain)

    def _get_volume(self):
        return _F32(2.0)


class IntervalSingleBoundaryPoint(BoundaryDomain):
    """one end point, e.g. for initial conditions (reference ``interval.py:152-191``)."""
    kind = _lib.TPX_SHAPE_INTERVAL_POINT

    def __init__(self, domain, side, normal_vec=-1):
        assert isinstance(domain, Interval)
        super().__init__(domain)
        self.side = float(_const(side, "side")[0])
        self.normal_vec = normal_vec

    def _params(self):
        return [self.side, float(self.normal_vec)]

    def _get_volume(self):
        return _F32(1.0)


class ParallelogramBoundary(BoundaryDomain):
    """the four sides (reference ``parallelogram.py:159-291``)."""
    kind = _lib.TPX_SHAPE_PARALLELOGRAM_BOUNDARY

    def __init__(self, domain):
        assert isinstance(domain, Parallelogram)
        super().__init__(domain)

    def _get_volume(self):
        d1, d2 = self.domain._dirs()
        l1 = np.sqrt(d1[0] * d1[0] + d1[1] * d1[1], dtype=_F32)
        l2 = np.sqrt(d2[0] * d2[0] + d2[1] * d2[1], dtype=_F32)
        return _F32(2.0) * (l1 + l2)


class CircleBoundary(BoundaryDomain):
    """the circle line (reference ``circle.py:111-175``)."""
    kind = _lib.TPX_SHAPE_CIRCLE_BOUNDARY

    def __init__(self, domain):
        assert isinstance(domain, Circle)
        super().__init__(domain)

    def _get_volume(self):
        return _F32(2 * np.pi) * _F32(self.domain.radius)


Interval.boundary = property(lambda self: IntervalBoundary(self))
Interval.boundary_left = property(lambda self: IntervalSingleBoundaryPoint(self, side=self.lower_bound))
Interval.boundary_right = property(lambda self: IntervalSingleBoundaryPoint(self, side=self.upper_bound, normal_vec=1))
Parallelogram.boundary = property(lambda self: ParallelogramBoundary(self))
Circle.boundary = property(lambda self: CircleBoundary(self))


# --------------------------------------------------------------------------------------
# boolean operations
# --------------------------------------------------------------------------------------
def _shift_ops(ops, by):
    return [o if (o & 0xFF) != _lib.TPX_OP_SHAPE else (_lib.TPX_OP_SHAPE | (((o >> 8) + by) << 8)) for o in ops]


def _combine(prog_a, prog_b, tail):
    sa, oa = prog_a
    sb, ob = prog_b
    return sa + sb, oa + _shift_ops(ob, len(sa)) + tail


class _BooleanDomain(Domain):
    def __init__(self, domain_a, domain_b):
        assert domain_a.space == domain_b.space
        super().__init__(space=domain_a.space, dim=domain_a.dim)
        self.domain_a, self.domain_b = domain_a, domain_b

    def _select(self, pts, keep_mask_of, n=None):
        """rows of ``pts`` (device tensor) selected by a boolean (rows,1) mask."""
        idx = torch.where(keep_mask_of.reshape(-1))[0]
        return pts[idx] if n is None else pts[idx[:n]]

    def _variable_count_with_density(self):
        return True


class CutDomain(_BooleanDomain):
    """domain_a minus domain_b (reference ``domainoperations/cut.py:14-108``)."""

    def __init__(self, domain_a, domain_b, contained=False):
        super().__init__(domain_a, domain_b)
        self.contained = contained

    def _program(self, col):
        return _combine(self.domain_a._program(col), self.domain_b._program(col), [_lib.TPX_OP_NOT, _lib.TPX_OP_AND])

    def _generator(self, col):
        gen = self.domain_a._generator(col)
        if gen is None:
            return None
        shape, shapes, ops = gen
        sb, ob = self.domain_b._program(col)
        if ops:
            shapes, ops = _combine((shapes, ops), (sb, ob), [_lib.TPX_OP_NOT, _lib.TPX_OP_AND])
        else:
            shapes, ops = sb, ob + [_lib.TPX_OP_NOT]
        return shape, shapes, ops

    def _get_volume(self):
        if not self.contained:
            warnings.warn("Exact volume of this cut is not known, will use the estimate: volume = "
                          "domain_a.volume. If you need the exact volume for sampling, use domain.set_volume()")
            return self.domain_a._volume_value()
        return self.domain_a._volume_value() - self.domain_b._volume_value()

    def _box(self):
        return self.domain_a._box()

    def _fill_random(self, out, col, n):
        if self._generator(col) is not None:
            return super()._fill_random(out, col, n)
        # generator needs host composition (e.g. a Union): candidates from A, membership of B
        # and the compaction are still kernels; torch only gathers the kept rows.
        filled, m, rounds = 0, n, 0
        while filled < n:
            rounds += 1
            if rounds > 20 or m > 2_000_000_000:      # same limits as Domain._fill_random: an empty cut must not grow until OOM
                raise RuntimeError(f"rejection sampling of {type(self).__name__} found {filled} of {n} points after {rounds - 1} "
                                   "rounds: is domain_a contained in domain_b?")
            cand = torch.zeros((int(m), out.shape[1]), dtype=torch.float32, device=out.device)
            self.domain_a._fill_random(cand, col, int(m))
            keep = ~self.domain_b._contains_rows(cand, col, out.device)
            good = self._select(cand, keep, n - filled)
            out[filled:filled + len(good), col:col + self.space.dim] = good[:, col:col + self.space.dim]
            got = len(good)
            filled += got
            m = 5 * m if got == 0 else int(m * (n - filled) / got * 1.1) + 16

    def _sample_random_with_d(self, d, params, device):
        pts = self.domain_a.sample_random_uniform(d=d, params=params, device=_compute_device(device))
        t = pts.as_tensor
        keep = ~self.domain_b._contains_rows(t, 0, t.device)
        return Points(self._select(t, keep).to(device), self.space)

    def _grid(self, n, dev, exact=True):
        # sampler_helper.py:79-96
        a = self.domain_a._grid(n, dev)
        keep = ~self.domain_b._contains_rows(a, 0, dev)
        inside = int(keep.sum())
        if inside == n:
            return a
        scaled = int(n ** 2 / max(inside, 1))
        a = self.domain_a._grid(scaled, dev)
        a = self._select(a, ~self.domain_b._contains_rows(a, 0, dev))
        if len(a) >= n:
            return a[:n]
        extra = torch.empty((n - len(a), self.space.dim), dtype=torch.float32, device=dev)
        self._fill_random(extra, 0, n - len(a))
        return torch.cat([a, extra], dim=0)

    def _sample_grid_with_d(self, d, params, device):
        pts = self.domain_a.sample_grid(d=d, params=params, device=_compute_device(device))
        t = pts.as_tensor
        keep = ~self.domain_b._contains_rows(t, 0, t.device)
        return Points(self._select(t, keep).to(device), self.space)


class CutBoundaryDomain(Domain):
    """Boundary of ``A - B`` for two basic shapes (reference ``domainoperations/cut.py:111-201``): the part of
    A's boundary outside B plus the part of B's boundary inside A.  Membership is one region program over the
    boundary shapes; random points alternate between the two boundaries with rejection by that program
    (``sampler_helper.py:129-200``), appended in candidate order by the stable-compaction kernels."""

    def __init__(self, domain):
        a, b = domain.domain_a, domain.domain_b
        for d in (a, b):
            if isinstance(d, BoundaryDomain) or not isinstance(d, (_ShapeDomain, _BooleanDomain)):
                raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "boundary of a boolean domain: the operands must be basic shapes "
                                    "or cuts / unions / intersections of them")
        super().__init__(space=domain.space, dim=domain.dim - 1)
        self.domain = domain
        self.necessary_variables = domain.necessary_variables
        # two basic shapes: membership is ONE region program and the rejection sampler one kernel per round; nested operands
        # (a cut of a cut, ...) compose their operands' membership / sampling / normal calls on the device, as the reference
        # does on the host (cut.py:111-201 with domain_a.boundary itself a boolean boundary)
        self._basic = isinstance(a, _ShapeDomain) and isinstance(b, _ShapeDomain)

    def _mask(self, in_a, in_b, on_a, on_b):
        return (on_a & ~in_b) | (on_b & in_a & ~on_a)

    def _contains_rows(self, rows, col, out_device):
        if self._basic:
            return super()._contains_rows(rows, col, out_device)
        a, b = self.domain.domain_a, self.domain.domain_b
        dev = _compute_device(rows.device)
        rows = rows.to(device=dev, dtype=torch.float32).contiguous()
        return self._mask(a._contains_rows(rows, col, dev), b._contains_rows(rows, col, dev),
                          a.boundary._contains_rows(rows, col, dev), b.boundary._contains_rows(rows, col, dev)).to(out_device)

    def _fill_random_composed(self, out, col, n):
        """sampler_helper.py:166-200 with the operands' own samplers: candidates alternate between the two boundaries (in
        proportion to their lengths), kept where they lie on this boundary, appended in candidate order."""
        a, b = self.domain.domain_a.boundary, self.domain.domain_b.boundary
        vm = _F32(self._volume_value())
        scaled = [int(_F32(n) * _F32(a._volume_value()) / vm) + 1, int(_F32(n) * _F32(b._volume_value()) / vm) + 1]
        filled, use_b, rounds = 0, 0, 0
        while filled < n:
            cand = torch.zeros((scaled[use_b], out.shape[1]), dtype=torch.float32, device=out.device)
            (b if use_b else a)._fill_random(cand, col, scaled[use_b])
            good = self._masked(cand, self._contains_rows(cand, col, out.device))[:n - filled]
            out[filled:filled + len(good), col:col + self.space.dim] = good[:, col:col + self.space.dim]
            filled += len(good)
            use_b = 1 - use_b
            rounds += 1
            if rounds > 4000:
                raise RuntimeError("Rejection sampling on the boundary of a boolean domain does not find enough points.")

    def __call__(self, **data):
        return self

    def _box(self):
        return self.domain._box()

    def dim_cols(self):
        return self.space.dim

    def _program(self, col):
        a, b = self.domain.domain_a, self.domain.domain_b
        S, AND, OR, NOT = _lib.TPX_OP_SHAPE, _lib.TPX_OP_AND, _lib.TPX_OP_OR, _lib.TPX_OP_NOT
        shapes = [a.boundary._shape(col), b._shape(col), b.boundary._shape(col), a._shape(col)]
        # (on dA and not in B) or (on dB and in A and not on dA)
        ops = [S | (0 << 8), S | (1 << 8), NOT, AND, S | (2 << 8), S | (3 << 8), AND, S | (0 << 8), NOT, AND, OR]
        return shapes, ops

    def _get_volume(self):
        if not self.domain.contained:
            warnings.warn("Exact volume of this domain is not known, will use the estimate: volume = "
                          "domain_a.volume + domain_b.volume. If you need the exact volume for sampling, use "
                          "domain.set_volume().")
        return _F32(self.domain.domain_a.boundary._volume_value()) + _F32(self.domain.domain_b.boundary._volume_value())

    def _variable_count_with_density(self):
        return True

    def _fill_random(self, out, col, n):
        if n == 0:
            return
        if not self._basic:
            return self._fill_random_composed(out, col, n)
        a, b = self.domain.domain_a.boundary, self.domain.domain_b.boundary
        vm = _F32(self._volume_value())
        scaled = [int(_F32(n) * _F32(a._volume_value()) / vm) + 1, int(_F32(n) * _F32(b._volume_value()) / vm) + 1]
        rg = _region_struct(*self._program(col))
        lib = _lib.lib()
        key, offset = _rng.next_key(), 0
        counters = torch.zeros(4, dtype=torch.int64, device=out.device)
        filled, use_b, rounds = 0, 0, 0
        while filled < n:
            m = scaled[use_b]
            nbytes = ctypes.c_size_t(0)
            _lib.check(lib.tpx_sample_reject_workspace_bytes(ctypes.c_int64(m), ctypes.byref(nbytes)))
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=out.device)
            gen = (b if use_b else a)._shape(col)
            _lib.check(lib.tpx_sample_uniform_reject(
                ctypes.byref(gen), ctypes.byref(rg), ctypes.c_uint64(key), ctypes.c_uint64(offset), ctypes.c_int64(m),
                ctypes.c_int64(n), _lib.ptr(out), ctypes.c_int32(out.shape[1]), _lib.ptr(counters), _lib.ptr(ws),
                ctypes.c_size_t(ws.numel()), _lib.stream_ptr()))
            offset += m
            filled = int(counters[0])
            use_b = 1 - use_b
            rounds += 1
            if rounds > 4000:
                raise RuntimeError("Rejection sampling on the cut boundary does not find enough points.")

    def _masked(self, pts, keep):
        return pts[keep.reshape(-1)]

    def _parts_with_d(self, sample, d, params, device):
        """density mode (cut.py:146-163 / 176-189): both boundaries at density d, A's part outside B, B's part
        strictly inside A; the count varies."""
        dev = _compute_device(device)
        a, b = self.domain.domain_a, self.domain.domain_b
        pa = sample(a.boundary, d, dev).as_tensor
        pa = self._masked(pa, ~b._contains_rows(pa, 0, dev))
        pb = sample(b.boundary, d, dev).as_tensor
        pb = self._masked(pb, a._contains_rows(pb, 0, dev) & ~a.boundary._contains_rows(pb, 0, dev))
        return Points(torch.cat([pa, pb], dim=0).to(device), self.space)

    def _sample_random_with_d(self, d, params, device):
        return self._parts_with_d(lambda dom, dd, dev: dom.sample_random_uniform(d=dd, device=dev), d, params, device)

    def _sample_grid_with_d(self, d, params, device):
        return self._parts_with_d(lambda dom, dd, dev: dom.sample_grid(d=dd, device=dev), d, params, device)

    def _grid(self, n, dev, exact=True):
        """sampler_helper.py:211-256: grids on both boundaries, kept where they lie on the cut's boundary; one
        rescaling by the estimated boundary length; random points if still short."""
        a, b = self.domain.domain_a.boundary, self.domain.domain_b.boundary

        def on_main(dom, m):
            g = dom._grid(m, dev)
            return self._masked(g, self._contains_rows(g, 0, dev))
        ka, kb = on_main(a, n), on_main(b, n)
        if len(ka) + len(kb) == n:
            return torch.cat([ka, kb], dim=0)
        sa, sb = _F32(a._volume_value()), _F32(b._volume_value())
        approx = sa * _F32(len(ka)) / _F32(n) + sb * _F32(len(kb)) / _F32(n)
        na, nb = int(_F32(n) * sa / approx) + 1, max(int(_F32(n) * sb / approx), 1)
        g = torch.cat([on_main(a, na), on_main(b, nb)], dim=0)
        if len(g) >= n:
            return g[:n]
        extra = torch.empty((n - len(g), self.space.dim), dtype=torch.float32, device=dev)
        self._fill_random(extra, 0, n - len(g))
        return torch.cat([g, extra], dim=0)

    def normal(self, points, params=Points.empty(), device="cpu"):
        if not isinstance(points, Points):
            points = Points(points, self.space)
        a, b = self.domain.domain_a.boundary, self.domain.domain_b.boundary
        on_a = a._contains(points)
        return torch.where(on_a, a.normal(points), -b.normal(points))


CutDomain.boundary = property(lambda self: CutBoundaryDomain(self))


class UnionBoundaryDomain(CutBoundaryDomain):
    """Boundary of ``A + B`` for two basic shapes (reference ``domainoperations/union.py:153-266``): the part of
    each boundary outside the other domain; a piece that lies on BOTH boundaries counts only where the outward
    normals agree (inner product >= 0.5), which a region program cannot express -- membership composes the
    shape / normal kernels on the device, random points alternate between the two boundaries with that mask."""

    overlap_tol = 0.5

    def _program(self, col):
        raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "the boundary of a union is not a region program")

    def _contains_rows(self, rows, col, out_device):
        a, b = self.domain.domain_a, self.domain.domain_b
        dev = _compute_device(rows.device)
        rows = rows.to(device=dev, dtype=torch.float32).contiguous()
        in_a, in_b = a._contains_rows(rows, col, dev), b._contains_rows(rows, col, dev)
        on_a, on_b = a.boundary._contains_rows(rows, col, dev), b.boundary._contains_rows(rows, col, dev)
        on_both = on_a & on_b
        pts = rows[:, col:col + self.space.dim]
        agree = (a.boundary.normal(pts) * b.boundary.normal(pts)).sum(dim=-1, keepdim=True) >= self.overlap_tol
        ok = torch.where(on_both, agree, torch.ones_like(on_both))
        return (((on_a & ~in_b) | (on_b & ~in_a) | on_both) & ok).to(out_device)

    def _get_volume(self):
        if not self.domain.disjoint:
            warnings.warn("Exact volume of this domain is not known, will use the estimate: volume = "
                          "domain_a.volume + domain_b.volume. If you need the exact volume for sampling, use "
                          "domain.set_volume().")
        return _F32(self.domain.domain_a.boundary._volume_value()) + _F32(self.domain.domain_b.boundary._volume_value())

    def _fill_random(self, out, col, n):
        if n > 0:
            self._fill_random_composed(out, col, n)

    def _parts_with_d(self, sample, d, params, device):
        """density mode (union.py:215-224 / 250-259): A's boundary outside B, B's boundary not strictly inside A."""
        dev = _compute_device(device)
        a, b = self.domain.domain_a, self.domain.domain_b
        pa = sample(a.boundary, d, dev).as_tensor
        pa = self._masked(pa, ~b._contains_rows(pa, 0, dev))
        pb = sample(b.boundary, d, dev).as_tensor
        pb = self._masked(pb, a.boundary._contains_rows(pb, 0, dev) | ~a._contains_rows(pb, 0, dev))
        return Points(torch.cat([pa, pb], dim=0).to(device), self.space)

    def normal(self, points, params=Points.empty(), device="cpu"):
        if not isinstance(points, Points):
            points = Points(points, self.space)
        a, b = self.domain.domain_a.boundary, self.domain.domain_b.boundary
        return torch.where(a._contains(points), a.normal(points), b.normal(points))


class IntersectionBoundaryDomain(CutBoundaryDomain):
    """Boundary of ``A & B`` for two basic shapes (reference ``domainoperations/intersection.py:112-206``): the part of A's
    boundary inside B plus the part of B's boundary inside A.  Same machinery as the boundary of a cut: one region program
    over the boundary shapes, random points alternating between the two boundaries with rejection by that program."""

    def _program(self, col):
        a, b = self.domain.domain_a, self.domain.domain_b
        S, AND, OR = _lib.TPX_OP_SHAPE, _lib.TPX_OP_AND, _lib.TPX_OP_OR
        shapes = [a.boundary._shape(col), b._shape(col), b.boundary._shape(col), a._shape(col)]
        # (on dA and in B) or (on dB and in A)
        ops = [S | (0 << 8), S | (1 << 8), AND, S | (2 << 8), S | (3 << 8), AND, OR]
        return shapes, ops

    def _mask(self, in_a, in_b, on_a, on_b):
        return (on_a & in_b) | (on_b & in_a)

    def _get_volume(self):
        warnings.warn("Exact volume of this intersection-boundary is not known, will use the estimate: volume = boundary_a + "
                      "boundary_b. If you need the exact volume for sampling, use domain.set_volume()")
        return _F32(self.domain.domain_a.boundary._volume_value()) + _F32(self.domain.domain_b.boundary._volume_value())

    def _parts_with_d(self, sample, d, params, device):
        """density mode (intersection.py:150-163 / 185-196): both boundaries at density d, each kept inside the other domain"""
        dev = _compute_device(device)
        a, b = self.domain.domain_a, self.domain.domain_b
        pa = sample(a.boundary, d, dev).as_tensor
        pa = self._masked(pa, b._contains_rows(pa, 0, dev))
        pb = sample(b.boundary, d, dev).as_tensor
        pb = self._masked(pb, a._contains_rows(pb, 0, dev))
        return Points(torch.cat([pa, pb], dim=0).to(device), self.space)

    def normal(self, points, params=Points.empty(), device="cpu"):
        if not isinstance(points, Points):
            points = Points(points, self.space)
        a, b = self.domain.domain_a.boundary, self.domain.domain_b.boundary
        on_a = a._contains(points)
        return torch.where(on_a, a.normal(points), b.normal(points))


class IntersectionDomain(_BooleanDomain):
    """domain_a and domain_b (reference ``domainoperations/intersection.py:14-109``)."""

    def _program(self, col):
        return _combine(self.domain_a._program(col), self.domain_b._program(col), [_lib.TPX_OP_AND])

    def _generator(self, col):
        gen = self.domain_a._generator(col)
        if gen is None:
            return None
        shape, shapes, ops = gen
        sb, ob = self.domain_b._program(col)
        if ops:
            shapes, ops = _combine((shapes, ops), (sb, ob), [_lib.TPX_OP_AND])
        else:
            shapes, ops = sb, ob
        return shape, shapes, ops

    def _get_volume(self):
        warnings.warn("Exact volume of this intersection is not known, will use the estimate: volume = "
                      "domain_a.volume. If you need the exact volume for sampling, use domain.set_volume()")
        return self.domain_a._volume_value()

    def _box(self):
        a, b = self.domain_a._box(), self.domain_b._box()
        out = []
        for i in range(self.space.dim):
            out += [max(a[2 * i], b[2 * i]), min(a[2 * i + 1], b[2 * i + 1])]
        return out

    def _sample_random_with_d(self, d, params, device):
        pts = self.domain_a.sample_random_uniform(d=d, params=params, device=_compute_device(device))
        t = pts.as_tensor
        return Points(self._select(t, self.domain_b._contains_rows(t, 0, t.device)).to(device), self.space)

    def _grid(self, n, dev, exact=True):
        a = self.domain_a._grid(n, dev)
        keep = self.domain_b._contains_rows(a, 0, dev)
        inside = int(keep.sum())
        if inside == n:
            return a
        scaled = int(n ** 2 / max(inside, 1))
        a = self.domain_a._grid(scaled, dev)
        a = self._select(a, self.domain_b._contains_rows(a, 0, dev))
        if len(a) >= n:
            return a[:n]
        extra = torch.empty((n - len(a), self.space.dim), dtype=torch.float32, device=dev)
        self._fill_random(extra, 0, n - len(a))
        return torch.cat([a, extra], dim=0)

    def _sample_grid_with_d(self, d, params, device):
        pts = self.domain_a.sample_grid(d=d, params=params, device=_compute_device(device))
        t = pts.as_tensor
        return Points(self._select(t, self.domain_b._contains_rows(t, 0, t.device)).to(device), self.space)


class UnionDomain(_BooleanDomain):
    """domain_a or domain_b (reference ``domainoperations/union.py:9-150``)."""

    def __init__(self, domain_a, domain_b, disjoint=False):
        super().__init__(domain_a, domain_b)
        self.disjoint = disjoint

    def _program(self, col):
        return _combine(self.domain_a._program(col), self.domain_b._program(col), [_lib.TPX_OP_OR])

    def _get_volume(self):
        if not self.disjoint:
            warnings.warn("Exact volume of this union is not known, will use the estimate: volume = "
                          "domain_a.volume + domain_b.volume. If you need the exact volume for sampling, "
                          "use domain.set_volume()")
        return self.domain_a._volume_value() + self.domain_b._volume_value()

    def _box(self):
        a, b = self.domain_a._box(), self.domain_b._box()
        out = []
        for i in range(self.space.dim):
            out += [min(a[2 * i], b[2 * i]), max(a[2 * i + 1], b[2 * i + 1])]
        return out

    def _fill_random(self, out, col, n):
        # union.py:72-93: n points in A, n points in B, keep A's unless B's point is outside A
        # and a coin with P(A) = vol_A / (vol_A + vol_B) says B
        stride = out.shape[1]
        a = torch.zeros((n, stride), dtype=torch.float32, device=out.device)
        b = torch.zeros((n, stride), dtype=torch.float32, device=out.device)
        self.domain_a._fill_random(a, col, n)
        self.domain_b._fill_random(b, col, n)
        va, vb = self.domain_a._volume_value(), self.domain_b._volume_value()
        ratio = float(_F32(va) / (_F32(va) + _F32(vb)))
        sa, oa = self.domain_a._program(col)
        rg = _region_struct(sa, oa)
        # pick writes whole rows of width stride; only our columns differ between a and b
        tmp = torch.empty_like(a)
        _lib.check(_lib.lib().tpx_union_pick(ctypes.byref(rg), _lib.ptr(a), _lib.ptr(b), ctypes.c_int64(n),
                                             ctypes.c_int32(stride), ctypes.c_int32(stride), ctypes.c_float(ratio),
                                             ctypes.c_uint64(_rng.next_key()), ctypes.c_uint64(0), _lib.ptr(tmp),
                                             _lib.stream_ptr()))
        out[:n, col:col + self.space.dim] = tmp[:, col:col + self.space.dim]

    def _append(self, ta, tb):
        keep = ~self.domain_a._contains_rows(tb, 0, tb.device)
        return torch.cat([ta, self._select(tb, keep)], dim=0)

    def _sample_random_with_d(self, d, params, device):
        dev = _compute_device(device)
        ta = self.domain_a.sample_random_uniform(d=d, params=params, device=dev).as_tensor
        tb = self.domain_b.sample_random_uniform(d=d, params=params, device=dev).as_tensor
        return Points(self._append(ta, tb).to(device), self.space)

    def _grid(self, n, dev, exact=True):
        # union.py:115-133
        va, vb = self.domain_a._volume_value(), self.domain_b._volume_value()
        scaled = int(math.ceil(_F32(n) * _F32(va) / (_F32(va) + _F32(vb))))
        a = self.domain_a._grid(scaled, dev)
        if n - scaled <= 0:
            return a
        a = self._select(a, ~self.domain_b._contains_rows(a, 0, dev))
        b = self.domain_b._grid(n - len(a), dev)
        return torch.cat([a, b], dim=0)

    def _sample_grid_with_d(self, d, params, device):
        dev = _compute_device(device)
        ta = self.domain_a.sample_grid(d=d, params=params, device=dev).as_tensor
        tb = self.domain_b.sample_grid(d=d, params=params, device=dev).as_tensor
        return Points(self._append(ta, tb).to(device), self.space)


class ProductDomain(Domain):
    """Cartesian product, columns of domain_a first (reference
    ``domainoperations/product.py:14-267``, constant case)."""

    def __init__(self, domain_a, domain_b):
        if not domain_a.space.keys().isdisjoint(domain_b.space):
            warnings.warn("Warning: The space of a ProductDomain will be the product of its factor domains "
                          "spaces. This may lead to unexpected behaviour.")
        super().__init__(space=domain_a.space * domain_b.space, dim=domain_a.dim + domain_b.dim)
        self.domain_a, self.domain_b = domain_a, domain_b
        self.bounds = None

    def _program(self, col):
        return _combine(self.domain_a._program(col), self.domain_b._program(col + self.domain_a.dim_cols()),
                        [_lib.TPX_OP_AND])

    def _get_volume(self):
        return _F32(self.domain_a._volume_value()) * _F32(self.domain_b._volume_value())

    def set_bounding_box(self, bounds):
        assert len(bounds) == 2 * self.dim, "Bounds dont fit the dimension."
        self.bounds = list(bounds)

    def _box(self):
        return self.bounds if self.bounds else self.domain_a._box() + self.domain_b._box()

    def _fill_random(self, out, col, n):
        self.domain_b._fill_random(out, col + self.domain_a.dim_cols(), n)
        self.domain_a._fill_random(out, col, n)

    def sample_random_uniform(self, n=None, d=None, params=Points.empty(), device="cpu"):
        if n is None:
            assert d is not None
            n = int(_F32(d) * _F32(self._volume_value()))
        return super().sample_random_uniform(n=n, params=params, device=device)

    def sample_grid(self, n=None, d=None, params=Points.empty(), device="cpu"):
        raise NotImplementedError("Grid sampling on a product domain is not implmented. Use a product "
                                  "sampler instead.")


UnionDomain.boundary = property(lambda self: UnionBoundaryDomain(self))
IntersectionDomain.boundary = property(lambda self: IntersectionBoundaryDomain(self))
# product.py:100-106: (dA x B) + (A x dB), no normals
ProductDomain.boundary = property(lambda self: UnionDomain(ProductDomain(self.domain_a.boundary, self.domain_b),
                                                           ProductDomain(self.domain_a, self.domain_b.boundary)))


# function sets / function samplers live in the same namespaces as in the reference
from .functionsets import FunctionSet, CustomFunctionSet, FunctionSetCollection  # noqa: E402,F401
