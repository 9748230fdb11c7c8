"""Named variable spaces and ``Points`` (a tensor whose last axis is labelled by a space).

Host-side glue with the semantics of the reference's ``problem/spaces/space.py`` and
``problem/spaces/points.py`` (pinned by the reference tests ``tests/test_spaces.py`` and
``tests/test_points.py``).  Pure bookkeeping -- no kernel is involved -- but it defines
which input column is which variable, and ``track_coord_gradients`` (reference
``points.py:380-384``) creates the coordinate leaves the differential operators are
called with; the native jet path recognises those leaves by identity
(``torchphysics_b200/jets.py``).
"""
import math
from collections import Counter, OrderedDict

import numpy as np
import torch


class Space(Counter, OrderedDict):
    """Ordered map variable name -> dimension (reference ``space.py:5-133``)."""

    def __init__(self, variables_dims):
        super().__init__(variables_dims)

    # -- algebra ---------------------------------------------------------------------
    def __mul__(self, other):
        assert isinstance(other, Space)
        return Space(self + other)

    def __contains__(self, item):
        if isinstance(item, str):
            return super().__contains__(item)
        if isinstance(item, Space):
            return (self & item) == item
        return False

    def __getitem__(self, key):
        if isinstance(key, slice):
            names = list(self.keys())
            lo = None if key.start is None else names.index(key.start)
            hi = None if key.stop is None else names.index(key.stop)
            return Space({k: self[k] for k in names[slice(lo, hi, key.step)]})
        if isinstance(key, (list, tuple)):
            return Space({k: self[k] for k in key})
        return super().__getitem__(key)

    def __eq__(self, other):
        return OrderedDict.__eq__(self, other)      # order sensitive

    def __ne__(self, other):
        return OrderedDict.__ne__(self, other)

    def __repr__(self):
        return f"{type(self).__name__}({dict(OrderedDict(self))!r})"

    def __reduce__(self):
        return self.__class__, (OrderedDict(self),)

    # -- queries ---------------------------------------------------------------------
    @property
    def dim(self):
        return sum(self.values())

    @property
    def variables(self):
        return set(self.keys())

    def cast_tensor_into_space(self, tensor):
        return tensor

    def check_values_in_space(self, values):
        assert values.shape[-1] == self.dim
        return values

    def column_slices(self):
        out, start = {}, 0
        for name in self:
            out[name] = slice(start, start + self[name])
            start += self[name]
        return out


class R1(Space):
    def __init__(self, variable_name):
        super().__init__({variable_name: 1})


class R2(Space):
    def __init__(self, variable_name):
        super().__init__({variable_name: 2})


class R3(Space):
    def __init__(self, variable_name):
        super().__init__({variable_name: 3})


class Rn(Space):
    def __init__(self, variable_name, n):
        super().__init__({variable_name: int(n)})


class IntegerSpace(Space):
    """values are rounded to integers, the dtype stays (reference ``space.py:189-220``)."""

    def __init__(self, variable_name, dim):
        super().__init__({variable_name: int(dim)})

    def cast_tensor_into_space(self, tensor):
        return torch.round(tensor)


class Z1(IntegerSpace):
    def __init__(self, variable_name):
        super().__init__(variable_name, 1)


class Z2(IntegerSpace):
    def __init__(self, variable_name):
        super().__init__(variable_name, 2)


class Z3(IntegerSpace):
    def __init__(self, variable_name):
        super().__init__(variable_name, 3)


class Zn(IntegerSpace):
    def __init__(self, variable_name, n):
        super().__init__(variable_name, n)


class NaturalNumberSpace(Space):
    """rounded absolute values (reference ``space.py:223-251``)."""

    def __init__(self, variable_name, dim):
        super().__init__({variable_name: int(dim)})

    def cast_tensor_into_space(self, tensor):
        return torch.round(tensor).abs()


class N1(NaturalNumberSpace):
    def __init__(self, variable_name):
        super().__init__(variable_name, 1)


class N2(NaturalNumberSpace):
    def __init__(self, variable_name):
        super().__init__(variable_name, 2)


class N3(NaturalNumberSpace):
    def __init__(self, variable_name):
        super().__init__(variable_name, 3)


class Nn(NaturalNumberSpace):
    def __init__(self, variable_name, n):
        super().__init__(variable_name, n)


class FunctionSpace:
    """Space of functions input_space -> output_space (reference ``functionspace.py:1-35``)."""

    def __init__(self, input_domain, output_space):
        self.input_domain = input_domain
        self.input_space = input_domain.space if hasattr(input_domain, "space") else input_domain
        self.output_space = output_space

    def __mul__(self, other):
        """product: outputs side by side; the input space is the common one, the larger one, or the product."""
        assert isinstance(other, FunctionSpace), "Can only multiply FunctionSpaces"
        assert not (self.output_space == other.output_space), "Output spaces must be different"
        out = self.output_space * other.output_space
        if self.input_space == other.input_space:
            return FunctionSpace(self.input_space, out)
        if self.input_space in other.input_space:
            return FunctionSpace(other.input_space, out)
        if other.input_space in self.input_space:
            return FunctionSpace(self.input_space, out)
        return FunctionSpace(self.input_space * other.input_space, out)


class Points:
    """``data`` (..., space.dim) + the space that names its columns."""

    def __init__(self, data, space, **kwargs):
        self.space = space
        self._t = space.cast_tensor_into_space(torch.as_tensor(data, **kwargs))
        assert self._t.dim() >= 2
        assert self._t.shape[-1] == space.dim, (
            "Data dimension does not fit dimension of the space " + str(list(space.keys())))

    # -- constructors ----------------------------------------------------------------
    @classmethod
    def empty(cls, **kwargs):
        return cls(torch.empty(0, 0, **kwargs), Space({}))

    @classmethod
    def joined(cls, *parts):
        tensors, space, shape = [], Space({}), parts[0].shape
        for pts in parts:
            if pts.isempty:
                continue
            assert space.keys().isdisjoint(pts.space)
            assert pts.shape == shape
            tensors.append(pts._t)
            space = space * pts.space
        return cls(torch.cat(tensors, dim=-1), space)

    @classmethod
    def from_coordinates(cls, coords):
        if coords == {}:
            return cls.empty()
        names = list(coords.keys())
        lead = coords[names[0]].shape[:-1]
        cols, dims = [], {}
        for name in names:
            coords[name] = torch.as_tensor(coords[name])
            assert coords[name].shape[:-1] == lead
            cols.append(coords[name])
            dims[name] = coords[name].shape[-1]
        return cls(torch.cat(cols, dim=-1), Space(dims))

    # -- views -----------------------------------------------------------------------
    @property
    def dim(self):
        return self.space.dim

    @property
    def variables(self):
        return self.space.variables

    @property
    def _variable_slices(self):
        return self.space.column_slices()

    @property
    def coordinates(self):
        sl = self.space.column_slices()
        width = self._t.shape[-1]
        # a variable that spans every column of a tensor with a graph (a model output) IS that tensor: no slice node, whose
        # backward would be a zero fill + copy of the batch.  Tensors without a graph keep the reference's fresh views
        # (``points.py:158-165``): callers may mark those as leaves.
        whole = self._t.requires_grad
        return {name: (self._t if whole and (sl[name].start, sl[name].stop) == (0, width) else self._t[..., sl[name]])
                for name in self.space}

    @property
    def as_tensor(self):
        return self._t

    @property
    def device(self):
        return self._t.device

    @property
    def shape(self):
        return self._t.shape[:-1]

    @property
    def isempty(self):
        return len(self) == 0 and self.space.dim == 0

    def __len__(self):
        return math.prod(self._t.shape[:-1])

    def __repr__(self):
        return f"{type(self).__name__}:\n{self.coordinates}"

    # -- indexing (points[1:3, ('x','t')], points[mask,], points[..., 'x']) ------------
    def _resolve_index(self, idx):
        if isinstance(idx, tuple):
            idx = list(idx)
        if isinstance(idx, (np.ndarray, torch.Tensor)) and idx.dtype in (bool, torch.bool):
            if idx.ndim == self._t.dim():
                raise IndexError("Boolean slicing in last dimension is not supported.")
        space = self.space
        if isinstance(idx, list):
            has_ellipsis = any(v is Ellipsis for v in idx)
            addresses_last = (len(idx) == self._t.dim()) or (has_ellipsis and idx[-1] is not Ellipsis)
            if addresses_last:
                sl = self.space.column_slices()
                last = idx[-1]
                if isinstance(last, str):
                    space = Space({last: self.space[last]})
                    idx[-1] = sl[last]
                else:
                    space = self.space[last]
                    cols = []
                    all_cols = list(range(self.dim))
                    for name in space:
                        cols += all_cols[sl[name]]
                    # a contiguous range of columns is a slice (a view whose backward is a slice too); a list is advanced
                    # indexing, and its backward an index_put with accumulation: 1 ms of a 2.4 ms PI-DeepONet step for the
                    # (256, 16384, 1) trunk points that FunctionSet._transform_locations selects by variable name
                    if cols and cols == list(range(cols[0], cols[0] + len(cols))):
                        idx[-1] = slice(cols[0], cols[0] + len(cols))
                    else:
                        idx[-1] = cols
            idx = tuple(idx)
        return idx, space

    def __getitem__(self, idx):
        if isinstance(idx, Space):
            if idx.variables == self.space.variables:
                return self
            return self[:, self.space.variables]
        idx, space = self._resolve_index(idx)
        out = self._t[idx]
        if out.dim() == 1:
            out = out.unsqueeze(0)
        return Points(out, space)

    def __setitem__(self, idx, points):
        idx, space = self._resolve_index(idx)
        assert space == points.space
        self._t[idx] = points._t

    def __iter__(self):
        for i in range(self._t.shape[0]):
            yield self[i]

    # -- arithmetic ------------------------------------------------------------------
    def _binary(self, other, op):
        assert isinstance(other, Points)
        assert other.space == self.space
        return Points(op(self._t, other._t), self.space)

    def __eq__(self, other):
        return self.space == other.space and torch.equal(self._t, other._t)

    def __add__(self, other):
        return self._binary(other, torch.add)

    def __sub__(self, other):
        return self._binary(other, torch.sub)

    def __mul__(self, other):
        return self._binary(other, torch.mul)

    def __pow__(self, other):
        return self._binary(other, torch.pow)

    def __truediv__(self, other):
        return self._binary(other, torch.div)

    def __or__(self, other):
        """row-wise concatenation of points of the same space."""
        assert isinstance(other, Points)
        if self.isempty:
            return other
        if other.isempty:
            return self
        assert other.space == self.space
        return Points(torch.cat([self._t, other._t], dim=0), self.space)

    def join(self, other):
        """column-wise concatenation: product space."""
        assert isinstance(other, Points)
        if self.isempty:
            return other
        if other.isempty:
            return self
        assert self.space.keys().isdisjoint(other.space)
        return Points(torch.cat([self._t, other._t], dim=-1), self.space * other.space)

    def repeat(self, *n):
        pad = (self._t.dim() - len(n)) * [1]
        return Points(self._t.repeat(*n, *pad), self.space)

    def unsqueeze(self, dim):
        assert dim < self._t.dim()
        if dim < 0:
            dim -= 1
        return Points(self._t.unsqueeze(dim), self.space)

    # -- torch interop ---------------------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = {} if kwargs is None else kwargs
        assert any(hasattr(a, "space") for a in args)
        return func(*[a._t if hasattr(a, "_t") else a for a in args], **kwargs)

    @property
    def requires_grad(self):
        return self._t.requires_grad

    @requires_grad.setter
    def requires_grad(self, value):
        self._t.requires_grad = value

    def cuda(self, *args, **kwargs):
        self._t = self._t.cuda(*args, **kwargs)
        return self

    def to(self, *args, **kwargs):
        self._t = self._t.to(*args, **kwargs)
        return self

    def track_coord_gradients(self):
        """Per-variable column views become leaves requiring grad; the returned Points is
        their concatenation (reference ``points.py:380-384``).  The new Points remembers the
        leaves so that the native jet path can identify derivative variables."""
        coords = self.coordinates
        for name in coords:
            coords[name].requires_grad = True
        tracked = Points.from_coordinates(coords)
        tracked._coord_leaves = coords
        return coords, tracked
