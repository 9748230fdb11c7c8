"""Bridges the tensor-in / tensor-out operator API of the reference onto the native jets.

The reference's operators receive *tensors* (``laplacian(u, x)``) and rely on the autograd
graph between them (reference ``utils/differentialoperators/differentialoperators.py``).
Here a model evaluated on tracked points returns a ``JetTensor``: an ordinary tensor that
remembers (a) the ``JetEvaluation`` that produced it and (b) which output columns it
holds.  The operators in ``operators.py`` look the requested derivatives up in the
evaluation's streams when the derivative variables are exactly the coordinate leaves
created by ``Points.track_coord_gradients``; every other use of the tensor behaves like
a plain tensor.

Which streams a model computes is learned per call site: the first call computes the
value only, an operator that needs more re-evaluates with the larger ``JetSpec`` and
records it as the hint for the next call from the same condition -- from the second
training step on, one condition = one forward launch + one reverse launch.
"""
import threading
import warnings

import torch

from .engine import JetSpec

_tls = threading.local()


def current_scope():
    return getattr(_tls, "scope", None)


class scope:
    """``with scope(key):`` marks which condition the enclosed model calls belong to."""

    def __init__(self, key):
        self.key = key

    def __enter__(self):
        self.prev = current_scope()
        _tls.scope = self.key
        return self

    def __exit__(self, *exc):
        _tls.scope = self.prev
        return False


class TpxFallbackWarning(UserWarning):
    pass


_warned = set()


def warn_fallback(reason):
    if reason not in _warned:
        _warned.add(reason)
        warnings.warn(f"native jet path not used: {reason}; evaluating with torch.autograd", TpxFallbackWarning)


class JetEvaluation:
    """One evaluation of a natively supported network on a batch of tracked points."""

    def __init__(self, engine, x2d, batch_shape, leaf_cols, hints, hint_key, autograd_fn, spec, params=None):
        self.engine = engine
        self.x = x2d                        # (N, d_in) detached, contiguous
        self.batch_shape = tuple(batch_shape)
        self.leaf_cols = leaf_cols          # id(leaf) -> (leaf, [input columns])
        self.hints, self.hint_key = hints, hint_key
        self.autograd_fn = autograd_fn      # () -> plain autograd output (generic path)
        self._plain = None
        self.spec = spec
        self.params = params                # None: the engine's own module parameters
        self.u, self.du, self.lap = engine.jet(x2d, spec, params)

    def columns_of(self, var):
        hit = self.leaf_cols.get(id(var))
        if hit is None or hit[0] is not var:
            return None
        return hit[1]

    def ensure(self, need):
        """make sure the computed streams cover ``need``; returns False if impossible."""
        if self.spec.covers(need):
            return True
        cols = list(self.spec.dcols)
        for c in need.dcols:
            if c not in cols:
                cols.append(c)
        cols.sort()
        lap_w = None
        if need.lap and self.spec.lap:
            # two different Laplacians requested from one evaluation: not one stream
            return False
        src = need if need.lap else self.spec
        if src.lap:
            wmap = dict(zip(src.dcols, src.lap_w))
            lap_w = [wmap.get(c, 0.0) for c in cols]
        if len(cols) > 3:
            return False
        new = JetSpec(cols, lap_w)
        u, du, lap = self.engine.jet(self.x, new, self.params)
        self.spec, self.du, self.lap = new, du, lap      # value stream: keep the first u
        self._u2 = u
        self.hints[self.hint_key] = new
        return True

    # streams reshaped back to the batch shape of the points
    def value(self):
        return self.u.reshape(*self.batch_shape, self.engine.d_out)

    @staticmethod
    def _columns(t, cols, axis):
        """t[..., cols] along ``axis`` without advanced indexing when the columns are a contiguous range
        (a view whose backward is a cheap slice instead of an index_put)."""
        cols = list(cols)
        if cols == list(range(t.shape[axis])):
            return t
        if cols == list(range(cols[0], cols[0] + len(cols))):
            return t.narrow(axis, cols[0], len(cols))
        return t.index_select(axis, torch.as_tensor(cols, device=t.device))

    def first(self, out_cols, in_col):
        k = self.spec.dcols.index(in_col)
        d = self.du.select(2, k)
        return self._columns(d, out_cols, 1).reshape(*self.batch_shape, len(out_cols))

    def second(self, out_cols):
        return self._columns(self.lap, out_cols, 1).reshape(*self.batch_shape, len(out_cols))

    def plain(self):
        """the same output through torch ops with a graph to the coordinate leaves."""
        if self._plain is None:
            self._plain = self.autograd_fn()
        return self._plain


class JoinedEvaluation:
    """Column-wise join of several native evaluations on the same tracked points (``tp.models.Parallel``,
    reference ``models/model.py:95-146``): output column c belongs to ``parts[i]``'s column ``c - offset_i``."""

    def __init__(self, parts):
        self.parts = parts                       # [(evaluation, [its columns])]
        self.index = [(ev, c) for ev, cols in parts for c in cols]
        self.batch_shape = parts[0][0].batch_shape

    def columns_of(self, var):
        for ev, _ in self.parts:
            cols = ev.columns_of(var)
            if cols is not None:
                return cols
        return None

    def ensure(self, need):
        return all(ev.ensure(need) for ev, _ in self.parts)

    def _gather(self, out_cols, fn):
        pieces, run_ev, run_cols = [], None, []
        for c in out_cols:                       # consecutive columns of one part -> one look-up
            ev, sub = self.index[c]
            if ev is not run_ev and run_cols:
                pieces.append(fn(run_ev, run_cols))
                run_cols = []
            run_ev = ev
            run_cols.append(sub)
        if run_cols:
            pieces.append(fn(run_ev, run_cols))
        return pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=-1)

    def value(self):
        return self._gather(range(len(self.index)), lambda ev, cols: JetEvaluation._columns(ev.value(), cols, ev.value().dim() - 1))

    def first(self, out_cols, in_col):
        return self._gather(out_cols, lambda ev, cols: ev.first(cols, in_col))

    def second(self, out_cols):
        return self._gather(out_cols, lambda ev, cols: ev.second(cols))

    def plain(self):
        return torch.cat([JetEvaluation._columns(ev.plain(), cols, ev.plain().dim() - 1) for ev, cols in self.parts], dim=-1)


class ConstrainedEvaluation:
    """Streams of ``g(x, y(x))`` for a pointwise output transform g of a native evaluation y (the reference's
    ``HardConstraint``, ``models/model.py:186-217``) -- same protocol as ``JetEvaluation``.

    Chain rule to second order along a coordinate direction e_k, with the partials of g taken by torch.autograd
    on the pointwise function (the network is never re-traversed):
        d_k out   = phi_k'(0),   phi_k(eps) = g(x + eps e_k, y + eps d_k y)
        d_kk out  = phi_k''(0) + g_y . d_kk y
    so the weighted Laplacian is  sum_k c_k phi_k''(0) + psi'(0),  psi(eps) = g(x, y + eps L y).
    Every result keeps its graph to the streams of y, hence to the parameters."""

    def __init__(self, inner, constraint, coords, in_space, out_space, inner_space):
        self.inner = inner                       # JetEvaluation of the wrapped model
        self.constraint = constraint             # UserFunction(coordinates + model outputs) -> tensor / tuple
        self.coords = coords                     # {variable: leaf tensor} of the tracked points
        self.in_space, self.out_space, self.inner_space = in_space, out_space, inner_space
        self.batch_shape = inner.batch_shape
        self._cache = {}

    # -- helpers -------------------------------------------------------------------------
    def _apply(self, x_shift, y):
        """g on coordinates shifted by ``x_shift`` (dict column -> eps tensor) and model output ``y``."""
        sl = self.in_space.column_slices()
        args = {}
        for name, leaf in self.coords.items():
            v = leaf.detach()
            if name in sl and x_shift:
                cols = range(sl[name].start, sl[name].stop)
                if any(c in x_shift for c in cols):
                    v = torch.cat([v[..., i:i + 1] + x_shift[c] if c in x_shift else v[..., i:i + 1]
                                   for i, c in enumerate(cols)], dim=-1)
            args[name] = v
        osl = self.inner_space.column_slices()
        for name in self.inner_space:
            args[name] = y[..., osl[name]]
        res = self.constraint(args)
        if isinstance(res, (tuple, list)):
            res = torch.column_stack(res)
        return res

    def _path(self, k_col, dy, order):
        """(phi'(0), phi''(0)) of phi(eps) = g(x + eps e_k, y + eps dy); k_col None keeps x fixed."""
        y = self.inner.value()
        eps = torch.zeros((*self.batch_shape, 1), device=y.device, dtype=y.dtype, requires_grad=True)
        out = self._apply({} if k_col is None else {k_col: eps}, y + eps * dy)
        d1 = torch.stack([torch.autograd.grad(out[..., m].sum(), eps, create_graph=True)[0][..., 0]
                          for m in range(out.shape[-1])], dim=-1)
        if order == 1:
            return d1, None
        d2 = torch.stack([_grad_or_zero(d1[..., m].sum(), eps)[..., 0] for m in range(out.shape[-1])], dim=-1)
        return d1, d2

    # -- JetEvaluation protocol ----------------------------------------------------------------
    def columns_of(self, var):
        return self.inner.columns_of(var)

    def ensure(self, need):
        self._cache.clear()
        return self.inner.ensure(need)

    @property
    def spec(self):
        return self.inner.spec

    def value(self):
        if "value" not in self._cache:
            self._cache["value"] = self._apply({}, self.inner.value())
        return self._cache["value"]

    def _inner_first(self, in_col):
        return self.inner.first(list(range(self.inner.engine.d_out)), in_col)

    def first(self, out_cols, in_col):
        key = ("d", in_col)
        if key not in self._cache:
            self._cache[key] = self._path(in_col, self._inner_first(in_col), 1)[0]
        return JetEvaluation._columns(self._cache[key], out_cols, self._cache[key].dim() - 1)

    def second(self, out_cols):
        if "lap" not in self._cache:
            spec = self.inner.spec
            acc = self._path(None, self.inner.second(list(range(self.inner.engine.d_out))), 1)[0]
            for c, w in zip(spec.dcols, spec.lap_w):
                if w != 0.0:
                    acc = acc + w * self._path(c, self._inner_first(c), 2)[1]
            self._cache["lap"] = acc
        return JetEvaluation._columns(self._cache["lap"], out_cols, self._cache["lap"].dim() - 1)

    def plain(self):
        raise NotImplementedError("a constrained native evaluation has no generic autograd graph to the coordinates")


def _grad_or_zero(scalar, eps):
    """d scalar / d eps with create_graph, zeros where the graph does not reach eps (g linear along the path)."""
    if not scalar.requires_grad:
        return torch.zeros_like(eps)
    g = torch.autograd.grad(scalar, eps, create_graph=True, allow_unused=True)[0]
    return torch.zeros_like(eps) if g is None else g


class JetTensor(torch.Tensor):
    """A tensor of model outputs that remembers its JetEvaluation and output columns."""

    @staticmethod
    def wrap(t, ev, cols):
        out = t.as_subclass(JetTensor)
        out._jet, out._cols = ev, list(cols)
        return out

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = {} if kwargs is None else kwargs
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **kwargs)
        if func is torch.Tensor.__getitem__ and len(args) == 2 and isinstance(args[0], JetTensor):
            cols = _column_selection(args[0], args[1])
            if cols is not None and isinstance(out, torch.Tensor):
                return JetTensor.wrap(out, args[0]._jet, cols)
        elif func in (torch.Tensor.narrow, torch.narrow) and isinstance(args[0], JetTensor):
            a = list(args) + [kwargs[k] for k in ("dim", "start", "length") if k in kwargs]
            if len(a) == 4:
                src, dim, start, length = a
                if dim in (-1, src.dim() - 1) and isinstance(start, int):
                    return JetTensor.wrap(out, src._jet, src._cols[start:start + length])
        return out


def _column_selection(src, idx):
    """columns kept by ``src[idx]`` if idx only slices the last axis, else None."""
    if not hasattr(src, "_cols"):
        return None
    if not isinstance(idx, tuple):
        if idx is Ellipsis or (isinstance(idx, slice) and idx == slice(None) and src.dim() > 1):
            return list(src._cols)
        return None
    idx = list(idx)
    if len(idx) == 0:
        return None
    last = idx[-1]
    lead = idx[:-1]
    if any(v is Ellipsis for v in lead):
        if len(lead) != 1:
            return None
    elif len(idx) != src.dim():
        return None
    if not all(v is Ellipsis or (isinstance(v, slice) and v == slice(None)) for v in lead):
        return None
    if isinstance(last, slice):
        return src._cols[last]
    if isinstance(last, int):
        return [src._cols[last]]
    if isinstance(last, (list, tuple)) and all(isinstance(v, int) for v in last):
        return [src._cols[v] for v in last]
    return None


def jet_of(t):
    """(evaluation, columns) if ``t`` is a live JetTensor, else (None, None)."""
    if isinstance(t, JetTensor) and getattr(t, "_jet", None) is not None:
        return t._jet, t._cols
    return None, None


def _derives_from_native_evaluation(t, limit=512):
    """True if the autograd graph of ``t`` contains a native jet evaluation (``_FcnJetFn``): such a tensor
    depends on the coordinates THROUGH the network, but its graph only reaches the parameters."""
    fn = getattr(t, "grad_fn", None)
    if fn is None:
        return False
    seen, todo = set(), [fn]
    while todo and len(seen) < limit:
        node = todo.pop()
        if node is None or node in seen:
            continue
        seen.add(node)
        if type(node).__name__ == "_FcnJetFnBackward":
            return True
        todo.extend(n for n, _ in node.next_functions)
    return False


def plain_of(t):
    """tensor with a torch.autograd graph to the coordinate leaves (generic operator path)."""
    ev, cols = jet_of(t)
    if ev is None:
        if isinstance(t, torch.Tensor) and _derives_from_native_evaluation(t):
            # e.g. laplacian(u ** 2, x) or a HardConstraint output: torch.autograd would silently drop the
            # dependence on x through the network
            from . import _lib
            raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED,
                                "a differential operator was applied to a FUNCTION of a native model output (only the "
                                "model output itself, or column slices of it, carry derivative streams); apply the "
                                "operator to the model output and combine the results (product / chain rule)")
        return t
    full = ev.plain()
    if cols == list(range(full.shape[-1])) and t.dim() == full.dim():
        return full
    sel = full[..., cols]
    return sel if t.dim() == full.dim() else sel.squeeze(-1)
