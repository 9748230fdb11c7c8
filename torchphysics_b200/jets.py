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
        # Streams that do not fit ONE jet (more than 3 derivative directions in total -- a Laplacian in three space dimensions
        # next to a time derivative -- or two different Laplacians) come from further evaluations of the same network on the
        # same points, created on demand; ``_hit`` is the evaluation that answered the last ``ensure``.
        self.extra = []
        self._hit = None                    # None: this evaluation itself (no reference cycle: the streams must die with the step)

    def columns_of(self, var):
        hit = self.leaf_cols.get(id(var))
        if hit is None or hit[0] is not var:
            return None
        return hit[1]

    def leaf_for_col(self, col):
        """(leaf tensor, index within it) of input column ``col``, or (None, None)."""
        for leaf, cols in self.leaf_cols.values():
            if col in cols:
                return leaf, cols.index(col)
        return None, None

    def ensure(self, need):
        """make sure the computed streams cover ``need``; returns False if impossible."""
        for ev in [self] + self.extra:
            if ev._ensure_own(need):
                self._hit = None if ev is self else ev
                return True
        if need.nd > 3:
            return False
        ev = JetEvaluation(self.engine, self.x, self.batch_shape, self.leaf_cols, {}, None, None,
                           JetSpec(need.dcols, need.lap_w, out_major=self.spec.out_major), self.params)
        self.extra.append(ev)
        self._hit = ev
        return True

    def _ensure_own(self, need):
        """extend THIS evaluation's jet to cover ``need`` if the union still is one jet."""
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
        new = JetSpec(cols, lap_w, out_major=self.spec.out_major)
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

    def part_with(self, in_col):
        """the evaluation (this one or a further one) whose jet carries direction ``in_col``"""
        hit = self._hit or self
        if in_col in hit.spec.dcols:
            return hit
        for ev in [self] + self.extra:
            if in_col in ev.spec.dcols:
                return ev
        raise KeyError(in_col)

    def lap_part(self):
        """the evaluation whose Laplacian stream answers the last ``ensure``"""
        hit = self._hit or self
        return hit if hit.spec.lap else self

    def lap_spec(self):
        """directions and weights of the Laplacian stream that ``second`` returns"""
        return self.lap_part().spec

    def first(self, out_cols, in_col):
        ev = self.part_with(in_col)
        d = ev.du.select(2, ev.spec.dcols.index(in_col))
        return self._columns(d, out_cols, 1).reshape(*self.batch_shape, len(out_cols))

    def second(self, out_cols):
        return self._columns(self.lap_part().lap, out_cols, 1).reshape(*self.batch_shape, len(out_cols))

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

    def leaf_for_col(self, col):
        return self.parts[0][0].leaf_for_col(col)

    @property
    def spec(self):
        return self.parts[0][0].spec

    def lap_spec(self):
        return self.parts[0][0].lap_spec()

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

    def leaf_for_col(self, col):
        return self.inner.leaf_for_col(col)

    @property
    def spec(self):
        return self.inner.spec

    def lap_spec(self):
        return self.inner.lap_spec()

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
            spec = self.inner.lap_spec()
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


class ExprEvaluation:
    """Streams of a POINTWISE torch expression of native model outputs (``u * g``, ``u ** 2 + v``, ``torch.sin(u) * x`` ...),
    same protocol as ``JetEvaluation``: what lets ``laplacian(u * g, x)`` work as in the reference, whose operators
    differentiate any autograd graph (``differentialoperators.py:11-44``).

    Operands are native evaluations (their derivative streams come from the kernel), tensors with a torch graph to the
    coordinate leaves (``torch.sin(x)``: differentiated by torch.autograd) and constants.  With every operand v_i moved
    along its own derivative, phi_k(eps) = f(v_1 + eps d_k v_1, ..., v_m + eps d_k v_m):
        d_k out  = phi_k'(0),      d_kk out = phi_k''(0) + sum_i f_{v_i} d_kk v_i
    so the weighted Laplacian is  sum_k c_k phi_k''(0) + psi'(0),  psi(eps) = f(v_i + eps L v_i); the derivatives of phi
    and psi are taken by torch.autograd on the pointwise expression (the network is never re-traversed) and keep their graph
    to the operands' streams, hence to the parameters."""

    def __init__(self, fn, operands, base):
        self.fn = fn                             # fn(*tensors) -> tensor (..., m)
        self.operands = operands                 # [("jet", evaluation, cols, value tensor) | ("plain", tensor)]
        self.base = base                         # a native evaluation among the operands (leaves, spec)
        self.batch_shape = base.batch_shape
        self._cache = {}

    def columns_of(self, var):
        return self.base.columns_of(var)

    def leaf_for_col(self, col):
        return self.base.leaf_for_col(col)

    @property
    def spec(self):
        return self.base.spec

    def lap_spec(self):
        return self.base.lap_spec()

    def ensure(self, need):
        self._cache.clear()
        seen, ok = set(), True
        for op in self.operands:
            if op[0] == "jet" and id(op[1]) not in seen:
                seen.add(id(op[1]))
                ok = op[1].ensure(need) and ok
        return ok

    # -- operand streams ---------------------------------------------------------------------
    @staticmethod
    def _value(op):
        return op[3] if op[0] == "jet" else op[1]

    def _first_of(self, op, in_col):
        if op[0] == "jet":
            return op[1].first(op[2], in_col).reshape(op[3].shape)
        t = op[1]
        leaf, idx = self.leaf_for_col(in_col)
        if not (isinstance(t, torch.Tensor) and t.requires_grad and leaf is not None):
            return None
        if t.dim() == 0 or t.shape[:len(self.batch_shape)] != tuple(self.batch_shape):
            return None                          # not a per-point quantity: constant along the coordinates
        cols = []
        for m in range(t.shape[-1]) if t.dim() > len(self.batch_shape) else [None]:
            comp = t if m is None else t[..., m]
            g = torch.autograd.grad(comp.sum(), leaf, create_graph=True, allow_unused=True)[0]
            cols.append(torch.zeros(self.batch_shape, device=t.device, dtype=t.dtype) if g is None else g[..., idx])
        return cols[0] if t.dim() == len(self.batch_shape) else torch.stack(cols, dim=-1)

    def _second_of(self, op):
        if op[0] == "jet":
            return op[1].second(op[2]).reshape(op[3].shape)
        acc = None
        spec = self.lap_spec()
        for c, w in zip(spec.dcols, spec.lap_w or ()):
            if w == 0.0:
                continue
            d1 = self._first_of(op, c)
            if d1 is None or not d1.requires_grad:
                continue
            leaf, idx = self.leaf_for_col(c)
            cols = []
            for m in range(d1.shape[-1]) if d1.dim() > len(self.batch_shape) else [None]:
                comp = d1 if m is None else d1[..., m]
                g = torch.autograd.grad(comp.sum(), leaf, create_graph=True, allow_unused=True)[0]
                cols.append(torch.zeros(self.batch_shape, device=d1.device, dtype=d1.dtype) if g is None else g[..., idx])
            d2 = cols[0] if d1.dim() == len(self.batch_shape) else torch.stack(cols, dim=-1)
            acc = w * d2 if acc is None else acc + w * d2
        return acc

    def _path(self, dirs, order):
        """(phi'(0), phi''(0)) of phi(eps) = fn(v_i + eps * dirs_i) (dirs_i None: the operand stays fixed)"""
        vals = [self._value(op) for op in self.operands]
        ref = vals[0]
        eps = torch.zeros((*self.batch_shape, 1), device=ref.device, dtype=ref.dtype, requires_grad=True)
        moved = []
        for v, d in zip(vals, dirs):
            if d is None:
                moved.append(v)
            else:
                e = eps if v.dim() > len(self.batch_shape) else eps[..., 0]
                moved.append(v + e * d)
        out = self.fn(*moved)
        d1 = torch.stack([_grad_or_zero(out[..., m].sum(), eps)[..., 0] for m in range(out.shape[-1])], dim=-1)
        if order == 1:
            return d1, None
        d2 = torch.stack([_grad_or_zero(d1[..., m].sum(), eps)[..., 0] for m in range(out.shape[-1])], dim=-1)
        return d1, d2

    # -- JetEvaluation protocol ----------------------------------------------------------------
    def value(self):
        if "value" not in self._cache:
            self._cache["value"] = self.fn(*[self._value(op) for op in self.operands])
        return self._cache["value"]

    def first(self, out_cols, in_col):
        key = ("d", in_col)
        if key not in self._cache:
            self._cache[key] = self._path([self._first_of(op, in_col) for op in self.operands], 1)[0]
        return JetEvaluation._columns(self._cache[key], out_cols, self._cache[key].dim() - 1)

    def second(self, out_cols):
        if "lap" not in self._cache:
            spec = self.lap_spec()
            acc = self._path([self._second_of(op) for op in self.operands], 1)[0]
            for c, w in zip(spec.dcols, spec.lap_w):
                if w != 0.0:
                    acc = acc + w * self._path([self._first_of(op, c) for op in self.operands], 2)[1]
            self._cache["lap"] = acc
        return JetEvaluation._columns(self._cache["lap"], out_cols, self._cache["lap"].dim() - 1)

    def plain(self):
        raise NotImplementedError("an expression of native evaluations has no generic autograd graph to the coordinates")


def _pointwise_functions():
    T = torch.Tensor
    fns = {torch.add, torch.sub, torch.mul, torch.div, torch.true_divide, torch.neg, torch.pow, torch.square, torch.sqrt,
           torch.exp, torch.log, torch.sin, torch.cos, torch.tanh, torch.sigmoid, torch.abs,
           T.add, T.sub, T.mul, T.div, T.true_divide, T.neg, T.pow, T.square, T.sqrt, T.exp, T.log, T.sin, T.cos, T.tanh,
           T.sigmoid, T.abs, T.__add__, T.__radd__, T.__sub__, T.__rsub__, T.__mul__, T.__rmul__, T.__truediv__,
           T.__rtruediv__, T.__neg__, T.__pow__, T.__rpow__}
    return fns


_POINTWISE = _pointwise_functions()
_JOINS = {torch.cat, torch.concat, torch.column_stack, torch.hstack}


def _derive_expression(func, args, kwargs, out):
    """JetTensor of ``out = func(*args, **kwargs)`` if func is pointwise (or a join along the last axis) and the operands
    are native outputs of one point set, plain tensors and numbers; None otherwise."""
    if not isinstance(out, torch.Tensor) or out.dim() < 1:
        return None
    join = func in _JOINS
    if not join and func not in _POINTWISE:
        return None
    if kwargs and (set(kwargs) - {"dim", "alpha"}):
        return None
    seq = list(args[0]) if join else list(args)
    if join:
        dim = kwargs.get("dim", args[1] if len(args) > 1 else (0 if func in (torch.cat, torch.concat) else -1))
        if func in (torch.cat, torch.concat) and dim not in (-1, out.dim() - 1):
            return None
        if func in (torch.column_stack, torch.hstack) and out.dim() != 2:
            return None
    operands, slots, base = [], [], None
    for a in seq:
        ev, cols = jet_of(a)
        if ev is not None:
            if base is not None and tuple(ev.batch_shape) != tuple(base.batch_shape):
                return None
            base = base or ev
            operands.append(("jet", ev, cols, a.as_subclass(torch.Tensor)))
            slots.append(len(operands) - 1)
        elif isinstance(a, torch.Tensor):
            operands.append(("plain", a.as_subclass(torch.Tensor)))
            slots.append(len(operands) - 1)
        elif isinstance(a, (int, float, bool)):
            slots.append(("const", a))
        else:
            return None
    if base is None or tuple(out.shape[:len(base.batch_shape)]) != tuple(base.batch_shape) or out.dim() != len(base.batch_shape) + 1:
        return None
    kw = dict(kwargs or {})

    def fn(*tensors):
        rebuilt = [tensors[sl] if isinstance(sl, int) else sl[1] for sl in slots]
        if join:
            return func(rebuilt, *args[1:], **kw)
        return func(*rebuilt, **kw)
    return JetTensor.wrap(out, ExprEvaluation(fn, operands, base), range(out.shape[-1]))


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
        else:
            derived = _derive_expression(func, args, kwargs, out)
            if derived is not None:
                return derived
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
            # e.g. laplacian(torch.cumsum(u, 0), x): not a pointwise expression (those carry streams, ExprEvaluation);
            # torch.autograd would silently drop the dependence on x through the network
            from . import _lib
            raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED,
                                "a differential operator was applied to a function of a native model output that is not a "
                                "pointwise arithmetic expression (+ - * / ** neg, sin cos exp log tanh sqrt ..., cat along "
                                "the last axis: those carry derivative streams); apply the operator to the model output and "
                                "combine the results")
        return t
    full = ev.plain()
    if cols == list(range(full.shape[-1])) and t.dim() == full.dim():
        return full
    sel = full[..., cols]
    return sel if t.dim() == full.dim() else sel.squeeze(-1)
