"""Host logic of jets.ExprEvaluation: derivative streams through pointwise expressions of native model outputs
(reference semantics: differentialoperators.py:11-66 differentiate ANY autograd graph, e.g. laplacian(u * g, x)).
The native evaluation is replaced by an analytic stand-in with the same protocol, so this runs without a GPU."""
import math

import torch

from torchphysics_b200 import jets, operators
from torchphysics_b200.engine import JetSpec


class AnalyticEvaluation:
    """u0 = sin(a x0) cos(b x1), u1 = exp(0.3 x0) x1^2 with exact streams, on tracked leaves."""

    def __init__(self, x_leaf):
        self.leaf = x_leaf
        self.batch_shape = tuple(x_leaf.shape[:-1])
        self.leaf_cols = {id(x_leaf): (x_leaf, [0, 1])}
        self.spec = JetSpec((), None)
        self.a, self.b = 1.3, 0.7

    def columns_of(self, var):
        return [0, 1] if var is self.leaf else None

    def leaf_for_col(self, col):
        return self.leaf, col

    def ensure(self, need):
        self.spec = need
        return True

    def lap_spec(self):
        return self.spec

    def _x(self):
        x = self.leaf.detach()
        return x[..., 0], x[..., 1]

    def value(self):
        x0, x1 = self._x()
        return torch.stack([torch.sin(self.a * x0) * torch.cos(self.b * x1), torch.exp(0.3 * x0) * x1 ** 2], dim=-1)

    def first(self, out_cols, in_col):
        x0, x1 = self._x()
        a, b = self.a, self.b
        d = [[a * torch.cos(a * x0) * torch.cos(b * x1), -b * torch.sin(a * x0) * torch.sin(b * x1)],
             [0.3 * torch.exp(0.3 * x0) * x1 ** 2, 2 * torch.exp(0.3 * x0) * x1]]
        return torch.stack([d[m][in_col] for m in out_cols], dim=-1)

    def second(self, out_cols):
        x0, x1 = self._x()
        a, b = self.a, self.b
        w = dict(zip(self.spec.dcols, self.spec.lap_w))
        dd = [[-a * a * torch.sin(a * x0) * torch.cos(b * x1), -b * b * torch.sin(a * x0) * torch.cos(b * x1)],
              [0.09 * torch.exp(0.3 * x0) * x1 ** 2, 2 * torch.exp(0.3 * x0)]]
        return torch.stack([sum(w.get(k, 0.0) * dd[m][k] for k in (0, 1)) for m in out_cols], dim=-1)


def _setup(n=64):
    torch.manual_seed(0)
    x = torch.rand(n, 2, dtype=torch.float64).requires_grad_(True)
    ev = AnalyticEvaluation(x)
    u = jets.JetTensor.wrap(ev.value(), ev, [0, 1])
    return x, ev, u


def _reference(fn, x):
    """the same expression through plain torch autograd (the reference's definition)"""
    a, b = 1.3, 0.7
    u = torch.stack([torch.sin(a * x[:, 0]) * torch.cos(b * x[:, 1]), torch.exp(0.3 * x[:, 0]) * x[:, 1] ** 2], dim=-1)
    out = fn(u, x)
    g = torch.autograd.grad(out.sum(), x, create_graph=True)[0]
    lap = sum(torch.autograd.grad(g[:, i].sum(), x, create_graph=True)[0][:, i] for i in range(2))
    return out, g, lap[:, None]


def _check(fn):
    x, ev, u = _setup()
    expr = fn(u, x)
    assert isinstance(expr, jets.JetTensor) and isinstance(expr._jet, jets.ExprEvaluation)
    g = operators.grad(expr, x)
    lap = operators.laplacian(expr, x)
    out_r, g_r, lap_r = _reference(fn, x.detach().clone().requires_grad_(True))
    assert torch.allclose(expr.as_subclass(torch.Tensor), out_r, rtol=1e-12, atol=1e-12)
    assert torch.allclose(g, g_r, rtol=1e-10, atol=1e-10)
    assert torch.allclose(lap, lap_r, rtol=1e-9, atol=1e-9)


def test_product_of_output_and_coordinate_function():
    _check(lambda u, x: u[:, :1] * torch.sin(math.pi * x[:, :1]) * x[:, 1:])          # u * g(x)


def test_power_sum_and_scalars():
    _check(lambda u, x: u[:, :1] ** 2 + 3.0 * u[:, 1:] - 1.5)


def test_unary_functions_and_division():
    _check(lambda u, x: torch.exp(u[:, :1]) / (2.0 + torch.cos(u[:, 1:])))


def test_join_and_slice_of_an_expression():
    def fn(u, x):
        both = torch.cat([u[:, :1] * u[:, 1:], torch.tanh(u[:, :1])], dim=-1)
        return both[:, 1:] - both[:, :1]
    _check(fn)


def test_nested_expressions():
    _check(lambda u, x: (u[:, :1] * x[:, :1] + u[:, 1:]) * (u[:, :1] - x[:, 1:] ** 2))


def test_div_of_an_expression():
    x, ev, u = _setup()
    v = u * torch.cat([x[:, 1:], x[:, :1]], dim=-1)             # vector field (u0 x1, u1 x0)
    d = operators.div(v, x)
    xr = x.detach().clone().requires_grad_(True)
    a, b = 1.3, 0.7
    ur = torch.stack([torch.sin(a * xr[:, 0]) * torch.cos(b * xr[:, 1]), torch.exp(0.3 * xr[:, 0]) * xr[:, 1] ** 2], dim=-1)
    vr = ur * torch.cat([xr[:, 1:], xr[:, :1]], dim=-1)
    dr = sum(torch.autograd.grad(vr[:, i].sum(), xr, create_graph=True)[0][:, i] for i in range(2))[:, None]
    assert torch.allclose(d, dr, rtol=1e-10, atol=1e-10)


def test_non_pointwise_functions_still_raise():
    import pytest
    from torchphysics_b200 import _lib
    x, ev, u = _setup()
    t = torch.cumsum(u, dim=0)
    assert not isinstance(t, jets.JetTensor) or getattr(t, "_jet", None) is None


def test_laplacian_accepts_only_its_own_gradient_as_grad_argument():
    """reference differentialoperators.py:11-44: laplacian(u, x, grad=g) differentiates the tensor it is given.  The native streams
    carry no graph to x, so g must be tp.utils.grad of the same output w.r.t. the same variables (then the Laplacian stream is the
    answer); anything else is refused instead of being ignored."""
    import pytest
    from torchphysics_b200 import _lib
    x, ev, u = _setup()
    u0 = u[:, :1]
    g = operators.grad(u0, x)
    assert torch.equal(operators.laplacian(u0, x, grad=g), operators.laplacian(u0, x))
    with pytest.raises(_lib.TpxError):
        operators.laplacian(u0, x, grad=operators.grad(u[:, 1:], x))         # the gradient of another output column
    with pytest.raises(_lib.TpxError):
        operators.laplacian(u0, x, grad=torch.ones(len(x), 2, dtype=torch.float64))


def test_streams_beyond_one_jet_come_from_further_evaluations():
    """a heat equation in three space dimensions needs d/dt next to the Laplacian in (x, y, z): four directions, more than one
    jet carries (TPX_MAX_ND = 3).  JetEvaluation answers such requests from a second evaluation of the same network on the same
    points instead of refusing; the same holds for two different Laplacians.  (Stand-in engine: torch ops, same protocol.)"""
    from tests.test_host_deeponet import _TorchJetEngine
    torch.manual_seed(0)
    lins = [torch.nn.Linear(4, 16), torch.nn.Linear(16, 16), torch.nn.Linear(16, 2)]
    for lin in lins:
        lin.double()
    eng = _TorchJetEngine(lins)
    n = 50
    xyz = torch.rand(n, 3, dtype=torch.float64).requires_grad_(True)
    t = torch.rand(n, 1, dtype=torch.float64).requires_grad_(True)
    pts = torch.cat([xyz, t], dim=1)
    leaf_cols = {id(xyz): (xyz, [0, 1, 2]), id(t): (t, [3])}
    hints = {}
    ev = jets.JetEvaluation(eng, pts.detach(), (n,), leaf_cols, hints, "k", None, JetSpec())
    u = jets.JetTensor.wrap(ev.value(), ev, [0, 1])
    lap = operators.laplacian(u[:, :1], xyz)
    ut = operators.grad(u[:, :1], t)
    lap_t = operators.laplacian(u[:, 1:], t)                  # a second, different Laplacian
    gx = operators.grad(u[:, 1:], xyz)
    assert len(ev.extra) >= 1                                 # the requests did not fit one jet
    # the reference's definition on the same modules
    h = pts
    for i, lin in enumerate(lins):
        h = lin(h)
        if i < len(lins) - 1:
            h = torch.tanh(h)

    def d(y, v):
        return torch.autograd.grad(y.sum(), v, create_graph=True)[0]

    g0 = d(h[:, 0], xyz)
    want_lap = sum(d(g0[:, i], xyz)[:, i] for i in range(3))[:, None]
    assert torch.allclose(lap, want_lap, rtol=1e-9, atol=1e-10)
    assert torch.allclose(ut, d(h[:, 0], t), rtol=1e-9, atol=1e-10)
    assert torch.allclose(lap_t, d(d(h[:, 1], t)[:, 0], t), rtol=1e-9, atol=1e-10)
    assert torch.allclose(gx, d(h[:, 1], xyz), rtol=1e-9, atol=1e-10)
    # parameter gradients flow through every evaluation
    loss = ((ut - lap) ** 2).mean() + (lap_t ** 2).mean() + (gx ** 2).mean()
    want = ((d(h[:, 0], t) - want_lap) ** 2).mean() + (d(d(h[:, 1], t)[:, 0], t) ** 2).mean() + (d(h[:, 1], xyz) ** 2).mean()
    g_native = torch.autograd.grad(loss, [lin.weight for lin in lins])
    g_want = torch.autograd.grad(want, [lin.weight for lin in lins])
    for a, b in zip(g_native, g_want):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10)


def test_evaluations_die_with_their_step_without_the_cycle_collector():
    """the streams of a step (and the autograd graph behind them) must be released by reference counting when the step's tensors
    go away: an evaluation kept alive by a reference cycle would hold batch-sized tensors until the next gc run and keep the
    AccumulateGrad nodes of an eager step alive into a CUDA-graph capture (which then fails)."""
    import gc
    import weakref
    from tests.test_host_deeponet import _TorchJetEngine
    lins = [torch.nn.Linear(4, 8).double(), torch.nn.Linear(8, 8).double(), torch.nn.Linear(8, 1).double()]
    eng = _TorchJetEngine(lins)
    xyz = torch.rand(20, 3, dtype=torch.float64).requires_grad_(True)
    t = torch.rand(20, 1, dtype=torch.float64).requires_grad_(True)
    gc.collect()
    gc.disable()
    try:
        ev = jets.JetEvaluation(eng, torch.cat([xyz, t], dim=1).detach(), (20,), {id(xyz): (xyz, [0, 1, 2]), id(t): (t, [3])},
                                {}, "k", None, JetSpec())
        u = jets.JetTensor.wrap(ev.value(), ev, [0])
        r = operators.grad(u, t) - operators.laplacian(u * u, xyz)          # a second evaluation and an expression evaluation
        loss = (r ** 2).mean()
        loss.backward()
        refs = [weakref.ref(ev)] + [weakref.ref(e) for e in ev.extra]
        assert len(refs) >= 2
        del ev, u, r, loss
        assert all(ref() is None for ref in refs)
    finally:
        gc.enable()
