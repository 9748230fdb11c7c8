"""The hot path through the reference-facing API (tp.models / tp.utils / tp.conditions /
tp.solver) on the GPU, against the golden outputs of the real reference and the oracle.
Tolerance 1e-5 max-norm relative (BASELINE.json)."""
import math

import numpy as np
import pytest
import torch

from oracle import jet_numpy, ref_autograd
from tests.golden_io import load, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _load_params(model, params):
    lins = [m for m in model.sequential if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(torch.as_tensor(params[2 * i]))
            lin.bias.copy_(torch.as_tensor(params[2 * i + 1]))


def _grads(model):
    return [None if p.grad is None else p.grad.cpu().numpy() for p in model._native_engine().params()]


def test_pinn_condition_poisson_matches_reference():
    import torchphysics_b200 as tp
    G = load("poisson_4x64")
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=(64,) * 4).cuda()
    _load_params(model, G["params"])
    sampler = tp.samplers.DataSampler(tp.spaces.Points(torch.from_numpy(G["x"]), X))
    cond = tp.conditions.PINNCondition(model, sampler, lambda u, x, f: tp.utils.laplacian(u, x) + f,
                                       data_functions={"f": ref_autograd.poisson_source})
    for it in range(3):                       # 1st call learns the jet spec, later calls are steady state
        model.zero_grad(set_to_none=True)
        loss = cond(device="cuda")
        loss.backward()
        assert abs(float(loss) - float(G["loss"])) < TOL * abs(float(G["loss"]))
        g = _grads(model)
        assert g[-1] is None and G["dtheta"][-1] is None      # output bias untouched, as in the reference
        for a, w in zip(g[:-1], G["dtheta"][:-1]):
            assert relerr(a, w) < TOL


@pytest.mark.parametrize("name,width", [("ns_3x32", 32), ("ns_3x128", 128)])     # 128: wide tensor-core path
def test_operators_on_model_output_match_reference(name, width):
    import torchphysics_b200 as tp
    G = load(name)
    X, Y = tp.spaces.R2("x"), tp.spaces.R2("u") * tp.spaces.R1("p")
    model = tp.models.FCN(X, Y, hidden=(width,) * 3).cuda()
    _load_params(model, G["params"])
    pts = tp.spaces.Points(torch.from_numpy(G["x"]).cuda(), X)
    coords, pts = pts.track_coord_gradients()
    out = model(pts)
    u, p = out.coordinates["u"], out.coordinates["p"]
    x = coords["x"]
    assert relerr(out.as_tensor.detach().cpu().numpy(), G["y"]) < TOL
    assert relerr(tp.utils.jac(out.as_tensor, x).detach().cpu().numpy(), G["jac"]) < TOL
    assert relerr(tp.utils.div(u, x).detach().cpu().numpy(), G["div"]) < TOL
    assert relerr(tp.utils.convective(u, u, x).detach().cpu().numpy(), G["conv"]) < TOL
    lap = torch.cat([tp.utils.laplacian(u[:, :1], x), tp.utils.laplacian(u[:, 1:2], x)], dim=1)
    assert relerr(lap.detach().cpu().numpy(), G["lap"]) < TOL
    gp = tp.utils.grad(p, x)
    assert relerr(gp.detach().cpu().numpy(), G["jac"][:, 2, :]) < TOL
    # whole Navier-Stokes residual loss and its parameter gradient
    r = torch.cat([tp.utils.convective(u, u, x) - 0.01 * lap + gp, tp.utils.div(u, x)], dim=1)
    loss = tp.conditions.SquaredError()(r).mean()
    assert abs(float(loss) - float(G["loss"])) < TOL * abs(float(G["loss"]))
    loss.backward()
    for a, w in zip(_grads(model), G["dtheta"]):
        if w is None:
            assert a is None or np.abs(a).max() == 0
        else:
            assert relerr(a, w) < TOL


@pytest.mark.parametrize("name,width", [("heat_3x48", 48), ("heat_3x128", 128)])   # 128: wide tensor-core path
def test_heat_condition_with_product_space(name, width):
    import torchphysics_b200 as tp
    G = load(name)
    X, T, U = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("u")
    model = tp.models.FCN(X * T, U, hidden=(width,) * 3).cuda()
    _load_params(model, G["params"])
    sampler = tp.samplers.DataSampler(tp.spaces.Points(torch.from_numpy(G["xt"]), X * T))
    cond = tp.conditions.PINNCondition(model, sampler, lambda u, x, t: tp.utils.grad(u, t) - tp.utils.laplacian(u, x))
    for it in range(2):
        model.zero_grad(set_to_none=True)
        loss = cond(device="cuda")
        loss.backward()
    assert abs(float(loss) - float(G["loss"])) < TOL * abs(float(G["loss"]))
    for a, w in zip(_grads(model)[:-1], G["dtheta"][:-1]):
        assert relerr(a, w) < TOL


@pytest.mark.parametrize("name,hidden", [("ritz_4x40", (40,) * 4), ("ritz_2x256", (256, 256))])   # 256: wide path, two blocks
def test_deep_ritz_condition(name, hidden):
    import torchphysics_b200 as tp
    G = load(name)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=hidden).cuda()
    _load_params(model, G["params"])
    sampler = tp.samplers.DataSampler(tp.spaces.Points(torch.from_numpy(G["x"]), X))
    cond = tp.conditions.DeepRitzCondition(
        model, sampler, lambda u, x: 0.5 * (tp.utils.grad(u, x) ** 2).sum(dim=1, keepdim=True) - u)
    for it in range(2):
        model.zero_grad(set_to_none=True)
        loss = cond(device="cuda")
        loss.backward()
    assert abs(float(loss) - float(G["loss"])) < TOL * abs(float(G["loss"]))
    for a, w in zip(_grads(model), G["dtheta"]):
        assert relerr(a, w) < TOL


@pytest.mark.parametrize("hidden", [(32,) * 3, (128,) * 3, (160, 256)])     # width-64, wide (1 block), wide (2 blocks) kernels
def test_solver_trains_poisson(hidden):
    """a few Adam steps through Solver/Trainer reduce the PINN loss; steady state uses one
    forward and one reverse launch per condition (width <= 64) or one launch per hidden matrix block and direction."""
    import torchphysics_b200 as tp
    torch.manual_seed(0)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=hidden)
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    f = lambda x: 2 * math.pi ** 2 * torch.sin(math.pi * x[:, :1]) * torch.sin(math.pi * x[:, 1:])
    pde = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq, n_points=4096),
                                      lambda u, x, f: tp.utils.laplacian(u, x) + f, data_functions={"f": f})
    solver = tp.solver.Solver([pde], optimizer_setting=tp.solver.OptimizerSetting(torch.optim.Adam, lr=2e-3))
    trainer = tp.solver.Trainer(max_steps=60)
    trainer.fit(solver)
    first = float(torch.stack(trainer.losses[:5]).mean())
    last = float(torch.stack(trainer.losses[-5:]).mean())
    assert last < 0.5 * first


def test_cpu_tensors_and_unsupported_nets_fail_loudly():
    import torchphysics_b200 as tp
    from torchphysics_b200 import _lib
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=(16, 16))
    with pytest.raises(_lib.TpxError):
        model(tp.spaces.Points(torch.rand(4, 2), X))
    relu = tp.models.FCN(X, U, hidden=(16, 16), activations=torch.nn.ReLU()).cuda()
    with pytest.raises(_lib.TpxError):
        relu(tp.spaces.Points(torch.rand(4, 2).cuda(), X))


def test_generic_operator_semantics_on_plain_tensors():
    """operators on hand-written torch functions keep the reference's autograd definition
    (closed-form cases of the reference's tests/tests_utils/test_differentialoperators.py)."""
    import torchphysics_b200 as tp
    a = torch.tensor([[1.0, 1.0], [2.0, 0.5]], requires_grad=True, device="cuda")
    u = (a ** 2).sum(dim=1, keepdim=True)
    assert torch.allclose(tp.utils.laplacian(u, a), torch.full((2, 1), 4.0, device="cuda"))
    assert torch.allclose(tp.utils.grad(u, a), 2 * a)
    v = torch.stack([a[:, 0] ** 2, a[:, 0] * a[:, 1]], dim=1)
    assert torch.allclose(tp.utils.div(v, a).reshape(-1), 2 * a[:, 0] + a[:, 0])
    j = tp.utils.jac(v, a)
    assert j.shape == (2, 2, 2) and torch.allclose(j[:, 1, 0], a[:, 1])


def test_sequential_normalization_layer_is_folded_into_the_native_network():
    """tp.models.Sequential(NormalizationLayer, FCN) against the real reference's outputs,
    including the gradient of the trainable normalisation parameters."""
    import torchphysics_b200 as tp
    from torchphysics_b200 import jets
    G = load("heat_norm_3x48")
    X, T, U = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("u")
    dom = tp.domains.Parallelogram(X, [0, 0], [2, 0], [0, 3]) * tp.domains.Interval(T, 0, 5)
    norm = tp.models.NormalizationLayer(dom)
    fcn = tp.models.FCN(X * T, U, hidden=(48,) * 3)
    model = tp.models.Sequential(norm, fcn).cuda()
    with torch.no_grad():
        for q, w in zip(model.parameters(), G["params"]):
            assert tuple(q.shape) == tuple(w.shape)
            q.copy_(torch.as_tensor(w))
    # the layer's own initialisation equals the reference's (bounding box -> (-1, 1)^d)
    fresh = tp.models.NormalizationLayer(dom)
    assert relerr(fresh.normalize.weight.detach().numpy(), G["params"][0]) < 1e-6
    assert relerr(fresh.normalize.bias.detach().numpy(), G["params"][1]) < 1e-6
    sampler = tp.samplers.DataSampler(tp.spaces.Points(torch.from_numpy(G["xt"]), X * T))
    captured = {}

    def residual(u, x, t):
        captured.update(u=u, ut=tp.utils.grad(u, t), gx=tp.utils.grad(u, x), lap=tp.utils.laplacian(u, x))
        return captured["ut"] - 0.1 * captured["lap"]

    cond = tp.conditions.PINNCondition(model, sampler, residual)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error", jets.TpxFallbackWarning)       # the native path must answer everything
        for it in range(2):
            model.zero_grad(set_to_none=True)
            loss = cond(device="cuda")
            loss.backward()
    for name in ("u", "ut", "gx", "lap"):
        assert relerr(captured[name].detach().cpu().numpy(), G[name]) < TOL, name
    assert abs(float(loss) - float(G["loss"])) < TOL * abs(float(G["loss"]))
    for q, w in zip(model.parameters(), G["dtheta"]):
        if w is None:
            assert q.grad is None or float(q.grad.abs().max()) == 0.0
        else:
            assert relerr(q.grad.cpu().numpy(), w) < TOL


def test_native_adam_equals_torch_adam():
    """the one-launch multi-tensor Adam against torch.optim.Adam, including a parameter without
    gradient (left untouched) and weight decay."""
    import torchphysics_b200 as tp
    torch.manual_seed(0)
    shapes = [(64, 2), (64,), (64, 64), (64,), (1, 64), (1,), (3, 5, 7)]
    for wd in (0.0, 0.01):
        a = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
        b = [torch.nn.Parameter(p.detach().clone()) for p in a]
        oa = tp.optim.Adam(a, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
        ob = torch.optim.Adam(b, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
        for it in range(25):
            for i, (p, q) in enumerate(zip(a, b)):
                if i == 5 and it < 10:                       # no gradient for this parameter in the first steps
                    p.grad, q.grad = None, None
                    continue
                g = torch.randn_like(p) * (1.0 + it)
                p.grad, q.grad = g.clone(), g.clone()
            oa.step()
            ob.step()
        for p, q in zip(a, b):
            assert relerr(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < 2e-6
        assert int(oa.state[a[5]]["step"]) == 15 and int(oa.state[a[0]]["step"]) == 25
        sa, sb = oa.state[a[2]], ob.state[b[2]]
        assert relerr(sa["exp_avg"].cpu().numpy(), sb["exp_avg"].cpu().numpy()) < 2e-6
        assert relerr(sa["exp_avg_sq"].cpu().numpy(), sb["exp_avg_sq"].cpu().numpy()) < 2e-6
    # the Trainer picks it for OptimizerSetting(torch.optim.Adam, ...)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=(16, 16))
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    cond = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq, n_points=256),
                                       lambda u, x: tp.utils.laplacian(u, x) + 1.0)
    solver = tp.solver.Solver([cond], optimizer_setting=tp.OptimizerSetting(torch.optim.Adam, lr=1e-3))
    trainer = tp.solver.Trainer(max_steps=3)
    trainer.fit(solver)
    assert isinstance(trainer.optimizer, tp.optim.Adam)


def test_operator_on_a_function_of_the_model_output():
    """grad(x * u, x), laplacian(u * u, x): pointwise expressions of the model output carry derivative streams (product /
    chain rule on the native streams, jets.ExprEvaluation) as in the reference, where they are ordinary autograd graphs; a
    function that is NOT pointwise (torch.autograd would silently miss du/dx: the native evaluation has no graph to the
    coordinates) is refused instead of returning a wrong derivative."""
    import torchphysics_b200 as tp
    from torchphysics_b200 import _lib
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=(16, 16)).cuda()
    coords, pts = tp.spaces.Points(torch.rand(64, 2).cuda(), X).track_coord_gradients()
    u = model(pts).as_tensor
    x = coords["x"]
    g = tp.utils.grad(u, x)
    want = torch.cat([u + x[:, :1] * g[:, :1], x[:, :1] * g[:, 1:]], dim=1)       # product rule by hand
    got = tp.utils.grad(x[:, :1] * u, x)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    lap_u = tp.utils.laplacian(u, x)
    want2 = 2.0 * (g * g).sum(dim=1, keepdim=True) + 2.0 * u * lap_u            # Lap(u^2) = 2 |grad u|^2 + 2 u Lap u
    assert torch.allclose(tp.utils.laplacian(u * u, x), want2, rtol=1e-4, atol=1e-5)
    with pytest.raises(_lib.TpxError):
        tp.utils.grad(torch.cumsum(u, dim=0), x)


def test_hard_constraint_streams_by_chain_rule_match_reference():
    """tp.models.HardConstraint(FCN, g) -- value, gradient and Laplacian of g(x, u(x)) from the native
    streams of u by the chain rule, loss and dL/dtheta against the real reference's autograd."""
    import warnings
    import torchphysics_b200 as tp
    from torchphysics_b200 import jets
    G = load("poisson_hard_3x32")
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    inner = tp.models.FCN(X, U, hidden=(32,) * 3).cuda()
    _load_params(inner, G["params"])
    model = tp.models.HardConstraint(
        inner, lambda u, x: x[:, :1] * (1 - x[:, :1]) * x[:, 1:] * (1 - x[:, 1:]) * u + torch.sin(x[:, :1]))
    sampler = tp.samplers.DataSampler(tp.spaces.Points(torch.from_numpy(G["x"]), X))
    cap = {}

    def residual(u, x):
        cap.update(u=u, grad=tp.utils.grad(u, x), lap=tp.utils.laplacian(u, x))
        return cap["lap"] + 1.0

    cond = tp.conditions.PINNCondition(model, sampler, residual)
    with warnings.catch_warnings():
        warnings.simplefilter("error", jets.TpxFallbackWarning)
        for _ in range(2):
            inner.zero_grad(set_to_none=True)
            loss = cond(device="cuda")
            loss.backward()
    for k in ("u", "grad", "lap"):
        assert relerr(cap[k].detach().cpu().numpy(), G[k]) < TOL, k
    assert abs(float(loss) - float(G["loss"])) < TOL * abs(float(G["loss"]))
    for a, w in zip(_grads(inner), G["dtheta"]):
        assert relerr(a, w) < TOL


def test_inverse_problem_parameter_and_parallel_models():
    """a learnable tp.models.Parameter in the residual (reference models/parameter.py, condition.py:268,295-299) and
    tp.models.Parallel of two native networks (model.py:95-146): loss and every gradient against torch autograd on
    the same weights (oracle call sites)."""
    import torchphysics_b200 as tp
    torch.manual_seed(21)
    X, U, V, Dsp = tp.spaces.R2("x"), tp.spaces.R1("u"), tp.spaces.R1("v"), tp.spaces.R1("D")
    net_u = tp.models.FCN(X, U, hidden=(32, 32)).cuda()
    net_v = tp.models.FCN(X, V, hidden=(24, 24, 24)).cuda()
    model = tp.models.Parallel(net_u, net_v)
    D = tp.models.Parameter(init=0.7, space=Dsp)
    x_np = np.random.default_rng(3).uniform(0, 1, size=(300, 2)).astype(np.float32)
    sampler = tp.samplers.DataSampler(tp.spaces.Points(torch.from_numpy(x_np), X))

    def residual(u, v, x, D):
        return D * tp.utils.laplacian(u, x) + tp.utils.grad(v, x)[:, :1] * u - v

    cond = tp.conditions.PINNCondition(model, sampler, residual, parameter=D, name="inverse")
    assert any(n.endswith("inverse_params") for n, _ in cond.named_parameters())
    cond = cond.cuda()
    for _ in range(2):
        cond.zero_grad(set_to_none=True)
        loss = cond(device="cuda")
        loss.backward()
    # oracle: the same two networks and D with torch autograd on the CPU
    su = ref_autograd.fcn_from_params([p.detach().cpu().numpy() for p in net_u._native_engine().params()])
    sv = ref_autograd.fcn_from_params([p.detach().cpu().numpy() for p in net_v._native_engine().params()])
    Dc = torch.tensor([[0.7]], requires_grad=True)
    cols, xin = ref_autograd.track(torch.from_numpy(x_np), [("x", 2)])
    u, v = su(xin), sv(xin)
    r = Dc * ref_autograd.laplacian(u, cols["x"]) + ref_autograd.grad(v, cols["x"])[:, :1] * u - v
    want = ref_autograd.pinn_loss(r)
    ps = ref_autograd.linear_params(su) + ref_autograd.linear_params(sv) + [Dc]
    gs = torch.autograd.grad(want, ps)
    assert abs(float(loss) - float(want)) < TOL * abs(float(want))
    got = [p.grad for p in net_u._native_engine().params()] + [p.grad for p in net_v._native_engine().params()]
    for a, w in zip(got, gs[:-1]):
        assert relerr(a.cpu().numpy(), w.numpy()) < TOL
    dgrad = [p.grad for n, p in cond.named_parameters() if n.endswith("inverse_params")][0]
    assert relerr(dgrad.cpu().numpy(), gs[-1].numpy()) < TOL


def test_solver_sums_weighted_conditions():
    """Solver.training_step = sum_c weight_c * loss_c (reference solver.py:100-109) over a PDE and a boundary condition."""
    import torchphysics_b200 as tp
    torch.manual_seed(5)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=(16, 16)).cuda()
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    pde = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq, n_points=200).make_static(),
                                      lambda u, x: tp.utils.laplacian(u, x) + 1.0, weight=0.5, name="pde")
    bc = tp.conditions.PINNCondition(model, tp.samplers.GridSampler(sq.boundary, n_points=64).make_static(),
                                     lambda u: u, weight=3.0, name="bc")
    solver = tp.solver.Solver([pde, bc])
    solver.to("cuda")
    solver._device = torch.device("cuda")
    solver.on_train_start()
    total = solver.training_step()
    lp, lb = pde(device="cuda"), bc(device="cuda")
    assert abs(float(total) - (0.5 * float(lp) + 3.0 * float(lb))) < 1e-6 * abs(float(total))
    assert set(solver.logged) >= {"train/pde", "train/bc", "train/loss"}
    total.backward()
    assert all(p.grad is not None for p in model.parameters())


def test_reference_flagship_example_runs_unchanged():
    """examples/heat_equation.py = the code cells of the reference's examples/pinn/heat-equation.ipynb with only the
    import changed (product domains with density sampling, adaptive sampler, boundary / boundary_left,
    Sequential(NormalizationLayer, FCN), data functions, three conditions, Adam): it trains."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("heat_equation", os.path.join(root, "examples", "heat_equation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(0)
    losses = mod.main(60)
    assert len(losses) == 60 and losses[-1] < 0.2 * losses[0]


@pytest.mark.parametrize("kind", ["pinn", "ritz"])
def test_reference_poisson_examples_run_unchanged(kind):
    """examples/poisson_equation.py = the reference's examples/pinn/poisson-equation.ipynb and
    examples/deepritz/poisson-equation.ipynb (disc with a square hole, `D.boundary`, density sampling)."""
    import importlib.util
    import os
    import warnings
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("poisson_equation", os.path.join(root, "examples", "poisson_equation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        losses, err = mod.main(kind, 150)
    assert losses[-1] < 0.3 * losses[0] and err < 1.0


@pytest.mark.parametrize("shape", [(1000, 1), (4097, 3), (7, 333, 2), (1, 1), (100003, 1)])
def test_native_squared_error_mean(shape):
    """tpx_squared_error_mean / _backward (the PINN loss reduction, reference condition.py:11-27 + 488-489) against the torch
    expression the reference uses, values and gradients, incl. a residual that is not 16-byte aligned."""
    from torchphysics_b200.conditions import SquaredError, _SquaredErrorMean
    torch.manual_seed(0)
    base = torch.randn(int(np.prod(shape)) + 1, device="cuda")
    for off in (0, 1):
        r = base[off:off + int(np.prod(shape))].reshape(shape).clone() if off == 0 else base[1:].reshape(shape)
        a = r.detach().clone().requires_grad_(True)
        b = r.detach().requires_grad_(True) if off == 0 else None
        ref = torch.mean(SquaredError()(a))
        (3.0 * ref).backward()
        src = b if b is not None else base.detach().clone().requires_grad_(True)
        rr = src if b is not None else src[1:].reshape(shape)
        out = _SquaredErrorMean.apply(rr)
        (3.0 * out).backward()
        assert out.shape == ref.shape and out.dtype == torch.float32
        assert abs(float(out) - float(ref)) <= 2e-6 * abs(float(ref))
        g = src.grad if b is not None else src.grad[1:].reshape(shape)
        assert torch.allclose(g, a.grad, rtol=1e-6, atol=0)


def test_operators_on_expressions_of_a_native_output():
    """laplacian(u * g, x), grad(u ** 2, x): the reference differentiates any autograd graph
    (differentialoperators.py:11-66); here pointwise expressions of a native output carry derivative streams
    (jets.ExprEvaluation).  Oracle: the same residual through torch.autograd in float64 on the CPU."""
    import torchphysics_b200 as tp
    G = load("poisson_4x64")
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=(64,) * 4).cuda()
    _load_params(model, G["params"])
    xs = torch.from_numpy(G["x"])
    sampler = tp.samplers.DataSampler(tp.spaces.Points(xs, X))

    def weight(x):
        return x[:, :1] * (1.0 - x[:, :1]) * torch.sin(math.pi * x[:, 1:])

    def residual(u, x):
        v = u * weight(x)                                  # hard-constraint style ansatz written inline
        return tp.utils.laplacian(v, x) + (tp.utils.grad(u ** 2, x) ** 2).sum(dim=1, keepdim=True)

    cond = tp.conditions.PINNCondition(model, sampler, residual)
    for it in range(2):
        model.zero_grad(set_to_none=True)
        loss = cond(device="cuda")
        loss.backward()
    # oracle
    seq = ref_autograd.fcn_from_params(G["params"], dtype=torch.float64)
    x = xs.double().requires_grad_(True)
    u = seq(x)
    v = u * weight(x)
    r = ref_autograd.laplacian(v, x) + (ref_autograd.grad(u ** 2, x) ** 2).sum(dim=1, keepdim=True)
    loss_ref = (r ** 2).sum(dim=1).mean()
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) < TOL * abs(float(loss_ref))
    ref_grads = [p.grad.numpy() for p in ref_autograd.linear_params(seq)]
    for a, w in zip(_grads(model), ref_grads):
        assert a is not None and relerr(a, w) < TOL


def test_adaptive_activation_function_is_native():
    """tp.models.AdaptiveActivationFunction (reference activation_fn.py:5-38): tanh(scaling * a * z) with a trainable a.
    Parity of loss, dLoss/dtheta and dLoss/da with the same network through torch.autograd in float64."""
    import torchphysics_b200 as tp
    torch.manual_seed(11)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    act = tp.models.AdaptiveActivationFunction(torch.nn.Tanh(), inital_a=0.8, scaling=1.5)
    model = tp.models.FCN(X, U, hidden=(48, 48, 48), activations=act).cuda()
    xs = torch.rand(311, 2)
    cond = tp.conditions.PINNCondition(model, tp.samplers.DataSampler(tp.spaces.Points(xs, X)),
                                       lambda u, x: tp.utils.laplacian(u, x) + u * u - tp.utils.grad(u, x)[:, :1])
    for it in range(2):
        model.zero_grad(set_to_none=True)
        loss = cond(device="cuda")
        loss.backward()
    assert model._native_engine() is not False
    # oracle: the same layers in float64 on the CPU
    lins = [m for m in model.sequential if isinstance(m, torch.nn.Linear)]
    Ws = [l.weight.detach().cpu().double().requires_grad_(True) for l in lins]
    bs = [l.bias.detach().cpu().double().requires_grad_(True) for l in lins]
    a = torch.tensor(float(act.a), dtype=torch.float64, requires_grad=True)
    x = xs.double().requires_grad_(True)
    h = x
    for i in range(len(lins) - 1):
        h = torch.tanh(1.5 * a * (h @ Ws[i].T + bs[i]))
    u = h @ Ws[-1].T + bs[-1]
    r = ref_autograd.laplacian(u, x) + u * u - ref_autograd.grad(u, x)[:, :1]
    loss_ref = (r ** 2).sum(dim=1).mean()
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) < TOL * abs(float(loss_ref))
    for l, W, b in zip(lins, Ws, bs):
        assert relerr(l.weight.grad.cpu().numpy(), W.grad.numpy()) < TOL
        assert relerr(l.bias.grad.cpu().numpy(), b.grad.numpy()) < TOL
    assert abs(float(act.a.grad) - float(a.grad)) < TOL * abs(float(a.grad))


def test_trainer_replays_the_whole_step_as_one_cuda_graph():
    """tp.solver.Trainer(cuda_graph=True): sampler (device-resident Philox key) -> jets -> residual -> loss -> reverse ->
    Adam (device-resident step count) captured once and replayed (reference solver.py:100-136 runs ~1,600 ATen ops per
    step).  The replayed training must behave like the eager one: same loss trajectory up to the sampling noise, the
    parameters change, the optimizer's step counts are written back."""
    import torchphysics_b200 as tp

    def train(graph):
        torch.manual_seed(5)
        X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
        sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
        model = tp.models.FCN(X, U, hidden=(32, 32, 32))
        pde = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq, n_points=4000),
                                          lambda u, x, f: tp.utils.laplacian(u, x) + f,
                                          data_functions={"f": ref_autograd.poisson_source})
        bc = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq.boundary, n_points=400), lambda u: u, weight=10.0)
        solver = tp.solver.Solver([pde, bc], optimizer_setting=tp.solver.OptimizerSetting(torch.optim.Adam, lr=2e-3))
        trainer = tp.solver.Trainer(max_steps=150, cuda_graph=graph)
        trainer.fit(solver)
        return trainer, [float(l) for l in trainer.losses]

    eager, le = train(False)
    graphed, lg = train(True)
    assert graphed.graphed and not eager.graphed and len(lg) == len(le) == 150
    assert lg[:3] == pytest.approx(le[:3], rel=1e-5)          # the eager prefix is the same computation
    assert lg[-1] < 0.2 * lg[0] and le[-1] < 0.2 * le[0]
    tail_e, tail_g = np.mean(le[-20:]), np.mean(lg[-20:])
    assert abs(tail_g - tail_e) < 0.5 * tail_e                 # same training up to the sampling noise
    st = graphed.optimizer.state[next(iter(graphed.optimizer.state))]
    assert int(st["step"]) == 150


def test_micro_batched_condition_equals_the_single_batch(monkeypatch):
    """A batch whose checkpoints exceed the HBM budget is evaluated slice by slice (forward saving -> reverse per slice,
    gradients accumulated): same loss and dLoss/dtheta as the batch in one piece."""
    import torchphysics_b200 as tp
    from torchphysics_b200 import conditions, engine
    G = load("heat_3x128")
    X, T, U = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("u")
    model = tp.models.FCN(X * T, U, hidden=(128,) * 3).cuda()
    _load_params(model, G["params"])
    pts = torch.rand(3001, 3)
    sampler = tp.samplers.DataSampler(tp.spaces.Points(pts, X * T))
    cond = tp.conditions.PINNCondition(model, sampler, lambda u, x, t: tp.utils.grad(u, t) - tp.utils.laplacian(u, x))

    def run():
        out = None
        for _ in range(2):                        # the first call learns the jet spec
            model.zero_grad(set_to_none=True)
            loss = cond(device="cuda")
            loss.backward()
            out = (float(loss.detach()), [None if p.grad is None else p.grad.clone() for p in model.parameters()])
        return out

    loss1, g1 = run()
    monkeypatch.setattr(engine, "save_budget_bytes", lambda device: 4 << 20)          # forces ~5 slices
    monkeypatch.setattr(engine, "_ALWAYS_SAVED", 0)                                   # (small buffers normally skip the budget query)
    cond.__dict__.pop("_micro_batch_cache", None)                                     # (the slicing is decided once per batch shape)
    assert cond._micro_batches(sampler.sample_points(device="cuda")) > 1
    loss2, g2 = run()
    assert abs(loss1 - loss2) < 1e-6 * abs(loss1)
    for a, b in zip(g1, g2):
        assert (a is None) == (b is None)                 # the output bias keeps grad None (reference parity)
        if a is not None:
            assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max())


@pytest.mark.parametrize("hidden", [(48,) * 3, (128,) * 3])            # width-64 kernels / wide path
def test_heat_equation_in_three_space_dimensions(hidden):
    """u_t = Lap_xyz u needs four derivative directions, one more than a jet carries (TPX_MAX_ND = 3): the time derivative
    comes from a second native evaluation of the same network on the same points (jets.JetEvaluation.extra) instead of a
    TpxError.  Loss and dLoss/dtheta against the reference's definition with torch autograd on the same modules."""
    import torchphysics_b200 as tp
    torch.manual_seed(2)
    X, T, U = tp.spaces.R3("x"), tp.spaces.R1("t"), tp.spaces.R1("u")
    model = tp.models.FCN(X * T, U, hidden=hidden).cuda()
    pts = torch.rand(2000, 4)
    sampler = tp.samplers.DataSampler(tp.spaces.Points(pts, X * T))
    cond = tp.conditions.PINNCondition(model, sampler, lambda u, x, t: tp.utils.grad(u, t) - 0.3 * tp.utils.laplacian(u, x))
    for it in range(3):                                                  # the first call learns the jet spec
        model.zero_grad(set_to_none=True)
        loss = cond(device="cuda")
        loss.backward()
    got = _grads(model)
    model.zero_grad(set_to_none=True)
    x = pts[:, :3].cuda().requires_grad_(True)
    t = pts[:, 3:].cuda().requires_grad_(True)
    u = model.sequential(torch.cat([x, t], dim=1))
    g = torch.autograd.grad(u.sum(), x, create_graph=True)[0]
    lap = sum(torch.autograd.grad(g[:, i].sum(), x, create_graph=True)[0][:, i:i + 1] for i in range(3))
    ut = torch.autograd.grad(u.sum(), t, create_graph=True)[0]
    want = ((ut - 0.3 * lap) ** 2).mean()
    want.backward()
    assert abs(float(loss) - float(want)) < TOL * abs(float(want))
    for a, w in zip(got[:-1], _grads(model)[:-1]):
        assert relerr(a, w) < TOL
