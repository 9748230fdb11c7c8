"""DeepONet host logic without a GPU: module construction and
initialisation parity, function sets / function samplers, and the PIDeepONetCondition
orchestration + stream contraction with a stand-in trunk engine (torch restatement of the jet
propagation on CPU -- test infrastructure, the product engine only runs on CUDA)."""
import math

import numpy as np
import pytest
import torch

from tests.golden_io import load, relerr


class _TorchJetEngine:
    """stand-in for FcnEngine.jet: value / d_k / weighted-Laplacian streams with torch ops."""

    def __init__(self, lins, d_out=None):
        self.lins, self.d_out = lins, lins[-1].out_features if d_out is None else d_out

    def params(self):
        return [p for lin in self.lins for p in (lin.weight, lin.bias)]

    def with_outputs(self, n_cols, device):
        """FcnEngine.with_outputs: same hidden layers, the caller's output layer of n_cols columns"""
        return _TorchJetEngine(self.lins, d_out=n_cols)

    def jet(self, x, spec, params=None):
        if params is not None:
            class _L:
                def __init__(self, w, b):
                    self.weight, self.bias = w, b
            return _TorchJetEngine([_L(params[2 * i], params[2 * i + 1]) for i in range(len(params) // 2)], self.d_out).jet(x, spec)
        n, nd = x.shape[0], spec.nd
        h = x
        d = [torch.zeros_like(x) for _ in range(nd)]
        for k, c in enumerate(spec.dcols):
            d[k][:, c] = 1.0
        L = torch.zeros_like(x) if spec.lap else None
        for i, lin in enumerate(self.lins):
            z = h @ lin.weight.t() + lin.bias
            dz = [dk @ lin.weight.t() for dk in d]
            Lz = L @ lin.weight.t() if spec.lap else None
            if i == len(self.lins) - 1:
                h, d, L = z, dz, Lz
                break
            t = torch.tanh(z)
            s1 = 1 - t * t
            s2 = -2 * t * s1
            if spec.lap:
                L = s2 * sum(w * dk * dk for w, dk in zip(spec.lap_w, dz)) + s1 * Lz
            d = [s1 * dk for dk in dz]
            h = t
        du = torch.stack(d, dim=-1) if nd else x.new_zeros((n, self.d_out, 0))
        return h, du, (L if spec.lap else x.new_zeros((n, 0)))


def _set_params(seq, params):
    lins = [m for m in seq if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(torch.as_tensor(params[2 * i]))
            lin.bias.copy_(torch.as_tensor(params[2 * i + 1]))
    return lins


class _FixedParams:
    """parameter sampler stand-in: always the golden k values."""

    def __init__(self, k, space):
        self.k, self.space, self.n_points = k, space, len(k)

    def sample_points(self, device="cpu"):
        import torchphysics_b200 as tp
        return tp.spaces.Points(self.k.to(device), self.space)


def build_deeponet(tp, G, device="cpu"):
    T, U, Fs, K = tp.spaces.R1("t"), tp.spaces.R1("u"), tp.spaces.R1("f"), tp.spaces.R1("k")
    fspace = tp.spaces.FunctionSpace(T, Fs)
    grid = tp.spaces.Points(torch.from_numpy(G["sensors"]), T)
    trunk = tp.models.FCTrunkNet(T, hidden=tuple(int(h) for h in G["trunk_hidden"]), default_trunk_input=grid)
    branch = tp.models.FCBranchNet(fspace, hidden=tuple(int(h) for h in G["branch_hidden"]), grid=grid)
    net = tp.models.DeepONet(trunk, branch, U, output_neurons=int(G["p"]))
    _set_params(trunk.sequential, G["trunk_params"])
    _set_params(branch.sequential, G["branch_params"])
    net = net.to(device)
    fset = tp.domains.CustomFunctionSet(fspace, _FixedParams(torch.from_numpy(G["k"]), K),
                                        lambda k, t: k * torch.sin(math.pi * t))

    class AllFunctions(tp.samplers.FunctionSampler):
        def sample_functions(self, device="cpu"):
            self._check_recreate_functions(device=device)
            self.current_indices = torch.arange(self.n_functions)
            return self.function_set.get_function(self.current_indices)

    fsampler = AllFunctions(len(G["k"]), fset, function_creation_interval=1)
    tsampler = tp.samplers.DataSampler(tp.spaces.Points(torch.from_numpy(G["t"]), T))
    return net, trunk, branch, fsampler, tsampler


def check_deeponet_step(tp, G, net, trunk, branch, fsampler, tsampler, device, tol):
    captured = {}

    def residual(u, t, f):
        captured.update(u=u, ut=tp.utils.grad(u, t), lap=tp.utils.laplacian(u, t), f=f)
        return captured["lap"] + f

    cond = tp.conditions.PIDeepONetCondition(net, fsampler, tsampler, residual)
    for it in range(2):                      # 1st call learns the jet spec
        net.zero_grad(set_to_none=True)
        loss = cond(device=device)
        loss.backward()
        assert relerr(captured["f"].detach().cpu().numpy(), G["fvals"]) < 2e-6
        assert relerr(branch.current_out.detach().cpu().numpy(), G["branch_out"]) < tol
        assert captured["u"].shape == tuple(G["u"].shape)
        assert relerr(captured["u"].detach().cpu().numpy(), G["u"]) < tol
        assert relerr(captured["ut"].detach().cpu().numpy(), G["ut"]) < tol
        assert relerr(captured["lap"].detach().cpu().numpy(), G["lap"]) < tol
        assert abs(float(loss) - float(G["loss"])) < tol * abs(float(G["loss"]))
        for mod, want in ((trunk, G["trunk_dtheta"]), (branch, G["branch_dtheta"])):
            for q, w in zip(mod.sequential.parameters(), want):
                if w is None:
                    assert q.grad is None or float(q.grad.abs().max()) == 0.0
                else:
                    assert relerr(q.grad.cpu().numpy(), w) < tol


@pytest.mark.parametrize("name", ["deeponet_2x32", "deeponet_4x64"])
def test_pideeponet_condition_orchestration_cpu(name):
    import torchphysics_b200 as tp
    G = load(name)
    net, trunk, branch, fsampler, tsampler = build_deeponet(tp, G)
    lins = [m for m in trunk.sequential if isinstance(m, torch.nn.Linear)]
    object.__setattr__(trunk, "_engine", _TorchJetEngine(lins))
    check_deeponet_step(tp, G, net, trunk, branch, fsampler, tsampler, "cpu", 1e-5)


def test_deeponet_initialisation_matches_reference():
    """same seed, same construction order -> the reference's parameters bit for bit."""
    import torchphysics_b200 as tp
    G = load("deeponet_2x32")
    torch.manual_seed(8)
    T, U, Fs = tp.spaces.R1("t"), tp.spaces.R1("u"), tp.spaces.R1("f")
    grid = tp.spaces.Points(torch.from_numpy(G["sensors"]), T)
    trunk = tp.models.FCTrunkNet(T, hidden=(32, 32), default_trunk_input=grid)
    branch = tp.models.FCBranchNet(tp.spaces.FunctionSpace(T, Fs), hidden=(24, 24), grid=grid)
    tp.models.DeepONet(trunk, branch, U, output_neurons=16)
    for q, w in zip(trunk.parameters(), G["trunk_params"]):
        assert np.array_equal(q.detach().numpy(), w)
    for q, w in zip(branch.parameters(), G["branch_params"]):
        assert np.array_equal(q.detach().numpy(), w)
    assert list(trunk.state_dict().keys())[:2] == ["sequential.0.weight", "sequential.0.bias"]
    assert "grid_buffer" in branch.state_dict()


def test_trunk_linear_equals_linear():
    """reference tests/tests_models/test_deep_o_net.py:165-184."""
    import torchphysics_b200 as tp
    a = tp.models.TrunkLinear(30, 20, bias=True)
    b = torch.nn.Linear(30, 20, bias=True)
    b.bias = torch.nn.Parameter(a.bias[:])
    b.weight = torch.nn.Parameter(a.weight[:])
    x = torch.randn(2, 4, 30).expand(5, -1, -1, -1)
    x.requires_grad = True
    ya, yb = a(x), b(x)
    ga = torch.autograd.grad(ya.sum(), x, retain_graph=True)[0]
    gb = torch.autograd.grad(yb.sum(), x, retain_graph=True)[0]
    assert torch.norm(ga - gb) < 1e-5
    ya.sum().backward()
    yb.sum().backward()
    assert torch.norm(a.weight.grad - b.weight.grad) < 1e-2
    assert torch.norm(a.bias.grad - b.bias.grad) < 1e-5


def test_function_samplers_and_collection():
    import torchphysics_b200 as tp
    T, Fs, K = tp.spaces.R1("t"), tp.spaces.R1("f"), tp.spaces.R1("k")
    fspace = tp.spaces.FunctionSpace(T, Fs)
    k = torch.arange(1.0, 7.0).reshape(6, 1)
    a = tp.domains.CustomFunctionSet(fspace, _FixedParams(k, K), lambda k, t: k * t)
    b = tp.domains.CustomFunctionSet(fspace, _FixedParams(k, K), lambda k, t: k * t ** 2)
    both = a.append(b)
    assert both.function_set_size == 12
    both.create_functions()
    loc = tp.spaces.Points(torch.tensor([[[0.5], [2.0]]]), T)
    out = both.get_function([1, 7])(loc).as_tensor
    assert out.shape == (2, 2, 1)
    assert torch.allclose(out[0, :, 0], torch.tensor([1.0, 4.0]))        # 2 * t
    assert torch.allclose(out[1, :, 0], torch.tensor([0.5, 8.0]))        # 2 * t^2
    ordered = tp.samplers.FunctionSamplerOrdered(3, a)
    ordered.sample_functions()
    assert ordered.current_indices.tolist() == [0, 0, 0]                 # reference function_sampler.py:78-85
    ordered.sample_functions()
    assert ordered.current_indices.tolist() == [3, 3, 3]
    rnd = tp.samplers.FunctionSamplerRandomUniform(4, a, function_creation_interval=2)
    f = rnd.sample_functions()
    assert f(loc).as_tensor.shape == (4, 2, 1)
    assert len(set(rnd.current_indices.tolist())) == 4
    coupled = tp.samplers.FunctionSamplerCoupled(b, rnd)
    coupled.sample_functions()
    assert coupled.current_indices is rnd.current_indices
    with pytest.raises(AssertionError):
        tp.samplers.FunctionSamplerRandomUniform(7, a)


def test_deeponet_refuses_cpu_evaluation():
    import torchphysics_b200 as tp
    from torchphysics_b200 import _lib
    G = load("deeponet_2x32")
    net, trunk, branch, fsampler, tsampler = build_deeponet(tp, G)
    net.fix_branch_input(lambda t: 20 * t)
    assert branch.current_out.shape == (1, 1, 16)
    with pytest.raises(_lib.TpxError):
        net(tp.spaces.Points(torch.tensor([[[2.0], [0.0]]]), tp.spaces.R1("t")))


def test_deeponet_step_leaves_no_evaluation_behind(monkeypatch):
    """every native evaluation of a step (trunk jets, the DeepONet wrapper around them) is released by reference counting when
    the step's loss goes away -- no cycle keeps batch-sized streams or the step's autograd graph alive."""
    import gc
    import weakref
    import torchphysics_b200 as tp
    from torchphysics_b200 import deeponet, jets
    created = []
    for cls in (jets.JetEvaluation, deeponet.DeepONetEvaluation):
        orig = cls.__init__

        def recording(self, *a, _orig=orig, **k):
            _orig(self, *a, **k)
            created.append(weakref.ref(self))
        monkeypatch.setattr(cls, "__init__", recording)
    G = load("deeponet_2x32")
    net, trunk, branch, fsampler, tsampler = build_deeponet(tp, G)
    object.__setattr__(trunk, "_engine", _TorchJetEngine([m for m in trunk.sequential if isinstance(m, torch.nn.Linear)]))
    cond = tp.conditions.PIDeepONetCondition(net, fsampler, tsampler, lambda u, t, f: tp.utils.laplacian(u, t) + tp.utils.grad(u, t) + f)
    gc.collect()
    gc.disable()
    try:
        for _ in range(2):
            net.zero_grad(set_to_none=True)
            loss = cond(device="cpu")
            loss.backward()
            del loss
        assert len(created) >= 4
        # what the modules legitimately keep is a tensor (the branch coefficients of the last fix_input), not an evaluation
        assert all(ref() is None for ref in created)
    finally:
        gc.enable()
