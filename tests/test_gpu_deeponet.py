"""SURVEY 8(a) rows a12/a13 on the GPU: PIDeepONetCondition through the native trunk jets
(value / d_t / Laplacian streams of the trunk, one GEMM per stream against the branch
coefficients, one native reverse launch) against the golden outputs of the real reference.
Tolerance 1e-5 max-norm relative (BASELINE.json)."""
import pytest
import torch

from tests.golden_io import load, relerr
from tests.test_host_deeponet import build_deeponet, check_deeponet_step

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("name", ["deeponet_2x32", "deeponet_4x64"])
def test_pideeponet_condition_matches_reference(name):
    import torchphysics_b200 as tp
    G = load(name)
    net, trunk, branch, fsampler, tsampler = build_deeponet(tp, G, device="cuda")
    check_deeponet_step(tp, G, net, trunk, branch, fsampler, tsampler, "cuda", TOL)
    from torchphysics_b200.engine import FcnEngine
    assert isinstance(trunk._native_engine(), FcnEngine)        # the kernels ran, not a torch fallback


def test_deeponet_forward_shapes_and_fixed_branch():
    """reference tests/tests_models/test_deep_o_net.py:113-154 (shapes), on the native trunk."""
    import torchphysics_b200 as tp
    T = tp.spaces.R1("t")
    fspace = tp.spaces.FunctionSpace(T, tp.spaces.R1("f"))
    grid = torch.rand((1, 100, 1))
    for out_space, width in ((tp.spaces.R1("u"), 1), (tp.spaces.R2("u"), 2)):
        trunk = tp.models.FCTrunkNet(T, default_trunk_input=grid)
        branch = tp.models.FCBranchNet(fspace, grid=grid)
        net = tp.models.DeepONet(trunk, branch, output_space=out_space, output_neurons=20).cuda()
        pts = tp.spaces.Points(torch.tensor([[[2.0], [0.0], [3.4], [2.9]]]).cuda(), T)
        out = net(pts, lambda t: 20 * t, device="cuda")
        assert "u" in out.space and out.as_tensor.shape == (1, 4, width)
        # value equals the torch-op definition of the same modules
        want = net._generic(pts)
        assert relerr(out.as_tensor.detach().cpu().numpy(), want.detach().cpu().numpy()) < TOL
        net.fix_branch_input(lambda t: torch.sin(t), device="cuda")
        assert net(pts).as_tensor.shape == (1, 4, width)
        assert net().as_tensor.shape == (1, 100, width)            # default trunk input
        assert trunk(pts).shape == (1, 4, width, 20)


def test_deeponet_uncopied_trunk_input():
    """trunk_input_copied=False: every function has its own trunk points (plain nn.Linear trunk)."""
    import torchphysics_b200 as tp
    torch.manual_seed(3)
    T, U = tp.spaces.R1("t"), tp.spaces.R1("u")
    fspace = tp.spaces.FunctionSpace(T, tp.spaces.R1("f"))
    grid = torch.rand((1, 12, 1))
    trunk = tp.models.FCTrunkNet(T, default_trunk_input=grid, hidden=(16, 16), trunk_input_copied=False)
    branch = tp.models.FCBranchNet(fspace, grid=grid, hidden=(8,))
    net = tp.models.DeepONet(trunk, branch, U, output_neurons=8).cuda()
    net.fix_branch_input(torch.rand(3, 12, 1).cuda())
    pts = tp.spaces.Points(torch.rand(3, 7, 1).cuda(), T)
    coords, pts = pts.track_coord_gradients()
    out = net(pts).as_tensor
    lap = tp.utils.laplacian(out, coords["t"])
    gen = net._generic(pts)
    g = torch.autograd.grad(gen.sum(), coords["t"], create_graph=True)[0]
    lap_ref = torch.autograd.grad(g.sum(), coords["t"])[0]
    assert relerr(out.detach().cpu().numpy(), gen.detach().cpu().numpy()) < TOL
    assert relerr(lap.detach().cpu().numpy(), lap_ref.cpu().numpy()) < TOL


def test_deeponet_full_size_equals_generic_autograd_definition():
    """config 4 at BASELINE.json's size (256 functions x 16,384 trunk points, p = 64, trunk 4x64): the native
    path (trunk jets once per point + one GEMM per stream) against the reference's definition evaluated with
    torch autograd on the SAME modules on the GPU (the trunk on all F x N copies, two create_graph sweeps)."""
    import math
    import torchphysics_b200 as tp
    torch.manual_seed(0)
    F, N, p = 256, 16384, 64
    T, U, Fs, K = tp.spaces.R1("t"), tp.spaces.R1("u"), tp.spaces.R1("f"), tp.spaces.R1("k")
    fspace = tp.spaces.FunctionSpace(T, Fs)
    ival = tp.domains.Interval(T, 0, 1)
    grid = tp.samplers.GridSampler(ival, n_points=50).sample_points(device="cuda")
    trunk = tp.models.FCTrunkNet(T, hidden=(64,) * 4, default_trunk_input=grid)
    branch = tp.models.FCBranchNet(fspace, hidden=(64,) * 4, grid=grid)
    net = tp.models.DeepONet(trunk, branch, U, output_neurons=p).cuda()
    fset = tp.domains.CustomFunctionSet(fspace, tp.samplers.RandomUniformSampler(tp.domains.Interval(K, -1, 1), n_points=F),
                                        lambda k, t: k * torch.sin(math.pi * t))
    fsampler = tp.samplers.FunctionSamplerRandomUniform(F, fset, function_creation_interval=1000)
    tsampler = tp.samplers.RandomUniformSampler(ival, n_points=N).make_static()
    cond = tp.conditions.PIDeepONetCondition(net, fsampler, tsampler, lambda u, t, f: tp.utils.laplacian(u, t) + f)
    for _ in range(2):
        net.zero_grad(set_to_none=True)
        torch.manual_seed(1)                                  # same function indices in every call
        loss = cond(device="cuda")
        loss.backward()
    got = [q.grad.clone() for q in net.parameters() if q.grad is not None]
    # the reference's definition on the same modules
    net.zero_grad(set_to_none=True)
    t = tsampler.sample_points(device="cuda").as_tensor
    tt = t.unsqueeze(0).repeat(F, 1, 1).requires_grad_(True)
    pts = tp.spaces.Points(tt, T)
    net.fix_branch_input(fset.get_function(fsampler.current_indices)(net.branch.grid), device="cuda")
    u = net._generic(pts)
    g = torch.autograd.grad(u.sum(), tt, create_graph=True)[0]
    lap = torch.autograd.grad(g.sum(), tt, create_graph=True)[0]
    k = fset.param_samples.as_tensor[fsampler.current_indices.cuda()]
    f = k.unsqueeze(1) * torch.sin(math.pi * tt)
    want = ((lap + f) ** 2).sum(dim=(1, 2)).mean()
    want.backward()
    assert abs(float(loss) - float(want)) < TOL * abs(float(want))
    ref = [q.grad for q in net.parameters() if q.grad is not None]
    assert len(ref) == len(got)
    for a, w in zip(got, ref):
        assert relerr(a.cpu().numpy(), w.cpu().numpy()) < 2e-5      # two fp32 evaluations, 4.2e6-term sums


def test_folded_deeponet_with_two_outputs_and_2d_trunk_input():
    """the branch coefficients folded into the trunk's output layer (F * o = 80 (function, output) columns: the wide output
    layer of csrc/tcw_kernels.cu, output-major streams) for a DeepONet with TWO outputs on a 2-D trunk input: value, gradient
    and Laplacian streams and dLoss/dtheta against the reference's definition with torch autograd on the same modules, and
    against the unfolded native path (trunk with its own output layer + torch contraction)."""
    import torchphysics_b200 as tp
    from torchphysics_b200 import jets
    torch.manual_seed(3)
    F, N, p = 40, 3001, 24
    X, U, Fs = tp.spaces.R2("x"), tp.spaces.R2("u"), tp.spaces.R1("f")
    fspace = tp.spaces.FunctionSpace(X, Fs)
    grid = torch.rand(1, 30, 2)
    trunk = tp.models.FCTrunkNet(X, hidden=(48, 48, 48), default_trunk_input=grid)
    branch = tp.models.FCBranchNet(fspace, hidden=(32, 32), grid=grid)
    net = tp.models.DeepONet(trunk, branch, U, output_neurons=p).cuda()
    sensors = torch.randn(F, 30, 1, device="cuda")
    x = torch.rand(N, 2, device="cuda")

    def native():
        net.zero_grad(set_to_none=True)
        pts = tp.spaces.Points(x.unsqueeze(0).expand(F, N, 2), X)
        coords, pts = pts.track_coord_gradients()
        with jets.scope("t"):
            u = net(pts, tp.spaces.Points(sensors, Fs), device="cuda").as_tensor
            lap = tp.utils.laplacian(u[..., :1], coords["x"]) + 0.5 * tp.utils.laplacian(u[..., 1:], coords["x"])
            g = tp.utils.grad(u[..., 1:], coords["x"])
            r = lap + g.sum(dim=-1, keepdim=True) * u[..., :1].as_subclass(torch.Tensor) + u[..., 1:].as_subclass(torch.Tensor) ** 2
        loss = (r ** 2).mean()
        loss.backward()
        return (u.as_subclass(torch.Tensor).detach().clone(), lap.detach().clone(), float(loss),
                [q.grad.clone() for q in net.parameters() if q.grad is not None])

    native()                                                  # the first call learns the jet spec
    u1, lap1, loss1, g1 = native()
    folded = net._folded[F * 2]
    assert folded is not False and folded.out_major            # the folded tensor-core path ran
    net._folded[F * 2] = False                                 # unfolded: trunk output layer + torch contraction
    net.__dict__.pop("_jet_hints", None)
    native()
    u2, lap2, loss2, g2 = native()
    net._folded.pop(F * 2)
    assert relerr(u1.cpu().numpy(), u2.cpu().numpy()) < TOL and relerr(lap1.cpu().numpy(), lap2.cpu().numpy()) < TOL
    assert abs(loss1 - loss2) < TOL * abs(loss2)
    for a, b in zip(g1, g2):
        assert relerr(a.cpu().numpy(), b.cpu().numpy()) < TOL
    # the reference's definition: the same modules through torch ops and two create_graph sweeps
    net.zero_grad(set_to_none=True)
    xx = x.unsqueeze(0).repeat(F, 1, 1).requires_grad_(True)
    net.fix_branch_input(tp.spaces.Points(sensors, Fs), device="cuda")
    u = net._generic(tp.spaces.Points(xx, X))

    def lap_of(v):
        gv = torch.autograd.grad(v.sum(), xx, create_graph=True)[0]
        return sum(torch.autograd.grad(gv[..., i].sum(), xx, create_graph=True)[0][..., i:i + 1] for i in range(2)), gv

    l0, _ = lap_of(u[..., :1])
    l1, gv1 = lap_of(u[..., 1:])
    lap = l0 + 0.5 * l1
    r = lap + gv1.sum(dim=-1, keepdim=True) * u[..., :1] + u[..., 1:] ** 2
    want = (r ** 2).mean()
    want.backward()
    assert relerr(u1.cpu().numpy(), u.detach().cpu().numpy()) < TOL
    assert relerr(lap1.cpu().numpy(), lap.detach().cpu().numpy()) < TOL
    assert abs(loss1 - float(want)) < TOL * abs(float(want))
    ref = [q.grad for q in net.parameters() if q.grad is not None]
    assert len(ref) == len(g1)
    for a, w in zip(g1, ref):
        assert relerr(a.cpu().numpy(), w.cpu().numpy()) < TOL
