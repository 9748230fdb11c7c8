"""Host logic that needs no GPU: the C-ABI library loads and exports every symbol of
include/tpx.h, descriptor construction, spaces/Points/UserFunction semantics, condition
orchestration with a stand-in module, gradient all-reduce over gloo (world_size 2)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from torchphysics_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tpx.h")).read()
    declared = set(re.findall(r"\b(tpx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.tpx_version() == 100


def test_struct_layouts_match_header():
    from torchphysics_b200 import _lib
    assert ctypes.sizeof(_lib.FcnDesc) == 4 * (3 + 16 + 1)
    assert ctypes.sizeof(_lib.JetDesc) == 4 * (1 + 3 + 1 + 3 + 1)
    assert ctypes.sizeof(_lib.Shape) == 32
    assert ctypes.sizeof(_lib.Region) == 8 + 8 * 32 + 24 * 4


def test_argument_errors_without_gpu():
    from torchphysics_b200 import _lib
    lib = _lib.lib()
    d = _lib.FcnDesc()
    d.d_in, d.d_out, d.n_hidden = 2, 1, 0
    n = ctypes.c_size_t(0)
    assert lib.tpx_fcn_packed_floats(ctypes.byref(d), ctypes.byref(n)) == _lib.TPX_E_ARG
    assert b"n_hidden" in lib.tpx_last_error()
    d.n_hidden = 2
    d.widths[0], d.widths[1] = 64, 300
    assert lib.tpx_fcn_packed_floats(ctypes.byref(d), ctypes.byref(n)) == _lib.TPX_E_UNSUPPORTED
    d.widths[1] = 64
    assert lib.tpx_fcn_packed_floats(ctypes.byref(d), ctypes.byref(n)) == 0 and n.value > 64 * 64


def test_packed_layout_of_wide_output_and_deep_narrow_networks():
    """Layout rules of csrc/fcn_common.cuh (no compute): a narrow net the resident-weight kernels cannot hold (more than 5
    hidden layers, or more than 8 outputs) is laid out as one padded 128-neuron block per layer for the layer-major
    tensor-core kernels, and 9 or more outputs store the output layer like a hidden matrix (Wt | W | b, padded to 128)."""
    from torchphysics_b200 import _lib
    lib = _lib.lib()

    def floats(d_in, d_out, widths):
        d = _lib.FcnDesc()
        d.d_in, d.d_out, d.n_hidden = d_in, d_out, len(widths)
        for l, w in enumerate(widths):
            d.widths[l] = w
        d.activation = _lib.TPX_ACT_TANH
        n = ctypes.c_size_t(0)
        _lib.check(lib.tpx_fcn_packed_floats(ctypes.byref(d), ctypes.byref(n)))
        return n.value

    def fp32_image(hp, d_in, L, out_floats):             # (the tensor-core images behind it start 1 KB aligned)
        return (d_in * hp + hp + (L - 1) * (2 * hp * hp + hp) + out_floats + 255) // 256 * 256

    pad4 = lambda k: (k + 3) // 4 * 4                                             # noqa: E731
    # 6 x 64, one output: HP = 128, CUDA-core output layer (d_out x HP weights + padded bias), no width-64 tensor-core image
    assert floats(2, 1, (64,) * 6) == fp32_image(128, 2, 6, 128 + pad4(1))
    # DeepONet trunk 4 x 64 with 64 outputs: HP = 128, output layer as a 128 x 128 hidden-style block
    assert floats(1, 64, (64,) * 4) == fp32_image(128, 1, 4, 2 * 128 * 128 + 128)
    # 300 (function, output) columns: three blocks of 128 outputs
    assert floats(1, 300, (64,) * 4) == fp32_image(128, 1, 4, 2 * 128 * 384 + 384)
    # width 256: the same block layout with 256-wide rows (Wt [256][128] | W [128][256] | b [128] for 40 outputs)
    assert floats(2, 40, (256, 256)) == fp32_image(256, 2, 2, 2 * 256 * 128 + 128)
    assert floats(2, 8, (256, 256)) == fp32_image(256, 2, 2, 8 * 256 + pad4(8))          # up to 8 outputs: the plain output layer
    # the width-64 path appends its own images behind the (1 KB aligned) FP32 image: strictly more than the FP32 image
    assert floats(2, 1, (64,) * 4) > fp32_image(64, 2, 4, 64 + pad4(1))
    # checkpoint buffers of the layer-major kernels (host arithmetic only): per array = tiles x S x 16 x 128 floats; hidden layers
    # 1..L-1 + two adjoint arrays (+ one array of partial products for several output blocks, + layer 0's own checkpoints for
    # 5..8 inputs), then 160 per-CTA dW partial sums.  A size > 0 also says: this shape runs on the tensor-core kernels.
    def saved(d_in, d_out, widths, nd, lap, n):
        d = _lib.FcnDesc()
        d.d_in, d.d_out, d.n_hidden, d.activation = d_in, d_out, len(widths), _lib.TPX_ACT_TANH
        for l, w in enumerate(widths):
            d.widths[l] = w
        j = _lib.JetDesc()
        j.nd, j.lap = nd, lap
        for k in range(nd):
            j.dcol[k], j.lap_w[k] = k, 1.0
        nb = ctypes.c_size_t(0)
        _lib.check(lib.tpx_fcn_saved_bytes(ctypes.byref(d), ctypes.byref(j), ctypes.c_int64(n), ctypes.byref(nb)))
        return nb.value

    per = lambda n, S: (n + 15) // 16 * S * 16 * 128 * 4                               # noqa: E731
    tail = 160 * 128 * 128 * 4
    assert saved(2, 1, (64,) * 6, 2, 1, 1000) == (6 + 1) * per(1000, 4) + tail          # deep narrow net
    assert saved(1, 64, (64,) * 4, 1, 1, 1000) == (4 + 1) * per(1000, 3) + tail         # DeepONet trunk, one block of outputs
    assert saved(1, 300, (64,) * 4, 1, 1, 1000) == (4 + 1 + 1) * per(1000, 3) + tail    # three blocks: + partial products
    assert saved(6, 2, (64,) * 3, 3, 1, 1000) == (3 + 1 + 1) * per(1000, 5) + tail      # 6 inputs: + layer 0's checkpoints
    assert saved(6, 2, (200, 256), 2, 0, 1000) == ((2 + 1) * 2 + 1 + 2) * per(1000, 3) + tail   # ... with two 128-neuron blocks
    assert saved(2, 300, (256, 256), 2, 1, 1000) == ((2 + 1) * 2 + 1) * per(1000, 4) + tail     # 256-wide with three output blocks
    d = _lib.FcnDesc()
    d.d_in, d.d_out, d.n_hidden, d.activation = 2, _lib.TPX_MAX_DOUT + 1, 2, _lib.TPX_ACT_TANH
    d.widths[0] = d.widths[1] = 64
    n = ctypes.c_size_t(0)
    assert lib.tpx_fcn_packed_floats(ctypes.byref(d), ctypes.byref(n)) == _lib.TPX_E_UNSUPPORTED


def test_region_programs():
    import torchphysics_b200 as tp
    from torchphysics_b200 import _lib
    X, T = tp.spaces.R2("x"), tp.spaces.R1("t")
    c = tp.domains.Circle(X, [0, 0], 1.0)
    sq = tp.domains.Parallelogram(X, [-.25, -.25], [.25, -.25], [-.25, .25])
    iv = tp.domains.Interval(T, 0, 2)
    S, A, O, N = _lib.TPX_OP_SHAPE, _lib.TPX_OP_AND, _lib.TPX_OP_OR, _lib.TPX_OP_NOT
    shapes, ops = (c - sq)._program(0)
    assert [s.kind for s in shapes] == [3, 2] and ops == [S, S | 1 << 8, N, A]
    shapes, ops = ((c + sq) * iv)._program(0)
    assert [s.col for s in shapes] == [0, 0, 2] and ops == [S, S | 1 << 8, O, S | 2 << 8, A]
    gen = ((c - sq) - tp.domains.Circle(X, [.5, .5], .1))._generator(0)
    assert gen[0].kind == 3 and gen[2] == [S, N, S | 1 << 8, N, A]
    assert (c + sq)._generator(0) is None
    assert abs(float((c - sq).bounding_box()[0]) + 1) < 1e-7
    assert (sq * iv).dim == 3 and list((sq * iv).space.keys()) == ["x", "t"]
    assert abs(float((sq * iv).volume()) - 0.5) < 1e-6
    assert c.compute_n_from_density(10.0) == 32


def test_points_and_spaces():
    import torchphysics_b200 as tp
    X, T = tp.spaces.R2("x"), tp.spaces.R1("t")
    sp = X * T
    assert sp.dim == 3 and list(sp.keys()) == ["x", "t"] and "x" in sp and X in sp
    p = tp.spaces.Points(torch.arange(12.0).reshape(4, 3), sp)
    assert len(p) == 4 and p.coordinates["t"].shape == (4, 1)
    assert p[:, ["t", "x"]].as_tensor[0].tolist() == [2.0, 0.0, 1.0]
    assert p[1:3].as_tensor.shape == (2, 3)
    q = tp.spaces.Points(torch.ones(4, 1), tp.spaces.R1("u"))
    assert list(p.join(q).space.keys()) == ["x", "t", "u"]
    assert len(p | p) == 8
    coords, tracked = p.track_coord_gradients()
    assert coords["x"].requires_grad and tracked.as_tensor.requires_grad and set(tracked._coord_leaves) == {"x", "t"}
    with pytest.raises(AssertionError):
        tp.spaces.Points(torch.ones(4, 2), sp)


def test_integer_spaces_and_function_space_product():
    """reference spaces/space.py:189-251 (cast_tensor_into_space) and functionspace.py (product)."""
    import torchphysics_b200 as tp
    z = tp.spaces.Z2("k")
    assert z.dim == 2 and z.cast_tensor_into_space(torch.tensor([[-1.6, 2.4]])).tolist() == [[-2.0, 2.0]]
    n = tp.spaces.N1("n")
    assert n.cast_tensor_into_space(torch.tensor([[-1.6]])).tolist() == [[2.0]]
    T, X = tp.spaces.R1("t"), tp.spaces.R2("x")
    fa = tp.spaces.FunctionSpace(T, tp.spaces.R1("f"))
    fb = tp.spaces.FunctionSpace(T * X, tp.spaces.R1("g"))
    prod = fa * fb
    assert list(prod.output_space.keys()) == ["f", "g"] and list(prod.input_space.keys()) == ["t", "x"]
    with pytest.raises(AssertionError):
        fa * fa


def test_user_function_binds_by_name():
    import torchphysics_b200 as tp
    f = tp.utils.UserFunction(lambda u, x, k=2.0: u + k * x)
    assert f.necessary_args == ["u", "x"] and f.optional_args == ["k"]
    out = f({"x": torch.ones(2, 1), "u": torch.zeros(2, 1), "unused": 5})
    assert out.tolist() == [[2.0], [2.0]]
    with pytest.raises(AssertionError):
        f({"x": torch.ones(2, 1)})
    with pytest.raises(ValueError):
        tp.utils.UserFunction(lambda *args: 0)


def test_condition_orchestration_with_stand_in_module():
    """reference tests/test_conditions.py style: the module is a UserFunction, so the
    condition's sample -> track -> data -> module -> residual -> error -> reduce chain runs
    without a GPU (the generic autograd definition of the operators)."""
    import torchphysics_b200 as tp
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")

    class Sq(torch.nn.Module):
        def forward(self, points):
            x = points.coordinates["x"]
            return tp.spaces.Points((x ** 2).sum(dim=1, keepdim=True), U)

    pts = tp.spaces.Points(torch.tensor([[0.0, 0.0], [1.0, 0.5], [0.5, 2.0]]), X)
    sampler = tp.samplers.DataSampler(pts)
    cond = tp.conditions.PINNCondition(Sq(), sampler, lambda u, x, f: tp.utils.laplacian(u, x) - f,
                                       data_functions={"f": lambda x: 3.0 * torch.ones_like(x[:, :1])})
    loss = cond()
    assert abs(float(loss) - 1.0) < 1e-6           # (4 - 3)^2
    ritz = tp.conditions.DeepRitzCondition(Sq(), sampler, lambda u, x: 0.5 * (tp.utils.grad(u, x) ** 2).sum(1, keepdim=True))
    want = float((2.0 * (pts.as_tensor ** 2).sum(1)).mean())
    assert abs(float(ritz()) - want) < 1e-6
    assert tp.conditions.SquaredError()(torch.ones(3, 2, 2)).tolist() == [4.0, 4.0, 4.0]


def test_fcn_layout_matches_reference_state_dict():
    import torchphysics_b200 as tp
    from tests.golden_io import load
    G = load("poisson_4x64")
    torch.manual_seed(0)
    m = tp.models.FCN(tp.spaces.R2("x"), tp.spaces.R1("u"), hidden=(64,) * 4)
    keys = list(m.state_dict().keys())
    assert keys == [f"sequential.{2 * i}.{k}" for i in range(5) for k in ("weight", "bias")]
    for p, w in zip(m.parameters(), G["params"]):
        assert np.array_equal(p.detach().numpy(), w)        # same RNG draw order as the reference


def _gloo_worker(rank, size, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from torchphysics_b200 import parallel
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    unused = torch.nn.Parameter(torch.zeros(2))
    x = torch.arange(24.0).reshape(8, 3)
    n = parallel.shard_count(8)
    lo = sum(parallel.shard_count(8, r, size) for r in range(rank))
    loss = (lin(x[lo:lo + n]) ** 2).mean()
    loss.backward()
    red = parallel.FlatGradAllReduce([lin.weight, lin.bias, unused], n_scalars=1)
    (mean_loss,) = red([loss])
    out.put((rank, lin.weight.grad.clone(), float(mean_loss), unused.grad is None))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    x = torch.arange(24.0).reshape(8, 3)
    loss = (lin(x) ** 2).mean()
    loss.backward()
    for rank, g, l, unused_none in res:
        assert torch.allclose(g, lin.weight.grad, rtol=1e-5, atol=1e-5)
        assert abs(l - float(loss)) < 1e-4 * abs(float(loss))
        assert unused_none


def test_constrained_evaluation_chain_rule_cpu():
    """jets.ConstrainedEvaluation (the HardConstraint path) against the real reference's outputs, with a torch
    stand-in for the jet engine (the chain rule itself is host logic)."""
    import torchphysics_b200 as tp
    from torchphysics_b200 import jets
    from torchphysics_b200.engine import JetSpec
    from torchphysics_b200.userfn import UserFunction
    from tests.golden_io import load, relerr
    from tests.test_host_deeponet import _TorchJetEngine
    G = load("poisson_hard_3x32")
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    lins = []
    for i in range(len(G["params"]) // 2):
        W, b = torch.as_tensor(G["params"][2 * i]), torch.as_tensor(G["params"][2 * i + 1])
        lin = torch.nn.Linear(W.shape[1], W.shape[0])
        with torch.no_grad():
            lin.weight.copy_(W)
            lin.bias.copy_(b)
        lins.append(lin)
    x = torch.from_numpy(G["x"])
    leaf = x.clone().requires_grad_(True)
    inner = jets.JetEvaluation(_TorchJetEngine(lins), x, x.shape[:-1], {id(leaf): (leaf, [0, 1])}, {}, None, None,
                               JetSpec((0, 1), (1.0, 1.0)))
    g = UserFunction(lambda u, x: x[:, :1] * (1 - x[:, :1]) * x[:, 1:] * (1 - x[:, 1:]) * u + torch.sin(x[:, :1]))
    cev = jets.ConstrainedEvaluation(inner, g, {"x": leaf}, X, U, U)
    assert relerr(cev.value().detach().numpy(), G["u"]) < 1e-6
    grad = torch.cat([cev.first([0], 0), cev.first([0], 1)], dim=1)
    assert relerr(grad.detach().numpy(), G["grad"]) < 1e-6
    lap = cev.second([0])
    assert relerr(lap.detach().numpy(), G["lap"]) < 1e-6
    loss = ((lap + 1.0) ** 2).sum(dim=1).mean()
    loss.backward()
    for lin, (w, b) in zip(lins, zip(G["dtheta"][0::2], G["dtheta"][1::2])):
        assert relerr(lin.weight.grad.numpy(), w) < 5e-6 and relerr(lin.bias.grad.numpy(), b) < 5e-6


def test_points_column_selection_is_a_view_for_contiguous_names():
    """Points[..., names]: a contiguous column range is a slice (a view: its backward is a slice, not an index_put with
    accumulation), any other selection keeps advanced indexing; values as in the reference (points.py:150-215)."""
    import torchphysics_b200 as tp
    X, T, D = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("D")
    t = torch.arange(24.0).reshape(2, 3, 4).requires_grad_(True)
    p = tp.spaces.Points(t, X * T * D)
    sub = p[..., ["x", "t"]]
    assert sub.as_tensor.shape == (2, 3, 3) and sub.as_tensor._base is not None          # a view of t
    assert torch.equal(sub.as_tensor, t[..., :3])
    swapped = p[..., ["t", "x"]]                                                        # not a contiguous range of columns
    assert torch.equal(swapped.as_tensor, t[..., [2, 0, 1]])
    sub.as_tensor.sum().backward()
    assert torch.equal(t.grad[..., :3], torch.ones(2, 3, 3)) and float(t.grad[..., 3].abs().sum()) == 0.0


def test_convective_equals_batched_matrix_product():
    """(v . grad) u as a broadcast multiply + sum equals the reference's torch.bmm(jac, v) (differentialoperators.py:320-342)."""
    import torchphysics_b200 as tp
    x = torch.rand(50, 2, requires_grad=True)
    u = torch.stack([x[:, 0] ** 2 * x[:, 1], torch.sin(x[:, 0]) + x[:, 1] ** 3, x[:, 0] * x[:, 1]], dim=1)
    v = torch.rand(50, 2)
    got = tp.utils.convective(u, v, x)
    j = tp.utils.jac(u, x)
    want = torch.bmm(j, v.unsqueeze(2)).squeeze(2)
    assert got.shape == (50, 3) and torch.allclose(got, want, rtol=1e-6, atol=1e-7)


class _ToySolver(torch.nn.Module):
    """the slice of tp.solver.Solver that Trainer.fit drives, on CPU tensors (no native kernels involved)"""

    def __init__(self, seed):
        super().__init__()
        torch.manual_seed(seed)                       # every rank builds DIFFERENT initial weights on purpose
        self.net = torch.nn.Linear(3, 1)
        self.register_buffer("shift", torch.full((1,), float(seed)))
        self.n_training_step = 0

    def on_train_start(self):
        pass

    def configure_optimizers(self):
        return torch.optim.SGD(self.parameters(), lr=0.05)

    def training_step(self, batch=None, idx=None):
        from torchphysics_b200 import parallel
        rank, _ = parallel.world()
        g = torch.Generator().manual_seed(100 + rank)                 # every rank owns its own batch
        x = torch.rand(16, 3, generator=g)
        return ((self.net(x) - x.sum(dim=1, keepdim=True)) ** 2).mean()


def _ddp_worker(rank, size, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from torchphysics_b200 import solver
    s = _ToySolver(seed=rank + 1)
    solver.Trainer(max_steps=5, device="cpu").fit(s)
    # plain lists: a tensor in a Queue is shared through a file descriptor that dies with this process
    out.put((rank, s.net.weight.detach().reshape(-1).tolist(), s.net.bias.detach().tolist(), float(s.shift)))
    dist.destroy_process_group()


def test_trainer_broadcasts_rank0_state_gloo_world2():
    """Data-parallel training starts every replica from rank 0's parameters and buffers (what Lightning DDP does for the
    reference): ranks that built their models under different seeds hold identical weights after K steps."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    (_, w0, b0, s0), (_, w1, b1, s1) = res
    assert w0 == w1 and b0 == b1
    assert s0 == s1 == 1.0                                             # buffers follow rank 0 too
    torch.manual_seed(1)
    assert w0 != torch.nn.Linear(3, 1).weight.detach().reshape(-1).tolist()   # ... and they did train
