"""Generates tests/golden/*.npz by RUNNING THE REAL REFERENCE (boschresearch/torchphysics
v1.1.0) in the authoring container:

    python tests/golden/make_golden.py

It imports ``torchphysics`` from /root/reference/src (never copied into this repo) with
the two import stubs in oracle/stubs (matplotlib and pytorch_lightning are not
installed).  The .npz files are the parity pins of the oracle (tests/test_oracle_golden.py)
and of the CUDA path (tests/test_gpu_*.py).  /root/reference does not exist on the GPU
box, so only the committed .npz files travel.
"""
import math
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, "/root/reference/src")
warnings.filterwarnings("ignore")

import torchphysics as tp  # noqa: E402  (the reference)


def params_of(model):
    return [p.detach().numpy().copy() for p in model.parameters()]


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, (list, tuple)):
            out[k + "__n"] = np.array(len(v))
            for i, a in enumerate(v):
                out[f"{k}__{i}"] = np.zeros(0, np.float32) if a is None else np.asarray(a)
                out[f"{k}__{i}__none"] = np.array(a is None)
        else:
            out[k] = v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def grads_of(model):
    return [None if p.grad is None else p.grad.detach().numpy().copy() for p in model.parameters()]


# ---------------------------------------------------------------------------- config 1
def poisson(hidden, n, name, act=None):
    torch.manual_seed(0)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    kw = {} if act is None else {"activations": act}
    model = tp.models.FCN(X, U, hidden=hidden, **kw)
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    torch.manual_seed(1)
    pts = tp.samplers.RandomUniformSampler(sq, n_points=n).make_static()
    x = pts.sample_points().as_tensor.clone()

    def f(x):
        return 2 * math.pi ** 2 * torch.sin(math.pi * x[:, :1]) * torch.sin(math.pi * x[:, 1:])

    captured = {}

    def residual(u, x, f):
        captured["u"] = u
        captured["grad"] = tp.utils.grad(u, x)
        captured["lap"] = tp.utils.laplacian(u, x)
        return captured["lap"] + f

    cond = tp.conditions.PINNCondition(model, pts, residual, data_functions={"f": f})
    model.zero_grad()
    loss = cond(device="cpu")
    loss.backward()
    save(name, x=x, params=params_of(model), u=captured["u"], grad=captured["grad"],
         lap=captured["lap"], loss=loss, dtheta=grads_of(model), hidden=np.array(hidden))


# ---------------------------------------------------------------------------- config 2
def heat(hidden, n, name):
    torch.manual_seed(2)
    X, T, U = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("u")
    model = tp.models.FCN(X * T, U, hidden=hidden)
    dom = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1]) * tp.domains.Interval(T, 0, 1)
    torch.manual_seed(3)
    pts = tp.samplers.RandomUniformSampler(dom, n_points=n).make_static()
    P = pts.sample_points()
    xt = torch.cat([P.coordinates["x"], P.coordinates["t"]], dim=1).clone()
    captured = {}

    def residual(u, x, t):
        captured["ut"] = tp.utils.grad(u, t)
        captured["lap"] = tp.utils.laplacian(u, x)
        captured["u"] = u
        return captured["ut"] - captured["lap"]

    cond = tp.conditions.PINNCondition(model, pts, residual)
    model.zero_grad()
    loss = cond(device="cpu")
    loss.backward()
    save(name, xt=xt, params=params_of(model), u=captured["u"], ut=captured["ut"],
         lap=captured["lap"], loss=loss, dtheta=grads_of(model), hidden=np.array(hidden))


# ---------------------------------------------------------------------------- config 3
def navier_stokes(hidden, n, name):
    torch.manual_seed(4)
    X, U, Pp = tp.spaces.R2("x"), tp.spaces.R2("u"), tp.spaces.R1("p")
    model = tp.models.FCN(X, U * Pp, hidden=hidden)
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    torch.manual_seed(5)
    pts = tp.samplers.RandomUniformSampler(sq, n_points=n).make_static()
    x = pts.sample_points().as_tensor.clone()
    nu = 0.01
    captured = {}

    def residual(u, p, x):
        conv = tp.utils.convective(u, u, x)
        lap = torch.cat([tp.utils.laplacian(u[:, :1], x), tp.utils.laplacian(u[:, 1:], x)], dim=1)
        gp = tp.utils.grad(p, x)
        cont = tp.utils.div(u, x)
        captured.update(conv=conv, lap=lap, gp=gp, div=cont, u=u, p=p,
                        jac=tp.utils.jac(torch.cat([u, p], dim=1), x))
        return torch.cat([conv - nu * lap + gp, cont], dim=1)

    cond = tp.conditions.PINNCondition(model, pts, residual)
    model.zero_grad()
    loss = cond(device="cpu")
    loss.backward()
    save(name, x=x, params=params_of(model), y=torch.cat([captured["u"], captured["p"]], dim=1),
         jac=captured["jac"], lap=captured["lap"], conv=captured["conv"], div=captured["div"],
         loss=loss, dtheta=grads_of(model), hidden=np.array(hidden))


# ---------------------------------------------------------------------------- config 5
def deep_ritz(hidden, n, name):
    torch.manual_seed(6)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=hidden)
    dom = tp.domains.Circle(X, [0, 0], 1.0) - tp.domains.Parallelogram(
        X, [-0.25, -0.25], [0.25, -0.25], [-0.25, 0.25])
    torch.manual_seed(7)
    pts = tp.samplers.RandomUniformSampler(dom, n_points=n).make_static()
    x = pts.sample_points().as_tensor.clone()
    captured = {}

    def integrand(u, x):
        g = tp.utils.grad(u, x)
        captured.update(u=u, grad=g)
        return 0.5 * (g ** 2).sum(dim=1, keepdim=True) - u

    cond = tp.conditions.DeepRitzCondition(model, pts, integrand)
    model.zero_grad()
    loss = cond(device="cpu")
    loss.backward()
    save(name, x=x, params=params_of(model), u=captured["u"], grad=captured["grad"],
         loss=loss, dtheta=grads_of(model), hidden=np.array(hidden))


# ---------------------------------------------------------------------------- config 4
def deeponet(n_fn, n_pts, p, name, trunk_hidden=(32, 32), branch_hidden=(24, 24), n_sensors=10):
    """physics-informed DeepONet for u'' + f = 0 (reference deeponet_condition.py:52-151,
    deeponet.py:71-110, trunknets.py:145-154, branchnets.py:188-194)."""
    torch.manual_seed(8)
    T, U, Fs, K = tp.spaces.R1("t"), tp.spaces.R1("u"), tp.spaces.R1("f"), tp.spaces.R1("k")
    int_t = tp.domains.Interval(T, 0, 1)
    fspace = tp.spaces.FunctionSpace(T, Fs)
    grid = tp.samplers.GridSampler(int_t, n_points=n_sensors).sample_points()
    trunk = tp.models.FCTrunkNet(T, hidden=trunk_hidden, default_trunk_input=grid)
    branch = tp.models.FCBranchNet(fspace, hidden=branch_hidden, grid=grid)
    net = tp.models.DeepONet(trunk, branch, U, output_neurons=p)

    def fn(k, t):
        return k * torch.sin(math.pi * t)

    param_sampler = tp.samplers.RandomUniformSampler(tp.domains.Interval(K, -1, 1), n_points=n_fn)
    fset = tp.domains.CustomFunctionSet(fspace, param_sampler, fn)
    fsampler = tp.samplers.FunctionSamplerRandomUniform(n_fn, fset, function_creation_interval=1)
    torch.manual_seed(9)
    tsampler = tp.samplers.RandomUniformSampler(int_t, n_points=n_pts).make_static()
    tt = tsampler.sample_points().as_tensor.clone()
    captured = {}

    def residual(u, t, f):
        captured["u"] = u
        captured["ut"] = tp.utils.grad(u, t)
        captured["lap"] = tp.utils.laplacian(u, t)
        captured["f"] = f
        return captured["lap"] + f

    cond = tp.conditions.PIDeepONetCondition(net, fsampler, tsampler, residual)
    net.zero_grad()
    torch.manual_seed(10)
    loss = cond(device="cpu")
    loss.backward()
    k_used = fset.param_samples.as_tensor[fsampler.current_indices].clone()      # (F, 1) in sampled order
    branch_in = fn(k_used.unsqueeze(1), grid.as_tensor.unsqueeze(0))            # (F, sensors, 1)
    save(name, t=tt, k=k_used, sensors=grid.as_tensor, branch_in=branch_in, fvals=captured["f"], u=captured["u"],
         ut=captured["ut"], lap=captured["lap"], loss=loss,
         trunk_params=[q.detach().numpy().copy() for q in trunk.parameters()],
         branch_params=[q.detach().numpy().copy() for q in branch.parameters()],
         branch_out=net.branch.current_out.detach(),
         trunk_dtheta=[None if q.grad is None else q.grad.numpy().copy() for q in trunk.parameters()],
         branch_dtheta=[None if q.grad is None else q.grad.numpy().copy() for q in branch.parameters()],
         p=np.array(p), trunk_hidden=np.array(trunk_hidden), branch_hidden=np.array(branch_hidden))


# ---------------------------------------------------------------------------- normalised input
def heat_normalized(hidden, n, name):
    """heat residual with tp.models.Sequential(NormalizationLayer(domain), FCN): the trainable affine
    input map of reference models/model.py:57-92 in front of the network."""
    torch.manual_seed(12)
    X, T, U = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("u")
    dom = tp.domains.Parallelogram(X, [0, 0], [2, 0], [0, 3]) * tp.domains.Interval(T, 0, 5)
    model = tp.models.Sequential(tp.models.NormalizationLayer(dom), tp.models.FCN(X * T, U, hidden=hidden))
    torch.manual_seed(13)
    pts = tp.samplers.RandomUniformSampler(dom, n_points=n).make_static()
    P = pts.sample_points()
    xt = torch.cat([P.coordinates["x"], P.coordinates["t"]], dim=1).clone()
    captured = {}

    def residual(u, x, t):
        captured["ut"] = tp.utils.grad(u, t)
        captured["gx"] = tp.utils.grad(u, x)
        captured["lap"] = tp.utils.laplacian(u, x)
        captured["u"] = u
        return captured["ut"] - 0.1 * captured["lap"]

    cond = tp.conditions.PINNCondition(model, pts, residual)
    model.zero_grad()
    loss = cond(device="cpu")
    loss.backward()
    save(name, xt=xt, params=params_of(model), u=captured["u"], ut=captured["ut"], gx=captured["gx"],
         lap=captured["lap"], loss=loss, dtheta=grads_of(model), hidden=np.array(hidden))


# ---------------------------------------------------------------------------- boundaries
def boundaries():
    """boundary domains of Interval / Parallelogram / Circle: membership, normals, grids, volumes
    (reference interval.py:96-191, parallelogram.py:159-291, circle.py:111-175)."""
    X, T = tp.spaces.R2("x"), tp.spaces.R1("t")
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    ival = tp.domains.Interval(T, -0.3, 0.8)
    out = {}
    g = torch.Generator().manual_seed(21)
    for nm, d in [("para", para), ("circ", circ), ("ival", ival)]:
        b = d.boundary
        for n in (1, 7, 50):
            out[f"grid_{nm}_{n}"] = b.sample_grid(n).as_tensor
        on = b.sample_grid(200).as_tensor
        # candidates: exact boundary points, slightly perturbed ones (inside/outside isclose), far points
        eps = torch.rand(on.shape, generator=g) * 4e-5 - 2e-5
        far = torch.rand(on.shape, generator=g) * 3.0 - 1.5
        cand = torch.cat([on, on + eps, on * (1 + 1e-6), far], dim=0)
        out["cand_" + nm] = cand
        out["in_" + nm] = b._contains(tp.spaces.Points(cand.clone(), d.space)).reshape(-1)
        out["normal_" + nm] = b.normal(tp.spaces.Points(on.clone(), d.space))
        out["on_" + nm] = on
        out["vol_" + nm] = b.volume().reshape(-1)
        torch.manual_seed(22)
        r = b.sample_random_uniform(n=300)
        out["n_random_" + nm] = np.array(len(r))
        out["allin_random_" + nm] = np.array(bool(b._contains(r).all()))
    for nm, b in [("left", ival.boundary_left), ("right", ival.boundary_right)]:
        cand = out["cand_ival"]
        out["in_" + nm] = b._contains(tp.spaces.Points(cand.clone(), T)).reshape(-1)
        out["grid_" + nm] = b.sample_grid(5).as_tensor
        out["normal_" + nm] = b.normal(tp.spaces.Points(b.sample_grid(5).as_tensor, T))
        out["vol_" + nm] = b.volume().reshape(-1)
    # a boundary condition as the examples write it: u = 0 on the boundary of the square x time
    prod = para.boundary * ival
    cand3 = torch.cat([out["cand_para"][:600], torch.rand(600, 1, generator=g) * 1.5 - 0.5], dim=1)
    out["cand3"] = cand3
    out["in_prod"] = prod._contains(tp.spaces.Points(cand3.clone(), X * T)).reshape(-1)
    out["vol_prod"] = prod.volume().reshape(-1)
    save("boundaries", **out)


# ---------------------------------------------------------------------------- boundary of a cut
def cut_boundary():
    """boundary of Circle - Parallelogram (a domain with a hole, hole contained) and of a Parallelogram
    cut by an overlapping Circle (reference domainoperations/cut.py:111-201, sampler_helper.py:129-266)."""
    X = tp.spaces.R2("x")
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    sq = tp.domains.Parallelogram(X, [-0.25, -0.25], [0.25, -0.25], [-0.25, 0.25])
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    bite = tp.domains.Circle(X, [0.7, 0.1], 0.45)
    out = {}
    g = torch.Generator().manual_seed(31)
    for nm, dom in (("hole", circ - sq), ("bite", para - bite)):
        b = dom.boundary
        ga = dom.domain_a.boundary.sample_grid(150).as_tensor
        gb = dom.domain_b.boundary.sample_grid(150).as_tensor
        on = torch.cat([ga, gb], dim=0)
        far = torch.rand(300, 2, generator=g) * 3.0 - 1.5
        cand = torch.cat([on, on + (torch.rand(on.shape, generator=g) * 4e-5 - 2e-5), far], dim=0)
        inside = b._contains(tp.spaces.Points(cand.clone(), X)).reshape(-1)
        out["cand_" + nm], out["in_" + nm] = cand, inside
        onb = cand[inside]
        out["on_" + nm] = onb
        out["normal_" + nm] = b.normal(tp.spaces.Points(onb.clone(), X))
        out["vol_" + nm] = b.volume().reshape(-1)
        for n in (40, 200):
            torch.manual_seed(32)
            gr = b.sample_grid(n)
            out[f"ngrid_{nm}_{n}"] = np.array(len(gr))
            out[f"grid_{nm}_{n}"] = gr.as_tensor
        torch.manual_seed(33)
        r = b.sample_random_uniform(n=500)
        out["n_random_" + nm] = np.array(len(r))
    save("cut_boundary", **out)


# ---------------------------------------------------------------------------- boundary of a union
def union_boundary():
    """boundary of Parallelogram + Circle (overlapping) and of two touching squares (shared edge: the
    normal test of reference domainoperations/union.py:161-192)."""
    X = tp.spaces.R2("x")
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    circ = tp.domains.Circle(X, [0.7, 0.1], 0.45)
    sq1 = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    sq2 = tp.domains.Parallelogram(X, [1, 0], [2, 0], [1, 1])
    out = {}
    g = torch.Generator().manual_seed(41)
    for nm, dom in (("lens", para + circ), ("touch", sq1 + sq2)):
        b = dom.boundary
        ga = dom.domain_a.boundary.sample_grid(150).as_tensor
        gb = dom.domain_b.boundary.sample_grid(150).as_tensor
        on = torch.cat([ga, gb], dim=0)
        far = torch.rand(300, 2, generator=g) * 3.0 - 0.75
        cand = torch.cat([on, on + (torch.rand(on.shape, generator=g) * 4e-5 - 2e-5), far], dim=0)
        inside = b._contains(tp.spaces.Points(cand.clone(), X)).reshape(-1)
        out["cand_" + nm], out["in_" + nm] = cand, inside
        onb = cand[inside]
        out["on_" + nm] = onb
        out["normal_" + nm] = b.normal(tp.spaces.Points(onb.clone(), X))
        out["vol_" + nm] = b.volume().reshape(-1)
        for n in (40, 200):
            torch.manual_seed(42)
            out[f"ngrid_{nm}_{n}"] = np.array(len(b.sample_grid(n)))
        torch.manual_seed(43)
        out["n_random_" + nm] = np.array(len(b.sample_random_uniform(n=500)))
    save("union_boundary", **out)


# ---------------------------------------------------------------------------- boundary of an intersection
def intersection_boundary():
    """boundary of Parallelogram & Circle (a lens) and of a Circle & axis-aligned square (reference
    domainoperations/intersection.py:112-206, sampler_helper.py:129-266)."""
    X = tp.spaces.R2("x")
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    bite = tp.domains.Circle(X, [0.7, 0.1], 0.45)
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    sq = tp.domains.Parallelogram(X, [0.0, -0.5], [1.5, -0.5], [0.0, 1.0])
    out = {}
    g = torch.Generator().manual_seed(51)
    for nm, dom in (("lens", para & bite), ("corner", circ & sq)):
        b = dom.boundary
        ga = dom.domain_a.boundary.sample_grid(150).as_tensor
        gb = dom.domain_b.boundary.sample_grid(150).as_tensor
        on = torch.cat([ga, gb], dim=0)
        far = torch.rand(300, 2, generator=g) * 3.0 - 1.5
        cand = torch.cat([on, on + (torch.rand(on.shape, generator=g) * 4e-5 - 2e-5), far], dim=0)
        inside = b._contains(tp.spaces.Points(cand.clone(), X)).reshape(-1)
        out["cand_" + nm], out["in_" + nm] = cand, inside
        onb = cand[inside]
        out["on_" + nm] = onb
        out["normal_" + nm] = b.normal(tp.spaces.Points(onb.clone(), X))
        out["vol_" + nm] = b.volume().reshape(-1)
        for n in (40, 200):
            torch.manual_seed(52)
            gr = b.sample_grid(n)
            out[f"ngrid_{nm}_{n}"] = np.array(len(gr))
            out[f"grid_{nm}_{n}"] = gr.as_tensor
        torch.manual_seed(53)
        out["n_random_" + nm] = np.array(len(b.sample_random_uniform(n=500)))
        out["n_grid_d_" + nm] = np.array(len(b.sample_grid(d=30.0)))
        out["grid_d_" + nm] = b.sample_grid(d=30.0).as_tensor
    save("intersection_boundary", **out)


# ---------------------------------------------------------------------------- boundaries of nested boolean domains
def nested_boundary():
    """boundaries whose operands are boolean domains themselves: a cut of a cut, a cut of a union, a cut BY a union
    (reference domainoperations/cut.py:111-201 with domain_a.boundary / domain_b.boundary boolean boundaries)."""
    X = tp.spaces.R2("x")
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    sq = tp.domains.Parallelogram(X, [-0.25, -0.25], [0.25, -0.25], [-0.25, 0.25])
    bite = tp.domains.Circle(X, [0.7, 0.1], 0.45)
    sq3 = tp.domains.Parallelogram(X, [0.0, -0.8], [0.5, -0.8], [0.0, -0.4])
    sq4 = tp.domains.Parallelogram(X, [-0.5, 0.0], [1.0, 0.0], [-0.5, 0.5])
    prims = [circ, sq, bite, sq3, sq4]
    out = {}
    g = torch.Generator().manual_seed(61)
    on = torch.cat([p.boundary.sample_grid(120).as_tensor for p in prims], dim=0)
    far = torch.rand(300, 2, generator=g) * 3.0 - 1.5
    cand = torch.cat([on, on + (torch.rand(on.shape, generator=g) * 4e-5 - 2e-5), far], dim=0)
    out["cand"] = cand
    cases = (("cutcut", (circ - sq) - bite), ("unioncut", (sq4 + bite) - sq), ("cutunion", circ - (sq + sq3)))
    for nm, dom in cases:
        b = dom.boundary
        inside = b._contains(tp.spaces.Points(cand.clone(), X)).reshape(-1)
        out["in_" + nm] = inside
        onb = cand[inside]
        out["on_" + nm] = onb
        out["normal_" + nm] = b.normal(tp.spaces.Points(onb.clone(), X))
        out["vol_" + nm] = b.volume().reshape(-1)
        for n in (40, 200):
            torch.manual_seed(62)
            out[f"ngrid_{nm}_{n}"] = np.array(len(b.sample_grid(n)))
        torch.manual_seed(63)
        out["n_random_" + nm] = np.array(len(b.sample_random_uniform(n=500)))
    save("nested_boundary", **out)


# ---------------------------------------------------------------------------- hard constraint
def poisson_hard_constraint(hidden, n, name):
    """Poisson residual on tp.models.HardConstraint(FCN, g): the output transform g(u, x) = x1 (1 - x1) x2 (1 - x2) u
    + sin(x1) of reference models/model.py:186-217 (chain rule through the network AND the explicit x dependence)."""
    torch.manual_seed(14)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    inner = tp.models.FCN(X, U, hidden=hidden)

    def g(u, x):
        return x[:, :1] * (1 - x[:, :1]) * x[:, 1:] * (1 - x[:, 1:]) * u + torch.sin(x[:, :1])

    model = tp.models.HardConstraint(inner, g)
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    torch.manual_seed(15)
    pts = tp.samplers.RandomUniformSampler(sq, n_points=n).make_static()
    x = pts.sample_points().as_tensor.clone()
    captured = {}

    def residual(u, x):
        captured["u"] = u
        captured["grad"] = tp.utils.grad(u, x)
        captured["lap"] = tp.utils.laplacian(u, x)
        return captured["lap"] + 1.0

    cond = tp.conditions.PINNCondition(model, pts, residual)
    model.zero_grad()
    loss = cond(device="cpu")
    loss.backward()
    save(name, x=x, params=params_of(inner), u=captured["u"], grad=captured["grad"], lap=captured["lap"], loss=loss,
         dtheta=grads_of(inner), hidden=np.array(hidden))


# ---------------------------------------------------------------------------- domains
def domains():
    X, T = tp.spaces.R2("x"), tp.spaces.R1("t")
    g = torch.Generator().manual_seed(11)
    cand = (torch.rand(4000, 2, generator=g) * 3.0 - 1.5)
    cand1 = (torch.rand(2000, 1, generator=g) * 3.0 - 1.5)
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    ival = tp.domains.Interval(T, -0.3, 0.8)
    sq = tp.domains.Parallelogram(X, [-0.25, -0.25], [0.25, -0.25], [-0.25, 0.25])
    cut = circ - sq
    uni = para + circ
    out = {"cand": cand, "cand1": cand1}
    for nm, d in [("para", para), ("circ", circ), ("sq", sq), ("cut", cut), ("uni", uni)]:
        out["in_" + nm] = d._contains(tp.spaces.Points(cand.clone(), X)).reshape(-1)
        out["bbox_" + nm] = d.bounding_box()
        out["vol_" + nm] = d.volume().reshape(-1)
    out["in_ival"] = ival._contains(tp.spaces.Points(cand1.clone(), T)).reshape(-1)
    out["bbox_ival"] = ival.bounding_box()
    out["vol_ival"] = ival.volume().reshape(-1)
    prod = para * ival
    cand3 = torch.cat([cand[:2000], cand1], dim=1)
    out["cand3"] = cand3
    out["in_prod"] = prod._contains(tp.spaces.Points(cand3.clone(), X * T)).reshape(-1)
    out["vol_prod"] = prod.volume().reshape(-1)
    for n in (1, 7, 50):
        out[f"grid_ival_{n}"] = ival.sample_grid(n).as_tensor
        out[f"grid_circ_{n}"] = circ.sample_grid(n).as_tensor
    for n in (16, 50, 100):
        torch.manual_seed(0)
        gpts = para.sample_grid(n).as_tensor
        out[f"grid_para_{n}"] = gpts
    # sample counts / membership of the reference's random samplers (n_points mode)
    torch.manual_seed(12)
    for nm, d in [("para", para), ("circ", circ), ("cut", cut), ("uni", uni)]:
        s = tp.samplers.RandomUniformSampler(d, n_points=500).sample_points()
        out["n_random_" + nm] = np.array(len(s))
        out["allin_random_" + nm] = np.array(bool(d._contains(s).all()))
    save("domains", **out)


if __name__ == "__main__":
    JOBS = {
        "poisson_4x64": lambda: poisson((64, 64, 64, 64), 257, "poisson_4x64"),
        "poisson_3x20_sin": lambda: poisson((20, 20, 20), 100, "poisson_3x20_sin", act=tp.models.Sinus()),
        "heat_3x48": lambda: heat((48, 48, 48), 193, "heat_3x48"),
        "ns_3x32": lambda: navier_stokes((32, 32, 32), 130, "ns_3x32"),
        "ritz_4x40": lambda: deep_ritz((40, 40, 40, 40), 150, "ritz_4x40"),
        "heat_norm_3x48": lambda: heat_normalized((48, 48, 48), 161, "heat_norm_3x48"),
        "poisson_hard_3x32": lambda: poisson_hard_constraint((32, 32, 32), 143, "poisson_hard_3x32"),
        "deeponet_2x32": lambda: deeponet(6, 40, 16, "deeponet_2x32"),
        "deeponet_4x64": lambda: deeponet(5, 33, 64, "deeponet_4x64", trunk_hidden=(64,) * 4, branch_hidden=(64,) * 4, n_sensors=50),
        "domains": domains,
        "boundaries": boundaries,
        "cut_boundary": cut_boundary,
        "union_boundary": union_boundary,
        "intersection_boundary": intersection_boundary,
        "nested_boundary": nested_boundary,
        # hidden width 128 (configs 2 and 3 of BASELINE.json at reduced depth): the wide tensor-core path
        "heat_3x128": lambda: heat((128, 128, 128), 211, "heat_3x128"),
        "ns_3x128": lambda: navier_stokes((128, 128, 128), 147, "ns_3x128"),
        # hidden width 256 (config 5 at reduced depth): two 128-neuron blocks per layer on the wide path
        "ritz_2x256": lambda: deep_ritz((256, 256), 139, "ritz_2x256"),
        # BASELINE.json's configurations at full depth / the point count BASELINE.md section 3 names for the parity run
        "poisson_4x64_10k": lambda: poisson((64, 64, 64, 64), 10000, "poisson_4x64_10k"),
        "heat_6x128": lambda: heat((128,) * 6, 1000, "heat_6x128"),
        "ns_6x128": lambda: navier_stokes((128,) * 6, 1000, "ns_6x128"),
        "ritz_8x256": lambda: deep_ritz((256,) * 8, 500, "ritz_8x256"),
    }
    for job in (sys.argv[1:] or list(JOBS)):      # `make_golden.py NAME ...` regenerates only those fixtures
        JOBS[job]()
