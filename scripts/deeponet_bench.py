"""Development aid: config 4 (PI-DeepONet, 256 functions x 16,384 points, trunk/branch 4x64, p = 64) step time
through PIDeepONetCondition on the native path vs the generic torch-autograd definition on the same GPU."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402

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
fsampler = tp.samplers.FunctionSamplerRandomUniform(F, fset, function_creation_interval=0)
tsampler = tp.samplers.RandomUniformSampler(ival, n_points=N)
cond = tp.conditions.PIDeepONetCondition(net, fsampler, tsampler, lambda u, t, f: tp.utils.laplacian(u, t) + f)


def native():
    net.zero_grad(set_to_none=True)
    loss = cond(device="cuda")
    loss.backward()
    return loss


def generic():
    net.zero_grad(set_to_none=True)
    t = tsampler.sample_points(device="cuda").as_tensor
    tt = t.unsqueeze(0).repeat(F, 1, 1).requires_grad_(True)
    net.fix_branch_input(fsampler.sample_functions("cuda")(net.branch.grid), device="cuda")
    u = net._generic(tp.spaces.Points(tt, T))
    g = torch.autograd.grad(u.sum(), tt, create_graph=True)[0]
    lap = torch.autograd.grad(g.sum(), tt, create_graph=True)[0]
    loss = ((lap + torch.sin(math.pi * tt)) ** 2).sum(dim=(1, 2)).mean()
    loss.backward()
    return loss


RUNS = (("native", native, 20),) if os.environ.get("TPX_DEEPONET_PROFILE") else (("native", native, 20), ("torch autograd on the same GPU", generic, 5))
for name, fn, reps in RUNS:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {dt * 1e3:.2f} ms/step  {F * N / dt:.3e} (function, point) pairs/s  {N / dt:.3e} trunk points/s")
