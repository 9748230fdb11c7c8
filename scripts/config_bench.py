"""Development aid: step time (condition.forward + backward through the public API, sampler included) of the
parity configurations 2, 3 and 5 of BASELINE.json (width 128: wide tensor-core kernels; width 256: FP32 CUDA-core kernels).
usage: config_bench.py [N [config]]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
X, T, U, V, P = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("u"), tp.spaces.R2("v"), tp.spaces.R1("p")
sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])


def cfg2():
    model = tp.models.FCN(X * T, U, hidden=(128,) * 6).cuda()
    s = tp.samplers.RandomUniformSampler(sq * tp.domains.Interval(T, 0, 1), n_points=N)
    return model, tp.conditions.PINNCondition(model, s, lambda u, x, t: tp.utils.grad(u, t) - tp.utils.laplacian(u, x))


def cfg3():
    model = tp.models.FCN(X, V * P, hidden=(128,) * 6).cuda()
    s = tp.samplers.RandomUniformSampler(sq, n_points=N)

    def residual(v, p, x):
        lap = torch.cat([tp.utils.laplacian(v[:, :1], x), tp.utils.laplacian(v[:, 1:], x)], dim=1)
        mom = tp.utils.convective(v, v, x) - 0.01 * lap + tp.utils.grad(p, x)
        return torch.cat([mom, tp.utils.div(v, x)], dim=1)
    return model, tp.conditions.PINNCondition(model, s, residual)


def cfg5():
    model = tp.models.FCN(X, U, hidden=(256,) * 8).cuda()
    dom = tp.domains.Circle(X, [0, 0], 1.0) - tp.domains.Parallelogram(X, [-0.25, -0.25], [0.25, -0.25], [-0.25, 0.25])
    s = tp.samplers.RandomUniformSampler(dom, n_points=N)
    return model, tp.conditions.DeepRitzCondition(
        model, s, lambda u, x: 0.5 * (tp.utils.grad(u, x) ** 2).sum(dim=1, keepdim=True) - u)


ONLY = sys.argv[2] if len(sys.argv) > 2 else None      # "2", "3" or "5": one configuration only
for name, make in (("config 2 (heat, 6x128, S=5)", cfg2), ("config 3 (Navier-Stokes, 6x128, 3 outputs, S=4)", cfg3),
                   ("config 5 (Deep Ritz on Circle - square, 8x256, S=3)", cfg5)):
    if ONLY is not None and not name.startswith("config " + ONLY):
        continue
    model, cond = make()

    def step():
        model.zero_grad(set_to_none=True)
        loss = cond(device="cuda")
        loss.backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: N={N}  {dt * 1e3:.1f} ms/step  {N / dt:.3e} points/s")
    del model, cond
    torch.cuda.empty_cache()
