"""Development aid: config 4 (PI-DeepONet, 256 functions x 16,384 points) trained through tp.solver.Trainer eagerly and with
cuda_graph=True: step time and loss trajectory."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402

F, N, p = 256, 16384, 64
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 200


def run(graph):
    torch.manual_seed(0)
    T, U, Fs, K = tp.spaces.R1("t"), tp.spaces.R1("u"), tp.spaces.R1("f"), tp.spaces.R1("k")
    fspace = tp.spaces.FunctionSpace(T, Fs)
    ival = tp.domains.Interval(T, 0, 1)
    grid = tp.samplers.GridSampler(ival, n_points=50).sample_points(device="cuda")
    trunk = tp.models.FCTrunkNet(T, hidden=(64,) * 4, default_trunk_input=grid)
    branch = tp.models.FCBranchNet(fspace, hidden=(64,) * 4, grid=grid)
    net = tp.models.DeepONet(trunk, branch, U, output_neurons=p)
    fset = tp.domains.CustomFunctionSet(fspace, tp.samplers.RandomUniformSampler(tp.domains.Interval(K, -1, 1), n_points=F),
                                        lambda k, t: k * torch.sin(math.pi * t))
    fsampler = tp.samplers.FunctionSamplerRandomUniform(F, fset, function_creation_interval=0)
    tsampler = tp.samplers.RandomUniformSampler(ival, n_points=N)
    cond = tp.conditions.PIDeepONetCondition(net, fsampler, tsampler, lambda u, t, f: tp.utils.laplacian(u, t) + f)
    solver = tp.solver.Solver([cond], optimizer_setting=tp.solver.OptimizerSetting(torch.optim.Adam, lr=1e-3))
    trainer = tp.solver.Trainer(max_steps=STEPS, cuda_graph=graph)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    trainer.fit(solver)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    losses = [float(l) for l in trainer.losses]
    if trainer.graphed:
        print(f"  replayed steps alone: {trainer.graph_step_ms * 1e3:.0f} us/step = {F * N / trainer.graph_step_ms * 1e3:.3e} pairs/s")
    print(f"cuda_graph={graph} graphed={trainer.graphed}: {STEPS} steps of {F} x {N}: {dt / STEPS * 1e6:.0f} us/step "
          f"({F * N / (dt / STEPS):.3e} pairs/s incl. the first steps); loss {losses[0]:.4f} -> {losses[len(losses) // 2]:.5f} -> {losses[-1]:.5f}")


run(False)
run(True)
