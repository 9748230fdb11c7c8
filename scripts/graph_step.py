"""Development aid: the literal config 1 of BASELINE.json (10,000 collocation points per step) trained through
tp.solver.Trainer eagerly and with cuda_graph=True: step time and loss trajectory."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 300


def run(graph):
    torch.manual_seed(0)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    model = tp.models.FCN(X, U, hidden=(64,) * 4)
    f = lambda x: 2 * math.pi ** 2 * torch.sin(math.pi * x[:, :1]) * torch.sin(math.pi * x[:, 1:])   # noqa: E731
    pde = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq, n_points=N),
                                      lambda u, x, f: tp.utils.laplacian(u, x) + f, data_functions={"f": f})
    bc = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq.boundary, n_points=N // 10),
                                     lambda u: u, weight=10.0)
    solver = tp.solver.Solver([pde, bc], optimizer_setting=tp.solver.OptimizerSetting(torch.optim.Adam, lr=1e-3))
    trainer = tp.solver.Trainer(max_steps=STEPS, cuda_graph=graph)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    trainer.fit(solver)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    losses = [float(l) for l in trainer.losses]
    if trainer.graphed:
        print(f"  replayed steps alone: {trainer.graph_step_ms * 1e3:.0f} us/step = {N / trainer.graph_step_ms * 1e3:.3e} points/s")
    print(f"cuda_graph={graph} graphed={trainer.graphed}: {STEPS} steps of {N} + {N // 10} points: {dt / STEPS * 1e6:.0f} us/step "
          f"({N / (dt / STEPS):.3e} points/s incl. the first steps); loss {losses[0]:.3f} -> {losses[len(losses) // 2]:.4f} -> {losses[-1]:.4f}")


run(False)
run(True)
