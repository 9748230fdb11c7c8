"""Development aid: Trainer(cuda_graph=True) under torchrun (one rank per GPU): the captured step contains the NCCL all-reduce of
dLoss/dtheta; prints whether the step was captured, its time, and whether the replicas' parameters still agree."""
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402

rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, STEPS = 10000, 200
torch.manual_seed(rank)                                    # replicas start differently: the Trainer broadcasts rank 0's
X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
model = tp.models.FCN(X, U, hidden=(64,) * 4)
f = lambda x: 2 * math.pi ** 2 * torch.sin(math.pi * x[:, :1]) * torch.sin(math.pi * x[:, 1:])   # noqa: E731
pde = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq, n_points=N),
                                  lambda u, x, f: tp.utils.laplacian(u, x) + f, data_functions={"f": f})
bc = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(sq.boundary, n_points=N // 10), lambda u: u, weight=10.0)
solver = tp.solver.Solver([pde, bc], optimizer_setting=tp.solver.OptimizerSetting(torch.optim.Adam, lr=1e-3))
trainer = tp.solver.Trainer(max_steps=STEPS, cuda_graph=True, device=dev)
trainer.fit(solver)
torch.cuda.synchronize()
flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
ref = flat.clone()
dist.broadcast(ref, src=0)
same = bool(torch.equal(flat, ref))
losses = [float(l) for l in trainer.losses]
print(f"rank {rank}: graphed={trainer.graphed} step {getattr(trainer, 'graph_step_ms', float('nan')) * 1e3:.0f} us "
      f"loss {losses[0]:.3f} -> {losses[-1]:.4f} params equal to rank 0: {same}", flush=True)
dist.destroy_process_group()
