"""Development aid: where does one API-level step (PINNCondition.forward + backward) spend its time?"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
dev = torch.device("cuda", 0)
torch.manual_seed(0)
X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
model = tp.models.FCN(X, U, hidden=(64,) * 4).to(dev)


def source(x):
    return 2 * math.pi ** 2 * torch.sin(math.pi * x[:, :1]) * torch.sin(math.pi * x[:, 1:2])


host_x = [torch.rand(N, 2).pin_memory() for _ in range(3)]
cond = tp.conditions.PINNCondition(model, tp.samplers.HostStreamSampler(host_x, X), lambda u, x, f: tp.utils.laplacian(u, x) + f,
                                   data_functions={"f": source})
loss_host = torch.zeros(2).pin_memory()
done = [torch.cuda.Event(), torch.cuda.Event()]
it = {"i": 0}


def step():
    i = it["i"]
    model.zero_grad(set_to_none=True)
    loss = cond(device=dev)
    loss.backward()
    loss_host[i % 2:i % 2 + 1].copy_(loss.detach().reshape(1), non_blocking=True)
    done[i % 2].record()
    if i > 0:
        done[(i - 1) % 2].synchronize()
    it["i"] += 1


for _ in range(4):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=8, max_name_column_width=60))

import time  # noqa: E402
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    print("wall ms/step", (time.perf_counter() - t0) * 100)
