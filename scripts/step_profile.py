"""Development aid: kernel table (torch.profiler, CUDA activities) of the public-API step of a bench.py configuration."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torchphysics_b200 as tp  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg]["points"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model, cond, _ = bench.build_config(tp, cfg, n, dev)
params = [p for p in cond.parameters() if p.requires_grad]


def step():
    for p in params:
        p.grad = None
    loss = cond(device=dev)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
print(f"config {cfg}, {n} points, {steps} steps")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=90))
