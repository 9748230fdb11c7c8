"""Development aid: where the HOST time of a small-batch step goes (cProfile over eager steps of a bench.py configuration)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torchphysics_b200 as tp  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
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


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(30)
st.sort_stats("cumtime").print_stats(60)
