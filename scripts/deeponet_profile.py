"""Development aid: kernel table (torch.profiler, CUDA activities) of one config-4 step through PIDeepONetCondition."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.argv = [sys.argv[0]]
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ["TPX_DEEPONET_PROFILE"] = "1"
import deeponet_bench as db  # noqa: E402  (runs its timing loops on import)

for _ in range(3):
    db.native()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        db.native()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
