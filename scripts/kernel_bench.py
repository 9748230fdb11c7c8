"""Development aid: times of the raw kernels (CUDA events) for config 1: forward with / without
checkpoint saving, reverse from saved checkpoints / with in-kernel recompute."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402
from torchphysics_b200.engine import JetSpec  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda", 0)
torch.manual_seed(0)
w, _, depth = os.environ.get("HIDDEN", "64x4").partition("x")     # e.g. HIDDEN=64x6: six hidden layers of width 64
model = tp.models.FCN(tp.spaces.R2("x"), tp.spaces.R1("u"), hidden=(int(w),) * int(depth)).to(dev)
eng = model._native_engine()
params = eng.params()
spec = JetSpec((0, 1), (1.0, 1.0))
packed = eng.packed(params)
xs = [torch.rand(N, 2, device=dev) for _ in range(4)]
glap = torch.randn(N, 1, device=dev) / N


def timed(fn, reps=10):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_f = timed(lambda i: eng.forward_raw(packed, xs[i % 4], spec, None))
t_br = timed(lambda i: eng.backward_raw(packed, xs[i % 4], spec, None, None, glap, None))
msg = f"N={N} hidden={w}x{depth}: fwd {t_f:.3f} ms ({N/t_f*1e3:.3e} pts/s)  bwd(recompute) {t_br:.3f} ms ({N/t_br*1e3:.3e})  step(recompute) {N/(t_f+t_br)*1e3:.3e} pts/s"
try:                                         # (the FP32 CUDA-core kernels have no checkpoint-saving forward)
    saved = torch.empty(eng.saved_bytes(spec, N), dtype=torch.uint8, device=dev)
    t_fs = timed(lambda i: eng.forward_raw(packed, xs[i % 4], spec, saved))
    eng.forward_raw(packed, xs[0], spec, saved)
    t_bs = timed(lambda i: eng.backward_raw(packed, xs[0], spec, None, None, glap, saved))
    msg += f"  fwd+save {t_fs:.3f} ms  bwd(saved) {t_bs:.3f} ms ({N/t_bs*1e3:.3e})  step(save) {N/(t_fs+t_bs)*1e3:.3e}"
except Exception as e:                       # noqa: BLE001
    msg += f"  [no saved path: {e}]"
print(msg)
