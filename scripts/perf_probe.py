"""Quick timing probe of the raw kernels (development aid, not the benchmark)."""
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, ".")
from oracle import ref_autograd
from torchphysics_b200.engine import FcnEngine, JetSpec


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    cfgs = [
        ("cfg1 4x64 d2 lap", 2, 1, (64,) * 4, (0, 1), (1.0, 1.0), 1 << 20),
        ("cfg2 6x128 d3 lap", 3, 1, (128,) * 6, (0, 1, 2), (1.0, 1.0, 0.0), 1 << 18),
        ("cfg3 6x128 d2 out3", 2, 3, (128,) * 6, (0, 1), (1.0, 1.0), 1 << 18),
        ("cfg5 8x256 d2 grad", 2, 1, (256,) * 8, (0, 1), None, 1 << 17),
        ("trunk 4x64 d1 out64", 1, 64, (64,) * 4, (0,), (1.0,), 1 << 18),
    ]
    for name, d_in, d_out, hidden, dcols, lapw, N in cfgs:
        torch.manual_seed(0)
        seq = ref_autograd.build_fcn(d_in, d_out, hidden).cuda()
        eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
        spec = JetSpec(dcols, lapw)
        x = torch.rand(N, d_in, device="cuda")
        packed = eng.packed(eng.params())
        gu = torch.randn(N, d_out, device="cuda")
        gdu = torch.randn(N, d_out, len(dcols), device="cuda")
        glap = torch.randn(N, d_out, device="cuda") if lapw else None
        tf = timeit(lambda: eng.forward_raw(packed, x, spec))
        tb = timeit(lambda: eng.backward_raw(packed, x, spec, gu, gdu, glap))
        print(f"{name}: N={N} fwd {tf:.3f} ms ({N/tf*1e3:.3e} pts/s)  bwd {tb:.3f} ms ({N/tb*1e3:.3e} pts/s)  "
              f"fwd+bwd {N/(tf+tb)*1e3:.3e} pts/s", flush=True)


if __name__ == "__main__":
    main()
