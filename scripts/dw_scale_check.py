"""Development aid: accuracy of dL/dtheta at scale (2^20 points) of the tensor-core reverse kernels against a float64
reference assembled from 64 chunks of 2^14 points (each chunk's TMEM accumulation is short, the chunks are summed in float64).
usage: dw_scale_check.py [width]   (64: config-1 kernels, 128 / 256: wide path)"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_autograd  # noqa: E402
from torchphysics_b200.engine import FcnEngine, JetSpec  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
seq = ref_autograd.build_fcn(2, 1, (W,) * 4).cuda()
eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
spec = JetSpec((0, 1), (1.0, 1.0))
N = 1 << 20
x = torch.rand(N, 2, device="cuda")
packed = eng.packed(eng.params())
for name, glap in (("coherent adjoints (all ones)", torch.ones(N, 1, device="cuda") / N),
                   ("random-sign adjoints", torch.randn(N, 1, device="cuda") / N)):
    def sweep(lo, hi):
        xs, g = x[lo:hi].contiguous(), glap[lo:hi].contiguous()
        saved = torch.empty(eng.saved_bytes(spec, hi - lo), dtype=torch.uint8, device="cuda")
        eng.forward_raw(packed, xs, spec, saved)
        return eng.backward_raw(packed, xs, spec, None, None, g, saved)
    full = sweep(0, N)
    C = 1 << 14
    ref = sum(sweep(i, i + C) for i in range(0, N, C))
    den = ref.abs().max()
    print(f"width {W}, {name}: max |full - chunked| / max|ref| = {float((full - ref).abs().max() / den):.2e}")
