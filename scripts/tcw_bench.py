"""Development aid: kernel-only timing of the wide tensor-core path (csrc/tcw_kernels.cu) through the engine's raw
entry points (checkpoint-saving forward, layer-major reverse), CUDA events, inputs resident in HBM.
usage: tcw_bench.py [N [reps]]   (config-2 network 6x128, d_in 3, S = 5; config-3 network 6x128, 3 outputs, S = 4)"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_autograd  # noqa: E402
from torchphysics_b200.engine import FcnEngine, JetSpec  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for name, d_in, d_out, spec in (("config 2 (6x128, S=5)", 3, 1, JetSpec((0, 1, 2), (1.0, 1.0, 0.0))),
                                ("config 3 (6x128, 3 outputs, S=4)", 2, 3, JetSpec((0, 1), (1.0, 1.0)))):
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(d_in, d_out, (128,) * 6).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    S = 1 + spec.nd + spec.lap
    x = torch.rand(N, d_in, device="cuda")
    gu = torch.randn(N, d_out, device="cuda")
    gdu = torch.randn(N, d_out, spec.nd, device="cuda")
    glap = torch.randn(N, d_out, device="cuda")
    packed = eng.packed(eng.params())
    saved = torch.empty(eng.saved_bytes(spec, N), dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for r in range(REPS + 2):
        ev[0].record()
        eng.forward_raw(packed, x, spec, saved)
        ev[1].record()
        eng.backward_raw(packed, x, spec, gu, gdu, glap, saved)
        ev[2].record()
        torch.cuda.synchronize()
        if r >= 2:
            tf += ev[0].elapsed_time(ev[1]) / REPS
            tb += ev[1].elapsed_time(ev[2]) / REPS
    arr = N * S * 512                          # bytes of one activation array
    L = 6
    fwd_bytes = (2 * (L - 1) - 1 + 1) * arr    # layer 1 reads x only; output kernel reads one array
    bwd_bytes = (2 + 5 * (L - 1) - 2 - 1) * arr  # output reverse 2; per matrix dW 2 + adjoint 3; l = 1: no checkpoint reads, no Zbar write
    print(f"{name}: N={N}  forward+save {tf:.2f} ms ({fwd_bytes / tf / 1e6:.0f} GB/s)  reverse {tb:.2f} ms ({bwd_bytes / tb / 1e6:.0f} GB/s)"
          f"  step {N / (tf + tb) * 1e3:.3e} points/s")
