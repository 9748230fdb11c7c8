"""Development aid: clock64 timeline of four consecutive steps of CTA 0 of the fused reverse kernel of the wide path
(build with `make EXTRA=-DTPX_TIMELINE`).  usage: tcw_probe.py [N [S]]"""
import ctypes
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_autograd  # noqa: E402
from torchphysics_b200 import _lib  # noqa: E402
from torchphysics_b200.engine import FcnEngine, JetSpec  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
S = int(sys.argv[2]) if len(sys.argv) > 2 else 5                # 5: config-2 streams (3 directions + Laplacian), 4: config 3
torch.manual_seed(0)
d_in = 3 if S == 5 else 2
seq = ref_autograd.build_fcn(d_in, 1, (128,) * 3).cuda()
eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
spec = JetSpec((0, 1, 2), (1.0, 1.0, 0.0)) if S == 5 else JetSpec((0, 1), (1.0, 1.0))
x = torch.rand(N, d_in, device="cuda")
packed = eng.packed(eng.params())
gdu = torch.randn(N, 1, spec.nd, device="cuda")
glap = torch.randn(N, 1, device="cuda")
saved = torch.empty(eng.saved_bytes(spec, N), dtype=torch.uint8, device="cuda")
eng.forward_raw(packed, x, spec, saved)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
lib = _lib.lib()
lib.tpx_debug_set_timestamps(ctypes.c_void_p(buf.data_ptr()))
eng.backward_raw(packed, x, spec, None, gdu, glap, saved)      # launches l = 2 (general) then l = 1: the last writer wins
torch.cuda.synchronize()
lib.tpx_debug_set_timestamps(ctypes.c_void_p(0))
t = buf.cpu().numpy()
t0 = t[t > 0].min()
roles = {0: ("MMA", ["top", "fzn", "dempty", "adj_issued", "fzt", "fa", "dw_issued"]),
         1: ("Zld", ["loads_ok", "zn_free", "fzn_arrived", "zt_free", "fzt_arrived"]),
         2: ("AE0", ["loads_issued", "at_free", "fa_arrived", "d_ready", "d_read", "done"]),
         3: ("AE1", ["loads_issued", "at_free", "fa_arrived", "d_ready", "d_read", "done"])}
for step in range(4):
    for r, (name, ev) in roles.items():
        row = t[r * 64 + step * 8: r * 64 + step * 8 + 8]
        if (row > 0).any():
            print(f"step {40 + step} {name}: " + " ".join(f"{n}={int(v - t0)}" for n, v in zip(ev, row) if v > 0))
