"""Development check of the wide tensor-core path (width 65..128) against the float64 oracle; prints max-norm
relative errors per tensor.  usage: python scripts/tcw_check.py [case ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
from oracle import jet_numpy, ref_autograd
from tests.golden_io import relerr
from tests.test_gpu_fcn_jet import _run

CASES = {
    "a": (2, 1, (128,) * 2, (0, 1), (1.0, 1.0), 16),
    "b": (2, 1, (128,) * 3, (0, 1), (1.0, 1.0), 1000),
    "c": (3, 1, (128,) * 6, (0, 1, 2), (1.0, 1.0, 0.0), 333),
    "d": (2, 3, (100, 128, 96), (0, 1), (1.0, 1.0), 200),
    "e": (4, 8, (128,) * 3, (), None, 777),
    "f": (1, 2, (96,) * 4, (0,), None, 5000),
    "g": (2, 1, (256,) * 3, (0, 1), None, 300),
    "h": (2, 1, (256,) * 8, (0, 1), None, 150),
    "i": (3, 2, (200, 256, 130), (0, 1, 2), (1.0, 1.0, 0.0), 1000),
    "j": (2, 1, (256, 256), (0, 1), (1.0, 1.0), 5000),
    "k": (2, 1, (128,) * 3, (0, 1), (1.0, 1.0), 1000, "sin"),
    "l": (3, 2, (200, 256, 130), (0, 1, 2), (1.0, 1.0, 0.0), 1000, "sin"),
    "m": (2, 1, (256,) * 3, (0, 1), (1.0, 0.5), 100, "sin"),
}
for name in (sys.argv[1:] or list(CASES)):
    d_in, d_out, hidden, dcols, lap_w, N = CASES[name][:6]
    act = CASES[name][6] if len(CASES[name]) > 6 else "tanh"
    torch.manual_seed(1234)
    seq = ref_autograd.build_fcn(d_in, d_out, hidden, act)
    params = [p.detach().numpy() for p in ref_autograd.linear_params(seq)]
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, size=(N, d_in)).astype(np.float32)
    nd = len(dcols)
    gu = rng.standard_normal((N, d_out)).astype(np.float32)
    gdu = rng.standard_normal((N, d_out, nd)).astype(np.float32) if nd else None
    glap = rng.standard_normal((N, d_out)).astype(np.float32) if lap_w is not None else None
    u, du, lap, grads = _run(params, act, x, dcols, lap_w, gu=gu, gdu=gdu, glap=glap)
    c = np.zeros(d_in)
    if lap_w is not None:
        for col, w in zip(dcols, lap_w):
            c[col] = w
    ou, odu, olap = jet_numpy.jet_forward(params, x, c if lap_w is not None else None, act)
    msg = [f"u {relerr(u, ou):.2e}"]
    if nd:
        msg.append(f"du {relerr(du, odu[:, :, list(dcols)]):.2e}")
    if lap_w is not None:
        msg.append(f"lap {relerr(lap, olap):.2e}")
    full_gdu = np.zeros((N, d_out, d_in))
    if nd:
        full_gdu[:, :, list(dcols)] = gdu
    og = jet_numpy.jet_backward(params, x, gu, full_gdu, glap, c if lap_w is not None else None, act)
    msg.append("grads " + " ".join(f"{relerr(g, w):.1e}" for g, w in zip(grads, og)))
    print(name, act, hidden, "S", 1 + nd + (lap_w is not None), "|", " ".join(msg), flush=True)
