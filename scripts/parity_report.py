"""Development aid: measured parity margins of the CUDA path against the golden outputs of the real reference
(max-norm relative errors; the tests assert < 1e-5)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchphysics_b200 as tp  # noqa: E402
from oracle import ref_autograd  # noqa: E402
from tests.golden_io import load, relerr  # noqa: E402


def load_params(model, params):
    lins = [m for m in model.sequential if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for i, lin in enumerate(lins):
            lin.weight.copy_(torch.as_tensor(params[2 * i]))
            lin.bias.copy_(torch.as_tensor(params[2 * i + 1]))


G = load("poisson_4x64")
X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
model = tp.models.FCN(X, U, hidden=(64,) * 4).cuda()
load_params(model, G["params"])
cap = {}


def residual(u, x, f):
    cap.update(u=u, grad=tp.utils.grad(u, x), lap=tp.utils.laplacian(u, x))
    return cap["lap"] + f


cond = tp.conditions.PINNCondition(model, tp.samplers.DataSampler(tp.spaces.Points(torch.from_numpy(G["x"]), X)), residual,
                                   data_functions={"f": ref_autograd.poisson_source})
for _ in range(2):
    model.zero_grad(set_to_none=True)
    loss = cond(device="cuda")
    loss.backward()
print("config 1 (FCN 4x64 tanh, tcgen05 kernels) vs the real reference's fp32 autograd:")
for k in ("u", "grad", "lap"):
    print(f"  {k:5s} {relerr(cap[k].detach().cpu().numpy(), G[k]):.2e}")
print(f"  loss  {abs(float(loss) - float(G['loss'])) / abs(float(G['loss'])):.2e}")
errs = [relerr(p.grad.cpu().numpy(), w) for p, w in zip(model._native_engine().params(), G["dtheta"]) if w is not None]
print("  dL/dtheta per tensor: " + " ".join(f"{e:.1e}" for e in errs))
