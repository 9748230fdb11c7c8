"""Development aid: parity margins of the wide tensor-core path on the golden tensors of configs 2 and 5."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.golden_io import load, relerr
from tests.test_gpu_fcn_jet import _run
for name in ("heat_3x128", "heat_6x128"):
    G = load(name); xt = G["xt"]; N = xt.shape[0]
    u, du, lap, _ = _run(G["params"], "tanh", xt, (0, 1, 2), (1.0, 1.0, 0.0))
    r = du[:, 0, 2:3] - lap
    gdu = np.zeros((N, 1, 3), np.float32); gdu[:, 0, 2] = (2 * r / N)[:, 0]
    _, _, _, grads = _run(G["params"], "tanh", xt, (0, 1, 2), (1.0, 1.0, 0.0), gdu=gdu, glap=-2 * r / N)
    print(name, "u %.2e ut %.2e lap %.2e" % (relerr(u, G["u"]), relerr(du[:, 0, 2:3], G["ut"]), relerr(lap, G["lap"])),
          "dtheta", " ".join("%.1e" % relerr(g, w) for g, w in zip(grads[:-1], G["dtheta"][:-1])))
for name in ("ritz_2x256", "ritz_8x256"):
    G = load(name); x = G["x"]; N = x.shape[0]
    u, du, lap, _ = _run(G["params"], "tanh", x, (0, 1), None)
    _, _, _, grads = _run(G["params"], "tanh", x, (0, 1), None, gu=-np.ones((N, 1), np.float32) / N, gdu=du / N)
    print(name, "grad %.2e" % relerr(du[:, 0, :], G["grad"]), "dtheta", " ".join("%.1e" % relerr(g, w) for g, w in zip(grads, G["dtheta"])))
