"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    lists = {k[:-3] for k in z.files if k.endswith("__n")}
    for k in z.files:
        if "__" in k:
            continue
        out[k] = z[k]
    for base in lists:
        n = int(z[base + "__n"])
        out[base] = [None if bool(z[f"{base}__{i}__none"]) else z[f"{base}__{i}"] for i in range(n)]
    return out


def relerr(a, b):
    """max-norm relative error ||a-b||_inf / ||b||_inf."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
