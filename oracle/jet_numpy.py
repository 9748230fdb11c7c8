"""Oracle part 2: the mathematics of the hot path restated in float64 numpy.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

What the reference obtains with 1 + d reverse(-over-reverse) sweeps of
``torch.autograd.grad`` (reference ``utils/differentialoperators/
differentialoperators.py:11-66``) through ``tp.models.FCN`` (reference
``models/fcn.py:9-29,85-87``) is, mathematically, the propagation of the streams

    value h, first derivatives d_k h (k < d_in), weighted Laplacian L h = sum_k c_k d_kk h

through every layer ``z = W h + b`` (linear in every stream) and activation ``s``:

    h+ = s(z),   d_k h+ = s'(z) d_k z,   L h+ = s''(z) sum_k c_k (d_k z)^2 + s'(z) L z.

``jet_forward`` does exactly that; ``jet_backward`` is the hand-derived reverse sweep
(SURVEY 8(a) "Math the kernel restates").  This file is the executable specification
the CUDA kernels in ``torchphysics_b200/csrc`` are written against, and an independent
(non-autograd, float64) check of ``oracle/ref_autograd.py``.
"""
import numpy as np


def _act(name, z):
    """returns s, s', s'', s''' at z."""
    if name == "tanh":
        t = np.tanh(z)
        s1 = 1.0 - t * t
        s2 = -2.0 * t * s1
        s3 = -2.0 * s1 * (1.0 - 3.0 * t * t)
        return t, s1, s2, s3
    if name == "sin":
        s, c = np.sin(z), np.cos(z)
        return s, c, -s, -c
    raise ValueError(name)


def jet_forward(params, x, lap_weights=None, activation="tanh", keep=False):
    """params = [W0, b0, ..., WL, bL] with W (out, in); x (N, d_in).

    Returns u (N, o), du (N, o, d_in), lap (N, o)  (lap is zeros if lap_weights is None).
    With keep=True also returns the per-layer pre-activation streams for the reverse.
    """
    P = [np.asarray(p, dtype=np.float64) for p in params]
    x = np.asarray(x, dtype=np.float64)
    N, d = x.shape
    c = np.zeros(d) if lap_weights is None else np.asarray(lap_weights, dtype=np.float64)
    nl = len(P) // 2
    h = x                                     # (N, w)
    dh = np.zeros((d, N, d))
    for k in range(d):
        dh[k, :, k] = 1.0                     # d_k x = e_k
    lh = np.zeros((N, d))                     # L x = 0
    saved = []
    for l in range(nl):
        W, b = P[2 * l], P[2 * l + 1]
        z = h @ W.T + b
        dz = dh @ W.T                         # (d, N, out)
        lz = lh @ W.T
        if l == nl - 1:
            if keep:
                saved.append((h, dh, lh, None, None, None))
            h, dh, lh = z, dz, lz
            break
        s, s1, s2, s3 = _act(activation, z)
        q = np.einsum("k,knj->nj", c, dz * dz)
        if keep:
            saved.append((h, dh, lh, z, dz, lz))
        h = s
        dh = s1[None] * dz
        lh = s2 * q + s1 * lz
    u = h
    du = np.transpose(dh, (1, 2, 0))          # (N, o, d)
    lap = lh
    if keep:
        return u, du, lap, saved
    return u, du, lap


def jet_backward(params, x, gu, gdu, glap, lap_weights=None, activation="tanh"):
    """Adjoint of jet_forward: given dLoss/du (N,o), dLoss/ddu (N,o,d), dLoss/dlap (N,o)
    (any may be None) returns [dW0, db0, ..., dWL, dbL] (float64)."""
    P = [np.asarray(p, dtype=np.float64) for p in params]
    x = np.asarray(x, dtype=np.float64)
    N, d = x.shape
    c = np.zeros(d) if lap_weights is None else np.asarray(lap_weights, dtype=np.float64)
    u, du, lap, saved = jet_forward(P, x, c, activation, keep=True)
    o = u.shape[1]
    nl = len(P) // 2
    zb = np.zeros((N, o)) if gu is None else np.asarray(gu, dtype=np.float64)
    dzb = np.zeros((d, N, o)) if gdu is None else np.transpose(np.asarray(gdu, dtype=np.float64), (2, 0, 1))
    lzb = np.zeros((N, o)) if glap is None else np.asarray(glap, dtype=np.float64)
    grads = [None] * (2 * nl)
    for l in range(nl - 1, -1, -1):
        W = P[2 * l]
        h, dh, lh, z, dz, lz = saved[l]
        # zb, dzb, lzb are adjoints of this layer's pre-activation streams
        grads[2 * l] = zb.T @ h + np.einsum("knj,kni->ji", dzb, dh) + lzb.T @ lh
        grads[2 * l + 1] = zb.sum(axis=0)
        if l == 0:
            break
        hb = zb @ W                          # adjoints of the layer input streams
        dhb = dzb @ W
        lhb = lzb @ W
        # through the activation of layer l-1
        _, _, _, zp, dzp, lzp = saved[l - 1]
        s, s1, s2, s3 = _act(activation, zp)
        q = np.einsum("k,knj->nj", c, dzp * dzp)
        zb = hb * s1 + np.einsum("knj,knj->nj", dhb, dzp) * s2 + lhb * (s3 * q + s2 * lzp)
        dzb = dhb * s1[None] + 2.0 * (lhb * s2)[None] * c[:, None, None] * dzp
        lzb = lhb * s1
    return grads


def poisson_loss_and_grads(params, x, f, activation="tanh"):
    """loss = mean((lap u + f)^2) and its parameter gradients (config 1)."""
    d = x.shape[1]
    u, du, lap = jet_forward(params, x, np.ones(d), activation)
    r = lap + f
    loss = np.mean(np.sum(r * r, axis=1))
    glap = 2.0 * r / x.shape[0]
    grads = jet_backward(params, x, None, None, glap, np.ones(d), activation)
    return loss, grads
