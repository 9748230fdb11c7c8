"""Development aid: operand-rounding emulation of the split products of the tensor-core kernels, on the float64
closed form of oracle/jet_numpy.py (the accumulation itself is kept exact: this isolates what the operand formats cost).

  tf32x3   x = hi + lo, hi = RN tf32, lo truncated to tf32 by the MMA; hi*hi + hi*lo + lo*hi      (round-1 kernels)
  f16x3    x * 2^e = hi + lo, both fp16 (RN), e per row (forward) / per tile (reverse); same three products

usage: python scripts/fp16_split_emul.py [L H N]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import jet_numpy as J  # noqa: E402


def tf32_rn(x):
    u = np.asarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_trunc(x):
    u = np.asarray(x, np.float32).view(np.uint32)
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


def split_tf32(x):
    x = np.asarray(x, np.float32)
    hi = tf32_rn(x)
    lo = tf32_trunc(x - hi)
    return hi.astype(np.float64), lo.astype(np.float64), 1.0


def pow2_scale(m, target=12):
    """power of two s with s * m in [2^target, 2^(target+1))  (m > 0), else 1"""
    m = np.asarray(m, np.float64)
    e = np.floor(np.log2(np.where(m > 0, m, 1.0)))
    return np.where(m > 0, 2.0 ** (target - e), 1.0)


def split_f16(x, scale):
    xs = (np.asarray(x, np.float64) * scale).astype(np.float32)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64), scale


def make_mm(mode, rowwise):
    def mm(a, w):
        """a (..., N, K) @ w.T, w (J, K)"""
        if mode == "exact":
            return a @ w.T
        if mode == "tf32x3":
            ah, al, _ = split_tf32(a)
            wh, wl, _ = split_tf32(w)
            return (ah @ wh.T + al @ wh.T + ah @ wl.T).astype(np.float32).astype(np.float64)
        sw = pow2_scale(np.abs(w).max(), 10)
        if rowwise:
            sa = pow2_scale(np.abs(a).max(axis=-1, keepdims=True), 12)
        else:
            sa = pow2_scale(np.abs(a).max(axis=(-1, -2), keepdims=True), 12)
        ah, al, _ = split_f16(a, sa)
        wh, wl, _ = split_f16(w, sw)
        return ((ah @ wh.T + al @ wh.T + ah @ wl.T) / (sa * sw)).astype(np.float32).astype(np.float64)
    return mm


def forward(P, x, c, mm, keep=False):
    N, d = x.shape
    nl = len(P) // 2
    h = x
    dh = np.zeros((d, N, d))
    for k in range(d):
        dh[k, :, k] = 1.0
    lh = np.zeros((N, d))
    saved = []
    for l in range(nl):
        W, b = P[2 * l], P[2 * l + 1]
        inner = 0 < l < nl - 1
        f = mm if inner else (lambda a, w: a @ w.T)
        z = f(h, W) + b
        dz = f(dh, W)
        lz = f(lh, W)
        if l == nl - 1:
            saved.append((h, dh, lh, None, None, None))
            h, dh, lh = z, dz, lz
            break
        s, s1, s2, s3 = J._act("tanh", z)
        q = np.einsum("k,knj->nj", c, dz * dz)
        saved.append((h, dh, lh, z, dz, lz))
        h, dh, lh = s, s1[None] * dz, s2 * q + s1 * lz
    return h, np.transpose(dh, (1, 2, 0)), lh, saved


def backward(P, x, glap, c, mm_adj, mm_dw):
    N, d = x.shape
    u, du, lap, saved = forward(P, x, c, make_mm("exact", True))
    nl = len(P) // 2
    o = u.shape[1]
    zb, dzb, lzb = np.zeros((N, o)), np.zeros((d, N, o)), glap
    grads = [None] * (2 * nl)
    for l in range(nl - 1, -1, -1):
        W = P[2 * l]
        h, dh, lh, z, dz, lz = saved[l]
        inner = 0 < l < nl - 1
        if inner:
            # dW[j][i] = sum over (stream, point) rows of Zbar[r][j] a[r][i]: one product with K = rows
            Z = np.concatenate([zb, dzb.reshape(-1, zb.shape[1]), lzb], axis=0)       # (R, J)
            A = np.concatenate([h, dh.reshape(-1, h.shape[1]), lh], axis=0)           # (R, I)
            grads[2 * l] = mm_dw(Z.T, A.T)
        else:
            grads[2 * l] = zb.T @ h + np.einsum("knj,kni->ji", dzb, dh) + lzb.T @ lh
        grads[2 * l + 1] = zb.sum(axis=0)
        if l == 0:
            break
        f = mm_adj if inner else (lambda a, w: a @ w.T)
        hb, dhb, lhb = f(zb, W.T), f(dzb, W.T), f(lzb, W.T)
        _, _, _, zp, dzp, lzp = saved[l - 1]
        s, s1, s2, s3 = J._act("tanh", zp)
        q = np.einsum("k,knj->nj", c, dzp * dzp)
        zb = hb * s1 + np.einsum("knj,knj->nj", dhb, dzp) * s2 + lhb * (s3 * q + s2 * lzp)
        dzb = dhb * s1[None] + 2.0 * (lhb * s2)[None] * c[:, None, None] * dzp
        lzb = lhb * s1
    return grads


def dw_mm(mode):
    """dW product: both operands share ONE scale over the tile (K = rows cannot carry per-row scales)"""
    def mm(zt, at):
        if mode == "exact":
            return zt @ at.T
        if mode == "tf32x3":
            zh, zl, _ = split_tf32(zt)
            ah, al, _ = split_tf32(at)
            return zh @ ah.T + zl @ ah.T + zh @ al.T + zl @ al.T
        out = 0.0
        T = 128 * 4
        for r0 in range(0, zt.shape[1], T):             # per-tile scales, fp32 flush per tile
            z, a = zt[:, r0:r0 + T], at[:, r0:r0 + T]
            sz, sa = pow2_scale(np.abs(z).max(), 12), pow2_scale(np.abs(a).max(), 12)
            zh, zl, _ = split_f16(z, sz)
            ah, al, _ = split_f16(a, sa)
            out = out + ((zh @ ah.T + zl @ ah.T + zh @ al.T + zl @ al.T) / (sz * sa)).astype(np.float32).astype(np.float64)
        return out
    return mm


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    L, H, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (4, 64, 2048)
    rng = np.random.default_rng(0)
    d = 2
    widths = [d] + [H] * L + [1]
    P = []
    for i in range(len(widths) - 1):
        bound = (5.0 / 3.0) * np.sqrt(6.0 / (widths[i] + widths[i + 1]))
        P += [rng.uniform(-bound, bound, (widths[i + 1], widths[i])), rng.uniform(-0.1, 0.1, widths[i + 1])]
    P = [p.astype(np.float32).astype(np.float64) for p in P]
    x = rng.uniform(0, 1, (N, d)).astype(np.float32).astype(np.float64)
    c = np.ones(d)
    u0, du0, lap0, _ = forward(P, x, c, make_mm("exact", True))
    f = np.sin(3 * x[:, :1])
    glap = 2.0 * (lap0 + f) / N
    g0 = backward(P, x, glap, c, make_mm("exact", True), dw_mm("exact"))
    for mode, rowwise in (("tf32x3", True), ("f16x3", True), ("f16x3", False)):
        mm = make_mm(mode, rowwise)
        u, du, lap, _ = forward(P, x, c, mm)
        g = backward(P, x, glap, c, mm, dw_mm(mode))
        print(f"{mode:7s} {'row' if rowwise else 'tile'} scales: u {rel(u, u0):.2e} du {rel(du, du0):.2e} lap {rel(lap, lap0):.2e}  dtheta "
              + " ".join(f"{rel(a, b):.1e}" for a, b in zip(g, g0)))


if __name__ == "__main__":
    main()
