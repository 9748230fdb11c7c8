"""Oracle part 3: domain membership, volumes, bounding boxes and deterministic grids.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  float32 numpy restatement of the
constant-parameter case of the reference's domains (paths relative to
``/root/reference/src/torchphysics/problem/domains``).  Membership is evaluated with the
same float32 operation order as the reference so that the answers agree bit-for-bit.

A domain is described by a small tuple tree:
    ("interval", lo, hi) | ("parallelogram", origin, c1, c2) | ("circle", center, r)
    ("cut", A, B) | ("union", A, B) | ("product", A, B)   (product: A columns first)
"""
import numpy as np

f32 = np.float32


def _norm2(q):
    """torch.linalg.norm(q, dim=1) of float32 rows of length 2 as torch's CPU kernel evaluates it:
    sqrt(fma(y, y, x * x)) -- pinned on 1800 borderline candidates of tests/golden/cut_boundary.npz
    (float64 holds the fused product-sum exactly before the single rounding to float32)."""
    x2 = (q[:, 0] * q[:, 0]).astype(np.float64)
    y = q[:, 1].astype(np.float64)
    return np.sqrt((y * y + x2).astype(f32), dtype=f32)


def dim(dom):
    kind = dom[0]
    if kind == "interval":
        return 1
    if kind in ("parallelogram", "circle"):
        return 2
    if kind in ("cut", "union"):
        return dim(dom[1])
    if kind == "intersection":
        # domainoperations/intersection.py:34-37
        return contains(dom[1], pts) & contains(dom[2], pts)
    if kind == "product":
        return dim(dom[1]) + dim(dom[2])
    raise ValueError(kind)


def contains(dom, pts):
    """pts (N, dim) float32 -> bool (N,)."""
    pts = np.asarray(pts, dtype=f32)
    kind = dom[0]
    if kind == "interval":
        # domain1D/interval.py:37-43 -- closed interval
        lo, hi = f32(dom[1]), f32(dom[2])
        return (pts[:, 0] >= lo) & (pts[:, 0] <= hi)
    if kind == "parallelogram":
        # domain2D/parallelogram.py:83-103 -- Cramer solve for barycentric coordinates
        o = np.asarray(dom[1], f32)
        d1 = np.asarray(dom[2], f32) - o
        d2 = np.asarray(dom[3], f32) - o
        p = pts - o
        det = d1[0] * d2[1] - d1[1] * d2[0]
        xd = d2[1] * p[:, 0] - d2[0] * p[:, 1]
        yd = d1[0] * p[:, 1] - d1[1] * p[:, 0]
        bx = xd / det
        by = yd / det
        return (bx >= 0) & (bx <= 1) & (by >= 0) & (by <= 1)
    if kind == "circle":
        # domain2D/circle.py:34-40 -- ||p - c||_2 <= r
        c = np.asarray(dom[1], f32)
        return _norm2(pts - c) <= f32(dom[2])
    if kind == "cut":
        # domainoperations/cut.py:39-42
        return contains(dom[1], pts) & ~contains(dom[2], pts)
    if kind == "union":
        # domainoperations/union.py:45-48
        return contains(dom[1], pts) | contains(dom[2], pts)
    if kind == "product":
        # domainoperations/product.py:108-111
        da = dim(dom[1])
        return contains(dom[1], pts[:, :da]) & contains(dom[2], pts[:, da:])
    raise ValueError(kind)


def volume(dom):
    kind = dom[0]
    if kind == "interval":
        return f32(dom[2]) - f32(dom[1])                        # interval.py:72-75
    if kind == "parallelogram":
        o = np.asarray(dom[1], f32)
        d1 = np.asarray(dom[2], f32) - o
        d2 = np.asarray(dom[3], f32) - o
        return d1[0] * d2[1] - d1[1] * d2[0]                    # parallelogram.py:54-58
    if kind == "circle":
        return f32(np.pi) * f32(dom[2]) ** 2                    # circle.py:101-104
    if kind == "cut":
        return volume(dom[1])                                   # cut.py:44-55 (not contained)
    if kind == "union":
        return volume(dom[1]) + volume(dom[2])                  # union.py:26-42
    if kind == "product":
        return volume(dom[1]) * volume(dom[2])                  # product.py:145-147
    raise ValueError(kind)


def bounding_box(dom):
    kind = dom[0]
    if kind == "interval":
        return [float(dom[1]), float(dom[2])]
    if kind == "parallelogram":
        o, c1, c2 = (np.asarray(v, f32) for v in dom[1:4])
        c3 = c1 + c2 - o
        cs = np.stack([o, c1, c2, c3])
        return [float(cs[:, 0].min()), float(cs[:, 0].max()), float(cs[:, 1].min()), float(cs[:, 1].max())]
    if kind == "circle":
        c, r = np.asarray(dom[1], f32), f32(dom[2])
        return [float(c[0] - r), float(c[0] + r), float(c[1] - r), float(c[1] + r)]
    if kind == "cut":
        return bounding_box(dom[1])
    if kind == "union":
        a, b = bounding_box(dom[1]), bounding_box(dom[2])
        out = []
        for i in range(len(a) // 2):
            out += [min(a[2 * i], b[2 * i]), max(a[2 * i + 1], b[2 * i + 1])]
        return out
    if kind == "product":
        return bounding_box(dom[1]) + bounding_box(dom[2])
    raise ValueError(kind)


def grid(dom, n):
    """Deterministic part of ``sample_grid(n)`` for the three primitives."""
    kind = dom[0]
    if kind == "interval":
        # interval.py:57-65 -- n interior points of linspace(0,1,n+2)
        lo, hi = f32(dom[1]), f32(dom[2])
        t = np.linspace(0, 1, n + 2, dtype=f32)[1:-1]
        return ((hi - lo) * t + lo).reshape(-1, 1)
    if kind == "parallelogram":
        # parallelogram.py:119-152 -- n1 x n2 interior barycentric grid (may be < n points;
        # the reference tops up with random points, which are not part of this oracle)
        o = np.asarray(dom[1], f32)
        d1 = np.asarray(dom[2], f32) - o
        d2 = np.asarray(dom[3], f32) - o
        l1 = np.sqrt(d1[0] * d1[0] + d1[1] * d1[1], dtype=f32)
        l2 = np.sqrt(d2[0] * d2[0] + d2[1] * d2[1], dtype=f32)
        n1 = int(np.sqrt(f32(n) * l1 / l2, dtype=f32))
        n2 = int(np.sqrt(f32(n) * l2 / l1, dtype=f32))
        bx = np.linspace(0, 1, n1 + 2, dtype=f32)[1:-1]
        by = np.linspace(0, 1, n2 + 2, dtype=f32)[1:-1]
        # reference ordering: y index outer, x index inner
        B = np.stack([np.tile(bx, n2), np.repeat(by, n1)], axis=1)
        return B[:, :1] * d1 + B[:, 1:] * d2 + o
    if kind == "circle":
        # circle.py:70-99 -- sunflower/Vogel spiral
        c, r = np.asarray(dom[1], f32), f32(dom[2])
        gr = (np.sqrt(5) + 1) / 2.0
        k = np.arange(1, n + 1)
        phi = f32(2 * np.pi / gr) * k.astype(f32)      # torch: python scalar * int64 tensor -> float32 product
        rad = (np.sqrt((k - 0.5).astype(f32), dtype=f32) / f32(np.sqrt(n + 0.5))).astype(f32)
        P = np.stack([rad * np.cos(phi, dtype=f32), rad * np.sin(phi, dtype=f32)], axis=1)
        return r * P + c
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# boundaries of the three primitives: ("boundary", dom), ("left", interval),
# ("right", interval).  interval.py:96-191, parallelogram.py:159-291, circle.py:111-175
# --------------------------------------------------------------------------------------
def isclose(a, b):
    """torch.isclose(a, b) with the default rtol=1e-5, atol=1e-8, float32."""
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    return np.abs(a - b) <= (f32(1e-8) + f32(1e-5) * np.abs(b))


def _bary(dom, pts):
    o = np.asarray(dom[1], f32)
    d1 = np.asarray(dom[2], f32) - o
    d2 = np.asarray(dom[3], f32) - o
    p = np.asarray(pts, f32) - o
    det = d1[0] * d2[1] - d1[1] * d2[0]
    bx = (d2[1] * p[:, 0] - d2[0] * p[:, 1]) / det
    by = (d1[0] * p[:, 1] - d1[1] * p[:, 0]) / det
    return o, d1, d2, bx, by


def boundary_contains(bnd, pts):
    pts = np.asarray(pts, f32)
    kind, dom = bnd[0], bnd[1]
    if kind in ("left", "right"):
        return isclose(pts[:, 0], f32(dom[1] if kind == "left" else dom[2]))          # interval.py:161-165
    if dom[0] == "interval":
        return isclose(pts[:, 0], f32(dom[1])) | isclose(pts[:, 0], f32(dom[2]))      # interval.py:103-113
    if dom[0] == "parallelogram":                                                      # parallelogram.py:166-181
        _, _, _, bx, by = _bary(dom, pts)

        def edge(c1, c2):
            return (isclose(c1, f32(0)) | isclose(c1, f32(1))) & (c2 >= 0) & (c2 <= 1)
        return edge(bx, by) | edge(by, bx)
    if dom[0] == "circle":                                                             # circle.py:118-124
        return isclose(_norm2(pts - np.asarray(dom[1], f32)), f32(dom[2]))
    raise ValueError(bnd)


def boundary_volume(bnd):
    kind, dom = bnd[0], bnd[1]
    if kind in ("left", "right"):
        return f32(1)
    if dom[0] == "interval":
        return f32(2)
    if dom[0] == "parallelogram":
        o = np.asarray(dom[1], f32)
        d1, d2 = np.asarray(dom[2], f32) - o, np.asarray(dom[3], f32) - o
        return f32(2) * (_norm2(d1[None])[0] + _norm2(d2[None])[0])
    if dom[0] == "circle":
        return f32(2 * np.pi) * f32(dom[2])
    raise ValueError(bnd)


def _walk(dom, loc):
    """parallelogram.py:220-246: walk the four sides, ``loc`` = arc length from the origin."""
    o = np.asarray(dom[1], f32)
    d1, d2 = np.asarray(dom[2], f32) - o, np.asarray(dom[3], f32) - o
    s1, s2 = _norm2(d1[None])[0], _norm2(d2[None])[0]
    loc = np.asarray(loc, f32).copy()
    pts = np.zeros((len(loc), 2), f32)
    for d, s in ((d1, s1), (d2, s2), (-d1, s1), (-d2, s2)):
        scale = np.clip(loc / s, f32(0), f32(1))
        pts += scale[:, None] * d
        loc -= s
    return pts + o


def boundary_grid(bnd, n):
    kind, dom = bnd[0], bnd[1]
    if kind in ("left", "right"):
        return np.full((n, 1), f32(dom[1] if kind == "left" else dom[2]), f32)
    if dom[0] == "interval":                     # interval.py:128-137: upper, lower, upper, ...
        return np.where(np.arange(n) % 2 == 1, f32(dom[1]), f32(dom[2])).astype(f32).reshape(-1, 1)
    t = np.linspace(0, 1, n + 1, dtype=f32)[:-1]
    if dom[0] == "parallelogram":
        return _walk(dom, t * boundary_volume(bnd))
    if dom[0] == "circle":                       # circle.py:143-160
        phi = np.linspace(0, 2 * np.pi, n + 1, dtype=f32)[:-1]
        c, r = np.asarray(dom[1], f32), f32(dom[2])
        return np.stack([r * np.cos(phi, dtype=f32), r * np.sin(phi, dtype=f32)], axis=1) + c
    raise ValueError(bnd)


def boundary_normal(bnd, pts):
    pts = np.asarray(pts, f32)
    kind, dom = bnd[0], bnd[1]
    if kind == "left":
        return np.full((len(pts), 1), f32(-1))
    if kind == "right":
        return np.full((len(pts), 1), f32(1))
    if dom[0] == "interval":                     # interval.py:139-144
        return np.where(isclose(pts[:, 0], f32(dom[1])), f32(-1), f32(1)).reshape(-1, 1)
    if dom[0] == "circle":                       # circle.py:162-170
        return (pts - np.asarray(dom[1], f32)) / f32(dom[2])
    if dom[0] == "parallelogram":                # parallelogram.py:248-286
        _, d1, d2, bx, by = _bary(dom, pts)

        def unit_normal(d):
            v = np.array([-d[1], d[0]], f32)
            return v / _norm2(v[None])[0]
        n1, n2 = unit_normal(d1), -unit_normal(d2)
        out = np.zeros((len(pts), 2), f32)
        for i in (0.0, 1.0):
            sgn = f32(2 * i - 1)
            out += n1 * np.where(isclose(by, f32(i)), sgn, f32(0))[:, None]
            out += n2 * np.where(isclose(bx, f32(i)), sgn, f32(0))[:, None]
        with np.errstate(invalid="ignore", divide="ignore"):
            return out / _norm2(out)[:, None]
    raise ValueError(bnd)


# boundary of a cut A - B of two primitives: ("cutb", A, B)   (domainoperations/cut.py:111-201)
def cut_boundary_contains(A, B, pts):
    on_a, on_b = boundary_contains(("boundary", A), pts), boundary_contains(("boundary", B), pts)
    return (on_a & ~contains(B, pts)) | (on_b & contains(A, pts) & ~on_a)


def cut_boundary_normal(A, B, pts):
    """normal of A where the point lies on A's boundary, minus the normal of B elsewhere (cut.py:191-201)."""
    on_a = boundary_contains(("boundary", A), pts)
    with np.errstate(invalid="ignore", divide="ignore"):
        na, nb = boundary_normal(("boundary", A), pts), boundary_normal(("boundary", B), pts)
    return np.where(on_a[:, None], na, -nb)


def cut_boundary_volume(A, B):
    return boundary_volume(("boundary", A)) + boundary_volume(("boundary", B))


def cut_boundary_grid(A, B, n):
    """deterministic part of sampler_helper.py:211-256 (None when the reference would top up with random points)."""
    def on_main(g):
        return g[cut_boundary_contains(A, B, g)]
    ga, gb = boundary_grid(("boundary", A), n), boundary_grid(("boundary", B), n)
    ka, kb = on_main(ga), on_main(gb)
    if len(ka) + len(kb) == n:
        return np.concatenate([ka, kb])
    sa, sb = boundary_volume(("boundary", A)), boundary_volume(("boundary", B))
    approx = sa * f32(len(ka)) / f32(n) + sb * f32(len(kb)) / f32(n)
    na, nb = int(f32(n) * sa / approx) + 1, max(int(f32(n) * sb / approx), 1)
    out = np.concatenate([on_main(boundary_grid(("boundary", A), na)), on_main(boundary_grid(("boundary", B), nb))])
    return out[:n] if len(out) >= n else None


# boundary of a union A + B of two primitives (domainoperations/union.py:153-266)
def union_boundary_contains(A, B, pts):
    pts = np.asarray(pts, f32)
    in_a, in_b = contains(A, pts), contains(B, pts)
    on_a, on_b = boundary_contains(("boundary", A), pts), boundary_contains(("boundary", B), pts)
    on_both = on_a & on_b
    ok = np.ones(len(pts), bool)
    if on_both.any():                       # a shared boundary piece counts only where the normals agree (:170-189)
        with np.errstate(invalid="ignore", divide="ignore"):
            na = boundary_normal(("boundary", A), pts[on_both])
            nb = boundary_normal(("boundary", B), pts[on_both])
        ok[on_both] = (na * nb).sum(axis=1, dtype=f32) >= f32(0.5)
    return ((on_a & ~in_b) | (on_b & ~in_a) | on_both) & ok


def union_boundary_normal(A, B, pts):
    on_a = boundary_contains(("boundary", A), pts)
    with np.errstate(invalid="ignore", divide="ignore"):
        na, nb = boundary_normal(("boundary", A), pts), boundary_normal(("boundary", B), pts)
    return np.where(on_a[:, None], na, nb)


# boundary of an intersection A & B of two primitives (domainoperations/intersection.py:112-206)
def intersection_boundary_contains(A, B, pts):
    pts = np.asarray(pts, f32)
    on_a, on_b = boundary_contains(("boundary", A), pts), boundary_contains(("boundary", B), pts)
    return (on_a & contains(B, pts)) | (on_b & contains(A, pts))


def intersection_boundary_normal(A, B, pts):
    """normal of A where the point lies on A's boundary, the normal of B elsewhere (intersection.py:197-206); the
    same selection as for a union"""
    return union_boundary_normal(A, B, pts)


def intersection_boundary_grid_d(A, B, d):
    """density grid (intersection.py:185-196): both boundary grids at density d, each kept inside the other domain"""
    ga = boundary_grid(("boundary", A), int(np.ceil(f32(d) * boundary_volume(("boundary", A)))))      # domain.py:273-287
    gb = boundary_grid(("boundary", B), int(np.ceil(f32(d) * boundary_volume(("boundary", B)))))
    return np.concatenate([ga[contains(B, ga)], gb[contains(A, gb)]])


# boundaries of boolean domains whose operands may be boolean domains themselves (the reference recurses through
# domain_a.boundary / domain_b.boundary: cut.py:111-201, union.py:153-266, intersection.py:112-206)
_PRIMS = ("interval", "parallelogram", "circle")


def any_boundary_contains(dom, pts):
    pts = np.asarray(pts, f32)
    if dom[0] in _PRIMS:
        return boundary_contains(("boundary", dom), pts)
    A, B = dom[1], dom[2]
    in_a, in_b = contains(A, pts), contains(B, pts)
    on_a, on_b = any_boundary_contains(A, pts), any_boundary_contains(B, pts)
    if dom[0] == "cut":
        return (on_a & ~in_b) | (on_b & in_a & ~on_a)
    if dom[0] == "intersection":
        return (on_a & in_b) | (on_b & in_a)
    if dom[0] == "union":
        on_both = on_a & on_b
        ok = np.ones(len(pts), bool)
        if on_both.any():
            na, nb = any_boundary_normal(A, pts[on_both]), any_boundary_normal(B, pts[on_both])
            ok[on_both] = (na * nb).sum(axis=1, dtype=f32) >= f32(0.5)
        return ((on_a & ~in_b) | (on_b & ~in_a) | on_both) & ok
    raise ValueError(dom)


def any_boundary_normal(dom, pts):
    pts = np.asarray(pts, f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        if dom[0] in _PRIMS:
            return boundary_normal(("boundary", dom), pts)
        A, B = dom[1], dom[2]
        on_a = any_boundary_contains(A, pts)
        na, nb = any_boundary_normal(A, pts), any_boundary_normal(B, pts)
    return np.where(on_a[:, None], na, -nb if dom[0] == "cut" else nb)


def any_boundary_volume(dom):
    """the reference's estimate: the operands' boundary lengths add up at every level"""
    if dom[0] in _PRIMS:
        return boundary_volume(("boundary", dom))
    return any_boundary_volume(dom[1]) + any_boundary_volume(dom[2])
