"""Sampler / domain kernels (through the C ABI, via tp.domains / tp.samplers) against the
numpy oracle and the golden outputs of the real reference: exact point counts, bit-exact
membership, deterministic grids, seeded reproducibility."""
import numpy as np
import pytest
import torch

from oracle import geometry, ref_autograd
from tests.golden_io import load, relerr
from tests.test_oracle_golden import DOMS

pytestmark = pytest.mark.gpu


def _tp():
    import torchphysics_b200 as tp
    return tp


def _build(tp, dom, names=None):
    """oracle tuple tree -> tp.domains object."""
    kind = dom[0]
    if kind == "interval":
        return tp.domains.Interval(tp.spaces.R1("t"), dom[1], dom[2])
    if kind == "parallelogram":
        return tp.domains.Parallelogram(tp.spaces.R2("x"), dom[1], dom[2], dom[3])
    if kind == "circle":
        return tp.domains.Circle(tp.spaces.R2("x"), dom[1], dom[2])
    a, b = _build(tp, dom[1]), _build(tp, dom[2])
    return {"cut": lambda: a - b, "union": lambda: a + b, "product": lambda: a * b}[kind]()


@pytest.mark.parametrize("name", ["para", "circ", "sq", "cut", "uni", "ival", "prod"])
def test_membership_matches_reference_bit_for_bit(name):
    tp = _tp()
    G = load("domains")
    dom = _build(tp, DOMS[name])
    cand = {"ival": "cand1", "prod": "cand3"}.get(name, "cand")
    pts = tp.spaces.Points(torch.from_numpy(G[cand]).cuda(), dom.space)
    got = dom._contains(pts).cpu().numpy().reshape(-1)
    assert np.array_equal(got, G["in_" + name])
    assert np.array_equal(got, geometry.contains(DOMS[name], G[cand]))


@pytest.mark.parametrize("name", ["para", "circ", "sq", "cut", "uni", "ival", "prod"])
@pytest.mark.parametrize("n", [1, 500, 100003])
def test_random_uniform_counts_and_membership(name, n):
    tp = _tp()
    dom = _build(tp, DOMS[name])
    pts = tp.samplers.RandomUniformSampler(dom, n_points=n).sample_points(device="cuda")
    assert len(pts) == n and pts.as_tensor.shape == (n, dom.dim) and pts.as_tensor.is_cuda
    x = pts.as_tensor.cpu().numpy()
    assert geometry.contains(DOMS[name], x).all()          # oracle membership, float32 op order
    assert bool(dom._contains(pts).all())
    bb = geometry.bounding_box(DOMS[name])
    for c in range(dom.dim):
        assert x[:, c].min() >= bb[2 * c] - 1e-6 and x[:, c].max() <= bb[2 * c + 1] + 1e-6


def test_seeded_reproducibility_and_streams():
    tp = _tp()
    dom = _build(tp, DOMS["cut"])
    torch.manual_seed(7)
    a = dom.sample_random_uniform(n=20000, device="cuda").as_tensor.clone()
    b = dom.sample_random_uniform(n=20000, device="cuda").as_tensor.clone()
    torch.manual_seed(7)
    a2 = dom.sample_random_uniform(n=20000, device="cuda").as_tensor
    b2 = dom.sample_random_uniform(n=20000, device="cuda").as_tensor
    assert torch.equal(a, a2) and torch.equal(b, b2) and not torch.equal(a, b)
    # the points do not depend on the acceptance estimate the domain has learned
    dom._accept_rate = 0.2
    torch.manual_seed(7)
    assert torch.equal(a, dom.sample_random_uniform(n=20000, device="cuda").as_tensor)
    tp.domains.set_rank_stream(1)
    torch.manual_seed(7)
    c = dom.sample_random_uniform(n=20000, device="cuda").as_tensor
    tp.domains.set_rank_stream(0)
    assert not torch.equal(a, c)


def test_uniformity_moments():
    """size-independent property at scale: first/second moments of 4M points of the unit disc
    minus a square agree with the closed form."""
    tp = _tp()
    X = tp.spaces.R2("x")
    dom = tp.domains.Circle(X, [0, 0], 1.0) - tp.domains.Parallelogram(X, [-.25, -.25], [.25, -.25], [-.25, .25])
    n = 1 << 22
    x = dom.sample_random_uniform(n=n, device="cuda").as_tensor.double()
    area = np.pi - 0.25
    ex2 = (np.pi / 4 - 0.5 ** 4 / 12 * 1.0) / area      # int x^2 over disc = pi/4, over square side .5: s^4/12
    assert abs(float(x[:, 0].mean())) < 3e-3 and abs(float(x[:, 1].mean())) < 3e-3
    assert abs(float((x[:, 0] ** 2).mean()) - ex2) < 3e-3
    r = x.norm(dim=1)
    assert float(r.max()) <= 1.0 and float(x.abs().max(dim=1).values.min()) > 0.25 - 1e-7


def test_density_mode_counts():
    tp = _tp()
    G = load("domains")
    dom = _build(tp, DOMS["circ"])
    pts = tp.samplers.RandomUniformSampler(dom, density=40.0).sample_points(device="cuda")
    assert len(pts) == int(np.ceil(np.float32(40.0) * np.float32(G["vol_circ"])))
    cut = _build(tp, DOMS["cut"])
    pts = tp.samplers.RandomUniformSampler(cut, density=500.0).sample_points(device="cuda")
    assert 0 < len(pts) < 500 * float(G["vol_circ"]) and geometry.contains(DOMS["cut"], pts.as_tensor.cpu().numpy()).all()


@pytest.mark.parametrize("n", [1, 7, 50])
def test_grids_interval_circle(n):
    tp = _tp()
    G = load("domains")
    g = tp.samplers.GridSampler(_build(tp, DOMS["ival"]), n_points=n).sample_points(device="cuda")
    assert np.allclose(g.as_tensor.cpu().numpy(), G[f"grid_ival_{n}"], rtol=1e-6, atol=1e-7)
    g = tp.samplers.GridSampler(_build(tp, DOMS["circ"]), n_points=n).sample_points(device="cuda")
    assert np.allclose(g.as_tensor.cpu().numpy(), G[f"grid_circ_{n}"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("n", [16, 50, 100])
def test_grid_parallelogram(n):
    tp = _tp()
    G = load("domains")
    dom = _build(tp, DOMS["para"])
    g = tp.samplers.GridSampler(dom, n_points=n).sample_points(device="cuda")
    ref = G[f"grid_para_{n}"]
    want = geometry.grid(DOMS["para"], n)
    assert len(g) == n == len(ref)
    got = g.as_tensor.cpu().numpy()
    assert np.allclose(got[: len(want)], ref[: len(want)], rtol=1e-6, atol=1e-7)
    assert geometry.contains(DOMS["para"], got).all()


def test_grid_on_cut_and_union():
    tp = _tp()
    for name in ("cut", "uni"):
        dom = _build(tp, DOMS[name])
        g = tp.samplers.GridSampler(dom, n_points=200).sample_points(device="cuda")
        assert len(g) == 200
        assert geometry.contains(DOMS[name], g.as_tensor.cpu().numpy()).all()


class _nullcontext:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def test_lhs_sampler():
    tp = _tp()
    dom = _build(tp, DOMS["para"])
    n = 1000
    pts = tp.samplers.LHSSampler(dom, n).sample_points(device="cuda")
    assert len(pts) == n and geometry.contains(DOMS["para"], pts.as_tensor.cpu().numpy()).all()
    # on an axis-aligned box every one of the n strata of every axis holds exactly one point
    box = tp.domains.Parallelogram(tp.spaces.R2("x"), [0, 0], [2, 0], [0, 1])
    x = tp.samplers.LHSSampler(box, n).sample_points(device="cuda").as_tensor.cpu().numpy()
    assert sorted(np.floor(x[:, 0] / 2 * n).astype(int).tolist()) == list(range(n))
    assert sorted(np.floor(x[:, 1] * n).astype(int).tolist()) == list(range(n))


def test_filter_and_product_and_static_samplers():
    tp = _tp()
    X, T = tp.spaces.R2("x"), tp.spaces.R1("t")
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    s = tp.samplers.RandomUniformSampler(sq, n_points=300, filter_fn=lambda x: x[:, :1] <= 0.5)
    p = s.sample_points(device="cuda")
    assert len(p) == 300 and float(p.as_tensor[:, 0].max()) <= 0.5
    ps = tp.samplers.RandomUniformSampler(sq, n_points=10) * tp.samplers.GridSampler(tp.domains.Interval(T, 0, 1), n_points=4)
    q = ps.sample_points(device="cuda")
    assert len(q) == 40 and list(q.space.keys()) == ["x", "t"]
    st = tp.samplers.RandomUniformSampler(sq * tp.domains.Interval(T, 0, 2), n_points=64).make_static()
    a = st.sample_points(device="cuda")
    b = st.sample_points(device="cuda")
    assert a == b and a.as_tensor.shape == (64, 3) and float(a.as_tensor[:, 2].max()) <= 2.0


def test_empty_and_error_paths():
    tp = _tp()
    from torchphysics_b200 import _lib
    X = tp.spaces.R2("x")
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    assert len(sq.sample_random_uniform(n=0, device="cuda")) == 0
    with pytest.raises(_lib.TpxError):
        tp.domains.Circle(X, [0, 0], lambda t: t + 1)
    far = tp.domains.Circle(X, [10, 10], 0.1)
    with pytest.raises(RuntimeError):
        (sq & far).sample_random_uniform(n=5, device="cuda")


# ---------------------------------------------------------------------------- boundaries
def _boundary_domains(tp):
    X, T = tp.spaces.R2("x"), tp.spaces.R1("t")
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    ival = tp.domains.Interval(T, -0.3, 0.8)
    return X, T, para, circ, ival


def test_boundary_membership_normals_grids_match_reference():
    import torchphysics_b200 as tp
    from tests.test_oracle_golden import BNDS
    G = load("boundaries")
    X, T, para, circ, ival = _boundary_domains(tp)
    for nm, d in (("para", para), ("circ", circ), ("ival", ival)):
        b = d.boundary
        assert b.dim == d.dim - 1 and b.space == d.space
        cand = tp.spaces.Points(torch.from_numpy(G["cand_" + nm]).cuda(), d.space)
        assert np.array_equal(b._contains(cand).reshape(-1).cpu().numpy(), G["in_" + nm]), nm      # bit exact
        assert abs(float(b.volume()) - float(G["vol_" + nm])) < 1e-6 * float(G["vol_" + nm])
        for n in (1, 7, 50):
            got = b.sample_grid(n, device="cuda").as_tensor.cpu().numpy()
            assert np.allclose(got, G[f"grid_{nm}_{n}"], rtol=1e-6, atol=1e-6), (nm, n)
        nrm = b.normal(tp.spaces.Points(torch.from_numpy(G["on_" + nm]).cuda(), d.space)).cpu().numpy()
        want = G["normal_" + nm]
        assert np.array_equal(np.isnan(nrm), np.isnan(want)), nm
        assert np.allclose(nrm, want, rtol=1e-6, atol=1e-7, equal_nan=True), nm
        # random boundary points: right count, on the boundary (distance, not the reference's brittle isclose)
        torch.manual_seed(5)
        pts = b.sample_random_uniform(n=4000, device="cuda").as_tensor.cpu().numpy()
        assert pts.shape == (4000, d.space.dim)
        if nm == "circ":
            r = np.linalg.norm(pts - np.array([0.1, -0.2], np.float32), axis=1)
            assert np.abs(r - 0.9).max() < 1e-6
            ang = np.arctan2(pts[:, 1] + 0.2, pts[:, 0] - 0.1)
            assert np.histogram(ang, bins=8, range=(-np.pi, np.pi))[0].min() > 380            # uniform in angle
        elif nm == "ival":
            assert set(np.unique(pts).tolist()) == {np.float32(-0.3).item(), np.float32(0.8).item()}
            assert 1800 < (pts == np.float32(-0.3)).sum() < 2200
        else:
            _, _, _, bx, by = geometry._bary(BNDS["para"][1], pts)
            on_edge = np.minimum(np.minimum(np.abs(bx), np.abs(bx - 1)), np.minimum(np.abs(by), np.abs(by - 1)))
            assert on_edge.max() < 1e-6 and bx.min() > -1e-6 and bx.max() < 1 + 1e-6
    for nm, b in (("left", ival.boundary_left), ("right", ival.boundary_right)):
        cand = tp.spaces.Points(torch.from_numpy(G["cand_ival"]).cuda(), T)
        assert np.array_equal(b._contains(cand).reshape(-1).cpu().numpy(), G["in_" + nm])
        g = b.sample_grid(5, device="cuda")
        assert np.allclose(g.as_tensor.cpu().numpy(), G["grid_" + nm])
        assert np.allclose(b.normal(g).cpu().numpy(), G["normal_" + nm])
        assert float(b.volume()) == 1.0
    prod = para.boundary * ival
    cand3 = tp.spaces.Points(torch.from_numpy(G["cand3"]).cuda(), X * T)
    assert np.array_equal(prod._contains(cand3).reshape(-1).cpu().numpy(), G["in_prod"])
    assert abs(float(prod.volume()) - float(G["vol_prod"])) < 1e-5
    s = tp.samplers.RandomUniformSampler(prod, n_points=1000).sample_points(device="cuda")
    assert s.as_tensor.shape == (1000, 3) and bool(ival._contains(s).all())


def test_boundary_condition_trains_through_native_path():
    """a Dirichlet + Neumann boundary condition as the reference's examples write it (poisson-equation.ipynb):
    value and normal derivative of the native FCN on sampled boundary points, gradient to the parameters."""
    import torchphysics_b200 as tp
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    model = tp.models.FCN(X, U, hidden=(32, 32)).cuda()
    sampler = tp.samplers.RandomUniformSampler(sq.boundary, n_points=512).make_static()

    def residual(u, x):
        n = sq.boundary.normal(x)
        return u - 1.0 + tp.utils.normal_derivative(u, n, x)

    cond = tp.conditions.PINNCondition(model, sampler, residual)
    loss = cond(device="cuda")
    loss.backward()
    x = sampler.sample_points(device="cuda").as_tensor.detach().cpu()
    assert bool(((x == 0) | (x == 1)).any(dim=1).all())              # unit square: exact edges
    seq = ref_autograd.fcn_from_params([p.detach().cpu().numpy() for p in model._native_engine().params()])
    cols, xin = ref_autograd.track(x, [("x", 2)])
    u = seq(xin)
    nrm = torch.from_numpy(geometry.boundary_normal(("boundary", ("parallelogram", [0, 0], [1, 0], [0, 1])), x.numpy()))
    want = ref_autograd.pinn_loss(u - 1.0 + ref_autograd.normal_derivative(u, nrm, cols["x"]))
    assert abs(float(loss) - float(want)) < 1e-5 * abs(float(want))
    gs = torch.autograd.grad(want, ref_autograd.linear_params(seq))
    for p, w in zip(model._native_engine().params(), gs):
        assert relerr(p.grad.cpu().numpy(), w.numpy()) < 1e-5


def test_adaptive_rejection_samplers_through_a_condition():
    """points with a high per-point loss survive, the rest are redrawn uniformly; the condition
    hands the unreduced loss of the previous iteration to the sampler (reference condition.py:280-302)."""
    import torchphysics_b200 as tp
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    sq = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    torch.manual_seed(0)
    s = tp.samplers.AdaptiveThresholdRejectionSampler(sq, resample_ratio=0.25, n_points=4000)
    assert s.is_adaptive
    first = s.sample_points(device="cuda").as_tensor.clone()
    loss = first[:, 0].clone()                                   # loss grows with x1
    second = s.sample_points(unreduced_loss=loss, device="cuda").as_tensor
    thr = loss.min() + (loss.max() - loss.min()) * 0.25
    keep = loss >= thr
    assert torch.equal(second[keep], first[keep]) and not torch.equal(second[~keep], first[~keep])
    assert bool(sq._contains(tp.spaces.Points(second, X)).all())
    r = tp.samplers.AdaptiveRandomRejectionSampler(sq, n_points=20000)
    first = r.sample_points(device="cuda").as_tensor.clone()
    loss = first[:, 0].clone()
    second = r.sample_points(unreduced_loss=loss, device="cuda").as_tensor
    kept = (second == first).all(dim=1)
    hi, lo = kept[loss > 0.8].float().mean(), kept[loss < 0.2].float().mean()
    assert hi > 0.8 and lo < 0.2                                  # P(keep) = (loss - min) / (max - min)
    model = tp.models.FCN(X, U, hidden=(16, 16)).cuda()
    cond = tp.conditions.PINNCondition(model, tp.samplers.AdaptiveThresholdRejectionSampler(sq, 0.5, n_points=512),
                                       lambda u, x: tp.utils.laplacian(u, x) + 1.0)
    a = cond(device="cuda")
    assert cond.last_unreduced_loss is not None and cond.last_unreduced_loss.shape == (512,)
    b = cond(device="cuda")
    b.backward()
    assert torch.isfinite(a) and torch.isfinite(b)


def test_host_stream_sampler_delivers_batches_in_order():
    import torchphysics_b200 as tp
    X = tp.spaces.R2("x")
    batches = [torch.rand(1000, 2) + k for k in range(3)]
    s = tp.samplers.HostStreamSampler(batches, X)
    for i in range(7):
        pts = s.sample_points(device="cuda")
        assert pts.space == X and torch.equal(pts.as_tensor.cpu(), batches[i % 3])
    with pytest.raises(ValueError):
        s.sample_points(device="cpu")


def test_cut_boundary_matches_reference():
    """boundary of A - B (a hole, and an overlapping bite): membership bit for bit, normals, volume, grid and
    random point counts, every produced point on the boundary (reference cut.py:111-201)."""
    import torchphysics_b200 as tp
    from tests.test_oracle_golden import CUTS
    G = load("cut_boundary")
    X = tp.spaces.R2("x")
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    sq = tp.domains.Parallelogram(X, [-0.25, -0.25], [0.25, -0.25], [-0.25, 0.25])
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    bite = tp.domains.Circle(X, [0.7, 0.1], 0.45)
    import warnings
    for nm, dom in (("hole", circ - sq), ("bite", para - bite)):
        A, B = CUTS[nm]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            b = dom.boundary
            cand = tp.spaces.Points(torch.from_numpy(G["cand_" + nm]).cuda(), X)
            assert np.array_equal(b._contains(cand).reshape(-1).cpu().numpy(), G["in_" + nm]), nm
            assert abs(float(b.volume()) - float(G["vol_" + nm][0])) < 1e-5
            nrm = b.normal(tp.spaces.Points(torch.from_numpy(G["on_" + nm]).cuda(), X)).cpu().numpy()
            assert np.allclose(nrm, G["normal_" + nm], rtol=1e-6, atol=1e-7), nm
            for n in (40, 200):
                g = b.sample_grid(n, device="cuda").as_tensor.cpu().numpy()
                assert len(g) == int(G[f"ngrid_{nm}_{n}"]) == n
                if nm == "hole":
                    assert np.allclose(g, G[f"grid_{nm}_{n}"], rtol=1e-6, atol=1e-6)
            torch.manual_seed(7)
            pts = b.sample_random_uniform(n=3000, device="cuda")
            assert len(pts) == 3000
            p = pts.as_tensor.cpu().numpy()
            assert geometry.cut_boundary_contains(A, B, p).mean() > 0.97       # (skewed edges: brittle isclose)
            on_a = geometry.boundary_contains(("boundary", A), p)
            share_a = float(geometry.boundary_volume(("boundary", A)) / geometry.cut_boundary_volume(A, B))
            if nm == "hole":                                                    # both boundaries are complete
                assert abs(on_a.mean() - share_a) < 0.04
            d = b.sample_random_uniform(d=50.0, device="cuda")
            assert 0 < len(d) <= int(np.ceil(50.0 * float(b.volume()))) + 2
    # Neumann condition on the boundary of the domain with a hole: normals feed tp.utils.normal_derivative
    U = tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=(16, 16)).cuda()
    dom = circ - sq
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bnd = dom.boundary
        cond = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(bnd, n_points=256),
                                           lambda u, x: tp.utils.normal_derivative(u, bnd.normal(x), x))
        loss = cond(device="cuda")
        loss.backward()
    assert torch.isfinite(loss)


def test_intersection_boundary_matches_reference():
    """boundary of A & B (a lens, and a disc clipped by a square): membership bit for bit, normals, volume, grid / random
    counts and the density grid point for point (reference intersection.py:112-206)."""
    import warnings
    import torchphysics_b200 as tp
    from tests.test_oracle_golden import INTERSECTIONS
    G = load("intersection_boundary")
    X = tp.spaces.R2("x")
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    bite = tp.domains.Circle(X, [0.7, 0.1], 0.45)
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    sq = tp.domains.Parallelogram(X, [0.0, -0.5], [1.5, -0.5], [0.0, 1.0])
    for nm, dom in (("lens", para & bite), ("corner", circ & sq)):
        A, B = INTERSECTIONS[nm]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            b = dom.boundary
            cand = tp.spaces.Points(torch.from_numpy(G["cand_" + nm]).cuda(), X)
            assert np.array_equal(b._contains(cand).reshape(-1).cpu().numpy(), G["in_" + nm]), nm
            assert abs(float(b.volume()) - float(G["vol_" + nm][0])) < 1e-5
            nrm = b.normal(tp.spaces.Points(torch.from_numpy(G["on_" + nm]).cuda(), X)).cpu().numpy()
            assert np.allclose(nrm, G["normal_" + nm], rtol=1e-6, atol=1e-7), nm
            for n in (40, 200):
                g = b.sample_grid(n, device="cuda").as_tensor.cpu().numpy()
                assert len(g) == int(G[f"ngrid_{nm}_{n}"]) == n
                if nm == "corner":               # axis-aligned clip: the reference keeps its own grid points exactly
                    assert np.allclose(g, G[f"grid_{nm}_{n}"], rtol=1e-6, atol=1e-6)
            torch.manual_seed(11)
            p = b.sample_random_uniform(n=3000, device="cuda").as_tensor.cpu().numpy()
            assert len(p) == 3000
            assert geometry.intersection_boundary_contains(A, B, p).mean() > 0.97
            d = b.sample_grid(d=30.0, device="cuda").as_tensor.cpu().numpy()
            assert len(d) == int(G["n_grid_d_" + nm])
            assert np.allclose(d, G["grid_d_" + nm], rtol=1e-6, atol=1e-6)
            r = b.sample_random_uniform(d=50.0, device="cuda")
            assert 0 < len(r) <= int(np.ceil(50.0 * float(b.volume()))) + 2


def test_product_boundary_is_the_union_of_the_two_faces():
    """(A x B).boundary = (dA x B) + (A x dB) (reference product.py:100-106): every point lies on one of the two faces,
    the faces are drawn in proportion to their volumes, membership agrees with the oracle's shape tests."""
    import warnings
    import torchphysics_b200 as tp
    X, T = tp.spaces.R2("x"), tp.spaces.R1("t")
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    ival = tp.domains.Interval(T, -0.3, 0.8)
    A, B = ("circle", [0.1, -0.2], 0.9), ("interval", -0.3, 0.8)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = (circ * ival).boundary
        assert b.space == X * T
        vol_side, vol_caps = 2 * np.pi * 0.9 * 1.1, np.pi * 0.81 * 2.0
        assert abs(float(b.volume()) - (vol_side + vol_caps)) < 1e-4
        torch.manual_seed(5)
        pts = b.sample_random_uniform(n=6000, device="cuda")
        assert len(pts) == 6000 and pts.space == X * T
        p = pts.as_tensor.cpu().numpy()
        side = geometry.boundary_contains(("boundary", A), p[:, :2]) & geometry.contains(B, p[:, 2:])
        caps = geometry.contains(A, p[:, :2]) & geometry.boundary_contains(("boundary", B), p[:, 2:])
        assert (side | caps).all()
        assert abs(side.mean() - vol_side / (vol_side + vol_caps)) < 0.03
        assert b._contains(pts).all()
        off = tp.spaces.Points(torch.tensor([[0.1, -0.2, 0.2], [3.0, 0.0, 0.8]], device="cuda"), X * T)
        assert not b._contains(off).any()


def test_nested_boolean_boundaries_match_reference():
    """boundaries whose operands are boolean domains themselves -- a cut of a cut, a cut of a union, a cut BY a union (the
    reference recurses through domain_a.boundary, cut.py:111-201): membership bit for bit, normals, volume, counts, every
    sampled point on the boundary."""
    import warnings
    import torchphysics_b200 as tp
    from tests.test_oracle_golden import NESTED
    G = load("nested_boundary")
    X = tp.spaces.R2("x")
    circ = tp.domains.Circle(X, [0.1, -0.2], 0.9)
    sq = tp.domains.Parallelogram(X, [-0.25, -0.25], [0.25, -0.25], [-0.25, 0.25])
    bite = tp.domains.Circle(X, [0.7, 0.1], 0.45)
    sq3 = tp.domains.Parallelogram(X, [0.0, -0.8], [0.5, -0.8], [0.0, -0.4])
    sq4 = tp.domains.Parallelogram(X, [-0.5, 0.0], [1.0, 0.0], [-0.5, 0.5])
    cases = (("cutcut", (circ - sq) - bite), ("unioncut", (sq4 + bite) - sq), ("cutunion", circ - (sq + sq3)))
    for nm, dom in cases:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            b = dom.boundary
            cand = tp.spaces.Points(torch.from_numpy(G["cand"]).cuda(), X)
            assert np.array_equal(b._contains(cand).reshape(-1).cpu().numpy(), G["in_" + nm]), nm
            assert abs(float(b.volume()) - float(G["vol_" + nm][0])) < 1e-5
            nrm = b.normal(tp.spaces.Points(torch.from_numpy(G["on_" + nm]).cuda(), X)).cpu().numpy()
            assert np.allclose(nrm, G["normal_" + nm], rtol=1e-6, atol=1e-7), nm
            for n in (40, 200):
                assert len(b.sample_grid(n, device="cuda")) == int(G[f"ngrid_{nm}_{n}"]) == n
            torch.manual_seed(13)
            p = b.sample_random_uniform(n=2000, device="cuda").as_tensor.cpu().numpy()
            assert len(p) == 2000
            assert geometry.any_boundary_contains(NESTED[nm], p).mean() > 0.97      # (isclose membership is brittle by an ulp)
            d = b.sample_random_uniform(d=20.0, device="cuda")
            assert 0 < len(d) <= int(np.ceil(20.0 * float(b.volume()))) + 4
    # a Dirichlet condition on the boundary of the doubly cut disc trains through the native path
    U = tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=(16, 16)).cuda()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bnd = ((circ - sq) - bite).boundary
        cond = tp.conditions.PINNCondition(model, tp.samplers.RandomUniformSampler(bnd, n_points=300), lambda u: u - 1.0)
        loss = cond(device="cuda")
        loss.backward()
    assert torch.isfinite(loss)


def test_union_boundary_matches_reference():
    """boundary of A + B: overlapping shapes and two squares sharing an edge (normal-agreement test of reference
    union.py:161-192): membership bit for bit, normals, volume, counts."""
    import warnings
    import torchphysics_b200 as tp
    from tests.test_oracle_golden import UNIONS
    G = load("union_boundary")
    X = tp.spaces.R2("x")
    para = tp.domains.Parallelogram(X, [-0.25, -0.5], [0.75, -0.25], [0.0, 0.6])
    circ = tp.domains.Circle(X, [0.7, 0.1], 0.45)
    sq1 = tp.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    sq2 = tp.domains.Parallelogram(X, [1, 0], [2, 0], [1, 1])
    for nm, dom in (("lens", para + circ), ("touch", sq1 + sq2)):
        A, B = UNIONS[nm]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            b = dom.boundary
            cand = tp.spaces.Points(torch.from_numpy(G["cand_" + nm]).cuda(), X)
            assert np.array_equal(b._contains(cand).reshape(-1).cpu().numpy(), G["in_" + nm]), nm
            assert abs(float(b.volume()) - float(G["vol_" + nm][0])) < 1e-5
            nrm = b.normal(tp.spaces.Points(torch.from_numpy(G["on_" + nm]).cuda(), X)).cpu().numpy()
            assert np.allclose(nrm, G["normal_" + nm], rtol=1e-6, atol=1e-7), nm
            for n in (40, 200):
                assert len(b.sample_grid(n, device="cuda")) == int(G[f"ngrid_{nm}_{n}"]) == n
            torch.manual_seed(9)
            p = b.sample_random_uniform(n=2000, device="cuda").as_tensor.cpu().numpy()
            assert len(p) == 2000
            if nm == "touch":                    # axis-aligned: the isclose membership is exact
                assert geometry.union_boundary_contains(A, B, p).all()
                assert not np.any(np.isclose(p[:, 0], 1.0) & (p[:, 1] > 1e-6) & (p[:, 1] < 1 - 1e-6))   # no shared edge
            else:
                assert geometry.union_boundary_contains(A, B, p).mean() > 0.97
            d = b.sample_grid(d=30.0, device="cuda")
            assert 0 < len(d) <= int(np.ceil(30.0 * float(b.volume()))) + 2
