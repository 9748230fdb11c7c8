"""Parity of the CUDA jet kernels (through the C ABI) against the oracle and the golden
outputs of the real reference.  Tolerance: 1e-5 max-norm relative (BASELINE.json)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import jet_numpy, ref_autograd
from tests.golden_io import load, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _engine(params, act):
    from torchphysics_b200.engine import FcnEngine
    lins = []
    for i in range(len(params) // 2):
        W = torch.as_tensor(params[2 * i]).float()
        lin = nn.Linear(W.shape[1], W.shape[0])
        with torch.no_grad():
            lin.weight.copy_(W)
            lin.bias.copy_(torch.as_tensor(params[2 * i + 1]).float())
        lins.append(lin.cuda())
    return FcnEngine(lins, act), lins


def _run(params, act, x, dcols, lap_w, gu=None, gdu=None, glap=None):
    from torchphysics_b200.engine import JetSpec
    eng, lins = _engine(params, act)
    spec = JetSpec(dcols, lap_w)
    xs = torch.as_tensor(x).float().cuda()
    u, du, lap = eng.jet(xs, spec)
    loss = 0
    if gu is not None:
        loss = loss + (u * torch.as_tensor(gu).float().cuda()).sum()
    if gdu is not None:
        loss = loss + (du * torch.as_tensor(gdu).float().cuda()).sum()
    if glap is not None:
        loss = loss + (lap * torch.as_tensor(glap).float().cuda()).sum()
    grads = None
    if gu is not None or gdu is not None or glap is not None:
        ps = eng.params()
        grads = torch.autograd.grad(loss, ps, allow_unused=True)
        grads = [None if g is None else g.cpu().numpy() for g in grads]
    torch.cuda.synchronize()
    return u.detach().cpu().numpy(), du.detach().cpu().numpy(), lap.detach().cpu().numpy(), grads


# poisson_4x64_10k: the parity run BASELINE.md section 3 names (config 1 at 10,000 points)
@pytest.mark.parametrize("name,act", [("poisson_4x64", "tanh"), ("poisson_3x20_sin", "sin"), ("poisson_4x64_10k", "tanh")])
def test_poisson_golden(name, act):
    G = load(name)
    x = G["x"]
    N = x.shape[0]
    f = ref_autograd.poisson_source(torch.from_numpy(x)).numpy()
    u, du, lap, _ = _run(G["params"], act, x, (0, 1), (1.0, 1.0))
    assert relerr(u, G["u"]) < TOL
    assert relerr(du[:, 0, :], G["grad"]) < TOL
    assert relerr(lap, G["lap"]) < TOL
    r = lap + f
    loss = float(np.mean(r.astype(np.float64) ** 2))
    assert abs(loss - float(G["loss"])) < TOL * abs(float(G["loss"]))
    _, _, _, grads = _run(G["params"], act, x, (0, 1), (1.0, 1.0), glap=2 * r / N)
    assert grads[-1] is None and G["dtheta"][-1] is None     # reference: output bias has no grad
    for g, w in zip(grads[:-1], G["dtheta"][:-1]):
        assert relerr(g, w) < TOL


# 128: wide tensor-core path (reverse sweep); heat_6x128: BASELINE.json config 2 at full depth
@pytest.mark.parametrize("name", ["heat_3x48", "heat_3x128", "heat_6x128"])
def test_heat_golden(name):
    G = load(name)
    xt = G["xt"]
    N = xt.shape[0]
    u, du, lap, _ = _run(G["params"], "tanh", xt, (0, 1, 2), (1.0, 1.0, 0.0))
    assert relerr(u, G["u"]) < TOL
    assert relerr(du[:, 0, 2:3], G["ut"]) < TOL
    assert relerr(lap, G["lap"]) < TOL
    r = du[:, 0, 2:3] - lap
    gdu = np.zeros((N, 1, 3), np.float32)
    gdu[:, 0, 2] = (2 * r / N)[:, 0]
    _, _, _, grads = _run(G["params"], "tanh", xt, (0, 1, 2), (1.0, 1.0, 0.0), gdu=gdu, glap=-2 * r / N)
    for g, w in zip(grads[:-1], G["dtheta"][:-1]):
        assert relerr(g, w) < TOL


@pytest.mark.parametrize("name", ["ns_3x32", "ns_3x128", "ns_6x128"])     # ns_6x128: config 3 at full depth
def test_navier_stokes_golden(name):
    G = load(name)
    u, du, lap, _ = _run(G["params"], "tanh", G["x"], (0, 1), (1.0, 1.0))
    assert relerr(u, G["y"]) < TOL
    assert relerr(du, G["jac"]) < TOL
    assert relerr(lap[:, :2], G["lap"]) < TOL


@pytest.mark.parametrize("name", ["ritz_4x40", "ritz_2x256", "ritz_8x256"])   # ritz_8x256: config 5 at full depth
def test_deep_ritz_golden(name):
    G = load(name)
    x = G["x"]
    N = x.shape[0]
    u, du, lap, _ = _run(G["params"], "tanh", x, (0, 1), None)
    assert lap.size == 0
    assert relerr(du[:, 0, :], G["grad"]) < TOL
    _, _, _, grads = _run(G["params"], "tanh", x, (0, 1), None, gu=-np.ones((N, 1), np.float32) / N, gdu=du / N)
    for g, w in zip(grads, G["dtheta"]):
        assert relerr(g, w) < TOL


CASES = [
    # d_in, d_out, hidden, act, dcols, lap_w, N
    (2, 1, (64,) * 4, "tanh", (0, 1), (1.0, 1.0), 1000),
    (3, 1, (128,) * 6, "tanh", (0, 1, 2), (1.0, 1.0, 0.0), 333),
    (2, 3, (128,) * 6, "tanh", (0, 1), (1.0, 1.0), 200),
    (2, 1, (256,) * 8, "tanh", (0, 1), None, 150),
    (2, 1, (256,) * 3, "sin", (0, 1), (1.0, 0.5), 100),
    (1, 64, (64,) * 4, "tanh", (0,), (1.0,), 300),
    (4, 2, (20, 50, 30), "tanh", (3, 1), (0.0, 2.0), 77),
    (3, 1, (17,), "sin", (), None, 65),
    (2, 1, (32, 32), "tanh", (1,), None, 1),
    # more shapes of the tensor-core path (width <= 64, tanh): 3 streams incl. an idle quadrant, value only,
    # three directions without Laplacian, 6 hidden layers, 4 outputs, one input column, ragged tile counts
    (1, 1, (64,) * 6, "tanh", (0,), (1.0,), 500),            # 6 hidden layers: beyond the resident-weight budget -> FP32 kernels
    (1, 1, (64,) * 5, "tanh", (0,), (1.0,), 500),            # deepest tensor-core network (1 wide + 3 narrow dW accumulators)
    (2, 1, (64,) * 2, "tanh", (0, 1), (1.0, 1.0), 700),       # shallowest (1 hidden matrix)
    (6, 1, (32, 32), "tanh", (0, 5), None, 301),               # d_in > 4: tensor-core forward, FP32 reverse
    (2, 6, (48, 48), "tanh", (0, 1), (1.0, 1.0), 222),         # d_out > 4: tensor-core forward, FP32 reverse
    (8, 8, (64, 64, 64), "tanh", (7,), (3.0,), 64),
    (2, 4, (48,) * 3, "tanh", (0, 1), None, 129),
    (3, 2, (64, 64), "tanh", (), None, 95),
    (3, 1, (40,) * 5, "tanh", (0, 1, 2), None, 257),
    (4, 1, (64,) * 4, "tanh", (2, 0), (0.5, 1.5), 4099),
    # wide tensor-core path (hidden width 65..128, tanh, d_in <= 4, d_out <= 8; csrc/tcw_kernels.cu): ragged widths,
    # every stream count 1..5, one hidden matrix, 8 outputs, tile counts below / above the SM count, a single point
    (2, 1, (128,) * 2, "tanh", (0, 1), (1.0, 1.0), 16),
    (2, 3, (100, 128, 96), "tanh", (0, 1), (1.0, 1.0), 200),
    (4, 8, (128,) * 3, "tanh", (), None, 777),
    (1, 2, (96,) * 4, "tanh", (0,), None, 5000),
    (3, 2, (65, 80), "tanh", (2, 0), None, 1),
    (4, 1, (128, 72, 128, 90), "tanh", (3, 0, 1), None, 2371),
    (2, 4, (112,) * 3, "tanh", (1,), (0.75,), 2500),
    # 5..8 inputs: layer 0 has kernels of its own on the wide path (w_in_fwd_kernel / w_in_bwd_kernel); narrow nets are padded to it
    (5, 1, (128,) * 3, "tanh", (0, 4), (1.0, 1.0), 150),
    (6, 2, (64,) * 3, "tanh", (1, 5, 3), (1.0, 0.5, 0.0), 700),
    (8, 1, (200, 256), "tanh", (0, 7), None, 300),
    (5, 3, (96,) * 3, "sin", (4,), (1.0,), 1001),
    (7, 20, (48, 48), "tanh", (6, 2), (1.0, 1.0), 77),
    # widths 129..256: two 128-neuron blocks per layer on the wide path
    (3, 2, (200, 256, 130), "tanh", (0, 1, 2), (1.0, 1.0, 0.0), 1000),
    (2, 1, (256, 256), "tanh", (0, 1), (1.0, 1.0), 5000),
    (2, 3, (129,) * 4, "tanh", (1,), None, 333),
    # tp.models.Sinus on the width-64 tensor-core kernels (ACT = 1 instantiations) ...
    (2, 2, (64,) * 3, "sin", (0, 1), (1.0, 1.0), 999),
    (3, 1, (20, 50, 30), "sin", (2,), None, 130),
    # ... and on the wide path (csrc/tcw_sin.cu)
    (2, 1, (128,) * 3, "sin", (0, 1), (1.0, 1.0), 1000),
    (3, 2, (200, 256, 130), "sin", (0, 1, 2), (1.0, 1.0, 0.0), 1000),
    # narrow but deeper than the resident-weight kernels hold (> 5 hidden layers): one zero-padded 128-neuron block per
    # layer on the wide path instead of the FP32 CUDA-core kernels
    (2, 1, (64,) * 6, "tanh", (0, 1), (1.0, 1.0), 3000),
    (3, 2, (20,) * 8, "tanh", (0, 1, 2), (1.0, 1.0, 0.0), 500),
    (2, 1, (50,) * 7, "sin", (0, 1), (1.0, 1.0), 1000),
    # 9..128 outputs (DeepONet trunks, o * p columns): the output layer is one more UMMA product of the wide path
    (2, 64, (64,) * 4, "tanh", (0, 1), (1.0, 1.0), 1000),
    (1, 32, (32, 32), "tanh", (0,), None, 300),
    (3, 40, (128, 100, 128), "sin", (0, 1, 2), (1.0, 1.0, 0.0), 700),
    (2, 128, (64,) * 3, "tanh", (1,), (1.0,), 555),
    (2, 12, (96,) * 2, "tanh", (), None, 100),
    (2, 9, (20, 20), "tanh", (0, 1), (1.0, 1.0), 17),
    # ... more than 128 outputs: several 128-output blocks whose adjoint products meet in the partial-product array (the shape
    # of a DeepONet with the branch coefficients folded into the trunk's output layer: functions x output dim columns)
    (1, 256, (64,) * 4, "tanh", (0,), (1.0,), 900),
    (2, 300, (128, 96), "tanh", (0, 1), (1.0, 1.0), 401),
    (3, 200, (48,) * 3, "sin", (0, 2), None, 333),
    # ... on 256-wide hidden layers (two input blocks per output block, the K blocks meet in the partial-product array)
    (2, 40, (256, 200), "tanh", (0, 1), (1.0, 1.0), 500),
    (2, 300, (160, 256, 256), "tanh", (1,), None, 257),
    (6, 130, (256, 256), "sin", (0, 5), (1.0, 1.0), 100),
]


@pytest.mark.parametrize("d_in,d_out,hidden,act,dcols,lap_w,N", CASES)
def test_against_float64_oracle(d_in, d_out, hidden, act, dcols, lap_w, N):
    """Full-size networks of BASELINE.json's configs (and ragged shapes) vs the float64
    closed-form oracle on seeded inputs."""
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
    # oracle propagates all d_in directions; pick the requested ones
    c = np.zeros(d_in)
    if lap_w is not None:
        for col, w in zip(dcols, lap_w):
            c[col] = w
    ou, odu, olap = jet_numpy.jet_forward(params, x, c if lap_w is not None else None, act)
    assert relerr(u, ou) < TOL
    if nd:
        assert relerr(du, odu[:, :, list(dcols)]) < TOL
    if lap_w is not None:
        assert relerr(lap, olap) < TOL
    full_gdu = np.zeros((N, d_out, d_in))
    if nd:
        full_gdu[:, :, list(dcols)] = gdu
    ograds = jet_numpy.jet_backward(params, x, gu, full_gdu, glap, c if lap_w is not None else None, act)
    for g, w in zip(grads, ograds):
        assert relerr(g, w) < TOL


def test_empty_batch():
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(2, 1, (32, 32))
    params = [p.detach().numpy() for p in ref_autograd.linear_params(seq)]
    u, du, lap, _ = _run(params, "tanh", np.zeros((0, 2), np.float32), (0, 1), (1.0, 1.0))
    assert u.shape == (0, 1) and du.shape == (0, 1, 2) and lap.shape == (0, 1)


def test_linearity_of_reverse_at_scale():
    """Size-independent property at a BASELINE-sized batch: dtheta is linear in the adjoints and
    additive over a split of the points."""
    from torchphysics_b200.engine import FcnEngine, JetSpec
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(2, 1, (64,) * 4).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    spec = JetSpec((0, 1), (1.0, 1.0))
    N = 1 << 20
    x = torch.rand(N, 2, device="cuda")
    g = torch.randn(N, 1, device="cuda")
    packed = eng.packed(eng.params())
    full = eng.backward_raw(packed, x, spec, None, None, g)
    a = eng.backward_raw(packed, x[: N // 3].contiguous(), spec, None, None, g[: N // 3].contiguous())
    b = eng.backward_raw(packed, x[N // 3:].contiguous(), spec, None, None, g[N // 3:].contiguous())
    twice = eng.backward_raw(packed, x, spec, None, None, 2 * g)
    # random-sign adjoints are the worst case for a float32 sum over 2^20 points (the result is ~sqrt(N) while the summed
    # magnitudes are ~N).  The TMEM accumulators of dW (round toward zero) are flushed every 8 tiles into round-to-nearest
    # fp32 partial sums per CTA, float64 across CTAs: before that flush this deviation was ~1e-4
    den = full.abs().max()
    assert ((a + b - full).abs().max() / den) < TOL
    assert ((twice - 2 * full).abs().max() / den) < 1e-6


@pytest.mark.parametrize("hidden,dcols,lap_w,d_out", [((64,) * 4, (0, 1), (1.0, 1.0), 1), ((40,) * 3, (0, 1), None, 2),
                                                       ((20, 50, 30), (1,), (2.0,), 3)])
def test_saved_checkpoints_equal_recompute(hidden, dcols, lap_w, d_out):
    """tensor-core path: the reverse kernel fed with the checkpoints the forward kernel saved returns the same
    gradient as the reverse kernel that recomputes the forward sweep, and both agree with the FP32 kernels."""
    import os
    import subprocess
    import sys
    from torchphysics_b200.engine import FcnEngine, JetSpec
    torch.manual_seed(3)
    seq = ref_autograd.build_fcn(2, d_out, hidden).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    spec = JetSpec(dcols, lap_w)
    N = 5000                                   # not a multiple of the 32-point tile
    x = torch.rand(N, 2, device="cuda")
    packed = eng.packed(eng.params())
    gu = torch.randn(N, d_out, device="cuda")
    gdu = torch.randn(N, d_out, len(dcols), device="cuda")
    glap = torch.randn(N, d_out, device="cuda") if lap_w else None
    nbytes = eng.saved_bytes(spec, N)
    # checkpoints + the per-CTA fp32 partial sums of dW (160 CTAs x 5 matrices x 64 x 64) the reverse kernel flushes into
    assert nbytes == ((N + 31) // 32) * len(hidden) * 4 * 64 * 32 * 4 + 160 * 5 * 64 * 64 * 4
    saved = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    u1, du1, lap1 = eng.forward_raw(packed, x, spec, saved)
    u0, du0, lap0 = eng.forward_raw(packed, x, spec)
    assert torch.equal(u0, u1) and torch.equal(du0, du1)
    g_saved = eng.backward_raw(packed, x, spec, gu, gdu, glap, saved)
    g_recomp = eng.backward_raw(packed, x, spec, gu, gdu, glap)
    den = g_recomp.abs().max()
    assert float((g_saved - g_recomp).abs().max() / den) < 2e-6
    # the FP32 CUDA-core kernels (TPX_DISABLE_TC=1, separate process: the switch is read once per process)
    code = (
        "import sys, torch, torch.nn as nn; sys.path.insert(0, '.');"
        "from oracle import ref_autograd; from torchphysics_b200.engine import FcnEngine, JetSpec;"
        f"torch.manual_seed(3); seq = ref_autograd.build_fcn(2, {d_out}, {hidden!r}).cuda();"
        "eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], 'tanh');"
        f"spec = JetSpec({dcols!r}, {lap_w!r}); x = torch.rand({N}, 2, device='cuda');"
        f"gu = torch.randn({N}, {d_out}, device='cuda'); gdu = torch.randn({N}, {d_out}, {len(dcols)}, device='cuda');"
        + (f"glap = torch.randn({N}, {d_out}, device='cuda');" if lap_w else "glap = None;") +
        "packed = eng.packed(eng.params()); g = eng.backward_raw(packed, x, spec, gu, gdu, glap);"
        "torch.save(g.cpu(), sys.argv[1])"
    )
    out = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"tpx_fp32_{os.getpid()}.pt")
    env = dict(os.environ, TPX_DISABLE_TC="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code, out], check=True, env=env, cwd=root, timeout=300)
    g_fp32 = torch.load(out).cuda()
    os.remove(out)
    n32 = min(g_fp32.numel(), g_saved.numel())
    # both packed buffers start with the FP32-kernel image (same indexing)
    assert float((g_saved[:n32] - g_fp32[:n32]).abs().max() / den) < 5e-6


@pytest.mark.parametrize("hidden,dcols,lap_w,d_out,d_in", [((128,) * 4, (0, 1, 2), (1.0, 1.0, 0.0), 1, 3),
                                                            ((96, 128, 80), (0, 1), (1.0, 1.0), 3, 2),
                                                            ((256, 192, 256), (0, 1), None, 1, 2),
                                                            ((64,) * 4, (0, 1), (1.0, 1.0), 64, 2)])
def test_wide_path_saved_workspace_and_fp32(hidden, dcols, lap_w, d_out, d_in):
    """wide tensor-core path: (a) checkpoint-saving forward + reverse over the saved buffer (the training step), (b) the
    same kernels in chunks of 65536 points through the caller workspace (no saved buffer), (c) the FP32 CUDA-core kernels
    (TPX_DISABLE_TCW=1, separate process) agree on the same inputs.  The saved buffer holds the checkpoints of hidden
    layers 1..L-1 and the scratch arrays of the reverse sweep."""
    import os
    import subprocess
    import sys
    from torchphysics_b200.engine import FcnEngine, JetSpec
    torch.manual_seed(7)
    seq = ref_autograd.build_fcn(d_in, d_out, hidden).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    spec = JetSpec(dcols, lap_w)
    N = 70001                                  # not a multiple of the 16-point tile, more tiles than SMs, two workspace chunks
    S = 1 + len(dcols) + (lap_w is not None)
    x = torch.rand(N, d_in, device="cuda")
    packed = eng.packed(eng.params())
    gu = torch.randn(N, d_out, device="cuda")
    gdu = torch.randn(N, d_out, len(dcols), device="cuda")
    glap = torch.randn(N, d_out, device="cuda") if lap_w else None
    nbytes = eng.saved_bytes(spec, N)
    nblk = 2 if max(hidden) > 128 else 1       # 128-neuron blocks per layer; one more half-array of partial products for two
    # + the per-CTA fp32 partial sums of dW (160 CTAs x 128 x 128) the reverse kernel flushes its TMEM accumulator into
    assert nbytes == ((N + 15) // 16) * ((len(hidden) + 1) * nblk + (nblk > 1)) * S * 16 * 128 * 4 + 160 * 128 * 128 * 4
    saved = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    u1, du1, lap1 = eng.forward_raw(packed, x, spec, saved)
    u0, du0, lap0 = eng.forward_raw(packed, x, spec)          # chunks through the workspace: the same arithmetic per point
    assert torch.equal(u1, u0) and torch.equal(du1, du0) and (not lap_w or torch.equal(lap1, lap0))
    g_saved = eng.backward_raw(packed, x, spec, gu, gdu, glap, saved)
    g_ws = eng.backward_raw(packed, x, spec, gu, gdu, glap)
    den = g_saved.abs().max()
    assert float((g_saved - g_ws).abs().max() / den) < 2e-6    # dW sums are split differently over CTAs / chunks
    # a second reverse sweep over the same saved buffer (the checkpoints are not consumed) gives the same result up to
    # the order of the atomics (fp32 in shared memory for the output layer, float64 across CTAs)
    g_again = eng.backward_raw(packed, x, spec, gu, gdu, glap, saved)
    assert float((g_again - g_saved).abs().max() / den) < 1e-6
    # FP32 CUDA-core kernels in a separate process (the switch is read once per process)
    code = (
        "import sys, torch, torch.nn as nn; sys.path.insert(0, '.');"
        "from oracle import ref_autograd; from torchphysics_b200.engine import FcnEngine, JetSpec;"
        f"torch.manual_seed(7); seq = ref_autograd.build_fcn({d_in}, {d_out}, {hidden!r}).cuda();"
        "eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], 'tanh');"
        f"spec = JetSpec({dcols!r}, {lap_w!r}); x = torch.rand({N}, {d_in}, device='cuda');"
        f"gu = torch.randn({N}, {d_out}, device='cuda'); gdu = torch.randn({N}, {d_out}, {len(dcols)}, device='cuda');"
        + (f"glap = torch.randn({N}, {d_out}, device='cuda');" if lap_w else "glap = None;") +
        "packed = eng.packed(eng.params()); u, du, lap = eng.forward_raw(packed, x, spec);"
        "g = eng.backward_raw(packed, x, spec, gu, gdu, glap);"
        "torch.save({'g': g.cpu(), 'u': u.cpu(), 'du': du.cpu(), 'lap': None if lap is None else lap.cpu()}, sys.argv[1])"
    )
    out = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"tpx_fp32w_{os.getpid()}.pt")
    env = dict(os.environ, TPX_DISABLE_TCW="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code, out], check=True, env=env, cwd=root, timeout=300)
    ref = torch.load(out)
    os.remove(out)
    # two fp32 evaluations, each ~2e-6 from the truth
    assert float((u1.cpu() - ref["u"]).abs().max() / ref["u"].abs().max()) < 6e-6
    assert float((du1.cpu() - ref["du"]).abs().max() / ref["du"].abs().max()) < 6e-6
    if lap_w:
        assert float((lap1.cpu() - ref["lap"]).abs().max() / ref["lap"].abs().max()) < 6e-6
    assert float((g_saved.cpu() - ref["g"]).abs().max() / den) < 8e-6     # (64 outputs with random-sign adjoints: 6.1e-6)


def test_wide_path_reverse_is_linear_at_scale():
    """size-independent property at a BASELINE-sized batch (config 2 network, 2^18 points): dtheta is additive
    over a split of the points and linear in the adjoints."""
    from torchphysics_b200.engine import FcnEngine, JetSpec
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(3, 1, (128,) * 6).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    spec = JetSpec((0, 1, 2), (1.0, 1.0, 0.0))
    N = 1 << 18
    x = torch.rand(N, 3, device="cuda")
    gdu = torch.randn(N, 1, 3, device="cuda")
    glap = torch.randn(N, 1, device="cuda")
    packed = eng.packed(eng.params())

    def sweep(lo, hi, scale=1.0):
        xs = x[lo:hi].contiguous()
        saved = torch.empty(eng.saved_bytes(spec, hi - lo), dtype=torch.uint8, device="cuda")
        eng.forward_raw(packed, xs, spec, saved)
        return eng.backward_raw(packed, xs, spec, None, (scale * gdu[lo:hi]).contiguous(), (scale * glap[lo:hi]).contiguous(), saved)

    full = sweep(0, N)
    a, b = sweep(0, N // 3), sweep(N // 3, N)
    twice = sweep(0, N, 2.0)
    den = full.abs().max()
    # dW accumulates in the fp32 TMEM accumulator (round toward zero) for at most 64 steps of 8 points before it is moved
    # into round-to-nearest fp32 partial sums per CTA and float64 across CTAs
    assert float((a + b - full).abs().max() / den) < TOL
    assert float((twice - 2 * full).abs().max() / den) < 1e-6


@pytest.mark.parametrize("width,N", [(64, 1 << 20), (128, 1 << 19)])
def test_reverse_at_scale_equals_float64_sum_of_short_sweeps(width, N):
    """BASELINE-sized batch: dL/dtheta of ONE sweep over N points (hundreds of tiles per CTA, the TMEM dW accumulators are
    flushed periodically) against the float64 sum of 64 short sweeps (a few tiles per CTA each: no accumulation length to
    speak of).  Coherent adjoints (all ones) expose a round-toward-zero accumulator, random-sign ones cancellation."""
    from torchphysics_b200.engine import FcnEngine, JetSpec
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(2, 1, (width,) * 4).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    spec = JetSpec((0, 1), (1.0, 1.0))
    x = torch.rand(N, 2, device="cuda")
    packed = eng.packed(eng.params())

    def sweep(lo, hi, glap):
        xs, g = x[lo:hi].contiguous(), glap[lo:hi].contiguous()
        saved = torch.empty(eng.saved_bytes(spec, hi - lo), dtype=torch.uint8, device="cuda")
        eng.forward_raw(packed, xs, spec, saved)
        return eng.backward_raw(packed, xs, spec, None, None, g, saved)

    for glap in (torch.ones(N, 1, device="cuda") / N, torch.randn(N, 1, device="cuda") / N):
        full = sweep(0, N, glap)
        C = N // 64
        ref = sum(sweep(i, i + C, glap) for i in range(0, N, C))
        assert float((full - ref).abs().max() / ref.abs().max()) < TOL       # 1.1e-4 before the periodic flush


def test_round2_switch_is_active_in_the_child_run():
    """runs only inside the child pytest of the next test: with TPX_TCP=1 the checkpoint buffer of a 4 x 64 net has the layout of
    tcp_fwd / tce_bwd (128-point tiles + row statistics), not tc_fwd's -- i.e. the switch reached the library"""
    import os
    if os.environ.get("TPX_TCP") != "1":
        pytest.skip("only meaningful under TPX_TCP=1")
    from torchphysics_b200.engine import FcnEngine, JetSpec
    seq = ref_autograd.build_fcn(2, 1, (64,) * 4).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    n = 4096
    assert eng.saved_bytes(JetSpec((0, 1), (1.0, 1.0)), n) != ((n + 31) // 32) * 4 * 4 * 64 * 32 * 4 + 160 * 5 * 64 * 64 * 4


def test_round2_split_fp16_kernels_in_a_subprocess():
    """TPX_TCP=1 (a development A/B switch, read once per process) routes width <= 64 networks through the round-2
    kernels tcp_fwd_kernel (points as TMEM lanes, split-fp16 operands with power-of-two scales from bounds) and
    tce_bwd_kernel: the golden and float64-oracle parity tests must hold on that path too."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, TPX_TCP="1", TPX_TEST_SUBPROCESS="1")       # (conftest.py removes the switches otherwise)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "test_poisson_golden or test_heat_golden or test_navier_stokes_golden or test_deep_ritz_golden "
                        "or test_against_float64_oracle or test_linearity_of_reverse_at_scale or test_round2_switch_is_active"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "skipped" not in r.stdout.splitlines()[-1], r.stdout[-500:]           # the switch check ran, i.e. the child saw TPX_TCP


@pytest.mark.parametrize("d_in,d_out,hidden,dcols,lap_w,N", [(1, 256, (64,) * 4, (0,), (1.0,), 16384 + 5),
                                                             (2, 40, (128, 96), (0, 1), None, 333),
                                                             (2, 300, (32, 32, 32), (1,), (0.5,), 70001),
                                                             (2, 150, (256, 256), (0, 1), (1.0, 1.0), 4097)])
def test_output_major_streams_equal_the_transposed_standard_ones(d_in, d_out, hidden, dcols, lap_w, N):
    """tpx_jet_desc.out_major (the layout of a DeepONet's (function, point) outputs): the same kernels write u (d_out, n),
    du (nd, d_out, n), lap (d_out, n) and read the adjoints in that layout -- bit-identical streams, dtheta up to the order of
    the reductions; with and without a saved buffer (the second in chunks through the workspace: a column range per chunk)."""
    from torchphysics_b200.engine import FcnEngine, JetSpec
    torch.manual_seed(11)
    seq = ref_autograd.build_fcn(d_in, d_out, hidden).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    std, tr = JetSpec(dcols, lap_w), JetSpec(dcols, lap_w, out_major=True)
    x = torch.rand(N, d_in, device="cuda")
    packed = eng.packed(eng.params())
    gu = torch.randn(N, d_out, device="cuda")
    gdu = torch.randn(N, d_out, len(dcols), device="cuda")
    glap = torch.randn(N, d_out, device="cuda") if lap_w else None
    gu_t, gdu_t = gu.t().contiguous(), gdu.permute(2, 1, 0).contiguous()
    glap_t = glap.t().contiguous() if lap_w else None
    for use_saved in (True, False):
        s1 = torch.empty(eng.saved_bytes(std, N), dtype=torch.uint8, device="cuda") if use_saved else None
        s2 = torch.empty(eng.saved_bytes(tr, N), dtype=torch.uint8, device="cuda") if use_saved else None
        u, du, lap = eng.forward_raw(packed, x, std, s1)
        ut, dut, lapt = eng.forward_raw(packed, x, tr, s2)
        assert ut.shape == (d_out, N) and dut.shape == (len(dcols), d_out, N)
        assert torch.equal(ut.t(), u) and torch.equal(dut.permute(2, 1, 0), du)
        if lap_w:
            assert torch.equal(lapt.t(), lap)
        g1 = eng.backward_raw(packed, x, std, gu, gdu, glap, s1)
        g2 = eng.backward_raw(packed, x, tr, gu_t, gdu_t, glap_t, s2)
        assert float((g1 - g2).abs().max() / g1.abs().max()) < 1e-6
    small = FcnEngine([m for m in ref_autograd.build_fcn(2, 3, (32, 32)).cuda() if isinstance(m, nn.Linear)], "tanh")
    with pytest.raises(Exception):                       # 8 or fewer outputs: no output-major kernels, refused loudly
        small.forward_raw(small.packed(small.params()), torch.rand(10, 2, device="cuda"), JetSpec((0,), None, out_major=True))
