#!/usr/bin/env python
"""Benchmark of the PINN residual hot path (BASELINE.json metric): collocation-point residual+grad evaluations per
second.  The headline (no flags) is BASELINE.json's config 1: 2-D Poisson -Laplace(u) = f, FCN 4x64 tanh.

One "step" = for one batch of collocation points per GPU: [sampler] -> forward jet (u, grad u, Laplace u ...) ->
the condition's residual -> loss -> reverse sweep dLoss/dtheta, and (N GPUs) ONE all-reduce of the flat gradient.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--points P] [--impl native|reference]

`value`    config 1: points/s with the points already resident in HBM (C-ABI kernels; the data function f and the loss are
           evaluated inside the step).  Configs 2-5: the whole `Condition.forward() + backward()` step through the public
           API with the Philox sampler inside.
`e2e`      the same step through the public API with the points in pinned HOST memory: H2D of the batch every step, D2H
           of the loss every step (the last one inside the timed region too).
`roofline` the dominant kernel (reverse sweep) timed alone with CUDA events, in algorithmic GEMM FLOP.
`configs`  (default run) short runs of BASELINE.json's other configurations at their stated sizes, the literal 10k-point
           config 1 and config 1 with the sampler inside; with --gpus N, configs 3 and 5 as STRONG splits over the ranks.
`cpu_baseline` / `--impl reference`: the REAL reference (baseline/_ref, pip-installed from /root/reference, with the two
           import stubs of oracle/stubs) on the host cores: its own FCN / samplers / PINNCondition on a bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "collocation-point residual+grad evals/sec (PINN Laplacian, FCN 4x64)"


def metric_name(cfg):
    return METRIC if cfg == 1 else f"residual+grad evals/sec, BASELINE.json config {cfg}"


def f_alg(d_in, width, depth, d_out, streams):
    """algorithmic GEMM FLOP per point of a residual+grad step: 3 x forward"""
    return 3 * 2 * (d_in * width + streams * (depth - 1) * width * width + streams * width * d_out)


# BASELINE.json `configs`: what is built, at which size, how it splits over ranks
CONFIGS = {
    1: dict(name="2D Poisson -Lap u = f on the unit square, FCN 4x64 tanh, residual lap(u)+f, loss mean(r^2), dLoss/dtheta",
            points=1 << 21, scaling="weak", flop=f_alg(2, 64, 4, 1, 4), cpu_points=20000, unit="points/s"),
    2: dict(name="2D heat u_t = Lap_x u on Parallelogram x Interval, FCN 6x128 tanh, residual grad(u,t)-lap(u,x)",
            points=1 << 20, scaling="weak", flop=f_alg(3, 128, 6, 1, 5), cpu_points=10000, unit="points/s"),
    3: dict(name="steady Navier-Stokes on the unit square (lid-driven cavity), FCN 6x128, outputs (v in R2, p): convective + "
                 "2 x laplacian + grad(p) + div", points=1 << 22, scaling="strong", flop=f_alg(2, 128, 6, 3, 4),
            cpu_points=10000, unit="points/s"),
    4: dict(name="physics-informed DeepONet, trunk / branch FCN 4x64, 50 sensors, p = 64, residual lap(u,t)+f, "
                 "256 functions x 16,384 trunk points", points=256 * 16384, scaling="weak", flop=1922,
            cpu_points=64 * 1024, unit="(function, point) pairs/s"),
    5: dict(name="Deep Ritz on Circle - Parallelogram (rejection sampling), FCN 8x256, integrand 0.5|grad u|^2 - u",
            points=1 << 24, scaling="strong", flop=f_alg(2, 256, 8, 1, 3), cpu_points=10000, unit="points/s"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--points", type=int, default=0, help="collocation points per step (default: the configuration's)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# one builder for both arms: `tpm` is torchphysics_b200 (native) or the reference's torchphysics
# ----------------------------------------------------------------------------------------
def source(x):
    import torch
    return 2 * math.pi ** 2 * torch.sin(math.pi * x[:, :1]) * torch.sin(math.pi * x[:, 1:2])


def build_config(tpm, cfg, n, device, sampler=None):
    """returns (module whose parameters are trained, condition, input space).  `sampler` replaces the configuration's
    own random sampler (host-streamed batches for the e2e measurement)."""
    import torch
    X, T, U, V, P = tpm.spaces.R2("x"), tpm.spaces.R1("t"), tpm.spaces.R1("u"), tpm.spaces.R2("v"), tpm.spaces.R1("p")
    sq = tpm.domains.Parallelogram(X, [0, 0], [1, 0], [0, 1])
    if cfg == 1:
        model = tpm.models.FCN(X, U, hidden=(64,) * 4).to(device)
        s = sampler or tpm.samplers.RandomUniformSampler(sq, n_points=n)
        cond = tpm.conditions.PINNCondition(model, s, lambda u, x, f: tpm.utils.laplacian(u, x) + f, data_functions={"f": source})
        return model, cond, X
    if cfg == 2:
        model = tpm.models.FCN(X * T, U, hidden=(128,) * 6).to(device)
        s = sampler or tpm.samplers.RandomUniformSampler(sq * tpm.domains.Interval(T, 0, 1), n_points=n)
        cond = tpm.conditions.PINNCondition(model, s, lambda u, x, t: tpm.utils.grad(u, t) - tpm.utils.laplacian(u, x))
        return model, cond, X * T
    if cfg == 3:
        model = tpm.models.FCN(X, V * P, hidden=(128,) * 6).to(device)
        s = sampler or tpm.samplers.RandomUniformSampler(sq, n_points=n)

        def residual(v, p, x):
            lap = torch.cat([tpm.utils.laplacian(v[:, :1], x), tpm.utils.laplacian(v[:, 1:], x)], dim=1)
            mom = tpm.utils.convective(v, v, x) - 0.01 * lap + tpm.utils.grad(p, x)
            return torch.cat([mom, tpm.utils.div(v, x)], dim=1)
        return model, tpm.conditions.PINNCondition(model, s, residual), X
    if cfg == 5:
        model = tpm.models.FCN(X, U, hidden=(256,) * 8).to(device)
        dom = tpm.domains.Circle(X, [0, 0], 1.0) - tpm.domains.Parallelogram(X, [-0.25, -0.25], [0.25, -0.25], [-0.25, 0.25])
        s = sampler or tpm.samplers.RandomUniformSampler(dom, n_points=n)
        cond = tpm.conditions.DeepRitzCondition(model, s, lambda u, x: 0.5 * (tpm.utils.grad(u, x) ** 2).sum(dim=1, keepdim=True) - u)
        return model, cond, X
    if cfg == 4:
        n_fn = 256 if n >= 256 * 1024 else 64
        n_pts = n // n_fn
        Fs, K = tpm.spaces.R1("f"), tpm.spaces.R1("k")
        fspace = tpm.spaces.FunctionSpace(T, Fs)
        ival = tpm.domains.Interval(T, 0, 1)
        grid = tpm.samplers.GridSampler(ival, n_points=50).sample_points(device=device)
        trunk = tpm.models.FCTrunkNet(T, hidden=(64,) * 4, default_trunk_input=grid)
        branch = tpm.models.FCBranchNet(fspace, hidden=(64,) * 4, grid=grid)
        net = tpm.models.DeepONet(trunk, branch, U, output_neurons=64).to(device)
        fset = tpm.domains.CustomFunctionSet(fspace, tpm.samplers.RandomUniformSampler(tpm.domains.Interval(K, -1, 1), n_points=n_fn),
                                             lambda k, t: k * torch.sin(math.pi * t))
        fsampler = tpm.samplers.FunctionSamplerRandomUniform(n_fn, fset, function_creation_interval=0)
        tsampler = sampler or tpm.samplers.RandomUniformSampler(ival, n_points=n_pts)
        cond = tpm.conditions.PIDeepONetCondition(net, fsampler, tsampler, lambda u, t, f: tpm.utils.laplacian(u, t) + f)
        return net, cond, T
    raise ValueError(cfg)


# ----------------------------------------------------------------------------------------
# reference arm: the real reference on the host cores
# ----------------------------------------------------------------------------------------
def import_reference():
    """the pip-installed reference (baseline/_ref) with the import stubs for matplotlib / pytorch_lightning"""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "torchphysics")):
        return None, f"{ref}/torchphysics missing (python -m pip install --no-deps --target baseline/_ref /root/reference)"
    for p in (os.path.join(ROOT, "oracle", "stubs"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import torchphysics
        return torchphysics, None
    except Exception as e:                                  # pragma: no cover
        return None, f"import torchphysics failed: {e!r}"[:200]


def cpu_reference_rate(cfg, steps, warmup, n_points):
    """(rate, ms/step, cores, kind, what): the reference's own step on the host cores, sampler inside, as it runs it"""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref, why = import_reference()
    if ref is not None:
        torch.manual_seed(0)
        try:
            model, cond, _ = build_config(ref, cfg, n_points, "cpu")
            kind, what = "reference", "boschresearch/torchphysics (baseline/_ref) Condition.forward(device='cpu') + loss.backward()"
        except Exception as e:
            ref, why = None, f"reference config {cfg}: {e!r}"[:200]
    if ref is None:
        if cfg != 1:
            raise RuntimeError(why)
        from oracle import ref_autograd                     # the restatement of the reference call sites (config 1 only)
        torch.manual_seed(0)
        seq = ref_autograd.build_fcn(2, 1, (64,) * 4)
        gen = torch.Generator().manual_seed(1)
        kind, what = "port", "oracle/ref_autograd.poisson_step (reference call sites on the same torch CPU ops); " + why

        class _C:
            def __call__(self, device="cpu"):
                return ref_autograd.poisson_step(seq, torch.rand(n_points, 2, generator=gen))["loss"]
        model, cond = seq, _C()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        loss = cond(device="cpu")
        loss.backward()
        float(loss.detach())
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    return n_points * len(times) / total, 1e3 * total / len(times), cores, kind, what


def gpu_reference_rate(cfg, steps, warmup, n_points, device):
    """BASELINE.md section 3, optional second row: the UNMODIFIED reference (baseline/_ref) with device='cuda' on this B200 --
    stock PyTorch kernels through its own autograd graphs -- for the same step; (rate, ms/step) or raises."""
    import torch
    ref, why = import_reference()
    if ref is None:
        raise RuntimeError(why)
    torch.manual_seed(0)
    model, cond, _ = build_config(ref, cfg, n_points, device)
    model.to(device)
    dev = str(device)
    for _ in range(warmup):
        model.zero_grad(set_to_none=True)
        loss = cond(device=dev)
        loss.backward()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.zero_grad(set_to_none=True)
        loss = cond(device=dev)
        loss.backward()
    e1.record()
    e1.synchronize()
    float(loss.detach())
    ms = e0.elapsed_time(e1) / steps
    return n_points / (ms * 1e-3), ms


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    c = CONFIGS[args.config]
    n = c["cpu_points"]
    steps, warmup = args.steps, args.warmup
    rate, ms, cores, kind, what = cpu_reference_rate(args.config, steps, warmup, n)
    line = {
        "metric": metric_name(args.config), "value": rate, "unit": c["unit"], "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": c["name"], "config_id": args.config, "points_per_step": n,
                   "sample": f"bounded sample: {n} points per step instead of {c['points']} (CPU throughput is flat in N: BASELINE.md section 2)"},
        "cpu_baseline": {"value": rate, "unit": c["unit"], "cores": cores, "kind": kind,
                         "sample": f"{steps} steps of {n} points after {warmup} warm-ups, torch {torch.__version__}, {what}"},
        "e2e": {"value": rate, "unit": c["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------
class Ctx:
    """process-wide state of the native arm"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        for name in ("TPX_ALLOW_TORCH", "TPX_DISABLE_TC", "TPX_DISABLE_TCW", "TPX_TCW_SPLIT", "TPX_TCP", "TPX_SAVE_GB"):
            os.environ.pop(name, None)                      # development switches never apply to a benchmark
        from torchphysics_b200 import _lib
        _lib.lib()                                          # fail loudly if libtpx.so is missing

    def timed(self, fn, steps, warmup, finish=None):
        torch, dist = self.torch, self.dist
        for i in range(warmup):
            fn(i)
        if finish:
            finish()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if finish:
            finish()                                        # e.g. the host read of the last step's loss
        e1.record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps


def points_per_rank(cfg, world, override=0):
    c = CONFIGS[cfg]
    n = override or c["points"]
    if c["scaling"] == "strong":
        n = n // world
    return n


def api_step_rate(ctx, cfg, n, steps, warmup, host=False):
    """the public-API step of configuration `cfg` with n points on this rank: (ms/step, launches/step estimate,
    h2d bytes/step).  host=True: the points come from pinned host memory (HostStreamSampler)."""
    import torchphysics_b200 as tp
    torch = ctx.torch
    torch.manual_seed(0)
    sampler, h2d = None, 0
    if host:
        _, cond0, space = build_config(tp, cfg, n if cfg != 4 else n, ctx.dev)
        s0 = cond0.sampler if cfg != 4 else cond0.trunk_points_sampler
        batches = [s0.sample_points(device=ctx.dev).as_tensor.float().cpu().pin_memory() for _ in range(3)]
        sampler = tp.samplers.HostStreamSampler(batches, space)
        h2d = batches[0].numel() * 4
        del cond0
        torch.manual_seed(0)
    model, cond, _ = build_config(tp, cfg, n, ctx.dev, sampler=sampler)
    reducer = tp.parallel.FlatGradAllReduce([p for p in model.parameters() if p.requires_grad], n_scalars=1)
    # D2H read of every step's loss, pipelined by one step: the copy of step i is enqueued behind its kernels and waited for
    # after step i+1 has been enqueued; `finish` reads the last one inside the timed region
    loss_host = torch.zeros(2).pin_memory()
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"last": -1, "read": []}

    def step(i):
        model.zero_grad(set_to_none=True)
        loss = cond(device=ctx.dev)
        loss.backward()
        (mean_loss,) = reducer([loss])
        loss_host[i % 2:i % 2 + 1].copy_(mean_loss.reshape(1), non_blocking=True)
        loss_done[i % 2].record()
        if state["last"] >= 0:
            loss_done[state["last"] % 2].synchronize()
            state["read"].append(float(loss_host[state["last"] % 2]))
        state["last"] = i

    def finish():
        if state["last"] >= 0:
            loss_done[state["last"] % 2].synchronize()
            state["read"].append(float(loss_host[state["last"] % 2]))
            state["last"] = -1
    ms = ctx.timed(step, steps, warmup, finish)
    loss = state["read"][-1] if state["read"] else float("nan")
    del model, cond
    torch.cuda.empty_cache()
    return ms, h2d, loss


def reverse_kernel_ms(ctx, cfg, n):
    """the reverse sweep of configuration cfg alone on min(n, 2^20) points (CUDA events), or None (DeepONet: several nets)"""
    import torchphysics_b200 as tp
    from torchphysics_b200.engine import JetSpec
    torch = ctx.torch
    if cfg == 4:
        return None, None
    spec = {1: JetSpec((0, 1), (1.0, 1.0)), 2: JetSpec((0, 1, 2), (1.0, 1.0, 0.0)), 3: JetSpec((0, 1), (1.0, 1.0)),
            5: JetSpec((0, 1), None)}[cfg]
    torch.manual_seed(0)
    model, _, _ = build_config(tp, cfg, 1024, ctx.dev)
    eng = model._native_engine()
    m = min(n, 1 << 20) if cfg != 1 else n
    x = torch.rand(m, eng.d_in, device=ctx.dev)
    gu = torch.randn(m, eng.d_out, device=ctx.dev) / m
    gdu = torch.randn(m, eng.d_out, spec.nd, device=ctx.dev) / m
    glap = torch.randn(m, eng.d_out, device=ctx.dev) / m if spec.lap else None
    packed = eng.packed(eng.params())
    nbytes = eng.saved_bytes(spec, m)
    saved = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev) if nbytes else None
    eng.forward_raw(packed, x, spec, saved)
    ms_bwd = ctx.timed(lambda i: eng.backward_raw(packed, x, spec, gu, gdu, glap, saved), 5, 3)
    ms_fwd = ctx.timed(lambda i: eng.forward_raw(packed, x, spec, saved), 5, 3)
    del saved, x, gu, gdu, glap, model
    torch.cuda.empty_cache()
    return ms_bwd * (n / m), ms_fwd * (n / m)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_block(cfg, n, ms_bwd, ms_fwd, peaks):
    c = CONFIGS[cfg]
    if ms_bwd is None:
        return None
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = n * (2.0 / 3.0) * c["flop"] / (ms_bwd * 1e-3) / 1e12       # the reverse sweep is 2 of the 3 forward-equivalents
    # dram bytes per launch from the ncu captures under profiles/ (per Mi points), where one exists for the kernel
    traffic = {1: 4.342e9}.get(cfg)
    kernels = {1: "tc_bwd_kernel (tcgen05 reverse sweep: split-tf32 adjoint product with A from TMEM + split-tf32 gradient product "
                  "from transposed shared-memory operands, saved checkpoints)",
               2: "w_rev_kernel sweep (tcgen05, width 128: one launch per hidden matrix)",
               3: "w_rev_kernel sweep (tcgen05, width 128: one launch per hidden matrix)",
               5: "w_rev_kernel sweep (tcgen05, width 256 as 128 x 128 blocks)"}
    return {"bound": "tensor", "kernel": kernels.get(cfg), "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": None if traffic is None else traffic * n / 2 ** 20,
            "traffic_unit": "bytes/launch",
            "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained",
            "algorithmic_flop_per_point": (2.0 / 3.0) * c["flop"], "ms_per_launch": ms_bwd, "forward_kernel_ms": ms_fwd,
            "note": "FLOP = algorithmic GEMM FLOP of SURVEY 8(d) (reverse = 2 x forward).  The kernels issue 3 split products per "
                    "algorithmic product to hold fp32 accuracy, in tf32 at half the bf16 rate, so the ceiling of frac is 1/6; what bounds "
                    "them is in profiles/README.md"}


def run_config1_resident(ctx, args, N):
    """config 1 with the points resident in HBM, through the C-ABI kernels: the headline `value`"""
    import ctypes
    import torchphysics_b200 as tp
    from torchphysics_b200 import _lib
    from torchphysics_b200.engine import JetSpec
    torch, dist, dev, world = ctx.torch, ctx.dist, ctx.dev, ctx.world
    torch.manual_seed(0)
    model, _, _ = build_config(tp, 1, 1024, dev)
    eng = model._native_engine()
    params = eng.params()
    spec = JetSpec((0, 1), (1.0, 1.0))
    n_theta = sum(p.numel() for p in params)
    # rotating batches so that consecutive steps never reuse L2-resident inputs
    bytes_per_batch = N * 2 * 4
    n_batches = max(2, int(math.ceil(160e6 / bytes_per_batch)) + 1)
    gen = torch.Generator(device=dev).manual_seed(1 + ctx.rank)
    xs = [torch.rand(N, 2, device=dev, generator=gen) for _ in range(n_batches)]
    flat = torch.zeros(n_theta + 1, dtype=torch.float32, device=dev)
    saved_bytes = eng.saved_bytes(spec, N)
    saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev) if saved_bytes else None
    loss_acc = torch.zeros(1, dtype=torch.float64, device=dev)
    glap_buf = torch.empty(N, 1, dtype=torch.float32, device=dev)

    def step(i):
        x = xs[i % n_batches]
        packed = eng.packed(params)
        u, du, lap = eng.forward_raw(packed, x, spec, saved)
        r = lap + source(x)                                 # the data function is evaluated inside the step
        loss_acc.zero_()
        _lib.check(_lib.lib().tpx_squared_error_mean(_lib.ptr(r), ctypes.c_int64(N), ctypes.c_int64(N), _lib.ptr(loss_acc),
                                                     _lib.stream_ptr()))
        _lib.check(_lib.lib().tpx_squared_error_mean_backward(_lib.ptr(r), ctypes.c_int64(N), ctypes.c_int64(N), None,
                                                              _lib.ptr(glap_buf), _lib.stream_ptr()))
        gacc = eng.backward_raw(packed, x, spec, None, None, glap_buf, saved)
        grads = eng.unpack(gacc, params, skip_last_bias=True)
        off = 0
        for g in grads:
            if g is not None:
                flat[off:off + g.numel()].copy_(g.reshape(-1))
                off += g.numel()
        flat[-1] = loss_acc[0]
        if world > 1:
            dist.all_reduce(flat)
    ms = ctx.timed(step, args.steps, args.warmup)
    del saved, xs
    torch.cuda.empty_cache()
    return ms, n_batches, bytes_per_batch, n_theta


def run_native(args):
    ctx = Ctx()
    torch, world, rank = ctx.torch, ctx.world, ctx.rank
    cfg = args.config
    c = CONFIGS[cfg]
    N = points_per_rank(cfg, world, args.points)
    peaks = load_peaks()
    clocks = ClockSampler(ctx.local)
    if rank == 0:
        clocks.start()
    extra = {}
    if cfg == 1:
        ms_step, n_batches, bytes_per_batch, n_theta = run_config1_resident(ctx, args, N)
        launches = 11          # fwd, 5 elementwise kernels of f, residual add, loss, loss adjoint, bwd, dW reduction, unpack (6 of them libtpx's)
        extra = {"l2": f"{n_batches} rotating input batches of {bytes_per_batch/1e6:.0f} MB (> 126 MB L2 in total)", "theta": n_theta}
    else:
        ms_step, _, _ = api_step_rate(ctx, cfg, N, args.steps, args.warmup)
        launches = None
        extra = {"l2": "every step samples a fresh batch (Philox sampler inside the step); inputs + activations exceed the 126 MB L2"}
    clk = clocks.stop() if rank == 0 else None
    value = N * world / (ms_step * 1e-3)

    # ---- end to end through the public API, host buffers ---------------------------------
    ms_e2e, h2d, _ = api_step_rate(ctx, cfg, N, args.steps, max(5, args.warmup), host=True)   # (3 were too few at 4 ranks: pinned buffers, NCCL streams)
    e2e = N * world / (ms_e2e * 1e-3)
    # ---- dominant kernel alone (roofline) ---------------------------------------------------
    ms_bwd, ms_fwd = reverse_kernel_ms(ctx, cfg, N)

    line = None
    if rank == 0:
        line = {
            "metric": metric_name(cfg), "value": value, "unit": c["unit"],
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": c["name"], "config_id": cfg, "points_per_gpu_per_step": N,
                            "global_points_per_step": N * world, "parallelism": f"dp{world}"}, **extra),
            "e2e": {"value": e2e, "unit": c["unit"], "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "path": "Condition.forward() + loss.backward() through the public API, the step's points from pinned host memory "
                            "(tp.samplers.HostStreamSampler: H2D of the next batch overlaps the step); the loss of every step is read "
                            "back to the host one step behind the launch queue, the last one inside the timed region"},
            "gpu_launches": launches,
            "roofline": roofline_block(cfg, N, ms_bwd, ms_fwd, peaks),
            "clocks": clk,
        }
        if line["gpu_launches"] is None:
            # libtpx kernel launches per step, counted in the torch.profiler tables of profiles/r02b_*_step_kernels.txt (layer-major
            # sweeps: forward, reverse and dW-reduction launch per hidden matrix; config 5: 13 micro-batches of 4 blocks per matrix)
            line["gpu_launches"] = {2: 23, 3: 23, 4: 25, 5: 1150}.get(cfg, 11)
    # ---- the other configurations (short runs) ----------------------------------------------
    if cfg == 1 and not args.no_other_configs:
        blocks = {}
        plan = [("1_sampler", 1, N), ("1_10k", 1, 10000)]
        plan += [(str(k), k, points_per_rank(k, world)) for k in ((2, 3, 4, 5) if world == 1 else (3, 5))]
        for key, k, n in plan:
            try:
                short = key in ("1_10k", "4")              # millisecond steps, host bound: many of them, or the number is noise
                ms, _, _ = api_step_rate(ctx, k, n, 3 if k == 5 else (50 if short else 5), 5 if short else 3)
                blk = {"workload": CONFIGS[k]["name"], "points_per_gpu_per_step": n, "global_points_per_step": n * world,
                       "scaling": CONFIGS[k]["scaling"] if key not in ("1_sampler", "1_10k") else "weak",
                       "value": n * world / (ms * 1e-3), "unit": CONFIGS[k]["unit"], "ms_per_step": ms,
                       "what": "Condition.forward() + loss.backward() through the public API, Philox sampler inside the step"
                               + (", one all-reduce of dLoss/dtheta" if world > 1 else "")}
                if world == 1 and key in ("2", "3", "5"):
                    mb, mf = reverse_kernel_ms(ctx, k, n)
                    blk["roofline"] = roofline_block(k, n, mb, mf, peaks)
                if world == 1 and key in ("2", "4"):
                    me, hb, _ = api_step_rate(ctx, k, n, 50 if short else 5, 5 if short else 3, host=True)
                    blk["e2e"] = {"value": n / (me * 1e-3), "unit": CONFIGS[k]["unit"], "ms_per_step": me, "h2d_bytes_per_step": hb,
                                  "d2h_bytes_per_step": 4}
                blocks[key] = blk
            except Exception as e:                           # never lose the headline line over the extras
                blocks[key] = {"error": repr(e)[:300]}
        if world == 1:
            # the literal config 1 (10,000 points) and config 4 (host-bound when run eagerly: ~100 small launches per step) once
            # more with the whole training step (samplers -> jets -> residual -> loss -> reverse -> Adam) captured in ONE CUDA
            # graph (tp.solver.Trainer(cuda_graph=True))
            import torchphysics_b200 as tp
            for key, k, n in (("1_10k_graph", 1, 10000), ("4_graph", 4, points_per_rank(4, 1))):
                try:
                    torch.manual_seed(0)
                    model, cond, _ = build_config(tp, k, n, ctx.dev)
                    solver = tp.solver.Solver([cond], optimizer_setting=tp.solver.OptimizerSetting(torch.optim.Adam, lr=1e-3))
                    tr = tp.solver.Trainer(max_steps=303 if k == 1 else 103, cuda_graph=True, device=ctx.dev)
                    tr.fit(solver)
                    if tr.graphed:
                        blocks[key] = {"workload": CONFIGS[k]["name"] + " + Adam update", "points_per_gpu_per_step": n,
                                       "value": n / (tr.graph_step_ms * 1e-3), "unit": CONFIGS[k]["unit"], "ms_per_step": tr.graph_step_ms,
                                       "what": "one CUDA-graph launch per training step (replays timed with CUDA events): "
                                               "device-resident Philox key, device-resident Adam step count",
                                       "final_loss": float(tr.losses[-1])}
                    else:
                        blocks[key] = {"error": "the step was not captured"}
                except Exception as e:
                    blocks[key] = {"error": repr(e)[:300]}
        if rank == 0:
            line["configs"] = blocks
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                rate, ms, cores, kind, what = cpu_reference_rate(cfg, 8, 2, c["cpu_points"])
                line["cpu_baseline"] = {"value": rate, "unit": c["unit"], "cores": cores, "kind": kind,
                                        "sample": f"8 steps of {c['cpu_points']} points after 2 warm-ups, {what}"}
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)[:300]}
            if cfg == 1:
                # optional second row of BASELINE.md section 3: the unmodified reference on THIS GPU (stock PyTorch kernels)
                try:
                    npts = 1 << 18                      # its autograd graphs of a 2 Mi-point step do not fit comfortably; flat in N
                    rate, ms = gpu_reference_rate(cfg, 5, 3, npts, ctx.dev)
                    line["reference_on_gpu"] = {"value": rate, "unit": c["unit"], "ms_per_step": ms,
                                                "sample": f"5 steps of {npts} points after 3 warm-ups, boschresearch/torchphysics "
                                                          "(baseline/_ref) Condition.forward(device='cuda') + loss.backward() on the same B200"}
                except Exception as e:
                    line["reference_on_gpu"] = {"error": repr(e)[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
