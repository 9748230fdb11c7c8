#!/usr/bin/env python
"""Benchmark of the PINN residual hot path (BASELINE.json metric): collocation-point
residual+grad evaluations per second, 2-D Poisson -Laplace(u)=f, FCN 4x64 tanh.

One "step" = for one batch of N collocation points per GPU: forward jet (u, grad u,
Laplace u), residual r = Laplace u + f, loss = mean(r^2), reverse sweep dLoss/dtheta,
and (N GPUs) ONE all-reduce of the flat gradient.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl native|reference]

`value`  : points/s with the points already resident in HBM (C-ABI kernels + residual).
`e2e`    : the same step through the public API (tp.conditions.PINNCondition.forward +
           backward) with the points in pinned HOST memory: H2D of x every step, D2H of the
           loss every step.
`roofline`: dominant kernel (reverse sweep) timed alone with CUDA events.
`cpu_baseline` / `--impl reference`: the reference's torch.autograd path (oracle/ref_autograd,
           a restatement of the reference call sites on the same torch CPU kernels) on the
           host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN = (64, 64, 64, 64)
D_IN, D_OUT, S_STREAMS = 2, 1, 4
F_FWD = 2 * (D_IN * 64 + S_STREAMS * 3 * 64 * 64 + S_STREAMS * 64 * D_OUT)   # 99,072
F_ALG = 3 * F_FWD                                                             # 297,216 FLOP / point
METRIC = "collocation-point residual+grad evals/sec (PINN Laplacian, FCN 4x64)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--points", type=int, default=1 << 21, help="collocation points per GPU per step")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# reference arm: torch CPU autograd restatement of the reference path (oracle/)
# ----------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, n_points):
    import torch
    from oracle import ref_autograd
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(D_IN, D_OUT, HIDDEN)
    gen = torch.Generator().manual_seed(1)
    times = []
    for i in range(warmup + steps):
        x = torch.rand(n_points, D_IN, generator=gen)
        t0 = time.perf_counter()
        out = ref_autograd.poisson_step(seq, x)
        float(out["loss"].detach())
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    return n_points * len(times) / total, 1e3 * total / len(times), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 20000
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 3))
    rate, ms, cores = cpu_reference_rate(steps, warmup, n)
    line = {
        "metric": METRIC, "value": rate, "unit": "points/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": "2D Poisson -Lap u = f on the unit square, FCN 4x64 tanh, residual lap(u)+f, "
                               "loss mean(r^2), dLoss/dtheta", "points_per_step": n},
        "cpu_baseline": {"value": rate, "unit": "points/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of {n} points, torch {__import__('torch').__version__} CPU autograd "
                                   "(oracle/ref_autograd.poisson_step: reference call sites on the same torch ops)"},
        "e2e": {"value": rate, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
def run_native(args):
    import torch
    import torch.distributed as dist
    import torchphysics_b200 as tp
    from torchphysics_b200 import _lib
    from torchphysics_b200.engine import JetSpec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()                                            # fail loudly if libtpx.so is missing

    N = args.points
    torch.manual_seed(0)
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    model = tp.models.FCN(X, U, hidden=HIDDEN).to(dev)
    eng = model._native_engine()
    params = eng.params()
    spec = JetSpec((0, 1), (1.0, 1.0))
    n_theta = sum(p.numel() for p in params)

    import math
    def source(x):
        return 2 * math.pi ** 2 * torch.sin(math.pi * x[:, :1]) * torch.sin(math.pi * x[:, 1:2])

    # rotating batches so that consecutive steps never reuse L2-resident inputs
    bytes_per_batch = N * (D_IN + 1) * 4
    n_batches = max(2, int(math.ceil(160e6 / bytes_per_batch)) + 1)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    xs = [torch.rand(N, D_IN, device=dev, generator=gen) for _ in range(n_batches)]
    fs = [source(x) for x in xs]
    flat = torch.zeros(n_theta + 1, dtype=torch.float32, device=dev)
    launches = {"n": 0}
    # activation-jet checkpoints: the forward kernel saves them (4 KB / point), the reverse kernel reads them
    saved_bytes = eng.saved_bytes(spec, N)
    saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev) if saved_bytes else None

    import ctypes
    from torchphysics_b200 import _lib
    loss_acc = torch.zeros(1, dtype=torch.float64, device=dev)
    glap_buf = torch.empty(N, 1, dtype=torch.float32, device=dev)

    def step_resident(i):
        """forward jet + residual + loss + reverse + flat gradient (+ all-reduce)."""
        x, f = xs[i % n_batches], fs[i % n_batches]
        packed = eng.packed(params)
        u, du, lap = eng.forward_raw(packed, x, spec, saved)
        r = lap + f
        # loss = mean(r^2) and dLoss/dr = 2 r / N: the native reduction kernels of csrc/loss.cu (warp shuffles, float64 atomics)
        loss_acc.zero_()
        _lib.check(_lib.lib().tpx_squared_error_mean(_lib.ptr(r), ctypes.c_int64(N), ctypes.c_int64(N), _lib.ptr(loss_acc),
                                                     _lib.stream_ptr()))
        _lib.check(_lib.lib().tpx_squared_error_mean_backward(_lib.ptr(r), ctypes.c_int64(N), ctypes.c_int64(N), None,
                                                              _lib.ptr(glap_buf), _lib.stream_ptr()))
        loss, glap = loss_acc[0], glap_buf
        gacc = eng.backward_raw(packed, x, spec, None, None, glap, saved)
        grads = eng.unpack(gacc, params, skip_last_bias=True)
        off = 0
        for g in grads:
            if g is not None:
                flat[off:off + g.numel()].copy_(g.reshape(-1))
                off += g.numel()
        flat[-1] = loss
        if world > 1:
            dist.all_reduce(flat)
        launches["n"] += 5                                # fwd kernel, loss + loss-adjoint kernels, bwd kernel, unpack kernel
        return loss

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches["n"] = 0
    ms_step = timed(step_resident, args.steps, args.warmup)
    n_launch = launches["n"] * args.steps // (args.steps + args.warmup)
    clk = clocks.stop() if rank == 0 else None
    value = N * world / (ms_step * 1e-3)

    # ---- end to end through the public API, host buffers ---------------------------------
    # the step's points come from pinned host memory: tp.samplers.HostStreamSampler copies batch i+1 to the
    # device on a side stream while step i runs (both inside the timed region: one 8 N byte H2D per step)
    host_x = [x.cpu().pin_memory() for x in xs[:3]]
    host_batches = tp.samplers.HostStreamSampler(host_x, X)
    cond = tp.conditions.PINNCondition(model, host_batches, lambda u, x, f: tp.utils.laplacian(u, x) + f,
                                       data_functions={"f": source})
    reducer = tp.parallel.FlatGradAllReduce(list(model.parameters()), n_scalars=1)
    # D2H read of every step's loss, pipelined by one step: the copy of step i is enqueued behind its kernels and
    # waited for after step i+1 has been enqueued, so the host never drains the GPU queue (the last read completes
    # before the timed region's closing synchronize)
    loss_host = torch.zeros(2).pin_memory()
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]
    losses_read = []

    def step_e2e(i):
        model.zero_grad(set_to_none=True)
        loss = cond(device=dev)
        loss.backward()
        (mean_loss,) = reducer([loss])
        loss_host[i % 2:i % 2 + 1].copy_(mean_loss.reshape(1), non_blocking=True)
        loss_done[i % 2].record()
        if i > 0:
            loss_done[(i - 1) % 2].synchronize()
            losses_read.append(float(loss_host[(i - 1) % 2]))
        return loss_host

    ms_e2e = timed(step_e2e, args.steps, max(3, args.warmup))
    e2e = N * world / (ms_e2e * 1e-3)

    # ---- dominant kernel alone (roofline) ---------------------------------------------------
    packed = eng.packed(params)
    glap = torch.randn(N, 1, device=dev) / N
    def bwd_only(i):
        eng.backward_raw(packed, xs[i % n_batches], spec, None, None, glap, saved)
    def fwd_only(i):
        eng.forward_raw(packed, xs[i % n_batches], spec, saved)
    fwd_only(0)
    ms_bwd = timed(bwd_only, 10, 3)
    ms_fwd = timed(fwd_only, 10, 3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        achieved = N * 2 * F_FWD / (ms_bwd * 1e-3) / 1e12
        kernel_name = "tc_bwd_kernel (tcgen05 reverse sweep: adjoint + dW products, saved checkpoints)" if saved is not None \
            else "fcn_bwd_kernel (reverse sweep incl. forward recompute)"
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "2D Poisson -Lap u = f on the unit square, FCN 4x64 tanh, residual lap(u)+f, "
                                   "loss mean(r^2), dLoss/dtheta", "points_per_gpu_per_step": N,
                       "global_points_per_step": N * world, "parallelism": f"dp{world}",
                       "l2": f"{n_batches} rotating input batches of {bytes_per_batch/1e6:.0f} MB (> 126 MB L2 in total)",
                       "theta": n_theta},
            "e2e": {"value": e2e, "unit": "points/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": N * D_IN * 4, "d2h_bytes_per_step": 4,
                    "path": "tp.conditions.PINNCondition.forward + loss.backward(), x from pinned host memory "
                            "(tp.samplers.HostStreamSampler: H2D of the next batch overlaps the step); the loss of every "
                            "step is read back to the host, one step behind the launch queue"},
            "gpu_launches": n_launch,
            "roofline": {"bound": "tensor", "kernel": kernel_name,
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of profiles/r01j_tc_raw.csv (4.342 GB for
                         # 2^20 points: the 4 KB/point of saved activation jets + points + adjoints + the periodic
                         # flush of the dW accumulators), scaled to this N
                         "traffic": 4.342e9 * N / 2 ** 20, "traffic_unit": "bytes/launch",
                         "peak_source": peak_src,
                         "algorithmic_flop_per_point": 2 * F_FWD, "ms_per_launch": ms_bwd,
                         "forward_kernel_ms": ms_fwd,
                         "tf32_issue_factor": 3,
                         "note": "FLOP = algorithmic GEMM FLOP of SURVEY 8(d) (reverse = 2 x forward); the kernel "
                                 "issues 3 tf32 products per algorithmic product (4 for dW; fp32-accurate split) and tf32 "
                                 "runs at half the bf16 rate (1.10 PFLOP/s measured with scripts/umma_probe2.cu), so the "
                                 "ceiling of frac is ~0.13.  ncu (profiles/r01e, r01j): tensor pipe active 25 %, shared-memory "
                                 "LSU wavefronts 46 % of peak plus the UMMA operand reads of the gradient product "
                                 "(~180-240 KB/tile-layer; 496 KB per tile-layer in total = 3970 port cycles vs 4980 measured): the kernel is bound by shared-memory bandwidth (operand "
                                 "transposition for dW + stream exchange), HBM 1.6 TB/s of 6.5"},
            "clocks": clk,
        }
        if not args.no_cpu_baseline and world == 1:
            n_cpu = 20000
            rate, ms, cores = cpu_reference_rate(8, 2, n_cpu)
            line["cpu_baseline"] = {"value": rate, "unit": "points/s", "cores": cores, "kind": "port",
                                    "sample": f"8 steps of {n_cpu} points, torch CPU autograd restatement of the "
                                              "reference call sites (oracle/ref_autograd.poisson_step)"}
        if world == 1 and not args.no_other_configs:
            # informational: kernel-only forward + reverse rate of the parity configurations 2 / 3 / 5 of BASELINE.json
            # (hidden width 128 / 256: wide tensor-core kernels, csrc/tcw_kernels.cu), 2^18 points, inputs resident
            try:
                line["other_configs"] = other_configs_rate(dev)
            except Exception as e:                       # never lose the headline line over the extras
                line["other_configs"] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def other_configs_rate(dev, n=1 << 18, reps=3):
    import torch
    import torchphysics_b200 as tp
    from torchphysics_b200.engine import JetSpec
    X, T, U, V, Pr = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("u"), tp.spaces.R2("v"), tp.spaces.R1("p")
    out = []
    for name, inp, outp, hidden, spec in (
            ("config 2: heat, FCN 6x128, S=5", X * T, U, (128,) * 6, JetSpec((0, 1, 2), (1.0, 1.0, 0.0))),
            ("config 3: Navier-Stokes, FCN 6x128, 3 outputs, S=4", X, V * Pr, (128,) * 6, JetSpec((0, 1), (1.0, 1.0))),
            ("config 5: Deep Ritz, FCN 8x256, S=3", X, U, (256,) * 8, JetSpec((0, 1), None))):
        torch.manual_seed(0)
        model = tp.models.FCN(inp, outp, hidden=hidden).to(dev)
        eng = model._native_engine()
        d_in, d_out = eng.d_in, eng.d_out
        x = torch.rand(n, d_in, device=dev)
        gu = torch.randn(n, d_out, device=dev)
        gdu = torch.randn(n, d_out, spec.nd, device=dev)
        glap = torch.randn(n, d_out, device=dev) if spec.lap else None
        packed = eng.packed(eng.params())
        saved = torch.empty(eng.saved_bytes(spec, n), dtype=torch.uint8, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for r in range(reps + 2):
            if r == 2:
                ev[0].record()
            eng.forward_raw(packed, x, spec, saved)
            eng.backward_raw(packed, x, spec, gu, gdu, glap, saved)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        out.append({"config": name, "value": n / ms * 1e3, "unit": "points/s", "ms_per_step": ms, "points": n,
                    "what": "forward (saving checkpoints) + reverse sweep, kernels only"})
        del saved, x, gu, gdu, glap
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
