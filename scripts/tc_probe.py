"""Development aid: times the tensor-core forward / reverse kernels for config 1 (FCN 4x64, S=4);
`python scripts/tc_probe.py N dbgb` prints the clock64 timeline of one tile of the reverse kernel."""
import sys

import torch
import torch.nn as nn

sys.path.insert(0, ".")
from oracle import ref_autograd
from torchphysics_b200.engine import FcnEngine, JetSpec


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    what = sys.argv[2] if len(sys.argv) > 2 else "fb"
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(2, 1, (64,) * 4).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    spec = JetSpec((0, 1), (1.0, 1.0))
    x = torch.rand(N, 2, device="cuda")
    packed = eng.packed(eng.params())
    glap = torch.randn(N, 1, device="cuda")
    if "f" in what:
        tf = timeit(lambda: eng.forward_raw(packed, x, spec))
        print(f"fwd N={N}: {tf:.3f} ms  {N / tf * 1e3:.3e} pts/s", flush=True)
    if "b" in what:
        tb = timeit(lambda: eng.backward_raw(packed, x, spec, None, None, glap))
        print(f"bwd (recompute) N={N}: {tb:.3f} ms  {N / tb * 1e3:.3e} pts/s", flush=True)
    if "s" in what:
        nbytes = eng.saved_bytes(spec, N)
        saved = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        tf = timeit(lambda: eng.forward_raw(packed, x, spec, saved))
        tb = timeit(lambda: eng.backward_raw(packed, x, spec, None, None, glap, saved))
        print(f"saved mode N={N} ({nbytes / 2**30:.2f} GiB): fwd {tf:.3f} ms  bwd {tb:.3f} ms  "
              f"fwd+bwd {N / (tf + tb) * 1e3:.3e} pts/s", flush=True)


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[2].startswith("dbg")):
    main()


def debug_bwd(N=1 << 18):
    import ctypes
    from torchphysics_b200 import _lib
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(2, 1, (64,) * 4).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    spec = JetSpec((0, 1), (1.0, 1.0))
    x = torch.rand(N, 2, device="cuda")
    packed = eng.packed(eng.params())
    glap = torch.randn(N, 1, device="cuda")
    saved = torch.empty(eng.saved_bytes(spec, N), dtype=torch.uint8, device="cuda")
    eng.forward_raw(packed, x, spec, saved)
    buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
    lib = _lib.lib()
    lib.tpx_debug_set_timestamps(ctypes.c_void_p(buf.data_ptr()))
    eng.backward_raw(packed, x, spec, None, None, glap, saved)
    torch.cuda.synchronize()
    lib.tpx_debug_set_timestamps(ctypes.c_void_p(0))
    t = buf.cpu().numpy()
    t0 = t[t > 0].min()
    names = ["enter", "d_ready", "in_bar", "computed", "f_free", "stored", "out_bar", "arrived"]
    for w, off in ((0, 0), (7, 256)):
        for l in (3, 2, 1, 0):
            row = t[off + l * 16: off + l * 16 + 8]
            print(f"warp{w} l{l}: " + " ".join(f"{n}={int(v - t0)}" for n, v in zip(names, row) if v > 0))
    for l in (3, 2, 1):
        row = t[512 + l * 4: 512 + l * 4 + 4]
        print(f"mma l{l}: " + " ".join(f"{n}={int(v - t0)}" for n, v in zip(["a_w_ready", "adjoint_done", "dw_issued", "a_ready"], row) if v > 0))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "dbgb":
    debug_bwd()


def debug_fwd(N=1 << 18):
    import ctypes
    from torchphysics_b200 import _lib
    torch.manual_seed(0)
    seq = ref_autograd.build_fcn(2, 1, (64,) * 4).cuda()
    eng = FcnEngine([m for m in seq if isinstance(m, nn.Linear)], "tanh")
    spec = JetSpec((0, 1), (1.0, 1.0))
    x = torch.rand(N, 2, device="cuda")
    packed = eng.packed(eng.params())
    buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
    lib = _lib.lib()
    eng.forward_raw(packed, x, spec)
    lib.tpx_debug_set_timestamps(ctypes.c_void_p(buf.data_ptr()))
    eng.forward_raw(packed, x, spec)
    torch.cuda.synchronize()
    lib.tpx_debug_set_timestamps(ctypes.c_void_p(0))
    t = buf.cpu().numpy()
    t0 = t[t > 0].min()
    names = ["enter", "d_ready", "in_bar", "computed", "out_bar", "arrived"]
    for w, off in ((0, 0), (11, 256)):
        for l in range(4):
            row = t[off + l * 8: off + l * 8 + 6]
            print(f"warp{w} l{l}: " + " ".join(f"{n}={int(v - t0)}" for n, v in zip(names, row) if v > 0))
    for l in range(1, 4):
        for s in range(2):
            row = t[512 + (l * 2 + s) * 2: 512 + (l * 2 + s) * 2 + 2]
            print(f"mma l{l} slot{s}: a_ready={int(row[0] - t0)} issued={int(row[1] - t0)}")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "dbgf":
    debug_fwd()
