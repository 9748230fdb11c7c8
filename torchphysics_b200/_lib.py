"""ctypes binding of libtpx.so (C ABI declared in include/tpx.h).

The library is the product; there is no Python or CPU substitute for it.  ``lib()``
raises ``TpxLibraryError`` when the shared object is missing or does not export every
symbol of the header, and every wrapper raises ``TpxError`` on a non-zero return code.
"""
import ctypes
import os

TPX_MAX_HIDDEN = 16
TPX_MAX_DIN = 8
TPX_MAX_ND = 3
TPX_MAX_WIDTH = 256
TPX_MAX_DOUT = 4096
TPX_ACT_TANH = 0
TPX_ACT_SIN = 1

TPX_E_ARG = -1
TPX_E_UNSUPPORTED = -2
TPX_E_CUDA = -3
TPX_E_WORKSPACE = -4

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtpx.so")

EXPORTS = (
    "tpx_version",
    "tpx_last_error",
    "tpx_fcn_packed_floats",
    "tpx_fcn_pack",
    "tpx_fcn_jet_forward",
    "tpx_fcn_jet_forward_ws",
    "tpx_fcn_backward_workspace_bytes",
    "tpx_fcn_jet_backward",
    "tpx_fcn_unpack_grads",
    "tpx_fcn_saved_bytes",
    "tpx_fcn_jet_forward_save",
    "tpx_fcn_jet_backward_saved",
    "tpx_region_contains",
    "tpx_sample_uniform",
    "tpx_sample_reject_workspace_bytes",
    "tpx_sample_uniform_reject",
    "tpx_key_advance",
    "tpx_sample_uniform_dkey",
    "tpx_sample_grid",
    "tpx_boundary_normal",
    "tpx_sample_keys",
    "tpx_sample_lhs",
    "tpx_union_pick",
    "tpx_adam_step",
    "tpx_adam_step_dstep",
    "tpx_squared_error_mean",
    "tpx_squared_error_mean_backward",
)

TPX_SHAPE_INTERVAL = 1
TPX_SHAPE_PARALLELOGRAM = 2
TPX_SHAPE_CIRCLE = 3
TPX_SHAPE_INTERVAL_BOUNDARY = 4
TPX_SHAPE_INTERVAL_POINT = 5
TPX_SHAPE_PARALLELOGRAM_BOUNDARY = 6
TPX_SHAPE_CIRCLE_BOUNDARY = 7
TPX_MAX_TENSORS = 64
TPX_MAX_SHAPES = 8
TPX_MAX_OPS = 24
TPX_OP_SHAPE, TPX_OP_AND, TPX_OP_OR, TPX_OP_NOT = 0, 1, 2, 3


class TpxLibraryError(RuntimeError):
    pass


class TpxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libtpx error {code}: {message}")
        self.code = code


class FcnDesc(ctypes.Structure):
    _fields_ = [
        ("d_in", ctypes.c_int32),
        ("d_out", ctypes.c_int32),
        ("n_hidden", ctypes.c_int32),
        ("widths", ctypes.c_int32 * TPX_MAX_HIDDEN),
        ("activation", ctypes.c_int32),
    ]


class JetDesc(ctypes.Structure):
    _fields_ = [
        ("nd", ctypes.c_int32),
        ("dcol", ctypes.c_int32 * TPX_MAX_ND),
        ("lap", ctypes.c_int32),
        ("lap_w", ctypes.c_float * TPX_MAX_ND),
        ("out_major", ctypes.c_int32),
    ]


class Shape(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("col", ctypes.c_int32), ("p", ctypes.c_float * 6)]


class Region(ctypes.Structure):
    _fields_ = [
        ("n_shapes", ctypes.c_int32),
        ("n_ops", ctypes.c_int32),
        ("shapes", Shape * TPX_MAX_SHAPES),
        ("ops", ctypes.c_int32 * TPX_MAX_OPS),
    ]


class TensorList(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("param", ctypes.c_void_p * TPX_MAX_TENSORS),
        ("grad", ctypes.c_void_p * TPX_MAX_TENSORS),
        ("exp_avg", ctypes.c_void_p * TPX_MAX_TENSORS),
        ("exp_avg_sq", ctypes.c_void_p * TPX_MAX_TENSORS),
        ("numel", ctypes.c_int64 * TPX_MAX_TENSORS),
    ]


_lib = None


def lib():
    """Load libtpx.so once; fail loudly if it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TpxLibraryError(
            f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C torchphysics_b200/csrc`). There is no fallback path."
        )
    handle = ctypes.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(handle, s)]
    if missing:
        raise TpxLibraryError(f"{LIB_PATH} does not export {missing}; rebuild it")
    handle.tpx_version.restype = ctypes.c_int
    handle.tpx_last_error.restype = ctypes.c_char_p
    for name in EXPORTS[2:]:
        getattr(handle, name).restype = ctypes.c_int
    _lib = handle
    return handle


def check(rc):
    if rc != 0:
        raise TpxError(rc, lib().tpx_last_error().decode("utf-8", "replace"))


def ptr(t):
    """device pointer of a torch tensor (or NULL for None)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    """the current CUDA stream of the current device as a cudaStream_t"""
    import torch

    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)       # no Stream object per launch (10 us each on the hot path)
    if raw is not None:
        return ctypes.c_void_p(raw(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """context manager: launches inside go to ``device`` (a torch.device of type cuda).  The guard is skipped when that is
    the current device already -- the common case, and torch.cuda.device() costs ~15 us per use."""
    import torch

    if not isinstance(device, torch.device):
        device = torch.device(device)
    idx = device.index
    if device.type != "cuda" or idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)
