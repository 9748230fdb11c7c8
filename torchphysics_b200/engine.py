"""Host side of the native FCN jet engine: builds the C descriptors, keeps the packed
parameter image fresh, and exposes the forward/reverse kernels as ONE autograd node.

Replaces, for a stack of ``nn.Linear`` + tanh/sin layers, what the reference does with
``nn.Sequential.forward`` (reference ``models/fcn.py:85-87``) followed by 1 + d
``torch.autograd.grad(create_graph=True)`` sweeps (reference ``utils/differentialoperators/
differentialoperators.py:34,40,64``) and the final ``loss.backward()`` through all of them.
"""
import ctypes
import os

import torch

from . import _lib


class JetSpec:
    """Which derivative streams to propagate: ``dcols`` = input columns of the (<= 3)
    directions, ``lap_w`` = weights of the second-derivative sum (None: no Laplacian)."""

    __slots__ = ("dcols", "lap_w", "out_major")

    def __init__(self, dcols=(), lap_w=None, out_major=False):
        # out_major: the streams (and their adjoints) as u (d_out, n), du (nd, d_out, n), lap (d_out, n) -- tpx_jet_desc.out_major
        self.out_major = bool(out_major)
        self.dcols = tuple(int(c) for c in dcols)
        self.lap_w = None if lap_w is None else tuple(float(w) for w in lap_w)
        if len(self.dcols) > _lib.TPX_MAX_ND:
            raise ValueError(f"at most {_lib.TPX_MAX_ND} derivative directions")
        if self.lap_w is not None and len(self.lap_w) != len(self.dcols):
            raise ValueError("lap_w must have one weight per direction")

    @property
    def nd(self):
        return len(self.dcols)

    @property
    def lap(self):
        return self.lap_w is not None

    def covers(self, other):
        """True if a jet computed with self also answers every request of other."""
        if not set(other.dcols) <= set(self.dcols):
            return False
        if other.lap:
            if not self.lap:
                return False
            mine = {c: w for c, w in zip(self.dcols, self.lap_w)}
            theirs = {c: w for c, w in zip(other.dcols, other.lap_w)}
            return all(mine.get(c, 0.0) == theirs.get(c, 0.0) for c in set(mine) | set(theirs))
        return True

    def key(self):
        return (self.dcols, self.lap_w, self.out_major)

    def c_struct(self):
        j = _lib.JetDesc()
        j.nd = self.nd
        j.out_major = 1 if self.out_major else 0
        for k, c in enumerate(self.dcols):
            j.dcol[k] = c
        j.lap = 1 if self.lap else 0
        if self.lap:
            for k, w in enumerate(self.lap_w):
                j.lap_w[k] = w
        return j

    def __repr__(self):
        return f"JetSpec(dcols={self.dcols}, lap_w={self.lap_w}" + (", out_major=True)" if self.out_major else ")")


def _void_p_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class FcnEngine:
    """Native evaluator of one fully connected net (list of nn.Linear + one activation)."""

    def __init__(self, linears, activation, d_out=None):
        """``d_out`` replaces the output width of the last Linear: the caller then always passes its own parameter
        tensors (``jet(..., params=...)``), e.g. a DeepONet's trunk with the branch coefficients folded into the output
        layer (``deeponet.py``)."""
        self.linears = list(linears)
        self._own_params = d_out is None
        if activation not in ("tanh", "sin"):
            raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, f"activation {activation!r}")
        self.activation = activation
        d = _lib.FcnDesc()
        d.d_in = self.linears[0].in_features
        d.d_out = self.linears[-1].out_features if d_out is None else int(d_out)
        d.n_hidden = len(self.linears) - 1
        if d.n_hidden < 1 or d.n_hidden > _lib.TPX_MAX_HIDDEN:
            raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, f"{d.n_hidden} hidden layers")
        for l, lin in enumerate(self.linears[:-1]):
            d.widths[l] = lin.out_features
        d.activation = _lib.TPX_ACT_TANH if activation == "tanh" else _lib.TPX_ACT_SIN
        self.desc = d
        self.d_in, self.d_out = d.d_in, d.d_out
        n = ctypes.c_size_t(0)
        _lib.check(_lib.lib().tpx_fcn_packed_floats(ctypes.byref(d), ctypes.byref(n)))
        self.packed_floats = n.value
        self._packed = None
        self._packed_key = None
        self._ws = {}
        self._ckpt_pool = _CheckpointPool()

    def with_outputs(self, n_cols, device):
        """an engine on the same hidden layers whose output layer has ``n_cols`` columns and takes its parameters from
        the caller; raises TpxError(TPX_E_UNSUPPORTED) when no kernel pair (forward + reverse) takes that shape."""
        eng = FcnEngine(self.linears, self.activation, d_out=n_cols)
        try:                                   # output-major streams where the kernels offer them (9 or more columns)
            eng.workspace(JetSpec(out_major=True), device)
            eng.out_major = True
        except _lib.TpxError as err:
            if err.code != _lib.TPX_E_UNSUPPORTED:
                raise
            eng.workspace(JetSpec(), device)
            eng.out_major = False
        return eng

    # ---- parameters ------------------------------------------------------------------
    def params(self):
        if not self._own_params:
            raise _lib.TpxError(_lib.TPX_E_ARG, "this engine has no output layer of its own: pass params")
        out = []
        for lin in self.linears:
            if lin.bias is None:
                raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "Linear without bias")
            out += [lin.weight, lin.bias]
        return out

    def packed(self, params):
        with _lib.on_device(params[0].device):       # launches go to the stream of the tensors' device, not the current one
            return self._packed_impl(params)

    def _packed_impl(self, params):
        """Packed image of the current parameter values (re-packed when they changed)."""
        # (data_ptr, version) identifies the VALUE of a module parameter (in-place optimizer updates bump the version; writes
        # through ``.data`` do not and are not tracked).  A tensor that is not a leaf -- e.g. the first layer with a
        # NormalizationLayer folded in, models.Sequential -- is a fresh temporary every call: its version is always 0 and the
        # allocator may hand out the same address again, so such images are always re-packed.
        key = tuple((p.data_ptr(), p._version) for p in params)
        fresh = any(p.grad_fn is not None for p in params)
        if params[0].is_cuda and torch.cuda.is_current_stream_capturing():
            fresh = True                  # the captured step must contain the pack launch: the parameters change between replays
        if fresh or self._packed is None or key != self._packed_key or self._packed.device != params[0].device:
            dev = params[0].device
            buf = torch.empty(self.packed_floats, dtype=torch.float32, device=dev)
            srcs = [p.detach().contiguous() for p in params]
            for s in srcs:
                if s.dtype != torch.float32:
                    raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "parameters must be float32")
            arr = _void_p_array(srcs)
            _lib.check(_lib.lib().tpx_fcn_pack(ctypes.byref(self.desc), arr, _lib.ptr(buf), _lib.stream_ptr()))
            self._packed, self._packed_key = buf, key
        return self._packed

    def workspace(self, spec, device):
        k = (spec.nd, spec.lap, spec.out_major, str(device))
        ws = self._ws.get(k)
        if ws is None:
            nbytes = ctypes.c_size_t(0)
            j = spec.c_struct()
            _lib.check(_lib.lib().tpx_fcn_backward_workspace_bytes(ctypes.byref(self.desc), ctypes.byref(j), ctypes.byref(nbytes)))
            ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=device)
            self._ws[k] = ws
        return ws

    # ---- raw kernels -----------------------------------------------------------------
    def saved_bytes(self, spec, n):
        """bytes of activation-jet checkpoints the forward can save for the reverse (0: not available)."""
        nbytes = ctypes.c_size_t(0)
        j = spec.c_struct()
        _lib.check(_lib.lib().tpx_fcn_saved_bytes(ctypes.byref(self.desc), ctypes.byref(j), ctypes.c_int64(n),
                                                  ctypes.byref(nbytes)))
        return nbytes.value

    def forward_raw(self, packed, x, spec, saved=None):
        with _lib.on_device(x.device):
            return self._forward_raw_impl(packed, x, spec, saved)

    def _forward_raw_impl(self, packed, x, spec, saved=None):
        n = x.shape[0]
        if spec.out_major:
            u = torch.empty((self.d_out, n), dtype=torch.float32, device=x.device)
            du = torch.empty((spec.nd, self.d_out, n), dtype=torch.float32, device=x.device) if spec.nd else None
            lap = torch.empty((self.d_out, n), dtype=torch.float32, device=x.device) if spec.lap else None
        else:
            u = torch.empty((n, self.d_out), dtype=torch.float32, device=x.device)
            du = torch.empty((n, self.d_out, spec.nd), dtype=torch.float32, device=x.device) if spec.nd else None
            lap = torch.empty((n, self.d_out), dtype=torch.float32, device=x.device) if spec.lap else None
        j = spec.c_struct()
        if saved is None:
            try:                                    # widths 65..256: activation arrays of the tensor-core kernels
                ws = self.workspace(spec, x.device)
            except _lib.TpxError:                   # a jet without a reverse kernel: plain forward
                ws = None
            if ws is None:
                _lib.check(_lib.lib().tpx_fcn_jet_forward(
                    ctypes.byref(self.desc), ctypes.byref(j), _lib.ptr(packed), _lib.ptr(x), ctypes.c_int64(n),
                    _lib.ptr(u), _lib.ptr(du), _lib.ptr(lap), _lib.stream_ptr()))
            else:
                _lib.check(_lib.lib().tpx_fcn_jet_forward_ws(
                    ctypes.byref(self.desc), ctypes.byref(j), _lib.ptr(packed), _lib.ptr(x), ctypes.c_int64(n),
                    _lib.ptr(u), _lib.ptr(du), _lib.ptr(lap), _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_ptr()))
        else:
            _lib.check(_lib.lib().tpx_fcn_jet_forward_save(
                ctypes.byref(self.desc), ctypes.byref(j), _lib.ptr(packed), _lib.ptr(x), ctypes.c_int64(n),
                _lib.ptr(u), _lib.ptr(du), _lib.ptr(lap), _lib.ptr(saved), ctypes.c_size_t(saved.numel()),
                _lib.stream_ptr()))
        return u, du, lap

    def backward_raw(self, packed, x, spec, gu, gdu, glap, saved=None):
        with _lib.on_device(x.device):
            return self._backward_raw_impl(packed, x, spec, gu, gdu, glap, saved)

    def _backward_raw_impl(self, packed, x, spec, gu, gdu, glap, saved=None):
        """returns the float64 packed gradient accumulator."""
        n = x.shape[0]
        gacc = torch.zeros(self.packed_floats, dtype=torch.float64, device=x.device)
        j = spec.c_struct()
        if saved is not None:
            _lib.check(_lib.lib().tpx_fcn_jet_backward_saved(
                ctypes.byref(self.desc), ctypes.byref(j), _lib.ptr(packed), _lib.ptr(x), ctypes.c_int64(n),
                _lib.ptr(gu), _lib.ptr(gdu), _lib.ptr(glap), _lib.ptr(gacc), _lib.ptr(saved),
                ctypes.c_size_t(saved.numel()), _lib.stream_ptr()))
            return gacc
        ws = self.workspace(spec, x.device)
        _lib.check(_lib.lib().tpx_fcn_jet_backward(
            ctypes.byref(self.desc), ctypes.byref(j), _lib.ptr(packed), _lib.ptr(x), ctypes.c_int64(n),
            _lib.ptr(gu), _lib.ptr(gdu), _lib.ptr(glap), _lib.ptr(gacc), _lib.ptr(ws),
            ctypes.c_size_t(ws.numel()), _lib.stream_ptr()))
        return gacc

    def unpack(self, gacc, params, skip_last_bias=False):
        with _lib.on_device(gacc.device):
            return self._unpack_impl(gacc, params, skip_last_bias)

    def _unpack_impl(self, gacc, params, skip_last_bias=False):
        grads = [torch.empty_like(p, dtype=torch.float32) for p in params]
        if skip_last_bias:
            grads[-1] = None
        arr = _void_p_array(grads)
        _lib.check(_lib.lib().tpx_fcn_unpack_grads(ctypes.byref(self.desc), _lib.ptr(gacc), arr, 0, _lib.stream_ptr()))
        return grads

    # ---- autograd entry point --------------------------------------------------------
    def jet(self, x, spec, params=None):
        """(u, du, lap) as differentiable functions of the network parameters.  ``params`` replaces the
        modules' own tensors, e.g. a first layer with a NormalizationLayer folded in (``models.Sequential``)."""
        if not x.is_cuda:
            raise _lib.TpxError(_lib.TPX_E_ARG, "the native engine only runs on CUDA tensors")
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        if params is None:
            params = self.params()
        return _FcnJetFn.apply(self, x, spec, *params)


class _CheckpointPool:
    """Activation-jet checkpoint buffers, owned by the engine and recycled across steps.

    The buffers are large (8 KB / point: 16 GB for 2 Mi points) and live from a forward launch to
    the matching reverse launch.  Taking them from torch's caching allocator every step lets small
    allocations split the freed block, so the next request misses and the allocator falls back to
    cudaMalloc / cudaFree -- measured as a step time creeping from 13 to 23 ms.  Here a buffer goes
    back to the pool when its lease dies (after the reverse launch, or when the graph is dropped)."""

    def __init__(self):
        self.free = []

    def acquire(self, nbytes, device):
        best = None
        for i, b in enumerate(self.free):
            if b.device == device and nbytes <= b.numel() <= 2 * nbytes and (best is None or b.numel() < self.free[best].numel()):
                best = i
        buf = self.free.pop(best) if best is not None else torch.empty(nbytes, dtype=torch.uint8, device=device)
        if best is None:                      # drop buffers the new size has made useless
            self.free = [b for b in self.free if b.device != device or b.numel() >= nbytes]
        return _CheckpointLease(self, buf)


class _CheckpointLease:
    __slots__ = ("pool", "buf")

    def __init__(self, pool, buf):
        self.pool, self.buf = pool, buf

    def __del__(self):
        pool, buf = self.pool, self.buf
        if pool is not None and buf is not None and len(pool.free) < 4:
            pool.free.append(buf)


def save_budget_bytes(device):
    """how much HBM one evaluation may spend on activation-jet checkpoints (TPX_SAVE_GB overrides; 0 disables)."""
    import os
    env = os.environ.get("TPX_SAVE_GB")
    if env is not None:
        return int(float(env) * 2 ** 30)
    free, _total = torch.cuda.mem_get_info(device)
    return int(0.6 * free)


_ALWAYS_SAVED = 1 << 30


def within_save_budget(nbytes, device):
    """True if ``nbytes`` of checkpoints may be saved.  Up to 1 GiB the answer is yes without asking the driver:
    cudaMemGetInfo costs ~0.3 ms, more than a whole 10,000-point step."""
    return nbytes <= _ALWAYS_SAVED and "TPX_SAVE_GB" not in os.environ or nbytes <= save_budget_bytes(device)


class _FcnJetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, spec, *params):
        packed = engine.packed(params)
        saved = None
        if any(ctx.needs_input_grad[3:]) and x.shape[0] > 0:     # a reverse sweep will follow
            nbytes = engine.saved_bytes(spec, x.shape[0])
            if nbytes > 0:
                pooled = any(b.device == x.device and nbytes <= b.numel() <= 2 * nbytes for b in engine._ckpt_pool.free)
                if pooled or within_save_budget(nbytes, x.device):
                    saved = engine._ckpt_pool.acquire(nbytes, x.device)
        u, du, lap = engine.forward_raw(packed, x, spec, None if saved is None else saved.buf)
        ctx.engine, ctx.spec, ctx.x, ctx.packed, ctx.saved = engine, spec, x, packed, saved
        ctx.params = params
        ctx.set_materialize_grads(False)
        outs = [u]
        if spec.out_major:
            outs.append(du if du is not None else x.new_zeros((0, engine.d_out, x.shape[0])))
            outs.append(lap if lap is not None else x.new_zeros((0, x.shape[0])))
        else:
            outs.append(du if du is not None else x.new_zeros((x.shape[0], engine.d_out, 0)))
            outs.append(lap if lap is not None else x.new_zeros((x.shape[0], 0)))
        ctx.mark_non_differentiable(*[o for o in outs[1:] if o.numel() == 0 and x.shape[0] > 0])
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable      # raw-pointer adjoint: a double backward raises instead of returning zeros
    def backward(ctx, gu, gdu, glap):
        engine, spec = ctx.engine, ctx.spec
        needs = ctx.needs_input_grad[3:]
        if not any(needs):
            return (None, None, None) + (None,) * len(needs)

        def prep(g):
            if g is None:
                return None
            g = g.contiguous()
            return g if g.dtype == torch.float32 else g.float()

        gu, gdu, glap = prep(gu), prep(gdu), prep(glap)
        if spec.nd == 0:
            gdu = None
        if not spec.lap:
            glap = None
        if gu is None and gdu is None and glap is None:
            return (None, None, None) + (None,) * len(needs)
        gacc = engine.backward_raw(ctx.packed, ctx.x, spec, gu, gdu, glap, None if ctx.saved is None else ctx.saved.buf)
        ctx.saved = None                      # the lease returns the buffer to the engine's pool
        # reference parity: a parameter the loss does not depend on keeps grad None;
        # that is exactly the output bias when the value stream carries no adjoint.
        grads = engine.unpack(gacc, ctx.params, skip_last_bias=(gu is None))
        grads = [g if need else None for g, need in zip(grads, needs)]
        return (None, None, None) + tuple(grads)
