"""Native parameter update: ``Adam`` with torch.optim.Adam's constructor and state layout
(``exp_avg`` / ``exp_avg_sq`` / ``step`` per parameter, so ``state_dict``s interchange), whose ``step()`` is ONE
libtpx launch for all parameter tensors (``csrc/optim.cu``) instead of 4-5 elementwise launches per tensor.
The reference configures ``torch.optim.Adam`` through ``OptimizerSetting`` (reference ``solver.py:111-136``);
``solver.Trainer`` substitutes this class for it on CUDA when the requested options are covered."""
import ctypes

import torch

from . import _lib


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.5 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def use_device_step(self, enable=True):
        """Keep the step count of every parameter group in GPU memory (``tpx_adam_step_dstep``): ``step()`` then has no
        host-side state that changes from call to call and can be captured in a CUDA graph.  Switching it off writes the
        device counts back into ``state[p]['step']``."""
        if enable:
            self._dev_steps = {}
            for gi, group in enumerate(self.param_groups):          # created now: a capture cannot copy from the host
                ps = [p for p in group["params"] if p.is_cuda]
                if ps:
                    n = max([int(self.state[p]["step"]) for p in ps if p in self.state and len(self.state[p])] or [0])
                    self._dev_steps[gi] = torch.tensor([n], dtype=torch.int32, device=ps[0].device)
        else:
            for gi, t in getattr(self, "_dev_steps", {}).items():
                n = int(t.item())
                for p in self.param_groups[gi]["params"]:
                    if p in self.state and len(self.state[p]):
                        self.state[p]["step"] = n
            self._dev_steps = None

    @staticmethod
    def covers(optimizer_args):
        """True if torch.optim.Adam(**optimizer_args) is exactly what this class computes."""
        allowed = {"betas", "eps", "weight_decay", "amsgrad", "maximize", "foreach", "fused", "capturable", "differentiable"}
        if not set(optimizer_args) <= allowed:
            return False
        if any(optimizer_args.get(k) for k in ("amsgrad", "maximize", "capturable", "differentiable")):
            return False
        return optimizer_args.get("betas", (0.9, 0.999))[0] >= 0.5

    def _step_on_device(self, gi, group, lr, dev_steps):
        todo = []
        for p in group["params"]:
            if p.grad is None:
                continue
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "native Adam needs contiguous float32 CUDA parameters")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            todo.append((p, st))
        if not todo:
            return
        if len(todo) > _lib.TPX_MAX_TENSORS:
            raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, f"{len(todo)} parameter tensors in one group (max {_lib.TPX_MAX_TENSORS})")
        if gi not in dev_steps:
            dev_steps[gi] = torch.tensor([int(todo[0][1]["step"])], dtype=torch.int32, device=todo[0][0].device)
        tl = _lib.TensorList()
        tl.n = len(todo)
        for i, (p, st) in enumerate(todo):
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            tl.param[i], tl.grad[i] = p.data_ptr(), g.data_ptr()
            tl.exp_avg[i], tl.exp_avg_sq[i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            tl.numel[i] = p.numel()
        with _lib.on_device(todo[0][0].device):
            _lib.check(_lib.lib().tpx_adam_step_dstep(
                ctypes.byref(tl), ctypes.c_float(lr), ctypes.c_float(group["betas"][0]), ctypes.c_float(group["betas"][1]),
                ctypes.c_float(group["eps"]), ctypes.c_float(group["weight_decay"]), _lib.ptr(dev_steps[gi]), _lib.stream_ptr()))
        for p, _ in todo:
            torch.autograd.graph.increment_version(p)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        dev_steps = getattr(self, "_dev_steps", None)
        for gi, group in enumerate(self.param_groups):
            lr = group["lr"]
            if isinstance(lr, torch.Tensor):
                lr = float(lr)
            chunk, step_of_chunk = [], None
            if dev_steps is not None:
                self._step_on_device(gi, group, lr, dev_steps)
                continue

            def flush():
                if not chunk:
                    return
                tl = _lib.TensorList()
                tl.n = len(chunk)
                for i, (p, st) in enumerate(chunk):
                    tl.param[i], tl.grad[i] = p.data_ptr(), p.grad.data_ptr()
                    tl.exp_avg[i], tl.exp_avg_sq[i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                    tl.numel[i] = p.numel()
                with _lib.on_device(chunk[0][0].device):
                    _lib.check(_lib.lib().tpx_adam_step(
                        ctypes.byref(tl), ctypes.c_float(lr), ctypes.c_float(group["betas"][0]),
                        ctypes.c_float(group["betas"][1]), ctypes.c_float(group["eps"]),
                        ctypes.c_float(group["weight_decay"]), ctypes.c_int32(step_of_chunk), _lib.stream_ptr()))
                for p, _ in chunk:                      # written through raw pointers: tell autograd / the
                    torch.autograd.graph.increment_version(p)     # engine's packed-image cache about it
                chunk.clear()

            for p in group["params"]:
                if p.grad is None:                      # untouched, as in torch.optim (reference parity: SURVEY 8(a))
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                    raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "native Adam needs contiguous float32 CUDA parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                if chunk and (st["step"] != step_of_chunk or len(chunk) == _lib.TPX_MAX_TENSORS
                              or p.device != chunk[0][0].device):
                    flush()
                step_of_chunk = st["step"]
                chunk.append((p, st))
            flush()
        return loss
