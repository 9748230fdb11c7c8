"""Data parallelism over collocation points: every rank owns N/G points and a
replica of the parameters; the ONLY exchange per step is one all-reduce of the flat
parameter gradient with the scalar losses appended.  The reference has no distributed code
(it would delegate to Lightning DDP, one bucketed all-reduce set per step)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class FlatGradAllReduce:
    """Flattens the gradients of ``params`` (+ ``n_scalars`` trailing slots) into one fp32
    buffer, all-reduces it once (NCCL over NVLink on GPUs, gloo on CPU) and scatters the
    mean back into ``.grad``.  Parameters without a gradient (``.grad is None``, which the
    reference also leaves unset for parameters the loss does not depend on) contribute
    zeros and stay ``None`` unless some rank produced a gradient."""

    def __init__(self, params, n_scalars=0):
        self.params = [p for p in params]
        self.sizes = [p.numel() for p in self.params]
        self.n_scalars = n_scalars
        self.total = sum(self.sizes) + n_scalars + len(self.params)   # + has-grad flags
        self.buf = None

    def _buffer(self, device):
        if self.buf is None or self.buf.device != device:
            self.buf = torch.zeros(self.total, dtype=torch.float32, device=device)
        return self.buf

    def __call__(self, scalars=()):
        rank, size = world()
        if not self.params:
            return list(scalars)
        if size == 1:                       # nothing to exchange: gradients stay where backward left them
            return [s.detach().reshape(()).clone() if isinstance(s, torch.Tensor) else torch.tensor(float(s))
                    for s in scalars]
        device = self.params[0].device
        buf = self._buffer(device)
        buf.zero_()
        off = 0
        nparam = len(self.params)
        flags = buf[self.total - nparam:]
        for i, (p, n) in enumerate(zip(self.params, self.sizes)):
            if p.grad is not None:
                buf[off:off + n].copy_(p.grad.reshape(-1))
                flags[i] = 1.0
            off += n
        for k, s in enumerate(scalars):
            buf[off + k] = s.detach().reshape(()) if isinstance(s, torch.Tensor) else float(s)
        if size > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            buf[: self.total - nparam].mul_(1.0 / size)
        off = 0
        have = flags.tolist() if size > 1 else None
        for i, (p, n) in enumerate(zip(self.params, self.sizes)):
            if size > 1 and have[i] > 0:
                g = buf[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
            off += n
        return [buf[off + k].clone() for k in range(len(scalars))]


def shard_count(n_total, rank=None, size=None):
    """points owned by ``rank`` when ``n_total`` points are split over ``size`` ranks."""
    if rank is None or size is None:
        rank, size = world()
    base, rem = divmod(int(n_total), size)
    return base + (1 if rank < rem else 0)
