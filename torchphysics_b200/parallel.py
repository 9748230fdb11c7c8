"""Data parallelism over collocation points: every rank owns a replica of the parameters
(broadcast from rank 0 when training starts) and draws ITS OWN batch of the samplers' n_points from its own
Philox stream (weak scaling: the global batch is world_size x n_points; ``shard_count`` is the helper for
callers that want a fixed global batch instead and size their samplers with it).  The ONLY exchange per step
is one all-reduce of the flat parameter gradient with the scalar losses appended.  The reference has no distributed code
(it would delegate to Lightning DDP, one bucketed all-reduce set per step)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def broadcast_module_state(module, src=0):
    """parameters and buffers of ``module`` take rank ``src``'s values on every rank (one flat broadcast per dtype)."""
    rank, size = world()
    if size == 1:
        return
    tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t is not None and t.numel() > 0]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    with torch.no_grad():
        for group in by_dtype.values():
            flat = torch.cat([t.detach().reshape(-1) for t in group])
            dist.broadcast(flat, src=src)
            off = 0
            for t in group:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()


class FlatGradAllReduce:
    """Flattens the gradients of ``params`` (+ ``n_scalars`` trailing slots) into one fp32 buffer, all-reduces
    it once (NCCL over NVLink on GPUs, gloo on CPU) and writes the mean back into ``.grad`` -- one concatenation,
    one collective, one multi-tensor copy, no host synchronisation.  A parameter without a gradient
    (``.grad is None``, which the reference also leaves unset for parameters the loss does not depend on) is
    left out: which parameters receive gradients is a property of the program, identical on every rank."""

    def __init__(self, params, n_scalars=0):
        self.params = [p for p in params]
        self.n_scalars = n_scalars

    def __call__(self, scalars=()):
        rank, size = world()
        if not self.params:
            return list(scalars)
        if size == 1:                       # nothing to exchange: gradients stay where backward left them
            return [s.detach().reshape(()).clone() if isinstance(s, torch.Tensor) else torch.tensor(float(s))
                    for s in scalars]
        grads = [p.grad for p in self.params if p.grad is not None]
        device = self.params[0].device
        tail = [s.detach().reshape(1).to(device=device, dtype=torch.float32) if isinstance(s, torch.Tensor)
                else torch.full((1,), float(s), device=device) for s in scalars]
        flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads] + tail)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / size)
        sizes = [g.numel() for g in grads]
        chunks = flat.split(sizes + [1] * len(tail))
        if grads:
            torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(chunks[:len(grads)], grads)])
        return [c.reshape(()).clone() for c in chunks[len(grads):]]


def shard_count(n_total, rank=None, size=None):
    """points owned by ``rank`` when ``n_total`` points are split over ``size`` ranks."""
    if rank is None or size is None:
        rank, size = world()
    base, rem = divmod(int(n_total), size)
    return base + (1 if rank < rem else 0)
