"""``Solver`` / ``OptimizerSetting`` with the reference's constructor and ``training_step``
semantics (reference ``solver.py:9-136``), plus the minimal ``Trainer`` the reference gets
from pytorch_lightning (not installed in this image): device placement, the step loop,
``optimizer.step()`` and -- when launched with one process per GPU -- the single all-reduce
of dL/dtheta per step (``parallel.FlatGradAllReduce``)."""
import torch
import torch.nn as nn

from . import domains, optim, parallel


class OptimizerSetting:
    def __init__(self, optimizer_class, lr, optimizer_args={}, scheduler_class=None, scheduler_args={},
                 scheduler_frequency=1):
        self.optimizer_class = optimizer_class
        self.lr = lr
        self.optimizer_args = optimizer_args
        self.scheduler_class = scheduler_class
        self.scheduler_args = scheduler_args
        self.scheduler_frequency = scheduler_frequency


class Solver(nn.Module):
    def __init__(self, train_conditions, val_conditions=(), optimizer_setting=OptimizerSetting(torch.optim.Adam, 1e-3)):
        super().__init__()
        self.train_conditions = nn.ModuleList(train_conditions)
        self.val_conditions = nn.ModuleList(val_conditions)
        self.optimizer_setting = optimizer_setting
        self.n_training_step = 0
        self.logged = {}
        self._device = torch.device("cpu")

    @property
    def device(self):
        return self._device

    def log(self, name, value, **kwargs):
        self.logged[name] = value.detach() if isinstance(value, torch.Tensor) else value

    def on_train_start(self):
        for condition in list(self.train_conditions) + list(self.val_conditions):
            condition._move_static_data(self.device)
        self.n_training_step = 0

    def training_step(self, batch=None, batch_idx=None):
        loss = torch.zeros(1, requires_grad=True, device=self.device)
        for condition in self.train_conditions:
            cond_loss = condition(device=self.device, iteration=self.n_training_step)
            self.log(f"train/{condition.name}", cond_loss)
            loss = loss + condition.weight * cond_loss
        self.log("train/loss", loss, prog_bar=True)
        self.n_training_step += 1
        return loss

    def validation_step(self, batch=None, batch_idx=None):
        for condition in self.val_conditions:
            with torch.set_grad_enabled(condition.track_gradients is not False):
                self.log(f"val/{condition.name}", condition(device=self.device))

    def configure_optimizers(self):
        setting = self.optimizer_setting
        cls = setting.optimizer_class
        if (cls is torch.optim.Adam and self.device.type == "cuda" and optim.Adam.covers(setting.optimizer_args)
                and all(p.is_cuda for p in self.parameters())):
            # same update, one launch per step instead of 4-5 per parameter tensor
            args = {k: v for k, v in setting.optimizer_args.items() if k in ("betas", "eps", "weight_decay")}
            optimizer = optim.Adam(self.parameters(), lr=setting.lr, **args)
        else:
            optimizer = cls(self.parameters(), lr=setting.lr, **setting.optimizer_args)
        if setting.scheduler_class is None:
            return optimizer
        scheduler = setting.scheduler_class(optimizer, **setting.scheduler_args)
        return [optimizer], [{"scheduler": scheduler, "name": "learning_rate", "interval": "step",
                              "frequency": setting.scheduler_frequency}]


class _GraphCaptureFailed(RuntimeError):
    pass


class Trainer:
    """The slice of ``pl.Trainer`` the reference's examples use:
    ``Trainer(devices=1, accelerator='gpu', max_steps=K).fit(solver)``."""

    def __init__(self, max_steps=1000, accelerator="gpu", devices=1, device=None, cuda_graph=False, **unused):
        self.max_steps = max_steps
        # cuda_graph=True: after a few eager steps the WHOLE step (samplers with a device-resident Philox key -> forward jets
        # -> residual -> loss -> reverse sweep -> gradient all-reduce -> Adam with a device-resident step count) is captured
        # once and replayed: one graph launch per step instead of ~40 kernel launches (small batches are launch bound)
        self.cuda_graph = cuda_graph
        self.graphed = False
        if device is None:
            if accelerator in ("gpu", "cuda", "auto") and torch.cuda.is_available():
                device = torch.device("cuda", torch.cuda.current_device())
            else:
                device = torch.device("cpu")
        self.device = torch.device(device)
        self.losses = []

    def _fit_graphed(self, solver, optimizer, reducer, size, eager_steps=3):
        """runs all ``max_steps`` steps: ``eager_steps`` eagerly (the conditions learn their jet specs, buffers are pooled),
        then one captured step replayed; returns the number of steps done (0 if the step cannot be captured)."""
        import warnings

        def one_step():
            optimizer.zero_grad(set_to_none=True)
            loss = solver.training_step(None, 0)
            loss.backward()
            (mean_loss,) = reducer([loss])
            optimizer.step()
            return mean_loss if size > 1 else loss.detach()

        n_eager = min(eager_steps, self.max_steps)
        for _ in range(n_eager):
            self.losses.append(one_step())
        if n_eager == self.max_steps:
            return n_eager
        for m in solver.modules():                           # tensors kept from the eager steps must not carry their graphs
            release = getattr(m, "_release_graph", None)
            if release is not None:
                release()
        try:
            domains.use_device_key(self.device)
            optimizer.use_device_step(True)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = one_step()                     # (recorded, not run)
        except Exception as err:                             # e.g. a sampler that reads its counters on the host
            domains.use_device_key(None)
            optimizer.use_device_step(False)
            torch.cuda.synchronize(self.device)
            import os
            import traceback
            detail = traceback.format_exc() if os.environ.get("TPX_DEBUG") else ""
            warnings.warn(f"the training step could not be captured in a CUDA graph ({err!r}); continuing eagerly\n{detail}")
            return n_eager
        self.graphed = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_eager, self.max_steps):
            graph.replay()
            self.losses.append(static_loss.reshape(-1)[:1].clone().reshape(()))
        e1.record()
        e1.synchronize()
        self.graph_step_ms = e0.elapsed_time(e1) / max(1, self.max_steps - n_eager)
        solver.n_training_step = self.max_steps
        torch.cuda.synchronize(self.device)
        domains.use_device_key(None)
        optimizer.use_device_step(False)
        # the graph is not kept: its private memory pool (and, with several ranks, the captured NCCL kernels -- a process group
        # cannot be destroyed while a graph that holds them is alive) goes away here; the losses above are copies
        del graph, static_loss
        self._last = self.losses[-1]
        return self.max_steps

    def fit(self, solver):
        rank, size = parallel.world()
        if size > 1:
            domains.set_rank_stream(rank)
        solver.to(self.device)
        solver._device = self.device
        solver.trainer = self
        if size > 1:
            # every replica starts from rank 0's weights (what Lightning DDP does for the reference); without this, ranks
            # whose models were built under different seeds would apply the averaged gradient to diverging copies
            parallel.broadcast_module_state(solver)
        configured = solver.configure_optimizers()
        schedulers = []
        if isinstance(configured, tuple) or isinstance(configured, list):
            optimizer = configured[0][0]
            schedulers = configured[1]
        else:
            optimizer = configured
        self.optimizer = optimizer
        solver.on_train_start()
        params = [p for p in solver.parameters() if p.requires_grad]
        reducer = parallel.FlatGradAllReduce(params, n_scalars=1)
        closure_based = isinstance(optimizer, torch.optim.LBFGS)
        first = 0
        if (self.cuda_graph and self.device.type == "cuda" and not closure_based and not schedulers
                and isinstance(optimizer, optim.Adam)):
            first = self._fit_graphed(solver, optimizer, reducer, size)
        for step in range(first, self.max_steps):
            def closure():
                optimizer.zero_grad(set_to_none=True)
                loss = solver.training_step(None, step)
                loss.backward()
                (mean_loss,) = reducer([loss])
                self._last = mean_loss if size > 1 else loss.detach()
                return loss
            if closure_based:
                optimizer.step(closure)
            else:
                closure()
                optimizer.step()
            for sch in schedulers:
                if (step + 1) % sch["frequency"] == 0:
                    sch["scheduler"].step()
            self.losses.append(self._last)
        return solver
