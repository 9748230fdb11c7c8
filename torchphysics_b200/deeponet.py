"""DeepONet with the reference's interface (``models/deeponet/deeponet.py:9-128``,
``trunknets.py:9-154``, ``branchnets.py:12-194``, ``layers.py:5-109``) on the native jets.

SURVEY 8(a) rows a12/a13.  The reference evaluates the trunk on F copies of the N trunk
points (``deeponet_condition.py:63``) and gets every derivative of ``u[f,n,o] = sum_p
T[n,o,p] B[f,o,p]`` from autograd sweeps through that (F, N, o, p) tensor.  Here the trunk's
value / first-derivative / Laplacian streams are propagated ONCE per point by the sm_100a jet
kernels, and the inner product with the branch coefficients is folded into the trunk's output
layer: ``u[f,n,o] = (B W_L)[f,o,:] a(x_n) + (B b_L)[f,o]`` with ``a`` the last hidden layer, so
the trunk is evaluated as a network of F*o outputs whose output layer runs as tensor-core products
(``csrc/tcw_kernels.cu``, "wide output layer") and whose streams come back as (function, point)
arrays (``tpx_jet_desc.out_major``).  ``d_k u`` and ``Lap u`` are the other streams of the same
evaluation; the reverse sweep returns dL/d(B W_L), which autograd carries on to the branch net and
to W_L through the two parameter-sized products.  Trunks that see different points per function
(``trunk_input_copied=False``), or more than TPX_MAX_DOUT (function, output) columns, keep their
own output layer and contract each stream with the branch coefficients by a torch einsum.  The
branch net (F rows of sensor values, d_in = #sensors > TPX_MAX_DIN) is a caller of the path and
stays on torch ops.
"""
import torch
import torch.nn as nn

from . import _lib, jets
from .engine import JetSpec
from .functionsets import FunctionSet
from .models import Model, Sequential, _NativeFcnMixin, _fc_layers, strict_native
from .spaces import Points
from .userfn import UserFunction


class _CopiedLinear(torch.autograd.Function):
    """y[f] = x[0] W^T + b for every copy f, while each copy of the input receives its own
    gradient row g[f] W -- the derivative semantics of the reference's ``linear`` Function
    (``layers.py:5-39``), which the operators rely on for per-function derivatives."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x[0], weight)
        y0 = nn.functional.linear(x[0], weight, bias)
        return y0.expand(x.shape[0], *y0.shape)

    @staticmethod
    def backward(ctx, g):
        x0, weight = ctx.saved_tensors
        gsum = g.sum(dim=0).reshape(-1, weight.shape[0])
        return g.matmul(weight), gsum.t().matmul(x0.reshape(-1, weight.shape[1])), gsum.sum(dim=0)


class TrunkLinear(nn.Linear):
    """``nn.Linear`` for an input that is identical along its first axis (reference
    ``layers.py:42-109``): computes on ``input[0]`` and expands.  Same parameters, same default
    initialisation and RNG draws as the reference's class (``reset_parameters`` :86-94).  This
    torch forward is the generic-graph definition; the hot path never calls it (the trunk's
    weights go to the jet kernels through ``FcnEngine``)."""

    def forward(self, input):
        if input.dim() < 3:
            input = input.unsqueeze(0)
        return _CopiedLinear.apply(input, self.weight, self.bias)


def _fc_trunk_layers(hidden, input_dim, output_dim, activations, xavier_gains):
    """reference ``trunknets.py:72-90``: ``_fc_layers`` with TrunkLinear modules (same draw order)."""
    layers = _fc_layers(hidden, input_dim, output_dim, activations, xavier_gains, linear_cls=TrunkLinear)
    return layers


class TrunkNet(Model):
    """reference ``trunknets.py:9-69``."""

    def __init__(self, input_space, default_trunk_input, trunk_input_copied=True):
        super().__init__(input_space, output_space=None)
        self.output_neurons = 0
        self.trunk_input_copied = trunk_input_copied
        if torch.is_tensor(default_trunk_input):
            self.default_trunk_input = Points(default_trunk_input, input_space)
        elif isinstance(default_trunk_input, Points):
            self.default_trunk_input = default_trunk_input
        else:
            raise ValueError("Provided default input is not supported!")
        self.default_trunk_input = self._fix_points_order(self.default_trunk_input)

    def finalize(self, output_space, output_neurons):
        self.output_neurons = output_neurons
        self.output_space = output_space

    def _reshape_multidimensional_output(self, output):
        return output.reshape(*output.shape[:-1], self.output_space.dim, self.output_neurons)


class FCTrunkNet(TrunkNet, _NativeFcnMixin):
    """Fully connected trunk (reference ``trunknets.py:93-154``)."""

    def __init__(self, input_space, default_trunk_input, hidden=(20, 20, 20), activations=nn.Tanh(),
                 xavier_gains=5 / 3, trunk_input_copied=True):
        super().__init__(input_space, default_trunk_input, trunk_input_copied=trunk_input_copied)
        self.hidden = hidden
        self.activations = activations
        self.xavier_gains = xavier_gains

    def finalize(self, output_space, output_neurons):
        super().finalize(output_space, output_neurons)
        build = _fc_trunk_layers if self.trunk_input_copied else _fc_layers
        self.sequential = nn.Sequential(*build(self.hidden, self.input_space.dim,
                                               self.output_neurons * self.output_space.dim,
                                               self.activations, self.xavier_gains))
        self.__dict__.pop("_engine", None)

    def _points_or_default(self, points):
        if points:
            return self._fix_points_order(points)
        dev = next(self.sequential[0].parameters()).device
        self.default_trunk_input = self.default_trunk_input.to(dev)
        return self.default_trunk_input

    def unique_rows(self, x):
        """(rows the trunk has to evaluate as (n, d), batch shape of the full output)."""
        if self.trunk_input_copied:
            if x.dim() < 3:
                x = x.unsqueeze(0)
            return x[0].reshape(-1, x.shape[-1]), tuple(x.shape[:-1])
        return x.reshape(-1, x.shape[-1]), tuple(x.shape[:-1])

    def forward(self, points=None):
        """(..., output dim, output neurons) tensor of trunk values (value stream only)."""
        points = self._points_or_default(points)
        x = points.as_tensor
        if not x.is_cuda:
            raise _lib.TpxError(_lib.TPX_E_ARG, "trunk evaluation runs on the sm_100a kernels only: the points "
                                "must live on a CUDA device (there is no CPU path)")
        eng = self._native_engine()
        if eng is False:
            if strict_native():
                raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, self._engine_reason)
            jets.warn_fallback(self._engine_reason)
            return self._reshape_multidimensional_output(self.sequential(x))
        rows, batch = self.unique_rows(x.detach())
        u, _, _ = eng.jet(rows, JetSpec())
        if self.trunk_input_copied:
            u = u.reshape(1, *batch[1:], u.shape[-1]).expand(*batch, u.shape[-1])
        else:
            u = u.reshape(*batch, u.shape[-1])
        return self._reshape_multidimensional_output(u)


class BranchNet(Model):
    """reference ``branchnets.py:12-136``: discretises the input function on ``grid`` and keeps
    the branch coefficients of the last ``fix_input`` in ``current_out``."""

    def __init__(self, function_space, grid):
        super().__init__(function_space, output_space=None)
        if torch.is_tensor(grid):
            grid = Points(grid, function_space.input_space)
        self.output_neurons = 0
        self.register_buffer("grid_buffer", grid.as_tensor)
        self.current_out = torch.empty(0)

    @property
    def grid(self):
        return Points(self.grid_buffer, self.input_space.input_space)

    def _release_graph(self):
        """drop the autograd graph of the last ``fix_input`` (``Trainer`` calls this before it captures a CUDA graph: the kept
        coefficients would keep the eager step's AccumulateGrad nodes -- bound to the default stream -- alive)."""
        self.current_out = self.current_out.detach()

    def finalize(self, output_space, output_neurons):
        self.output_neurons = output_neurons
        self.output_space = output_space

    def _reshape_multidimensional_output(self, output):
        return output.reshape(*output.shape[:-1], self.output_space.dim, self.output_neurons)

    def forward(self, discrete_function_batch, device="cpu"):
        raise NotImplementedError

    def fix_input(self, function, device="cpu"):
        out_space = self.input_space.output_space
        if isinstance(function, FunctionSet):
            function.create_functions(device=device)
            fns = function.get_function(torch.arange(function.function_set_size))
            discrete_fn = fns(self.grid.to(device))
        elif callable(function):
            function = UserFunction(function)
            discrete_fn = function(self.grid.to(device))
            discrete_fn = Points(discrete_fn.unsqueeze(0), out_space)
        elif isinstance(function, Points):
            discrete_fn = Points(function._t.unsqueeze(0), out_space) if function._t.dim() < 3 else function
        elif isinstance(function, torch.Tensor):
            discrete_fn = Points(function.unsqueeze(0) if function.dim() < 3 else function, out_space)
        else:
            raise NotImplementedError("Function has to be callable, a FunctionSet, a tensor, or a tp.Point")
        self(discrete_fn)


class FCBranchNet(BranchNet):
    """Fully connected branch (reference ``branchnets.py:139-194``)."""

    def __init__(self, function_space, grid, hidden=(20, 20, 20), activations=nn.Tanh(), xavier_gains=5 / 3):
        super().__init__(function_space, grid)
        self.hidden = hidden
        self.activations = activations
        self.xavier_gains = xavier_gains
        self.input_neurons = len(self.grid) * self.input_space.output_space.dim

    def finalize(self, output_space, output_neurons):
        super().finalize(output_space, output_neurons)
        self.sequential = nn.Sequential(*_fc_layers(self.hidden, self.input_neurons,
                                                    self.output_neurons * self.output_space.dim,
                                                    self.activations, self.xavier_gains))

    def forward(self, discrete_function_batch):
        x = discrete_function_batch.as_tensor.reshape(-1, self.input_neurons)
        self.current_out = self._reshape_multidimensional_output(self.sequential(x))


class DeepONetEvaluation:
    """Streams of ``u = <trunk, branch>`` for the operators (same protocol as ``jets.JetEvaluation``).

    ``tev`` is the trunk's JetEvaluation on the rows the trunk really has to evaluate; every
    stream of u is the contraction of that trunk stream with the branch coefficients."""

    def __init__(self, tev, branch_out, batch_shape, copied, leaf_cols, plain_fn, folded=False):
        self.tev = tev
        self.B = branch_out                       # (Fb, o, p), graph to the branch parameters
        self.batch_shape = tuple(batch_shape)     # shape of the model output without its last axis
        self.copied = copied
        self.folded = folded                      # tev's columns are already (function, output): no contraction left
        self.leaf_cols = leaf_cols
        self.plain_fn = plain_fn
        self._plain = None
        self._cache = {}
        self._value = self._contract(tev.u)

    def _contract(self, stream):
        """(rows, o*p) trunk stream -> (..., o) stream of u."""
        Fb, o, p = self.B.shape
        if self.copied:
            if self.folded and self.tev.spec.out_major:
                out = stream.reshape(Fb, o, -1).permute(0, 2, 1)          # (Fb * o, n) -> (Fb, n, o): contiguous for o = 1
            elif self.folded:
                out = stream.reshape(-1, Fb, o).permute(1, 0, 2)          # (n, Fb * o) -> (Fb, n, o), a view
            else:
                T = stream.reshape(-1, o, p)                              # (n, o, p)
                if o == 1:
                    out = torch.matmul(self.B[:, 0, :], T[:, 0, :].t()).unsqueeze(-1)      # (Fb, n, 1)
                else:
                    out = torch.einsum("nop,fop->fno", T, self.B)
            lead = self.batch_shape[0]
            if lead != Fb:
                if Fb == 1:
                    out = out.expand(lead, *out.shape[1:])
                elif lead != 1:
                    raise ValueError(f"{Fb} branch inputs do not match trunk batch {lead}")
            return out.reshape(max(lead, Fb), *self.batch_shape[1:], o)
        T = stream.reshape(self.batch_shape[0], -1, o, p)                  # (F, n, o, p)
        out = torch.einsum("fnop,fop->fno", T, self.B.expand(T.shape[0], o, p))
        return out.reshape(*self.batch_shape, o)

    # -- JetEvaluation protocol ------------------------------------------------------------
    def columns_of(self, var):
        hit = self.leaf_cols.get(id(var))
        if hit is None or hit[0] is not var:
            return None
        return hit[1]

    def ensure(self, need):
        return self.tev.ensure(need)

    def lap_spec(self):
        return self.tev.lap_spec()

    def value(self):
        return self._value

    def first(self, out_cols, in_col):
        part = self.tev.part_with(in_col)
        k = part.spec.dcols.index(in_col)
        key = ("d", id(part.du), k)
        if key not in self._cache:
            self._cache[key] = self._contract(part.du[k] if part.spec.out_major else part.du[:, :, k])
        return jets.JetEvaluation._columns(self._cache[key], out_cols, self._cache[key].dim() - 1)

    def second(self, out_cols):
        lap = self.tev.lap_part().lap
        key = ("lap", id(lap))
        if key not in self._cache:
            self._cache[key] = self._contract(lap)
        return jets.JetEvaluation._columns(self._cache[key], out_cols, self._cache[key].dim() - 1)

    def plain(self):
        if self._plain is None:
            self._plain = self.plain_fn()
        return self._plain


class DeepONet(Model):
    """reference ``deeponet.py:9-128``."""

    def __init__(self, trunk_net, branch_net, output_space, output_neurons, constrain_fn=None):
        last = trunk_net.models[-1] if isinstance(trunk_net, Sequential) else trunk_net
        assert isinstance(last, TrunkNet)
        assert isinstance(branch_net, BranchNet)
        super().__init__(input_space=trunk_net.input_space, output_space=output_space)
        self.trunk = trunk_net
        self.branch = branch_net
        last.finalize(output_space, output_neurons)
        self.branch.finalize(output_space, output_neurons)
        self.constrain_fn = UserFunction(constrain_fn) if constrain_fn else None

    def _forward_branch(self, function_set, iteration_num=-1, device="cpu"):
        self.branch.fix_input(function_set, device)

    def fix_branch_input(self, function, device="cpu"):
        self.branch.fix_input(function, device=device)

    def _generic(self, trunk_inputs):
        """the reference's torch-op evaluation (``deeponet.py:91-100``); only reachable with
        TPX_ALLOW_TORCH=1 or through ``plain()`` for operators the jets do not cover."""
        if isinstance(self.trunk, Sequential):
            pts = self.trunk._fix_points_order(trunk_inputs)
            for m in self.trunk.models[:-1]:
                pts = m(pts)
            last = self.trunk.models[-1]
        else:
            last, pts = self.trunk, self.trunk._points_or_default(trunk_inputs)
        x = pts.as_tensor
        h = x
        for m in last.sequential:
            h = nn.functional.linear(h, m.weight, m.bias) if isinstance(m, nn.Linear) else m(h)
        trunk_out = last._reshape_multidimensional_output(h)
        cur = self.branch.current_out
        view = [1] * (trunk_out.dim() - 3)
        branch_out = cur.view(cur.shape[0], *view, self.branch.output_space.dim, self.branch.output_neurons)
        return torch.sum(trunk_out * branch_out, dim=-1)

    def forward(self, trunk_inputs=None, branch_inputs=None, device="cpu"):
        if branch_inputs is not None:
            self.fix_branch_input(branch_inputs, device=device)
        trunk = self.trunk
        native = isinstance(trunk, FCTrunkNet) and trunk._native_engine() is not False
        if not native:
            reason = getattr(trunk, "_engine_reason", "trunk is not a plain FCTrunkNet")
            if strict_native():
                raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, reason)
            jets.warn_fallback(reason)
            out = self._generic(trunk_inputs)
        else:
            out = self._native(trunk_inputs)
        points_out = Points(out, self.output_space)
        if self.constrain_fn is None:
            return points_out
        if trunk_inputs is None:
            trunk_inputs = trunk.default_trunk_input.repeat(self.branch.current_out.shape[0])
        return Points(self.constrain_fn(trunk_inputs.join(points_out).coordinates), self.output_space)

    def _folded_engine(self, n_cols):
        """Engine of the trunk's hidden layers with ``n_cols`` = functions x output dim output columns (cached per
        width; False when no kernel takes that shape, e.g. more than TPX_MAX_DOUT columns: the trunk then keeps its
        own output layer and the streams are contracted with the branch coefficients by torch)."""
        cache = self.__dict__.setdefault("_folded", {})
        if n_cols not in cache:
            base = self.trunk._native_engine()
            try:
                eng = base.with_outputs(n_cols, next(self.trunk.parameters()).device)
            except _lib.TpxError as err:
                if err.code != _lib.TPX_E_UNSUPPORTED:
                    raise
                eng = False
            cache[n_cols] = eng
        return cache[n_cols]

    def _native(self, trunk_inputs):
        trunk = self.trunk
        points = trunk._points_or_default(trunk_inputs)
        x = points.as_tensor                 # CPU tensors are refused by FcnEngine.jet: there is no CPU path
        eng = trunk._native_engine()
        rows, batch = trunk.unique_rows(x.detach())
        leaves = getattr(points, "_coord_leaves", None)
        leaf_cols = {}
        if leaves is not None:
            sl = self.input_space.column_slices()
            for name, leaf in leaves.items():
                if name in sl:
                    leaf_cols[id(leaf)] = (leaf, list(range(sl[name].start, sl[name].stop)))
        hints = self.__dict__.setdefault("_jet_hints", {})
        key = jets.current_scope()
        spec = hints.get(key) if leaves is not None else None
        if spec is None:
            spec = JetSpec()
        B = self.branch.current_out
        B = B.reshape(B.shape[0], self.branch.output_space.dim, self.branch.output_neurons)
        folded = self._folded_engine(B.shape[0] * B.shape[1]) if trunk.trunk_input_copied else False
        if folded is not False:
            # u[f, n, o] = sum_p B[f, o, p] (W_L a(x_n) + b_L)[o, p] = (B W_L)[f, o, :] a(x_n) + (B b_L)[f, o]: the inner product
            # with the branch coefficients IS an output layer of F * o columns on the trunk's last hidden layer.  The
            # two parameter-sized products below (F x p x h) are the only library calls; everything of size N runs in
            # the jet kernels (the output layer as UMMA products, 128 (function, output) columns per launch).
            Fb, o, p = B.shape
            last = [m for m in trunk.sequential if isinstance(m, nn.Linear)][-1]
            Wl, bl = last.weight.reshape(o, p, -1), last.bias.reshape(o, p)
            if o == 1:
                W_eff, b_eff = B[:, 0, :].matmul(Wl[0]), B[:, 0, :].matmul(bl[0])
            else:
                W_eff = torch.einsum("fop,oph->foh", B, Wl).reshape(Fb * o, -1)
                b_eff = torch.einsum("fop,op->fo", B, bl).reshape(Fb * o)
            params = eng.params()[:-2] + [W_eff, b_eff]
            want = bool(getattr(folded, "out_major", False))
            if spec.out_major != want:
                spec = JetSpec(spec.dcols, spec.lap_w, out_major=want)
            tev = jets.JetEvaluation(folded, rows.contiguous(), rows.shape[:-1], {}, hints, key, None, spec, params)
        else:
            if spec.out_major:
                spec = JetSpec(spec.dcols, spec.lap_w)
            tev = jets.JetEvaluation(eng, rows.contiguous(), rows.shape[:-1], {}, hints, key, None, spec)
        ev = DeepONetEvaluation(tev, B, batch, trunk.trunk_input_copied, leaf_cols,
                                lambda: self._generic(points), folded=folded is not False)
        return jets.JetTensor.wrap(ev.value(), ev, range(self.output_space.dim))
