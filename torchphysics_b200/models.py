"""Models with the reference's interface (``models/model.py``, ``models/fcn.py``,
``models/activation_fn.py``, ``models/parameter.py``).

``FCN`` keeps the reference's ``nn.Sequential`` parameter layout
(``sequential.{0,2,4,...}.{weight,bias}``, Xavier-normal gain 5/3, output gain 1 --
reference ``fcn.py:9-29``) so ``state_dict``s interchange.  Its forward is the native jet
engine (``engine.py``) and nothing else: CPU tensors and networks without a kernel raise.
"""
import os

import torch
import torch.nn as nn

from . import _lib, jets
from .engine import FcnEngine, JetSpec
from .spaces import Points, Space
from .userfn import UserFunction


class Sinus(nn.Module):
    """sin activation (reference ``activation_fn.py:76-85``)."""

    def forward(self, input):
        return torch.sin(input)


class AdaptiveActivationFunction(nn.Module):
    """activation_fn(scaling * a * x) with learnable a (reference ``activation_fn.py:5-38``)."""

    def __init__(self, activation_fn, inital_a=1.0, scaling=1.0):
        super().__init__()
        self.activation_fn = activation_fn
        self.a = nn.Parameter(torch.tensor(inital_a))
        self.scaling = scaling

    def forward(self, x):
        return self.activation_fn(self.scaling * self.a * x)


class Parameter(Points):
    """Learnable parameter of an inverse problem (reference ``parameter.py:6-32``)."""

    def __init__(self, init, space, **kwargs):
        init = torch.as_tensor(init).float().reshape(1, -1)
        super().__init__(nn.Parameter(init), space, **kwargs)


class Model(nn.Module):
    def __init__(self, input_space, output_space):
        super().__init__()
        self.input_space = input_space
        self.output_space = output_space

    def _fix_points_order(self, points):
        if points.space != self.input_space:
            if points.space.keys() != self.input_space.keys():
                raise ValueError(f"Points are in {points.space} but should lie in {self.input_space}.")
            leaves = getattr(points, "_coord_leaves", None)
            points = points[..., list(self.input_space.keys())]
            if leaves is not None:
                points._coord_leaves = leaves
        return points

    def save(self, path):
        torch.save(self.state_dict(), path)

    def load(self, path):
        self.load_state_dict(torch.load(path))


def _fc_layers(hidden, input_dim, output_dim, activations, xavier_gains, linear_cls=nn.Linear):
    if not isinstance(activations, (list, tuple)):
        activations = len(hidden) * [activations]
    if not isinstance(xavier_gains, (list, tuple)):
        xavier_gains = len(hidden) * [xavier_gains]
    widths = [input_dim] + list(hidden)
    layers = []
    for i in range(len(hidden)):
        lin = linear_cls(widths[i], widths[i + 1])
        nn.init.xavier_normal_(lin.weight, gain=xavier_gains[i])
        layers += [lin, activations[i]]
    out = linear_cls(widths[-1], output_dim)
    nn.init.xavier_normal_(out.weight, gain=1)
    layers.append(out)
    return layers


def _inner_activation(m):
    """the activation a module applies after an optional trainable input scale (``AdaptiveActivationFunction``)"""
    return m.activation_fn if isinstance(m, AdaptiveActivationFunction) or type(m).__name__ == "AdaptiveActivationFunction" else m


def _activation_kind(mods):
    kinds = set()
    for m in mods:
        m = _inner_activation(m)
        if isinstance(m, nn.Tanh):
            kinds.add("tanh")
        elif isinstance(m, Sinus) or type(m).__name__ == "Sinus":
            kinds.add("sin")
        else:
            return None
    return kinds.pop() if len(kinds) == 1 else None


def strict_native():
    """Default: a network the kernels do not cover raises.  TPX_ALLOW_TORCH=1 lets such a
    network be evaluated by torch (with a TpxFallbackWarning) -- never silently."""
    return os.environ.get("TPX_ALLOW_TORCH", "0") != "1"


class _NativeFcnMixin:
    """Shared by FCN and the DeepONet trunk: route CUDA evaluations through the engine."""

    def _native_engine(self):
        eng = getattr(self, "_engine", None)
        if eng is None:
            seq = self.sequential
            lins = [m for m in seq if isinstance(m, nn.Linear)]
            acts = [m for m in seq if not isinstance(m, nn.Linear)]
            ok = (len(seq) == 2 * len(lins) - 1 and len(lins) >= 2
                  and all(isinstance(seq[2 * i], nn.Linear) for i in range(len(lins))))
            kind = _activation_kind(acts) if ok else None
            eng = False
            if kind is not None:
                try:
                    eng = FcnEngine(lins, kind)
                except _lib.TpxError as err:
                    if err.code != _lib.TPX_E_UNSUPPORTED:
                        raise
                    self._engine_reason = str(err)
            else:
                self._engine_reason = "layer structure or activation has no native kernel"
            object.__setattr__(self, "_engine", eng)
        return eng

    def _evaluate(self, points, input_space=None, first_layer=None, generic=None):
        """points: Points in model input order -> tensor (..., d_out) (JetTensor if native).
        ``first_layer`` = (W0, b0) tensors replacing the first Linear's (an affine input map folded in),
        ``input_space`` the space whose columns ``points`` carries in that case."""
        x = points.as_tensor
        if not x.is_cuda:
            raise _lib.TpxError(_lib.TPX_E_ARG, "FCN evaluation runs on the sm_100a kernels only: the points "
                                "must live on a CUDA device (there is no CPU path)")
        eng = self._native_engine()
        seq = self.sequential
        if generic is None:
            generic = lambda: seq(x)  # noqa: E731
        if eng is False:
            if strict_native():
                raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, self._engine_reason)
            jets.warn_fallback(self._engine_reason)
            return generic()
        batch_shape = x.shape[:-1]
        x2d = x.detach().reshape(-1, x.shape[-1])
        leaves = getattr(points, "_coord_leaves", None)
        leaf_cols = {}
        if leaves is not None:
            sl = (self.input_space if input_space is None else input_space).column_slices()
            for name, leaf in leaves.items():
                if name in sl:
                    leaf_cols[id(leaf)] = (leaf, list(range(sl[name].start, sl[name].stop)))
        hints = self.__dict__.setdefault("_jet_hints", {})
        key = jets.current_scope()
        spec = hints.get(key) if leaves is not None else None
        if spec is None:
            spec = JetSpec()
        params = None
        if first_layer is not None:
            params = list(first_layer) + eng.params()[2:]
        # AdaptiveActivationFunction (reference activation_fn.py:5-38): s(scaling * a * (W h + b)) = s((c W) h + c b) with
        # c = scaling * a -- the trainable slope folded into the layer's parameters as a differentiable torch expression, so
        # the kernels see a plain tanh / sin network and autograd carries dLoss/d(cW), dLoss/d(cb) on to W, b and a
        acts = [seq[2 * i + 1] for i in range(len(eng.linears) - 1)]
        if any(_inner_activation(m) is not m for m in acts):
            params = list(eng.params()) if params is None else list(params)
            for i, m in enumerate(acts):
                if _inner_activation(m) is not m:
                    c = m.scaling * m.a
                    params[2 * i], params[2 * i + 1] = c * params[2 * i], c * params[2 * i + 1]
        ev = jets.JetEvaluation(eng, x2d, batch_shape, leaf_cols, hints, key, generic, spec, params)
        return jets.JetTensor.wrap(ev.value(), ev, range(eng.d_out))


class FCN(Model, _NativeFcnMixin):
    """Fully connected network (reference ``fcn.py:32-87``)."""

    def __init__(self, input_space, output_space, hidden=(20, 20, 20), activations=nn.Tanh(),
                 xavier_gains=5 / 3, activation_fn_output=None):
        super().__init__(input_space, output_space)
        layers = _fc_layers(hidden, self.input_space.dim, self.output_space.dim, activations, xavier_gains)
        if activation_fn_output is not None:
            layers.append(activation_fn_output)
        self.sequential = nn.Sequential(*layers)

    def forward(self, points):
        points = self._fix_points_order(points)
        return Points(self._evaluate(points), self.output_space)


class NormalizationLayer(Model):
    """Trainable affine map of a domain's bounding box onto (-1,1)^d (reference ``model.py:57-92``)."""

    def __init__(self, domain):
        super().__init__(input_space=domain.space, output_space=domain.space)
        self.normalize = nn.Linear(domain.space.dim, domain.space.dim)
        box = domain.bounding_box()
        lo, hi = box[::2], box[1::2]
        width = torch.tensor([float(hi[i] - lo[i]) for i in range(domain.dim)])
        centre = torch.tensor([float(hi[i] + lo[i]) / 2 for i in range(domain.dim)])
        scale = 2.0 / width
        with torch.no_grad():
            self.normalize.weight.copy_(torch.diag(scale))
            self.normalize.bias.copy_(-centre * scale)

    def forward(self, points):
        points = self._fix_points_order(points)
        return Points(self.normalize(points), self.output_space)


class Parallel(Model):
    def __init__(self, *models):
        input_space, output_space = Space({}), Space({})
        for m in models:
            assert output_space.keys().isdisjoint(m.output_space)
            input_space = input_space * Space(m.input_space - input_space)
            output_space = output_space * m.output_space
        super().__init__(input_space, output_space)
        self.models = nn.ModuleList(models)

    def forward(self, points):
        outs = []
        for m in self.models:
            sub = points[..., list(m.input_space.keys())]
            leaves = getattr(points, "_coord_leaves", None)
            if leaves is not None:
                sub._coord_leaves = leaves
            outs.append(m(sub))
        joined = Points.joined(*outs)
        parts = [jets.jet_of(o.as_tensor) for o in outs]
        leaves = getattr(points, "_coord_leaves", None) or {}
        same_layout = all(ev is not None for ev, _ in parts) and all(
            len({tuple(ev.columns_of(leaf) or ()) for ev, _ in parts}) == 1 and parts[0][0].columns_of(leaf) is not None
            for leaf in leaves.values())
        if same_layout:
            # every member is a native evaluation: the joined output keeps its derivative streams
            jev = jets.JoinedEvaluation(parts)
            joined = Points(jets.JetTensor.wrap(joined.as_tensor, jev, range(joined.as_tensor.shape[-1])), joined.space)
        return joined


class Sequential(Model):
    """Chain of models (reference ``model.py:149-183``).  ``Sequential(NormalizationLayer, FCN)`` -- the
    composition of the reference's examples -- is evaluated as ONE native network: the trainable affine
    map is folded into the first layer (W0 Wn, W0 bn + b0 as differentiable torch expressions, so its
    gradient still reaches ``normalize.weight`` / ``.bias``) and the jets are taken with respect to the
    un-normalised coordinates."""

    def __init__(self, *models):
        super().__init__(models[0].input_space, models[-1].output_space)
        self.models = nn.ModuleList(models)

    def _fused(self, points):
        if len(self.models) != 2:
            return None
        norm, fcn = self.models
        if not (isinstance(norm, NormalizationLayer) and isinstance(fcn, FCN)):
            return None
        if fcn.input_space.keys() != norm.output_space.keys() or not points.as_tensor.is_cuda:
            return None
        if fcn._native_engine() is False:
            return None
        # columns of the normalised point in the order the network wants them
        src = norm.output_space.column_slices()
        cols = [c for name in fcn.input_space for c in range(src[name].start, src[name].stop)]
        lin0 = fcn.sequential[0]
        Wn, bn = norm.normalize.weight[cols, :], norm.normalize.bias[cols]
        first = (lin0.weight @ Wn, lin0.bias + lin0.weight @ bn)
        x = points.as_tensor
        out = fcn._evaluate(points, input_space=norm.input_space, first_layer=first,
                            generic=lambda: fcn.sequential(norm.normalize(x)[..., cols]))
        return Points(out, fcn.output_space)

    def forward(self, points):
        points = self._fix_points_order(points)
        fused = self._fused(points)
        if fused is not None:
            return fused
        for m in self.models:
            points = m(points)
        return points


class HardConstraint(Model):
    def __init__(self, model, constraint):
        super().__init__(model.input_space, model.output_space)
        self.model = model
        self.constraint = UserFunction(constraint)

    def forward(self, points):
        out = self.model(points)
        ev, cols = jets.jet_of(out.as_tensor)
        leaves = getattr(points, "_coord_leaves", None)
        if ev is not None and leaves is not None and cols == list(range(out.as_tensor.shape[-1])):
            # native model: the streams of g(x, u(x)) follow from those of u by the chain rule (jets.ConstrainedEvaluation)
            cev = jets.ConstrainedEvaluation(ev, self.constraint, leaves, self.model.input_space, self.output_space,
                                             self.model.output_space)
            value = cev.value()
            return Points(jets.JetTensor.wrap(value, cev, range(value.shape[-1])), self.output_space)
        res = self.constraint(points.join(out).coordinates)
        if isinstance(res, torch.Tensor):
            return Points(res, self.output_space)
        if isinstance(res, (tuple, list)):
            return Points(torch.column_stack(res), self.output_space)
        raise AssertionError("Constraint function should return a tensor or a tuple of tensors.")
