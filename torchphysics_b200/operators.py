"""Differential operators with the reference's signatures
(``utils/differentialoperators/differentialoperators.py``: laplacian :11, grad :47,
normal_derivative :111, div :137, jac :233, partial :294, convective :320).

Two evaluations of the same definition:
  * native: ``model_out`` is a ``JetTensor`` and every derivative variable is one of the
    coordinate leaves of its evaluation -> the result is read from the value / first /
    Laplacian streams the sm_100a kernel propagated (no autograd sweep at all);
  * generic: anything else (hand-written torch functions, wrapped models, mixed or
    third-order partials) -> the textbook ``torch.autograd.grad(create_graph=True)``
    definition, which is the API's semantics for arbitrary graphs.
"""
import torch

from .engine import JetSpec
from .jets import jet_of, plain_of, warn_fallback


def _native(model_out, variables, need_lap):
    """(evaluation, out_cols, [input columns per variable]) or None."""
    ev, out_cols = jet_of(model_out)
    if ev is None:
        return None
    per_var = []
    for v in variables:
        cols = ev.columns_of(v)
        if cols is None:
            warn_fallback("derivative variable is not a coordinate leaf of the model input")
            return None
        per_var.append(cols)
    flat = sorted({c for cols in per_var for c in cols})
    if len(flat) > 3:
        warn_fallback("more than 3 derivative directions")
        return None
    spec = JetSpec(flat, [1.0] * len(flat) if need_lap else None)
    if not ev.ensure(spec):
        warn_fallback("requested streams do not fit one jet evaluation")
        return None
    return ev, out_cols, per_var


def _sum_last(t):
    """sum over the output columns, keeping the axis (a single column is returned as is)."""
    return t if t.shape[-1] == 1 else t.sum(dim=-1, keepdim=True)


def _d(scalar, var):
    return torch.autograd.grad(scalar, var, create_graph=True)[0]


def grad(model_out, *derivative_variable):
    hit = _native(model_out, derivative_variable, False)
    if hit is not None:
        ev, oc, per_var = hit
        pieces = [_sum_last(ev.first(oc, c)) for cols in per_var for c in cols]
        out = pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=-1)
        if len(pieces) > 1 or not hasattr(out, "_tpx_grad_of"):
            try:                                     # lets laplacian(u, x, grad=this) recognise its own gradient
                out._tpx_grad_of = (id(ev), tuple(oc), tuple(c for cols in per_var for c in cols))
            except AttributeError:
                pass
        return out
    u = plain_of(model_out)
    return torch.column_stack([_d(u.sum(), v) for v in derivative_variable])


def laplacian(model_out, *derivative_variable, grad=None):
    hit = _native(model_out, derivative_variable, True)
    if hit is not None:
        ev, oc, per_var = hit
        if grad is not None:
            # the reference differentiates the tensor it is given (differentialoperators.py:34-43); the native streams carry
            # no graph to the coordinates, so only tp.utils.grad of this very output w.r.t. these variables is accepted
            want = (id(ev), tuple(oc), tuple(c for cols in per_var for c in cols))
            if getattr(grad, "_tpx_grad_of", None) != want:
                from . import _lib
                raise _lib.TpxError(_lib.TPX_E_UNSUPPORTED, "laplacian(..., grad=g): g is not tp.utils.grad of this model output "
                                    "with respect to the same variables; drop the argument (the Laplacian stream needs no gradient)")
        return _sum_last(ev.second(oc))
    u = plain_of(model_out)
    out = torch.zeros((*u.shape[:-1], 1), device=u.device)
    for v in derivative_variable:
        if grad is None or len(derivative_variable) > 1:
            grad = _d(u.sum(), v)
        if grad.grad_fn is None:        # u linear in v
            continue
        for i in range(v.shape[-1]):
            out = out + _d(grad.narrow(-1, i, 1).sum(), v).narrow(-1, i, 1)
    return out


def div(model_out, *derivative_variable):
    hit = _native(model_out, derivative_variable, False)
    if hit is not None:
        ev, oc, per_var = hit
        acc, pos = None, 0
        for cols in per_var:
            for c in cols:
                term = ev.first([oc[pos]], c)
                acc = term if acc is None else acc + term
                pos += 1
        return acc
    u = plain_of(model_out)
    out = torch.zeros((*derivative_variable[0].shape[:-1], 1), device=derivative_variable[0].device)
    col = 0
    for v in derivative_variable:
        for i in range(v.shape[-1]):
            out = out + _d(u.narrow(-1, col + i, 1).sum(), v).narrow(-1, i, 1)
        col += v.shape[-1]
    return out


def jac(model_out, *derivative_variable):
    hit = _native(model_out, derivative_variable, False)
    if hit is not None:
        ev, oc, per_var = hit
        cols = [c for cs in per_var for c in cs]
        return torch.stack([ev.first(oc, c) for c in cols], dim=-1)
    u = plain_of(model_out)
    rows = []
    for i in range(u.shape[1]):
        rows.append(torch.cat([_d(u[..., i].sum(), v) for v in derivative_variable], dim=-1))
    return torch.stack(rows, dim=-2)


def partial(model_out, *derivative_variables):
    ev, _ = jet_of(model_out)
    if ev is not None:
        vs = derivative_variables
        if len(vs) == 1:
            return grad(model_out, vs[0])
        if len(vs) == 2 and vs[0] is vs[1] and vs[0].shape[-1] == 1:
            return laplacian(model_out, vs[0])
        warn_fallback("mixed or higher-order partial derivative")
    du = plain_of(model_out)
    for v in derivative_variables:
        if du.grad_fn is None:
            return torch.zeros_like(v)
        du = _d(du.sum(), v)
    return du


def normal_derivative(model_out, normals, *derivative_variable):
    return (grad(model_out, *derivative_variable) * normals).sum(dim=-1, keepdim=True)


def convective(deriv_out, convective_field, *derivative_variable):
    j = jac(deriv_out, *derivative_variable)
    # (v . grad) u = J v per point (reference differentialoperators.py:320-342 uses torch.bmm: a batch of N tiny products that
    # cuBLAS runs as ~16 launches per 2^20 points, forward and twice in backward: 4.8 ms of an 18.6 ms Navier-Stokes step);
    # the same contraction as one broadcast multiply + sum over the last axis
    return (j * convective_field.unsqueeze(dim=-2)).sum(dim=-1)


def rot(model_out, *derivative_variable):
    j = jac(model_out, *derivative_variable)
    out = torch.zeros((*j.shape[:-2], 3), device=j.device)
    out[..., 0] = j[..., 2, 1] - j[..., 1, 2]
    out[..., 1] = j[..., 0, 2] - j[..., 2, 0]
    out[..., 2] = j[..., 1, 0] - j[..., 0, 1]
    return out


def sym_grad(model_out, *derivative_variable):
    j = jac(model_out, *derivative_variable)
    return 0.5 * (j + torch.transpose(j, -2, -1))


def matrix_div(model_out, *derivative_variable):
    out = torch.zeros((len(model_out), model_out.shape[1]), device=model_out.device)
    for i in range(model_out.shape[1]):
        out[:, i:i + 1] = div(model_out.narrow(1, i, 1).squeeze(1), *derivative_variable)
    return out
