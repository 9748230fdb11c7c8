"""Oracle part 1: the reference's autograd path restated on torch CPU ops.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function cites the reference
lines it follows (paths relative to ``/root/reference/src/torchphysics``).  The
reference builds each differential operator from ``torch.autograd.grad(...,
create_graph=True)``; so does this file, on the same torch, which makes it the
executable statement of "what the reference computes" for given points and weights.
"""
import math

import torch
import torch.nn as nn


class Sin(nn.Module):
    """models/activation_fn.py:76-85 (``Sinus``): plain ``torch.sin``."""

    def forward(self, z):
        return torch.sin(z)


def make_activation(name):
    if name == "tanh":
        return nn.Tanh()
    if name == "sin":
        return Sin()
    raise ValueError(name)


def build_fcn(d_in, d_out, hidden, activation="tanh", gain=5.0 / 3.0):
    """models/fcn.py:9-29: Linear/act pairs, Xavier-normal weights (gain ``gain``,
    output layer gain 1), default ``nn.Linear`` bias init.  The RNG draw order equals
    the reference's (each Linear is constructed, then its weight re-drawn), so under the
    same ``torch.manual_seed`` the parameters are identical to ``tp.models.FCN``'s.
    """
    widths = [d_in] + list(hidden)
    mods = []
    for a, b in zip(widths[:-1], widths[1:]):
        lin = nn.Linear(a, b)
        nn.init.xavier_normal_(lin.weight, gain=gain)
        mods += [lin, make_activation(activation)]
    last = nn.Linear(widths[-1], d_out)
    nn.init.xavier_normal_(last.weight, gain=1)
    mods.append(last)
    return nn.Sequential(*mods)


def fcn_from_params(params, activation="tanh", dtype=torch.float32):
    """Build the same Sequential from a flat list [W0, b0, W1, b1, ...] (numpy/tensors)."""
    mods = []
    n_lin = len(params) // 2
    for i in range(n_lin):
        W = torch.as_tensor(params[2 * i]).to(dtype)
        b = torch.as_tensor(params[2 * i + 1]).to(dtype)
        lin = nn.Linear(W.shape[1], W.shape[0]).to(dtype)
        with torch.no_grad():
            lin.weight.copy_(W)
            lin.bias.copy_(b)
        mods.append(lin)
        if i < n_lin - 1:
            mods.append(make_activation(activation))
    return nn.Sequential(*mods)


def linear_params(seq):
    out = []
    for m in seq:
        if isinstance(m, nn.Linear):
            out += [m.weight, m.bias]
    return out


# --------------------------------------------------------------------------------------
# differential operators (utils/differentialoperators/differentialoperators.py)
# --------------------------------------------------------------------------------------
def _d(out_scalar, var):
    return torch.autograd.grad(out_scalar, var, create_graph=True)[0]


def grad(u, *variables):
    """:47-66 -- one reverse sweep of ``u.sum()`` per variable, columns stacked."""
    return torch.column_stack([_d(u.sum(), v) for v in variables])


def laplacian(u, *variables):
    """:11-44 -- sum over variables and their columns of the pure second derivatives;
    result is float32 of shape (..., 1) (:31); a variable u is linear in contributes 0 (:37).
    """
    acc = torch.zeros((*u.shape[:-1], 1), device=u.device)
    for v in variables:
        g = _d(u.sum(), v)
        if g.grad_fn is None:
            continue
        for i in range(v.shape[-1]):
            acc = acc + _d(g[..., i : i + 1].sum(), v)[..., i : i + 1]
    return acc


def div(u, *variables):
    """:137-165 -- sum_i d u_i / d x_i, output columns consumed in variable order."""
    acc = torch.zeros((*variables[0].shape[:-1], 1), device=variables[0].device)
    col = 0
    for v in variables:
        for i in range(v.shape[-1]):
            acc = acc + _d(u[..., col + i : col + i + 1].sum(), v)[..., i : i + 1]
        col += v.shape[-1]
    return acc


def jac(u, *variables):
    """:233-258 -- (b, m, n) Jacobian, one sweep per output component and variable."""
    rows = []
    for i in range(u.shape[1]):
        rows.append(torch.cat([_d(u[..., i].sum(), v) for v in variables], dim=-1))
    return torch.stack(rows, dim=-2)


def partial(u, *variables):
    """:294-317 -- chained first derivatives; zeros once the graph ends."""
    du = u
    for v in variables:
        if du.grad_fn is None:
            return torch.zeros_like(v)
        du = _d(du.sum(), v)
    return du


def normal_derivative(u, normals, *variables):
    """:111-134."""
    return (grad(u, *variables) * normals).sum(dim=-1, keepdim=True)


def convective(u, field, *variables):
    """:320-342 -- (v . nabla) u = J_u v."""
    return torch.bmm(jac(u, *variables), field.unsqueeze(2)).squeeze(2)


# --------------------------------------------------------------------------------------
# losses (problem/conditions/condition.py)
# --------------------------------------------------------------------------------------
def squared_error(r):
    """condition.py:11-27 -- sum of squares over every axis but the first."""
    return torch.sum(torch.square(r), dim=list(range(1, r.dim())))


def pinn_loss(residual):
    """condition.py:473-495 -- SquaredError then torch.mean."""
    return torch.mean(squared_error(residual))


def mean_loss(integrand):
    """condition.py:351-373 -- Identity then torch.mean (Deep Ritz)."""
    return torch.mean(integrand)


def track(x, splits):
    """problem/spaces/points.py:380-384 -- per-variable column views become leaves that
    require grad, and the model input is their re-concatenation."""
    cols, start = {}, 0
    for name, width in splits:
        leaf = x[..., start : start + width].detach().clone().requires_grad_(True)
        cols[name] = leaf
        start += width
    return cols, torch.cat([cols[n] for n, _ in splits], dim=-1)


# --------------------------------------------------------------------------------------
# whole steps, used for parity and for the CPU baseline timing
# --------------------------------------------------------------------------------------
def poisson_source(x):
    """f(x) = 2 pi^2 sin(pi x1) sin(pi x2)."""
    return 2 * math.pi ** 2 * torch.sin(math.pi * x[:, :1]) * torch.sin(math.pi * x[:, 1:2])


def poisson_step(seq, x, backward=True):
    """Config 1: residual = laplacian(u, x) + f, loss = mean(residual^2) and dL/dtheta.
    Mirrors SingleModuleCondition.forward (condition.py:278-304) + loss.backward()."""
    cols, xin = track(x, [("x", x.shape[1])])
    f = poisson_source(cols["x"])
    u = seq(xin)
    g = grad(u, cols["x"])
    lap = laplacian(u, cols["x"])
    r = lap + f
    loss = pinn_loss(r)
    out = {"u": u, "grad": g, "lap": lap, "residual": r, "loss": loss}
    if backward:
        params = linear_params(seq)
        gs = torch.autograd.grad(loss, params, allow_unused=True)
        out["dtheta"] = gs
    return out


def heat_step(seq, xt, backward=True):
    """Config 2: residual = u_t - laplacian_x(u); input columns (x1, x2, t)."""
    cols, xin = track(xt, [("x", 2), ("t", 1)])
    u = seq(xin)
    ut = grad(u, cols["t"])
    lap = laplacian(u, cols["x"])
    r = ut - lap
    loss = pinn_loss(r)
    out = {"u": u, "ut": ut, "lap": lap, "residual": r, "loss": loss}
    if backward:
        out["dtheta"] = torch.autograd.grad(loss, linear_params(seq), allow_unused=True)
    return out


def navier_stokes_step(seq, x, nu=0.01, backward=True):
    """Config 3: outputs (u in R^2, p); residual rows [(u.grad)u - nu lap u + grad p, div u]."""
    cols, xin = track(x, [("x", 2)])
    y = seq(xin)
    u, p = y[:, :2], y[:, 2:3]
    conv = convective(u, u, cols["x"])
    lap = torch.cat([laplacian(u[:, :1], cols["x"]), laplacian(u[:, 1:2], cols["x"])], dim=1)
    gp = grad(p, cols["x"])
    mom = conv - nu * lap + gp
    cont = div(u, cols["x"])
    r = torch.cat([mom, cont], dim=1)
    loss = pinn_loss(r)
    out = {"y": y, "jac": jac(y, cols["x"]), "lap": lap, "residual": r, "loss": loss}
    if backward:
        out["dtheta"] = torch.autograd.grad(loss, linear_params(seq), allow_unused=True)
    return out


def deep_ritz_step(seq, x, backward=True):
    """Config 5: integrand = 1/2 |grad u|^2 - u, loss = mean(integrand)."""
    cols, xin = track(x, [("x", 2)])
    u = seq(xin)
    g = grad(u, cols["x"])
    integrand = 0.5 * (g ** 2).sum(dim=1, keepdim=True) - u
    loss = mean_loss(integrand)
    out = {"u": u, "grad": g, "integrand": integrand, "loss": loss}
    if backward:
        out["dtheta"] = torch.autograd.grad(loss, linear_params(seq), allow_unused=True)
    return out


def deeponet_step(trunk, branch, t, branch_in, fvals, p, backward=True):
    """Config 4 (problem/conditions/deeponet_condition.py:52-108, models/deeponet/
    deeponet.py:91-102): the trunk runs on F copies of the N points, the output is
    sum_p trunk[f,n,p] * branch[f,p]; residual u'' + f; loss = mean_f sum_n r^2.
    ``t`` (N,1), ``branch_in`` (F, sensors), ``fvals`` (F, N, 1)."""
    F = branch_in.shape[0]
    tt = t.unsqueeze(0).repeat(F, 1, 1).detach().clone().requires_grad_(True)
    T = trunk(tt)                                  # (F, N, p)
    B = branch(branch_in).reshape(F, 1, p)         # (F, 1, p)
    u = (T * B).sum(dim=-1, keepdim=True)          # (F, N, 1)
    lap = laplacian(u, tt)
    r = lap + fvals
    loss = torch.mean(squared_error(r))
    out = {"u": u, "lap": lap, "residual": r, "loss": loss}
    if backward:
        ps = linear_params(trunk) + linear_params(branch)
        out["dtheta"] = torch.autograd.grad(loss, ps, allow_unused=True)
    return out
