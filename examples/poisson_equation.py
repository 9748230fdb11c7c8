"""The reference's examples/pinn/poisson-equation.ipynb and examples/deepritz/poisson-equation.ipynb on
torchphysics_b200: Laplace u = 0 on a disc with a square hole (Circle - Parallelogram), u = cos(x1) cos(x2) on the whole
boundary (outer circle and the hole's edges: `D.boundary`), density sampling; as a PINN and as a Deep Ritz functional.
Only the import line differs from the notebooks (plots left out).

    python examples/poisson_equation.py [pinn|ritz] [max_steps]
"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402


def main(kind="pinn", max_steps=300):
    X, U = tp.spaces.R2("x"), tp.spaces.R1("u")
    A = tp.domains.Circle(X, center=[0, 0], radius=1.0)
    B = tp.domains.Parallelogram(X, [-0.25, -0.25], [-0.25, 0.25], [0.25, -0.25])
    D = A - B                                              # cutting operation
    dens = 1e3 if kind == "pinn" else 1e4
    inner_sampler = tp.samplers.RandomUniformSampler(D, density=dens)
    boundary_sampler = tp.samplers.RandomUniformSampler(D.boundary, density=dens)
    inner_points, boundary_points = next(inner_sampler), next(boundary_sampler)

    def f(x):
        return torch.cos(x[:, :1]) * torch.cos(x[:, 1:])

    if kind == "pinn":
        model = tp.models.FCN(input_space=X, output_space=U, hidden=(20, 20, 20))
        pde_condition = tp.conditions.PINNCondition(module=model, sampler=inner_sampler,
                                                    residual_fn=lambda u, x: tp.utils.laplacian(u, x))
        boundary_condition = tp.conditions.PINNCondition(module=model, sampler=boundary_sampler,
                                                         residual_fn=lambda u, f: u - f, data_functions={"f": f})
    else:
        model = tp.models.FCN(input_space=X, output_space=U, hidden=(50, 50, 50))
        pde_condition = tp.conditions.DeepRitzCondition(
            module=model, sampler=inner_sampler, weight=5,
            integrand_fn=lambda u, x: 0.5 * (torch.sum(tp.utils.grad(u, x) ** 2, dim=1)))
        boundary_condition = tp.conditions.DeepRitzCondition(module=model, sampler=boundary_sampler,
                                                             integrand_fn=lambda u, f: (u - f) ** 2,
                                                             data_functions={"f": f})
    solver = tp.solver.Solver([pde_condition, boundary_condition])
    trainer = tp.solver.Trainer(devices=1, accelerator="gpu", max_steps=max_steps)
    trainer.fit(solver)
    losses = [float(v) for v in trainer.losses]
    # error against the harmonic function with these boundary values is not known in closed form; report the boundary fit
    with torch.no_grad():
        pts = boundary_sampler.sample_points(device="cuda")
        err = float((model(pts).as_tensor - f(pts.as_tensor)).abs().max())
    print(f"poisson ({kind}): {len(inner_points)} inner + {len(boundary_points)} boundary points/step, "
          f"loss {losses[0]:.4f} -> {losses[-1]:.4f} after {max_steps} steps, max boundary misfit {err:.3f}")
    return losses, err


if __name__ == "__main__":
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        main(sys.argv[1] if len(sys.argv) > 1 else "pinn", int(sys.argv[2]) if len(sys.argv) > 2 else 300)
