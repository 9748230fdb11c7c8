"""The reference's flagship example (examples/pinn/heat-equation.ipynb) on torchphysics_b200: parametric heat
equation u_t = D Laplace_x u on a 20 x 20 plate, time 0..40, D in [0.1, 1], with an adaptive sampler for the PDE,
Dirichlet boundary and an initial condition, normalised inputs.  Only the import line differs from the notebook
(validation against the FDM data set and the plots are left out: data conditions and plotting are out of scope).

    python examples/heat_equation.py [max_steps]
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchphysics_b200 as tp  # noqa: E402


def main(max_steps=200):
    X, T, D, U = tp.spaces.R2("x"), tp.spaces.R1("t"), tp.spaces.R1("D"), tp.spaces.R1("u")
    h, w = 20, 20
    A_x = tp.domains.Parallelogram(X, [0, 0], [w, 0], [0, h])
    A_t = tp.domains.Interval(T, 0, 40)
    A_D = tp.domains.Interval(D, 0.1, 1.0)

    inner_sampler = tp.samplers.AdaptiveRandomRejectionSampler(A_x * A_t * A_D, density=1)
    initial_v_sampler = tp.samplers.RandomUniformSampler(A_x * A_t.boundary_left * A_D, density=1)
    boundary_v_sampler = tp.samplers.RandomUniformSampler(A_x.boundary * A_t * A_D, density=1)

    model = tp.models.Sequential(
        tp.models.NormalizationLayer(A_x * A_t * A_D),
        tp.models.FCN(input_space=X * T * D, output_space=U, hidden=(50, 50, 50)),
    )

    def heat_residual(u, x, t, D):
        return D * tp.utils.laplacian(u, x) - tp.utils.grad(u, t)

    pde_condition = tp.conditions.PINNCondition(module=model, sampler=inner_sampler, residual_fn=heat_residual,
                                                name="pde_condition")

    def boundary_v_residual(u):
        return u

    boundary_v_condition = tp.conditions.PINNCondition(module=model, sampler=boundary_v_sampler,
                                                       residual_fn=boundary_v_residual, name="boundary_condition")

    def f(x):
        return torch.sin(math.pi / w * x[:, :1]) * torch.sin(math.pi / h * x[:, 1:])

    def initial_v_residual(u, f):
        return u - f

    initial_v_condition = tp.conditions.PINNCondition(module=model, sampler=initial_v_sampler,
                                                      residual_fn=initial_v_residual, data_functions={"f": f},
                                                      name="initial_condition")

    solver = tp.solver.Solver([pde_condition, boundary_v_condition, initial_v_condition])
    trainer = tp.solver.Trainer(devices=1, accelerator="gpu", max_steps=max_steps)
    trainer.fit(solver)
    losses = [float(v) for v in trainer.losses]
    n = [len(s.sample_points(device="cuda")) for s in (boundary_v_sampler, initial_v_sampler)]
    print(f"heat equation: {len(inner_sampler.last_points)} + {n[0]} + {n[1]} points/step, "
          f"loss {losses[0]:.4f} -> {losses[-1]:.4f} after {max_steps} steps")
    return losses


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200)
