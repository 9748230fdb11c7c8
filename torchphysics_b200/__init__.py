"""B200-native PINN residual engine behind the TorchPhysics API.

``import torchphysics_b200 as tp`` gives the reference's namespaces for the hot path:
``tp.spaces``, ``tp.domains``, ``tp.samplers``, ``tp.models``, ``tp.utils``,
``tp.conditions``, ``tp.solver``.
"""
__version__ = "0.1.0"

from . import spaces, userfn, domains, samplers, models, operators, utils, conditions, solver, parallel  # noqa: F401
from .solver import OptimizerSetting, Solver, Trainer  # noqa: F401
