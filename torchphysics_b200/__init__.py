"""B200-native PINN residual engine behind the TorchPhysics API.

``import torchphysics_b200 as tp`` gives the reference's namespaces for the hot path:
``tp.spaces``, ``tp.domains``, ``tp.samplers``, ``tp.models``, ``tp.utils``,
``tp.conditions``, ``tp.solver``.
"""
__version__ = "0.1.0"

from . import spaces, userfn, functionsets, domains, samplers, models, operators, utils, conditions, optim, solver, parallel  # noqa: F401
from . import deeponet as _deeponet

# the reference's tp.models namespace also holds the DeepONet classes (models/__init__.py)
for _n in ("TrunkLinear", "TrunkNet", "FCTrunkNet", "BranchNet", "FCBranchNet", "DeepONet"):
    setattr(models, _n, getattr(_deeponet, _n))
from .solver import OptimizerSetting, Solver, Trainer  # noqa: F401
