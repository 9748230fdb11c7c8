"""B200-native PINN residual engine behind the TorchPhysics API."""
__version__ = "0.1.0"
