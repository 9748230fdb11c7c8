"""Import stub (test infrastructure): lets the reference package import without Lightning."""
import torch


class LightningModule(torch.nn.Module):
    pass


from . import callbacks  # noqa: E402,F401
