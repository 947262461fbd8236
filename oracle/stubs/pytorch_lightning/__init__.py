"""Test-infrastructure stub of `pytorch_lightning` (not installed in this image).

`LightningModule` = nn.Module + the handful of hooks the reference hot path calls
(log / save_hyperparameters / optimizers / manual_backward / device / current_epoch).
"""
import torch
import torch.nn as nn


class LightningModule(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.logged = {}
        self._optimizers = None
        self.current_epoch = 0
        self.automatic_optimization = True

    def log(self, name, value, *args, **kwargs):
        self.logged[name] = value.detach().clone() if torch.is_tensor(value) else value

    def save_hyperparameters(self, *args, **kwargs):
        pass

    def optimizers(self):
        if self._optimizers is None:
            opts = self.configure_optimizers()
            self._optimizers = opts if isinstance(opts, (list, tuple)) else [opts]
        return self._optimizers

    def manual_backward(self, loss, *args, **kwargs):
        loss.backward(*args, **kwargs)

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


class Callback:
    pass


class Trainer:
    def __init__(self, *a, **k):
        pass


class LightningDataModule:
    pass


def seed_everything(seed):
    torch.manual_seed(seed)
