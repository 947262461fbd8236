"""`LightningModule` base: the real pytorch_lightning class when installed, else a minimal
stand-in with the hooks the reference modules use (log / save_hyperparameters / optimizers /
manual_backward / device / current_epoch), so the modules run under a plain training loop."""
import torch
import torch.nn as nn

try:  # pragma: no cover - pytorch_lightning is not in this image
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__()
            self.logged = {}
            self._optimizers = None
            self.current_epoch = 0
            self.automatic_optimization = True
            self.hparams = {}

        def log(self, name, value, *args, **kwargs):
            # keep the device tensor: no host sync on the hot path
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        def save_hyperparameters(self, *args, **kwargs):
            pass

        def optimizers(self):
            if self._optimizers is None:
                opts = self.configure_optimizers()
                self._optimizers = list(opts) if isinstance(opts, (list, tuple)) else [opts]
            return self._optimizers

        def manual_backward(self, loss, *args, **kwargs):
            loss.backward(*args, **kwargs)

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")
