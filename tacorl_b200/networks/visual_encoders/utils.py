"""Mirror of /root/reference/src/tacorl/networks/visual_encoders/utils.py:22-76."""
import torch
import torch.nn as nn


class SpatialSoftArgmax(nn.Module):
    """Holds the (learned) temperature; the op is fused into tacorl_lmp_encoder_{fwd,bwd}."""

    def __init__(self, temperature: float = None, normalize: bool = False):
        super().__init__()
        if normalize:
            raise NotImplementedError("normalize=True is not used by lmp_vision_encoder.yaml")
        if temperature is None:
            self.temperature = nn.Parameter(torch.ones(1))
        else:
            self.register_buffer("temperature", torch.tensor([float(temperature)]), persistent=False)
        self.normalize = normalize
