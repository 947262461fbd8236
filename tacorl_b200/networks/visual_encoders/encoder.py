"""Mirror of LMPVisionEncoder, /root/reference/src/tacorl/networks/visual_encoders/encoder.py:349-428.
Same ctor kwargs (config/networks/encoder/lmp_vision_encoder.yaml) and state_dict keys
(model.{0,2,4}.{weight,bias}, model.6.temperature, fc_layers.{0,3}.{weight,bias})."""
import torch
import torch.nn as nn

from ... import ops
from ..layers import Conv2dParams, Linear, Marker
from .utils import SpatialSoftArgmax


class LMPVisionEncoder(nn.Module):
    def __init__(self, input_channels: int = 3, latent_dim: int = 32, hidden_dim: int = 256,
                 activation_function: str = "ReLU", dropout: float = 0.0, temperature: float = None,
                 normalize_spatial_softmax: bool = False, normalize_output: bool = False, vib: bool = False):
        super().__init__()
        if input_channels != 3 or activation_function != "ReLU" or dropout != 0.0 or normalize_output or vib:
            raise NotImplementedError(
                "the B200 encoder kernels cover the shipped lmp_vision_encoder.yaml configuration "
                "(3 input channels, ReLU, dropout 0, no output LayerNorm, no VIB)")
        self.latent_dim = latent_dim
        # uint8 frames may be fed directly: ScaleImageTensor + Normalize(mean, std) of the reference's transform
        # pipeline (config/datamodule/transform_manager/rl_train.yaml:2-14) are then fused into the first kernel
        self.input_mean, self.input_std = 0.5, 0.5
        self.normalize_output = normalize_output
        self.vib = vib
        self.model = nn.Sequential(
            Conv2dParams(input_channels, 32, 8, 4), Marker("ReLU (fused)"),
            Conv2dParams(32, 64, 4, 2), Marker("ReLU (fused)"),
            Conv2dParams(64, 64, 3, 1), Marker("ReLU (fused)"),
            SpatialSoftArgmax(temperature, normalize_spatial_softmax), Marker("Flatten"))
        self.fc_layers = nn.Sequential(Linear(128, hidden_dim), Marker("ReLU (fused)"), Marker("Dropout(0)"),
                                       Linear(hidden_dim, latent_dim))

    def kernel_params(self):
        m, f = self.model, self.fc_layers
        return [m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[4].weight, m[4].bias, m[6].temperature,
                f[0].weight, f[0].bias, f[3].weight, f[3].bias]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.lmp_encoder(x, self.kernel_params(), (self.input_mean, self.input_std))
