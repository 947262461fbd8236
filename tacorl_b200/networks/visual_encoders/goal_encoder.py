"""Mirror of VisualGoalEncoder, /root/reference/src/tacorl/networks/visual_encoders/goal_encoder.py."""
import torch
import torch.nn as nn

from ... import ops
from ..layers import Linear, Marker


class VisualGoalEncoder(nn.Module):
    def __init__(self, in_features: int = 32, out_features: int = 32, hidden_size: int = 256,
                 activation_function: str = "ReLU", last_layer_activation: str = "Identity",
                 normalize_output: bool = False):
        super().__init__()
        if activation_function != "ReLU" or last_layer_activation != "Identity" or normalize_output:
            raise NotImplementedError("goal encoder kernels cover config/networks/goal_encoder/default.yaml")
        self.normalize_output = normalize_output
        self.mlp = nn.Sequential(Linear(in_features, hidden_size), Marker("ReLU (fused)"),
                                 Linear(hidden_size, hidden_size), Marker("ReLU (fused)"),
                                 Linear(hidden_size, out_features))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # the three layers in one fused launch each way (ops.mlp_chain)
        return ops.mlp_chain(x, [(self.mlp[i].weight, self.mlp[i].bias) for i in (0, 2, 4)], ("relu", "relu"))
