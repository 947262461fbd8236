"""Mirror of PlanRecognitionTanhNetwork,
/root/reference/src/tacorl/networks/plan_encoders/plan_recognition_tanh_net.py:10-52."""
from typing import Tuple

import torch
import torch.nn as nn

from ... import ops
from ...utils.distributions import TanhNormal
from ..layers import Linear, ReluRNN


class PlanRecognitionTanhNetwork(nn.Module):
    def __init__(self, state_dim: int, latent_plan_dim: int = 16, birnn_dropout_p: float = 0.0,
                 min_std: float = 0.0001, hidden_dim: int = 2048):
        super().__init__()
        self.latent_plan_dim = latent_plan_dim
        self.min_std = min_std
        self.state_dim = state_dim
        self.birnn_model = ReluRNN(state_dim, hidden_dim, num_layers=2, bidirectional=True,
                                   dropout=birnn_dropout_p)
        self.mean_fc = Linear(2 * hidden_dim, latent_plan_dim)
        self.variance_fc = Linear(2 * hidden_dim, latent_plan_dim)

    def forward(self, perceptual_emb: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        # only out[:, -1] is consumed (:43): the kernels skip the unused reverse-direction steps
        x, _ = self.birnn_model(perceptual_emb, last_only=True)
        w = torch.cat([self.mean_fc.weight, self.variance_fc.weight], dim=0)
        b = torch.cat([self.mean_fc.bias, self.variance_fc.bias], dim=0)
        return ops.softplus_head(ops.linear(x, w, b), self.min_std)

    def __call__(self, *args, **kwargs):
        mean, std = super().__call__(*args, **kwargs)
        return TanhNormal(mean, std)
