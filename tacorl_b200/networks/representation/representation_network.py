"""Mirror of LateFusion, /root/reference/src/tacorl/networks/representation/representation_network.py."""
from typing import List

import torch
import torch.nn as nn

from ...utils.config import instantiate


class LateFusion(nn.Module):
    """One encoder per image modality (config/networks/representation/lmp_encoder.yaml)."""

    def __init__(self, networks: dict = {}, modalities: List[str] = []):
        super().__init__()
        for modality in modalities:
            assert modality in networks.keys(), f"Network configuration for {modality} is missing"
        nets = {}
        self.visual_state_dim = 0
        for modality, encoder_cfg in networks.items():
            if modality in modalities:
                nets[modality] = instantiate(encoder_cfg)
                self.visual_state_dim += nets[modality].latent_dim
        self.networks = nn.ModuleDict(nets)

    def forward(self, inputs):
        return {m: self.networks[m](inputs[m]) for m in inputs.keys() if m in self.networks.keys()}

    def get_state_from_observation(self, observation: dict, modalities: List[str] = [], cat_output: bool = True):
        if not isinstance(observation, dict):
            return observation
        state = {}
        for modality in modalities:
            if "rgb" in modality or "depth" in modality:
                img = observation[modality]
                squeeze = img.ndim == 3
                if squeeze:
                    img = img.unsqueeze(0)
                out = self.networks[modality](img)
                state[modality] = out.squeeze(0)
            else:
                if isinstance(observation[modality], list):
                    observation[modality] = torch.stack(observation[modality], dim=-1)
                state[modality] = observation[modality].float()
        if cat_output:
            vals = list(state.values())
            state = vals[0] if len(vals) == 1 else torch.cat(vals, dim=-1)
        return state

    def calc_state_dim(self, modalities: List[str] = []):
        return sum(self.networks[m].latent_dim for m in modalities)
