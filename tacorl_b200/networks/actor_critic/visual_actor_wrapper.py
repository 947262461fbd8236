"""Mirror of VisualActorWrapper, /root/reference/src/tacorl/networks/actor_critic/visual_actor_wrapper.py."""
from typing import List, Optional, Union

import torch
import torch.nn as nn


class VisualActorWrapper(nn.Module):
    def __init__(self, actor: nn.Module, encoder: nn.Module, goal_encoder: Optional[nn.Module] = None,
                 env_modalities: List[str] = [], goal_modalities: List[str] = []):
        super().__init__()
        self.actor = actor
        self.action_dim = actor.action_dim
        self.discrete_gripper = actor.discrete_gripper
        self.encoder = encoder
        self.goal_encoder = goal_encoder
        self.env_modalities = env_modalities
        self.goal_modalities = goal_modalities

    def get_emb_obs_representation(self, obs: Union[dict, torch.Tensor]):
        if not isinstance(obs, dict):
            return obs
        obs_dict = obs["observation"] if (len(self.goal_modalities) > 0 and "goal" in obs) else obs
        return self.encoder.get_state_from_observation(observation=obs_dict, modalities=self.env_modalities)

    def get_emb_representation(self, obs: Union[dict, torch.Tensor]):
        if not isinstance(obs, dict):
            return obs
        if len(self.goal_modalities) > 0 and "goal" in obs:
            emb_obs = self.encoder.get_state_from_observation(observation=obs["observation"],
                                                              modalities=self.env_modalities)
            emb_goal = self.encoder.get_state_from_observation(observation=obs["goal"],
                                                               modalities=self.goal_modalities)
            if self.goal_encoder is not None:
                emb_goal = self.goal_encoder(emb_goal)
            return torch.cat([emb_obs, emb_goal], dim=-1)
        return self.encoder.get_state_from_observation(observation=obs, modalities=self.env_modalities)

    def forward(self, obs: Union[dict, torch.Tensor], *args, **kwargs):
        return self.actor(self.get_emb_representation(obs=obs), *args, **kwargs)

    def get_actions(self, observation, *args, **kwargs):
        return self.actor.get_actions(self.get_emb_representation(obs=observation), *args, **kwargs)

    def sample_n_with_log_prob(self, observation, *args, **kwargs):
        return self.actor.sample_n_with_log_prob(self.get_emb_representation(obs=observation), *args, **kwargs)

    def log_prob(self, observation, *args, **kwargs):
        return self.actor.log_prob(self.get_emb_representation(obs=observation), *args, **kwargs)
