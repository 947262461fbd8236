"""Mirror of Actor / MLPPolicy, /root/reference/src/tacorl/networks/actor_critic/actor.py:18-156, 217-270
(config/networks/actor_critic/actor/default.yaml: MLPPolicy, continuous actions; actor/discrete_gripper.yaml:
continuous arm + open/close gripper head, the flat-CQL baseline's actor)."""
from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ...utils.config import instantiate, to_container
from ...utils.distributions import GumbelSoftmax, TanhNormal
from ..layers import Linear

LOG_SIG_MAX = 2
LOG_SIG_MIN = -5
MEAN_MIN = -9.0
MEAN_MAX = 9.0


class MLPPolicy(nn.Module):
    def __init__(self, input_dim: int, action_dim: int, num_layers: int = 2, hidden_dim: int = 256,
                 init_w: float = 1e-3, discrete_gripper: bool = False):
        super().__init__()
        self.discrete_gripper = discrete_gripper
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        if discrete_gripper:                      # registered first, as in the reference (state_dict order)
            self.gripper_action = Linear(hidden_dim, 2)
            self.gripper_action.weight.data.uniform_(-init_w, init_w)
            self.gripper_action.bias.data.uniform_(-init_w, init_w)
            action_dim -= 1
        self.fc_layers = nn.ModuleList([Linear(input_dim, hidden_dim)] +
                                       [Linear(hidden_dim, hidden_dim) for _ in range(num_layers - 1)])
        self.fc_mean = Linear(hidden_dim, action_dim)
        self.fc_log_std = Linear(hidden_dim, action_dim)
        for lin in (self.fc_log_std, self.fc_mean):
            lin.weight.data.uniform_(-init_w, init_w)
            lin.bias.data.uniform_(-init_w, init_w)

    def get_last_hidden_state(self, policy_input):
        x = policy_input
        for fc in self.fc_layers:
            x = fc(x, act="silu")
        return x

    def forward(self, policy_input):
        """policy_input: one tensor, or a (state_emb, goal_emb) pair that the fused kernel concatenates itself.
        The SiLU trunk and the fc_mean | fc_log_std (| gripper_action) heads run as ONE launch each way
        (ops.mlp_chain)."""
        layers = [(fc.weight, fc.bias) for fc in self.fc_layers]
        heads = [(self.fc_mean.weight, self.fc_mean.bias), (self.fc_log_std.weight, self.fc_log_std.bias)]
        if self.discrete_gripper:
            heads.append((self.gripper_action.weight, self.gripper_action.bias))
        layers.append(heads)
        raw = ops.mlp_chain(policy_input, layers, ("silu",) * len(self.fc_layers))
        lead = raw.shape[:-1]
        raw = raw.reshape(-1, raw.shape[-1])
        if self.discrete_gripper:
            mean, std = ops.gauss_head(raw[:, :-2])
            return mean.view(*lead, -1), std.view(*lead, -1), raw[:, -2:].reshape(*lead, 2)
        mean, std = ops.gauss_head(raw)
        return mean.view(*lead, -1), std.view(*lead, -1)


class Actor(nn.Module):
    def __init__(self, state_dim: int, goal_dim: int = 0, action_dim: int = 16, policy: dict = {},
                 discrete_gripper: bool = False):
        super().__init__()
        self.discrete_gripper = discrete_gripper
        self.action_dim = action_dim
        self.state_dim = state_dim
        self.goal_dim = goal_dim
        policy_cfg = to_container(policy)
        policy_cfg.update({"input_dim": state_dim + goal_dim, "action_dim": action_dim,
                           "discrete_gripper": discrete_gripper})
        self.policy = instantiate(policy_cfg)

    def forward(self, state_emb: torch.Tensor, goal_emb: Optional[torch.Tensor] = None):
        return self.policy((state_emb, goal_emb) if goal_emb is not None else state_emb)

    def get_dist(self, state_emb, goal_emb=None):
        mean, std = self.forward(state_emb, goal_emb)[:2]
        return TanhNormal(mean, std)

    def get_actions(self, observation, deterministic: bool = False, reparameterize: bool = False):
        if self.discrete_gripper:                                # actor.py:72-98
            mean, std, logits = self.forward(observation)
            if deterministic:
                grip = torch.argmax(torch.softmax(logits, dim=-1), dim=-1).unsqueeze(-1).to(mean.dtype) * 2.0 - 1
                actions = torch.cat((torch.tanh(mean), grip), dim=-1)
                return actions, torch.zeros_like(actions)
            dist, grip_dist = TanhNormal(mean, std), GumbelSoftmax(temperature=0.5, logits=logits)
            if reparameterize:
                actions, log_pi = dist.rsample_and_logprob()
                index = grip_dist.rsample_index()
            else:
                actions, log_pi = dist.sample_and_logprob()
                index = grip_dist.sample()
            log_pi = log_pi + grip_dist.log_prob(index)
            return torch.cat((actions, index * 2.0 - 1), dim=-1), log_pi
        mean, std = self.forward(observation)
        if deterministic:
            actions = torch.tanh(mean)
            return actions, torch.zeros_like(actions)
        dist = TanhNormal(mean, std)
        return dist.rsample_and_logprob() if reparameterize else dist.sample_and_logprob()

    def sample_n_with_log_prob(self, observation, n_actions: int):
        if self.discrete_gripper:                                # actor.py:117-133
            mean, std, logits = self.forward(observation)
            dist, grip_dist = TanhNormal(mean, std), GumbelSoftmax(temperature=0.5, logits=logits)
            actions, z = dist.sample_n(n_actions, return_pre_tanh_value=True)
            log_pi = dist.log_prob(actions, pre_tanh_value=z)
            index = grip_dist.sample((n_actions,))
            return torch.cat((actions, index * 2 - 1), dim=-1), log_pi + grip_dist.log_prob(index)
        mean, std = self.forward(observation)
        dist = TanhNormal(mean, std)
        actions, z = dist.sample_n(n_actions, return_pre_tanh_value=True)
        return actions, dist.log_prob(actions, pre_tanh_value=z)

    def log_prob(self, observations, actions):
        if self.discrete_gripper:                                # actor.py:143-153
            mean, std, logits = self.forward(observations)
            log_pi = TanhNormal(mean, std).log_prob(value=actions[..., :-1])
            return log_pi + GumbelSoftmax(temperature=0.5, logits=logits).log_prob(actions[..., -1:] / 2 + 0.5)
        mean, std = self.forward(observations)
        return TanhNormal(mean, std).log_prob(value=actions)
