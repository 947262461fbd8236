"""Mirror of Critic / MLPQNetwork, /root/reference/src/tacorl/networks/actor_critic/critic.py:9-30, 73-97."""
import torch
import torch.nn as nn

from ... import ops
from ...utils.config import instantiate, to_container
from ..layers import Linear


class MLPQNetwork(nn.Module):
    def __init__(self, input_dim: int, hidden_dim: int = 256, num_layers: int = 2,
                 last_layer_activation: str = "Identity", init_w: float = 1e-3):
        super().__init__()
        if last_layer_activation != "Identity":
            raise NotImplementedError("q_network/default.yaml uses an Identity output activation")
        self.fc_layers = nn.ModuleList([Linear(input_dim, hidden_dim)] +
                                       [Linear(hidden_dim, hidden_dim) for _ in range(num_layers - 1)])
        self.out = Linear(hidden_dim, 1)
        self.out.weight.data.uniform_(-init_w, init_w)
        self.out.bias.data.uniform_(-init_w, init_w)

    def forward(self, q_input, detach_params: bool = False):
        """q_input: one tensor or an (embedding, action) pair.  detach_params: gradient w.r.t. the input only.
        One fused launch each way (ops.mlp_chain)."""
        f = (lambda t: t.detach()) if detach_params else (lambda t: t)
        layers = [(f(fc.weight), f(fc.bias)) for fc in self.fc_layers] + [(f(self.out.weight), f(self.out.bias))]
        return ops.mlp_chain(q_input, layers, ("silu",) * len(self.fc_layers))


class Critic(nn.Module):
    def __init__(self, state_dim: int, goal_dim: int = 0, action_dim: int = 16, q_network: dict = {}):
        super().__init__()
        q_cfg = to_container(q_network)
        q_cfg.update({"input_dim": state_dim + goal_dim + action_dim})
        self.Q = instantiate(q_cfg)

    def forward(self, obs: torch.Tensor, action: torch.Tensor):
        if len(action.shape) == 2 and action.shape[0] == 1 and len(obs.shape) == 1:
            obs = obs.unsqueeze(0)
        return self.Q((obs, action))
