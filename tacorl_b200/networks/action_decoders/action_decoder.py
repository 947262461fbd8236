"""Mirror of the ActionDecoder interface, /root/reference/src/tacorl/networks/action_decoders/action_decoder.py."""
from torch import nn


class ActionDecoder(nn.Module):
    def act(self, latent_plan, perceptual_emb, latent_goal):
        raise NotImplementedError

    def loss(self, latent_plan, perceptual_emb, latent_goal, actions):
        raise NotImplementedError

    def loss_and_act(self, latent_plan, perceptual_emb, latent_goal, actions):
        raise NotImplementedError

    def clear_hidden_state(self) -> None:
        raise NotImplementedError

    def _sample(self, *args, **kwargs):
        raise NotImplementedError

    def forward(self, latent_plan, perceptual_emb, latent_goal):
        raise NotImplementedError
