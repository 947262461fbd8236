"""Mirror of PlanRecognitionNetwork,
/root/reference/src/tacorl/networks/plan_encoders/plan_recognition_net.py:14-56
(returns the un-squashed diagonal Normal)."""
from ...utils.distributions import DiagNormal
from .plan_recognition_tanh_net import PlanRecognitionTanhNetwork


class PlanRecognitionNetwork(PlanRecognitionTanhNetwork):
    def __init__(self, state_dim: int, latent_plan_dim: int, birnn_dropout_p: float, min_std: float,
                 hidden_dim: int = 2048):
        super().__init__(state_dim, latent_plan_dim, birnn_dropout_p, min_std, hidden_dim)

    def __call__(self, *args, **kwargs):
        mean, std = super(PlanRecognitionTanhNetwork, self).__call__(*args, **kwargs)
        return DiagNormal(mean, std)
