"""Mirror of ActionDecoderLogistic,
/root/reference/src/tacorl/networks/action_decoders/action_decoder_logistic.py:21-300
(config/networks/action_decoder/logistic.yaml; discrete gripper, no goal conditioning)."""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from ... import ops
from ...utils import rng
from ..layers import Linear
from . import rnn_models
from .action_decoder import ActionDecoder

LOG_SIG_MIN = -5


class ActionDecoderLogistic(ActionDecoder):
    def __init__(self, state_dim: int = 32, goal_dim: int = 32, latent_plan_dim: int = 16,
                 hidden_size: int = 256, out_features: int = 7,
                 act_max_bound: List[float] = [1.0] * 7, act_min_bound: List[float] = [-1.0] * 7,
                 gripper_alpha: float = 1.0, policy_rnn_dropout_p: float = 0.0, num_layers: int = 2,
                 rnn_model: str = "rnn_decoder", discrete_gripper: bool = True, include_goal: bool = False,
                 num_classes: int = 10, n_mixtures: int = 10):
        super().__init__()
        if not discrete_gripper or include_goal or rnn_model != "rnn_decoder" or n_mixtures != 10:
            raise NotImplementedError("kernels cover logistic.yaml: discrete gripper, rnn_decoder, 10 mixtures, "
                                      "no goal conditioning")
        if len(set(act_max_bound[:-1])) != 1 or len(set(act_min_bound[:-1])) != 1:
            raise NotImplementedError("per-dimension action bounds are not used by the shipped configs")
        self.n_dist = n_mixtures
        self.discrete_gripper = discrete_gripper
        self.num_classes = num_classes
        self.latent_plan_dim = latent_plan_dim
        self.include_goal = include_goal
        in_features = state_dim + latent_plan_dim
        self.out_features = out_features - 1
        self.gripper_alpha = gripper_alpha
        self.rnn = getattr(rnn_models, rnn_model)(in_features, hidden_size, num_layers, policy_rnn_dropout_p)
        self.mean_fc = Linear(hidden_size, self.out_features * self.n_dist)
        self.log_scale_fc = Linear(hidden_size, self.out_features * self.n_dist)
        self.prob_fc = Linear(hidden_size, self.out_features * self.n_dist)
        self.register_buffer("one_hot_embedding_eye", torch.eye(self.n_dist))
        self.register_buffer("ones", torch.ones(1, 1, self.n_dist))
        self.register_buffer("gripper_bounds", torch.Tensor([act_min_bound[-1], act_max_bound[-1]]))
        amax = torch.Tensor(act_max_bound[:-1]).float().view(1, 1, -1, 1) * self.ones
        amin = torch.Tensor(act_min_bound[:-1]).float().view(1, 1, -1, 1) * self.ones
        self.register_buffer("action_max_bound", amax)
        self.register_buffer("action_min_bound", amin)
        self._act_min, self._act_max = float(act_min_bound[0]), float(act_max_bound[0])
        self._grip = (float(act_min_bound[-1]), float(act_max_bound[-1]))
        self.gripper_fc = Linear(hidden_size, 2)
        self.hidden_state = None

    # ---- fused internals -------------------------------------------------------------------
    def _head_params(self):
        """[prob | mean | log_scale | gripper] stacked: one (3*A*10+2, hidden) GEMM for all four heads."""
        w = torch.cat([self.prob_fc.weight, self.mean_fc.weight, self.log_scale_fc.weight, self.gripper_fc.weight], 0)
        b = torch.cat([self.prob_fc.bias, self.mean_fc.bias, self.log_scale_fc.bias, self.gripper_fc.bias], 0)
        return w, b

    def _logits_time_major(self, latent_plan, perceptual_emb, h_0=None, grad_rows=None):
        """(T*B, 3*A*10+2) head outputs, rows ordered (t, b)."""
        B, T = perceptual_emb.shape[:2]
        x = torch.cat([latent_plan.unsqueeze(1).expand(-1, T, -1), perceptual_emb], dim=-1)
        r, h_n = self.rnn.forward_time_major(x.transpose(0, 1), h_0, grad_rows=grad_rows)
        w, b = self._head_params()
        return ops.linear(r.reshape(T * B, -1), w, b), h_n, (B, T)

    @staticmethod
    def _tm(t):   # (B,T,...) -> (T*B, ...)
        return t.transpose(0, 1).reshape(t.shape[0] * t.shape[1], *t.shape[2:])

    def _draw_sample_noise(self, B, T, device):
        u1 = rng.rand((B, T, self.out_features, self.n_dist), device)   # :247
        u2 = rng.rand((B, T, self.out_features), device)                # :259
        return self._tm(u1), self._tm(u2)

    def _loss_from_logits(self, logits, actions):
        return ops.dlm_loss(logits, self._tm(actions), self.num_classes, self._act_min, self._act_max,
                            self.gripper_alpha)

    # ---- reference API ---------------------------------------------------------------------
    def clear_hidden_state(self) -> None:
        self.hidden_state = None

    def loss_and_act(self, latent_plan, perceptual_emb, actions, latent_goal=None):
        logits, _, (B, T) = self._logits_time_major(latent_plan, perceptual_emb)
        u1, u2 = self._draw_sample_noise(B, T, logits.device)
        pred, acc = ops.dlm_sample(logits, u1, u2, self._tm(actions), *self._grip)
        loss = self._loss_from_logits(logits, actions)
        self.last_gripper_accuracy = acc
        return loss, pred.view(T, B, -1).transpose(0, 1)

    def loss_and_act_two_plans(self, latent_plan, other_plan, perceptual_emb, actions, noise, other_noise):
        """One batched decoder pass for two latent plans over the same window (PlayLMP evaluates the sampled plan AND a
        random plan every step, play_lmp_for_rl.py:235-252).  The first plan carries gradients; the second is
        logging-only.  noise / other_noise = (u1, u2) drawn by the caller in the reference's order.
        Returns (loss, pred, acc), (other_loss, other_pred, other_acc)."""
        B = latent_plan.shape[0]
        plans = torch.cat([latent_plan, other_plan.detach()], dim=0)
        embs = torch.cat([perceptual_emb, perceptual_emb.detach()], dim=0)
        logits, _, (B2, T) = self._logits_time_major(plans, embs, grad_rows=B)      # the second plan is logging-only
        lg = logits.view(T, B2, -1)
        la = lg[:, :B].reshape(T * B, -1)
        lb = lg[:, B:].reshape(T * B, -1).detach()
        acts = self._tm(actions)
        pred, acc = ops.dlm_sample(la, noise[0], noise[1], acts, *self._grip)
        loss = ops.dlm_loss(la, acts, self.num_classes, self._act_min, self._act_max, self.gripper_alpha)
        with torch.no_grad():
            pred_b, acc_b = ops.dlm_sample(lb, other_noise[0], other_noise[1], acts, *self._grip)
            loss_b = ops.dlm_loss(lb, acts, self.num_classes, self._act_min, self._act_max, self.gripper_alpha)
        return ((loss, pred.view(T, B, -1).transpose(0, 1), acc),
                (loss_b, pred_b.view(T, B, -1).transpose(0, 1), acc_b))

    def act(self, latent_plan, perceptual_emb, latent_goal=None):
        with torch.no_grad():
            logits, self.hidden_state, (B, T) = self._logits_time_major(latent_plan, perceptual_emb,
                                                                          self.hidden_state)
            u1, u2 = self._draw_sample_noise(B, T, logits.device)
            pred, _ = ops.dlm_sample(logits, u1, u2, None, *self._grip)
        return pred.view(T, B, -1).transpose(0, 1)

    def loss(self, latent_plan, perceptual_emb, actions, latent_goal=None):
        logits, _, _ = self._logits_time_major(latent_plan, perceptual_emb)
        return self._loss_from_logits(logits, actions)

    def _split(self, logits, B, T):
        A, K = self.out_features, self.n_dist
        lg = logits.view(T, B, -1).transpose(0, 1)
        probs, means, ls, grip = lg[..., :A * K], lg[..., A * K:2 * A * K], lg[..., 2 * A * K:3 * A * K], lg[..., 3 * A * K:]
        return (probs.reshape(B, T, A, K), ls.reshape(B, T, A, K), means.reshape(B, T, A, K), grip)

    def forward(self, latent_plan, perceptual_emb, latent_goal=None, h_0=None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """API-parity entry (rollout / debugging): returns (logit_probs, log_scales, means, gripper_act, h_n)
        exactly like :268-300.  The training hot path uses loss / loss_and_act, which keep the heads fused."""
        logits, h_n, (B, T) = self._logits_time_major(latent_plan, perceptual_emb, h_0)
        logit_probs, log_scales, means, grip = self._split(logits, B, T)
        return logit_probs, torch.clamp(log_scales, min=LOG_SIG_MIN), means, grip, h_n

    def _join(self, logit_probs, log_scales, means, gripper_act):
        B, T = means.shape[:2]
        return self._tm(torch.cat([logit_probs.reshape(B, T, -1), means.reshape(B, T, -1),
                                   log_scales.reshape(B, T, -1), gripper_act], dim=-1)), B, T

    def _loss(self, logit_probs, log_scales, means, gripper_act, actions):
        logits, _, _ = self._join(logit_probs, log_scales, means, gripper_act)
        return self._loss_from_logits(logits, actions)

    def _sample(self, logit_probs, log_scales, means, gripper_act):
        logits, B, T = self._join(logit_probs, log_scales, means, gripper_act)
        u1, u2 = self._draw_sample_noise(B, T, logits.device)
        pred, _ = ops.dlm_sample(logits, u1, u2, None, *self._grip)
        return pred.view(T, B, -1).transpose(0, 1)
