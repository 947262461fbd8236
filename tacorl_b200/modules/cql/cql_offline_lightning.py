"""Mirror of CQL_Offline, /root/reference/src/tacorl/modules/cql/cql_offline_lightning.py:24-574: offline conservative
Q-learning with twin visual critics, Lagrange multiplier, BC warm-up epochs and Polyak targets.  Two users:

  * the flat-CQL baseline (config/module/cql_offline_goal_cond.yaml, experiment/cql_real_world.yaml): 7-D robot actions
    from a discrete-gripper MLPPolicy (continuous arm + GumbelSoftmax open/close head), entropy-regularised backup;
  * TACORL (modules/tacorl/tacorl.py), which subclasses it with latent plans as actions and a deterministic backup.

Same ctor kwargs, same `training_step(batch, batch_idx)` contract (manual optimisation, returns None, steps its own
optimisers in the reference's order), same state_dict layout.  The update is restructured for the device (SURVEY.md
section 0 finding 7-i, Appendix E):
  * each (network, image) pair is encoded ONCE and the embedding is repeated over the n_action_samples copies, instead
    of pushing identical image copies through the encoder (`expand_obs`, utils/misc.py:132-153);
  * gradients the reference computes and then throws away (q-network weights from the actor loss, log_alpha /
    log_alpha_prime deposits) are never computed;
  * every scalar loss (Bellman, conservative logsumexp, Lagrange, actor, alpha) comes out of two fused kernels, value
    and gradient in one pass; log values stay on the device (no host sync per metric).
DR3 / VIB regularisers (off in every shipped module config) and the simulator-backed constructor path (`real_world:
False` builds a CALVIN environment only to read its modality names) are not covered."""
import math
from typing import List

import torch
import torch.nn as nn

from ... import ops
from ...networks.actor_critic.visual_actor_wrapper import VisualActorWrapper
from ...networks.actor_critic.visual_critic_wrapper import VisualCriticWrapper
from ...optim import FlatAdam, FlatBuffer, polyak_update
from ...utils import rng
from ...utils.config import instantiate, to_container
from ...utils.distributions import GumbelSoftmax, TanhNormal
from ...utils.lightning import LightningModule


class CQL_Offline(LightningModule):
    def __init__(self, env: dict = {}, actor: dict = {}, critic: dict = {}, actor_encoder: dict = {},
                 critic_encoder: dict = {}, goal_encoder: dict = {}, transform_manager: dict = {},
                 discount: float = 0.99, tau: float = 0.005, actor_lr: float = 3e-4, critic_lr: float = 3e-4,
                 deterministic_backup: bool = False, reward_scale: float = 1.0, bc_epochs: int = 0,
                 clip_grad: bool = True, clip_grad_val: int = 1, conservative_weight: float = 1.0,
                 lagrange_thresh: float = 5.0, n_action_samples: int = 10, temp: float = 1.0,
                 with_lagrange: bool = False, with_dr3: bool = False, dr3_coefficient: float = 0.03,
                 with_vib: bool = False, vib_coefficient: float = 0.01, real_world: bool = False,
                 obs_modalities: List[str] = [], goal_modalities: List[str] = [], action_dim: int = 7,
                 *args, **kwargs):
        super().__init__()
        if with_dr3 or with_vib:
            raise NotImplementedError("DR3 / VIB regularisers are disabled in every shipped module config")
        self.real_world = real_world
        self.env = None
        self.env_cfg = env
        self.transform_manager = instantiate(transform_manager) if transform_manager else None
        self.deterministic_backup = deterministic_backup
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.actor_cfg, self.critic_cfg = actor, critic
        self.actor_encoder_cfg, self.critic_encoder_cfg = actor_encoder, critic_encoder
        self.goal_encoder_cfg = goal_encoder
        self.discount, self.reward_scale, self.tau = discount, reward_scale, tau
        self.bc_epochs = bc_epochs
        self.clip_grad, self.clip_grad_val = clip_grad, clip_grad_val
        self.obs_modalities, self.goal_modalities = list(obs_modalities), list(goal_modalities)
        self.action_dim = action_dim
        self.build_networks()
        # heuristic target entropy from the robot action dim (:93-98)
        self.target_entropy = -float(action_dim)
        self.log_alpha = nn.Parameter(torch.zeros(1), requires_grad=True)
        self.conservative_weight = conservative_weight
        self.n_action_samples = n_action_samples
        self.temp = temp
        self.with_lagrange = with_lagrange
        if with_lagrange:
            self.target_action_gap = lagrange_thresh
            self.log_alpha_prime = nn.Parameter(torch.zeros(1), requires_grad=True)
        self.automatic_optimization = False
        self._target_bufs = None
        self.save_hyperparameters()

    # ------------------------------------------------------------------------------ build (:150-227)
    def build_networks(self):
        if not self.real_world:
            raise NotImplementedError("real_world=False reads the modality names from a CALVIN simulator instance "
                                      "(utils/gym_utils.py); pass real_world=True with obs_modalities / goal_modalities "
                                      "/ action_dim as config/experiment/cql_real_world.yaml does")
        env_modalities, goal_modalities, action_dim = self.obs_modalities, self.goal_modalities, self.action_dim
        all_modalities = list(dict.fromkeys(env_modalities + goal_modalities))
        actor_encoder_cfg = to_container(self.actor_encoder_cfg)
        actor_encoder_cfg["modalities"] = all_modalities
        actor_encoder = instantiate(actor_encoder_cfg)
        state_dim = actor_encoder.calc_state_dim(modalities=env_modalities)
        goal_dim = actor_encoder.calc_state_dim(modalities=goal_modalities)
        goal_encoder_cfg = to_container(self.goal_encoder_cfg)
        goal_encoder_cfg["in_features"] = goal_dim
        goal_encoder_cfg["out_features"] = goal_dim
        update = {"state_dim": state_dim, "goal_dim": goal_dim, "action_dim": action_dim}
        actor_cfg = to_container(self.actor_cfg)
        actor_cfg.update(update)
        critic_cfg = to_container(self.critic_cfg)
        critic_cfg.update(update)
        self.actor = VisualActorWrapper(actor=instantiate(actor_cfg), encoder=actor_encoder,
                                        goal_encoder=instantiate(goal_encoder_cfg), env_modalities=env_modalities,
                                        goal_modalities=goal_modalities)
        critic_encoder_cfg = to_container(self.critic_encoder_cfg)
        critic_encoder_cfg["modalities"] = all_modalities

        def make_q():
            return VisualCriticWrapper(critic=instantiate(critic_cfg), encoder=instantiate(critic_encoder_cfg),
                                       goal_encoder=instantiate(goal_encoder_cfg), env_modalities=env_modalities,
                                       goal_modalities=goal_modalities)

        self.q1, self.q2, self.target_q1, self.target_q2 = make_q(), make_q(), make_q(), make_q()
        self.target_q1.load_state_dict(self.q1.state_dict())
        self.target_q2.load_state_dict(self.q2.state_dict())

    # ------------------------------------------------------------------------------ optimisers (:553-574)
    def configure_optimizers(self):
        clip = float(self.clip_grad_val) if self.clip_grad else None
        req = lambda mod: [p for p in mod.parameters() if p.requires_grad]
        opts = [FlatAdam([self.log_alpha], lr=self.actor_lr),
                FlatAdam(req(self.actor), lr=self.actor_lr, max_grad_norm=clip),
                FlatAdam(req(self.q1), lr=self.critic_lr, max_grad_norm=clip),
                FlatAdam(req(self.q2), lr=self.critic_lr, max_grad_norm=clip)]
        if self.with_lagrange:
            opts.append(FlatAdam([self.log_alpha_prime], lr=self.critic_lr))
        # Polyak pairs target.parameters() with source.parameters() by order (:229-232): same flat layout
        self._target_bufs = (FlatBuffer(list(self.target_q1.parameters())), FlatBuffer(list(self.target_q2.parameters())))
        return opts

    @staticmethod
    def soft_update_from_to(source_flat, target_buf, tau):
        polyak_update(target_buf, source_flat, tau)

    def overwrite_batch(self, batch):                            # :119-148
        rew, dones = batch["rewards"], batch["terminals"]
        dev = self.device
        rew = rew.reshape(-1, 1).to(device=dev, dtype=torch.float32)
        dones = dones.reshape(-1, 1).to(dev).int().to(torch.float32)
        return (batch["observations"], batch["actions"].to(device=dev, dtype=torch.float32), batch["next_observations"],
                rew, dones)

    # ------------------------------------------------------------------------------ embeddings
    def _emb(self, wrapper, obs_img, goal_img, goal_emb=None):
        """Visual*Wrapper.get_emb_representation with the goal embedding optionally re-used.  Observation and goal
        frames of a view go through that view's encoder as ONE batch (same weights, independent frames: identical
        values; half the kernel launches of these small 64-frame passes, one weight-gradient pass instead of two)."""
        enc = wrapper.encoder
        if goal_emb is not None:
            return torch.cat([enc.get_state_from_observation(obs_img, modalities=wrapper.env_modalities), goal_emb], dim=-1), goal_emb
        B = next(iter(obs_img.values())).shape[0]
        parts = {}
        for m in dict.fromkeys(list(wrapper.env_modalities) + list(wrapper.goal_modalities)):
            srcs = ([obs_img[m]] if m in wrapper.env_modalities else []) + ([goal_img[m]] if m in wrapper.goal_modalities else [])
            if len(srcs) == 2 and srcs[0].shape == srcs[1].shape and srcs[0].dtype == srcs[1].dtype:
                out = enc.networks[m](self._stack_frames(srcs))
                parts[(m, "obs")], parts[(m, "goal")] = out[:B], out[B:]
            else:
                if m in wrapper.env_modalities:
                    parts[(m, "obs")] = enc.networks[m](obs_img[m])
                if m in wrapper.goal_modalities:
                    parts[(m, "goal")] = enc.networks[m](goal_img[m])
        cat = lambda ts: ts[0] if len(ts) == 1 else torch.cat(ts, dim=-1)
        e = cat([parts[(m, "obs")] for m in wrapper.env_modalities])
        g = cat([parts[(m, "goal")] for m in wrapper.goal_modalities])
        goal_emb = wrapper.goal_encoder(g) if wrapper.goal_encoder is not None else g
        return torch.cat([e, goal_emb], dim=-1), goal_emb

    @staticmethod
    def _stack_frames(srcs):
        """cat along the batch dim.  uint8 frames are copied through an int32 view: torch's byte-wise cat / strided copy
        kernels run at 0.2-0.4 TB/s (78 us per 15 MB pair of 64-frame batches, measured), 4-byte elements at > 2 TB/s."""
        n = sum(t.shape[0] for t in srcs)
        out = torch.empty((n,) + tuple(srcs[0].shape[1:]), device=srcs[0].device, dtype=srcs[0].dtype)
        wide = srcs[0].dtype == torch.uint8 and srcs[0].shape[-1] % 4 == 0
        o = 0
        for t in srcs:
            dst = out[o:o + t.shape[0]]
            if wide and t.stride(-1) == 1 and all(st % 4 == 0 for st in t.stride()[:-1]) and t.storage_offset() % 4 == 0:
                dst.view(torch.int32).copy_(t.view(torch.int32))
            else:
                dst.copy_(t)
            o += t.shape[0]
        return out

    @staticmethod
    def _q_mlp(qnet, emb, action, detach_params=False):
        """MLPQNetwork.forward on cat(emb, action) (critic.py:24-30, 92-97), one fused launch each way;
        detach_params: gradient w.r.t. the input only."""
        return qnet((emb, action), detach_params=detach_params)

    # ------------------------------------------------------------------------------ policy heads
    # Actor.get_actions / sample_n_with_log_prob / log_prob (actor.py:66-156) on an already computed policy output,
    # so that one policy forward serves the actor loss, the BC term and the conservative samples, as the values are
    # identical to the reference's repeated forwards.
    @staticmethod
    def _policy_rsample(head):
        dist = TanhNormal(head[0], head[1])
        actions, z = dist.rsample_with_pretanh()
        log_pi = dist.log_prob(actions, z)
        if len(head) == 3:
            grip = GumbelSoftmax(temperature=0.5, logits=head[2])
            index = grip.rsample_index()
            log_pi = log_pi + grip.log_prob(index)
            actions = torch.cat((actions, index * 2.0 - 1), dim=-1)
        return actions, log_pi

    @staticmethod
    def _policy_sample(head):
        actions, log_pi = TanhNormal(head[0], head[1]).sample_and_logprob()
        if len(head) == 3:
            grip = GumbelSoftmax(temperature=0.5, logits=head[2])
            index = grip.sample()
            log_pi = log_pi + grip.log_prob(index)
            actions = torch.cat((actions, index * 2.0 - 1), dim=-1)
        return actions, log_pi

    @staticmethod
    def _policy_sample_n(head, n):
        mean, std = head[0], head[1]
        actions, z = TanhNormal(mean, std).sample_n(n, return_pre_tanh_value=True)
        log_pi = ops.tanh_logprob(mean, std, z, False)
        if len(head) == 3:
            grip = GumbelSoftmax(temperature=0.5, logits=head[2])
            index = grip.sample((n,))
            log_pi = log_pi + grip.log_prob(index)
            actions = torch.cat((actions, index * 2 - 1), dim=-1)
        return actions, log_pi

    @staticmethod
    def _policy_log_prob(head, actions):
        if len(head) == 3:
            log_pi = TanhNormal(head[0], head[1]).log_prob(value=actions[..., :-1])
            return log_pi + GumbelSoftmax(temperature=0.5, logits=head[2]).log_prob(actions[..., -1:] / 2 + 0.5)
        return TanhNormal(head[0], head[1]).log_prob(value=actions)

    # ------------------------------------------------------------------------------ the update (:284-542)
    def compute_update(self, batch, optimize: bool = True, log_type: str = "train"):
        states, data_actions, next_states, rewards, dones = batch
        obs, goal, nxt = states["observation"], states["goal"], next_states["observation"]
        opts = self.optimizers()
        alpha_opt, actor_opt, q1_opt, q2_opt = opts[:4]
        n = self.n_action_samples
        B, A = data_actions.shape
        log = lambda k, v: self.log(f"{log_type}/{k}", v, on_step=True)
        detach = lambda head: tuple(t.detach() for t in head)

        # ---- actor forward + alpha (:439-457)
        a_in, a_goal = self._emb(self.actor, obs, goal)
        head = self.actor.actor(a_in)
        curr_actions, curr_log_pi = self._policy_rsample(head)
        alpha_loss, d_log_alpha = ops.cql_alpha_loss(curr_log_pi, self.log_alpha, self.target_entropy)
        if optimize:
            alpha_opt.set_grad(d_log_alpha)
            alpha_opt.step(gathered=True)        # alpha is stepped BEFORE it is read for the actor loss (:451-456)

        # ---- critic embeddings: one encoder pass per (network, image)
        q1_e, _ = self._emb(self.q1, obs, goal)
        q2_e, _ = self._emb(self.q2, obs, goal)

        # ---- actor loss (:459-466)
        if self.current_epoch < self.bc_epochs:
            plp = self._policy_log_prob(head, data_actions)
            actor_loss, aout = ops.CqlActorLossFn.apply(1, curr_log_pi, plp, None, self.log_alpha)
        else:
            qa1 = self._q_mlp(self.q1.critic.Q, q1_e.detach(), curr_actions, True)
            qa2 = self._q_mlp(self.q2.critic.Q, q2_e.detach(), curr_actions, True)
            actor_loss, aout = ops.CqlActorLossFn.apply(2, curr_log_pi, qa1, qa2, self.log_alpha)
        log("alpha", aout[1])

        # ---- Bellman target (:284-308), no grad
        with torch.no_grad():
            an_in, _ = self._emb(self.actor, nxt, goal, goal_emb=a_goal.detach())
            head_n = self.actor.actor(an_in)
            next_actions, next_log_pi = self._policy_sample(head_n)
            t1_e, _ = self._emb(self.target_q1, nxt, goal)
            t2_e, _ = self._emb(self.target_q2, nxt, goal)
            tq1 = self._q_mlp(self.target_q1.critic.Q, t1_e, next_actions)
            tq2 = self._q_mlp(self.target_q2.critic.Q, t2_e, next_actions)
            if not self.deterministic_backup:    # min(Q'1, Q'2) - alpha * log pi(a'|s'), with the stepped alpha (:301-303)
                entropy_term = self.log_alpha.detach().exp() * next_log_pi
                tq1, tq2 = tq1 - entropy_term, tq2 - entropy_term
            # ---- sampled actions for the conservative term (:238-282); draw order = reference's
            rand_actions = rng.uniform((n * B, A), -1.0, 1.0, data_actions.device)
            if self.actor.discrete_gripper:
                rand_actions[..., -1] = torch.where(rand_actions[..., -1] >= 0, 1.0, -1.0)
            ac, lp_curr = self._policy_sample_n(detach(head), n)
            an, lp_next = self._policy_sample_n(head_n, n)
            acts_all = torch.cat([data_actions, rand_actions, ac.reshape(n * B, A), an.reshape(n * B, A)], dim=0)
        reps = 1 + 3 * n
        q1_all = self._q_mlp(self.q1.critic.Q, q1_e.repeat(reps, 1), acts_all)
        q2_all = self._q_mlp(self.q2.critic.Q, q2_e.repeat(reps, 1), acts_all)
        rand_density = math.log(0.5 ** A)
        q1_loss, q2_loss, scal, d_lap = ops.CqlCriticLossFn.apply(
            q1_all, q2_all, lp_curr, lp_next, tq1, tq2, rewards * 1.0, dones * 1.0,
            self.log_alpha_prime if self.with_lagrange else None, n, rand_density, self.discount, self.reward_scale,
            self.target_action_gap if self.with_lagrange else 0.0, self.conservative_weight, self.temp,
            self.with_lagrange)
        for i, k in enumerate(ops.CQL_SCALARS):
            if self.with_lagrange or k not in ("alpha_prime", "alpha_prime_loss"):
                log(k, scal[i])
        log("actor_loss", actor_loss)
        log("alpha_loss", alpha_loss[0])

        if not optimize:
            return
        if self.with_lagrange:                   # alpha' steps from alpha_prime_loss alone (:400-404)
            opts[4].set_grad(d_lap)
            opts[4].step(gathered=True)
        for o in (actor_opt, q1_opt, q2_opt):
            o.zero_grad(set_to_none=True)
        # actor, q1, q2 parameter sets are disjoint and every loss was built from pre-step values, so one
        # backward pass yields the three gradients the reference obtains from three retained passes (:519-538)
        torch.autograd.backward([actor_loss, q1_loss, q2_loss])
        actor_opt.step()
        q1_opt.step()
        q2_opt.step()
        self.soft_update_from_to(q1_opt.flat_params, self._target_bufs[0], self.tau)    # :541-542
        self.soft_update_from_to(q2_opt.flat_params, self._target_bufs[1], self.tau)

    def training_step(self, batch, batch_idx=0):                 # :544-551
        self.compute_update(self.overwrite_batch(batch), optimize=True, log_type="train")

    def validation_step(self, batch, *args, **kwargs):           # :234-236
        with torch.no_grad():
            self.compute_update(self.overwrite_batch(batch), optimize=False, log_type="validation")
