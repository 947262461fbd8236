"""Mirror of TACORL(CQL_Offline): /root/reference/src/tacorl/modules/tacorl/tacorl.py:21-300 and
modules/cql/cql_offline_lightning.py:24-574 (the parts config/module/tacorl.yaml exercises:
offline CQL with Lagrange, deterministic backup, BC warm-up epochs; DR3 / VIB are off).

Same ctor kwargs, same `training_step(batch)` contract (manual optimisation, returns None, steps its own
optimisers in the reference's order), same state_dict layout.  The update is
restructured for the device (SURVEY.md §0 finding 7-i, Appendix E):
  * each (network, image) pair is encoded ONCE and the 32-float embedding is repeated over the
    n_action_samples copies, instead of pushing 4 identical image copies through the encoder;
  * gradients the reference computes and then throws away (q-network weights from the actor loss,
    log_alpha / log_alpha_prime deposits) are never computed;
  * every scalar loss (Bellman, conservative logsumexp, Lagrange, actor, alpha) comes out of two fused
    kernels, value and gradient in one pass; log values stay on the device (no host sync per metric).
"""
import copy
import math
from pathlib import Path
from typing import List

import torch
import torch.nn as nn

from ... import ops
from ...networks.actor_critic.visual_actor_wrapper import VisualActorWrapper
from ...networks.actor_critic.visual_critic_wrapper import VisualCriticWrapper
from ...optim import FlatAdam, FlatBuffer, polyak_update
from ...utils import rng
from ...utils.config import instantiate, to_container
from ...utils.distributions import TanhNormal
from ...utils.lightning import LightningModule
from ...utils.misc import set_parameter_requires_grad
from ...utils.networks import load_pl_module_from_checkpoint


class TACORL(LightningModule):
    def __init__(self, play_lmp_dir: str = "~/tacorl/models/play_lmp", lmp_epoch_to_load: int = -1,
                 overwrite_lmp_cfg: dict = {}, finetune_action_decoder: bool = False,
                 action_decoder_lr: float = 1e-4,
                 # CQL_Offline kwargs (cql_offline_lightning.py:28-61)
                 env: dict = {}, actor: dict = {}, critic: dict = {}, actor_encoder: dict = {},
                 critic_encoder: dict = {}, goal_encoder: dict = {}, transform_manager: dict = {},
                 discount: float = 0.99, tau: float = 0.005, actor_lr: float = 3e-4, critic_lr: float = 3e-4,
                 deterministic_backup: bool = False, reward_scale: float = 1.0, bc_epochs: int = 0,
                 clip_grad: bool = True, clip_grad_val: int = 1, conservative_weight: float = 1.0,
                 lagrange_thresh: float = 5.0, n_action_samples: int = 10, temp: float = 1.0,
                 with_lagrange: bool = False, with_dr3: bool = False, dr3_coefficient: float = 0.03,
                 with_vib: bool = False, vib_coefficient: float = 0.01, real_world: bool = False,
                 obs_modalities: List[str] = [], goal_modalities: List[str] = [], action_dim: int = 7,
                 play_lmp: nn.Module = None, *args, **kwargs):
        super().__init__()
        if with_dr3 or with_vib:
            raise NotImplementedError("DR3 / VIB regularisers are disabled in config/module/tacorl.yaml")
        if not deterministic_backup:
            raise NotImplementedError("config/module/tacorl.yaml uses deterministic_backup: True")
        self.play_lmp_dir = Path(play_lmp_dir).expanduser()
        self.lmp_epoch_to_load = lmp_epoch_to_load
        self.overwrite_lmp_cfg = overwrite_lmp_cfg
        self._play_lmp_module = play_lmp
        self.finetune_action_decoder = finetune_action_decoder
        self.action_decoder_lr = action_decoder_lr
        self.real_world = real_world
        self.env = None
        self.transform_manager = instantiate(transform_manager) if transform_manager else None
        self.deterministic_backup = deterministic_backup
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.critic_cfg, self.critic_encoder_cfg = critic, critic_encoder
        self.discount, self.reward_scale, self.tau = discount, reward_scale, tau
        self.bc_epochs = bc_epochs
        self.clip_grad, self.clip_grad_val = clip_grad, clip_grad_val
        self.action_dim = action_dim
        self.build_networks()
        # heuristic target entropy from the LOW-LEVEL action dim (cql_offline_lightning.py:93-98; Appendix E.11)
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

    # ------------------------------------------------------------------------------ build (tacorl.py:44-126)
    def build_networks(self):
        play_lmp = self._play_lmp_module
        if play_lmp is None:
            play_lmp = load_pl_module_from_checkpoint(self.play_lmp_dir, epoch=self.lmp_epoch_to_load,
                                                      overwrite_cfg=self.overwrite_lmp_cfg)
        self._play_lmp_module = None
        self.action_decoder = play_lmp.action_decoder
        self.perceptual_encoder = play_lmp.perceptual_encoder
        self.plan_recognition = play_lmp.plan_recognition
        self.action_decoder_modalities = play_lmp.action_decoder_modalities
        self.plan_recognition_modalities = play_lmp.plan_recognition_modalities
        mods = self.action_decoder_modalities + self.plan_recognition_modalities
        self.all_modalities = sorted(set(mods), key=mods.index)
        env_modalities = play_lmp.plan_proposal_obs_modalities
        goal_modalities = play_lmp.plan_proposal_goal_modalities
        actor = play_lmp.plan_proposal
        self.actor = VisualActorWrapper(encoder=copy.deepcopy(self.perceptual_encoder),
                                        goal_encoder=copy.deepcopy(play_lmp.goal_encoder), actor=actor,
                                        env_modalities=env_modalities, goal_modalities=goal_modalities)
        critic_cfg = to_container(self.critic_cfg)
        critic_cfg["q_network"]["num_layers"] = actor.policy.num_layers
        critic_cfg["q_network"]["hidden_dim"] = actor.policy.hidden_dim
        critic_cfg["state_dim"] = actor.state_dim
        critic_cfg["goal_dim"] = actor.goal_dim
        critic_cfg["action_dim"] = actor.action_dim
        enc_cfg = to_container(self.critic_encoder_cfg)
        em = env_modalities + goal_modalities
        enc_cfg["modalities"] = sorted(set(em), key=em.index)
        for modality, net_cfg in enc_cfg["networks"].items():
            if "latent_dim" in net_cfg and modality in self.perceptual_encoder.networks:
                net_cfg["latent_dim"] = self.perceptual_encoder.networks[modality].latent_dim

        def make_q():
            return VisualCriticWrapper(critic=instantiate(critic_cfg), encoder=instantiate(enc_cfg),
                                       goal_encoder=copy.deepcopy(play_lmp.goal_encoder),
                                       env_modalities=env_modalities, goal_modalities=goal_modalities)

        self.q1, self.q2, self.target_q1, self.target_q2 = make_q(), make_q(), make_q(), make_q()
        self.target_q1.load_state_dict(self.q1.state_dict())
        self.target_q2.load_state_dict(self.q2.state_dict())
        set_parameter_requires_grad(self.perceptual_encoder, requires_grad=False)
        set_parameter_requires_grad(self.plan_recognition, requires_grad=False)

    # ------------------------------------------------------------------------------ optimisers
    def configure_optimizers(self):                              # cql…py:553-574 + tacorl.py:289-300
        clip = float(self.clip_grad_val) if self.clip_grad else None
        req = lambda mod: [p for p in mod.parameters() if p.requires_grad]
        opts = [FlatAdam([self.log_alpha], lr=self.actor_lr),
                FlatAdam(req(self.actor), lr=self.actor_lr, max_grad_norm=clip),
                FlatAdam(req(self.q1), lr=self.critic_lr, max_grad_norm=clip),
                FlatAdam(req(self.q2), lr=self.critic_lr, max_grad_norm=clip)]
        if self.with_lagrange:
            opts.append(FlatAdam([self.log_alpha_prime], lr=self.critic_lr))
        if self.finetune_action_decoder:
            opts.append(FlatAdam(req(self.action_decoder), lr=self.action_decoder_lr))
        # Polyak pairs target.parameters() with source.parameters() by order (:229-232): same flat layout
        self._target_bufs = (FlatBuffer(list(self.target_q1.parameters())), FlatBuffer(list(self.target_q2.parameters())))
        return opts

    @staticmethod
    def soft_update_from_to(source_flat, target_buf, tau):
        polyak_update(target_buf, source_flat, tau)

    # ------------------------------------------------------------------------------ frozen LMP (tacorl.py:128-252)
    def get_emb_states(self, states, modalities: List[str] = []):
        bs, seq_len = list(states.values())[0].shape[:2]
        flat = {k: v.reshape(bs * seq_len, *v.shape[2:]) for k, v in states.items()}
        emb = self.perceptual_encoder.get_state_from_observation(observation=flat, modalities=modalities,
                                                                 cat_output=False)
        return {k: v.view(bs, seq_len, -1) for k, v in emb.items()}

    @staticmethod
    def _cat(ts):
        return ts[0] if len(ts) == 1 else torch.cat(ts, dim=-1)

    def get_pr_latent_plan(self, batch, return_emb_states=True):
        with torch.no_grad():
            # the frozen LMP runs in eval mode (tacorl.py:237-238): no dropout in the transformer plan recogniser
            self.perceptual_encoder.eval()
            self.plan_recognition.eval()
            emb_states = self.get_emb_states(batch["states"], modalities=self.all_modalities)
            pr_states = self._cat([emb_states[k] for k in self.plan_recognition_modalities])
            latent_plan = self.plan_recognition(pr_states).sample()
        return (latent_plan, emb_states) if return_emb_states else latent_plan

    def get_rl_batch(self, batch, latent_plan):                  # tacorl.py:142-179, vectorised
        obs = {k: v[:, 0] for k, v in batch["states"].items()}
        nxt = {k: v[:, -1] for k, v in batch["states"].items()}
        goal = batch["goal"]
        success = (batch["disp"] == 1).to(torch.float32).unsqueeze(-1)
        states = {"observation": obs, "goal": goal}
        next_states = {"observation": nxt, "goal": goal}
        return states, latent_plan, next_states, success, success

    def compute_action_decoder_update(self, emb_states, actions, latent_plan, optimize=True, log_type="train"):
        ad_states = self._cat([emb_states[k] for k in self.action_decoder_modalities])
        if optimize:
            opt = self.optimizers()[-1]
            action_loss = self.action_decoder.loss(latent_plan=latent_plan, perceptual_emb=ad_states[:, :-1],
                                                   actions=actions[:, :-1])
            opt.zero_grad(set_to_none=True)
            action_loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                action_loss = self.action_decoder.loss(latent_plan=latent_plan, perceptual_emb=ad_states[:, :-1],
                                                       actions=actions[:, :-1])
        self.log(f"{log_type}/action_loss", action_loss, on_step=True, on_epoch=True, sync_dist=True)

    # ------------------------------------------------------------------------------ CQL update
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

    def compute_update(self, batch, optimize: bool = True, log_type: str = "train"):
        states, plan, next_states, rewards, dones = batch
        obs, goal, nxt = states["observation"], states["goal"], next_states["observation"]
        opts = self.optimizers()
        alpha_opt, actor_opt, q1_opt, q2_opt = opts[:4]
        n = self.n_action_samples
        B, Ld = plan.shape
        log = lambda k, v: self.log(f"{log_type}/{k}", v, on_step=True)

        # ---- actor forward + alpha (cql…py:439-457)
        a_in, a_goal = self._emb(self.actor, obs, goal)
        mean, std = self.actor.actor(a_in)
        dist_a = TanhNormal(mean, std)
        curr_actions, z = dist_a.rsample_with_pretanh()
        curr_log_pi = dist_a.log_prob(curr_actions, z)
        alpha_loss, d_log_alpha = ops.cql_alpha_loss(curr_log_pi, self.log_alpha, self.target_entropy)
        if optimize:
            alpha_opt.set_grad(d_log_alpha)
            alpha_opt.step(gathered=True)        # alpha is stepped BEFORE it is read for the actor loss (:451-456)

        # ---- critic embeddings: one encoder pass per (network, image)
        q1_e, _ = self._emb(self.q1, obs, goal)
        q2_e, _ = self._emb(self.q2, obs, goal)

        # ---- actor loss (:459-466)
        if self.current_epoch < self.bc_epochs:
            plp = dist_a.log_prob(value=plan)
            actor_loss, aout = ops.CqlActorLossFn.apply(1, curr_log_pi, plp, None, self.log_alpha)
        else:
            qa1 = self._q_mlp(self.q1.critic.Q, q1_e.detach(), curr_actions, True)
            qa2 = self._q_mlp(self.q2.critic.Q, q2_e.detach(), curr_actions, True)
            actor_loss, aout = ops.CqlActorLossFn.apply(2, curr_log_pi, qa1, qa2, self.log_alpha)
        log("alpha", aout[1])

        # ---- Bellman target (:284-308), no grad
        with torch.no_grad():
            an_in, _ = self._emb(self.actor, nxt, goal, goal_emb=a_goal.detach())
            mean_n, std_n = self.actor.actor(an_in)
            next_actions, _ = TanhNormal(mean_n, std_n).sample_and_logprob()
            t1_e, _ = self._emb(self.target_q1, nxt, goal)
            t2_e, _ = self._emb(self.target_q2, nxt, goal)
            tq1 = self._q_mlp(self.target_q1.critic.Q, t1_e, next_actions)
            tq2 = self._q_mlp(self.target_q2.critic.Q, t2_e, next_actions)
            # ---- sampled actions for the conservative term (:238-282); draw order = reference's
            rand_actions = rng.uniform((n * B, Ld), -1.0, 1.0, plan.device)
            ac, zc = TanhNormal(mean.detach(), std.detach()).sample_n(n, return_pre_tanh_value=True)
            lp_curr = ops.tanh_logprob(mean.detach(), std.detach(), zc, False)
            an, zn = TanhNormal(mean_n, std_n).sample_n(n, return_pre_tanh_value=True)
            lp_next = ops.tanh_logprob(mean_n, std_n, zn, False)
            acts_all = torch.cat([plan, rand_actions, ac.reshape(n * B, Ld), an.reshape(n * B, Ld)], dim=0)
        reps = 1 + 3 * n
        q1_all = self._q_mlp(self.q1.critic.Q, q1_e.repeat(reps, 1), acts_all)
        q2_all = self._q_mlp(self.q2.critic.Q, q2_e.repeat(reps, 1), acts_all)
        rand_density = math.log(0.5 ** Ld)
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

    def training_step(self, batch, batch_idx=0):                 # tacorl.py:254-273
        latent_plan, emb_states = self.get_pr_latent_plan(batch, return_emb_states=True)
        self.compute_action_decoder_update(emb_states, batch["actions"], latent_plan,
                                           optimize=self.finetune_action_decoder, log_type="train")
        rl_batch = self.get_rl_batch(batch, latent_plan)
        self.compute_update(rl_batch, optimize=True, log_type="train")

    def validation_step(self, batch, *args, **kwargs):           # tacorl.py:275-287
        latent_plan, emb_states = self.get_pr_latent_plan(batch, return_emb_states=True)
        self.compute_action_decoder_update(emb_states, batch["actions"], latent_plan, optimize=False,
                                           log_type="validation")
        rl_batch = self.get_rl_batch(batch, latent_plan)
        with torch.no_grad():
            self.compute_update(rl_batch, optimize=False, log_type="validation")
