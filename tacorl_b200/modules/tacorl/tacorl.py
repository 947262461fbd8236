"""Mirror of TACORL(CQL_Offline): /root/reference/src/tacorl/modules/tacorl/tacorl.py:21-300 on top of
modules/cql/cql_offline_lightning.py (config/module/tacorl.yaml: offline CQL over latent plans with Lagrange,
deterministic backup, BC warm-up epochs; DR3 / VIB are off).

Same ctor kwargs, same `training_step(batch)` contract (manual optimisation, returns None, steps its own
optimisers in the reference's order), same state_dict layout.  The CQL update itself
(`compute_update`, restructured for the device) is inherited from tacorl_b200.modules.cql.cql_offline_lightning.
"""
import copy
from pathlib import Path
from typing import List

import torch
import torch.nn as nn

from ... import ops
from ...networks.actor_critic.visual_actor_wrapper import VisualActorWrapper
from ...networks.actor_critic.visual_critic_wrapper import VisualCriticWrapper
from ...optim import FlatAdam
from ...utils.config import instantiate, to_container
from ...utils.misc import set_parameter_requires_grad
from ...utils.networks import load_pl_module_from_checkpoint
from ..cql.cql_offline_lightning import CQL_Offline


class TACORL(CQL_Offline):
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
        self._pre_init = dict(play_lmp_dir=Path(play_lmp_dir).expanduser(), lmp_epoch_to_load=lmp_epoch_to_load,
                              overwrite_lmp_cfg=overwrite_lmp_cfg, _play_lmp_module=play_lmp,
                              finetune_action_decoder=finetune_action_decoder, action_decoder_lr=action_decoder_lr)
        super().__init__(env=env, actor=actor, critic=critic, actor_encoder=actor_encoder, critic_encoder=critic_encoder,
                         goal_encoder=goal_encoder, transform_manager=transform_manager, discount=discount, tau=tau,
                         actor_lr=actor_lr, critic_lr=critic_lr, deterministic_backup=deterministic_backup,
                         reward_scale=reward_scale, bc_epochs=bc_epochs, clip_grad=clip_grad, clip_grad_val=clip_grad_val,
                         conservative_weight=conservative_weight, lagrange_thresh=lagrange_thresh,
                         n_action_samples=n_action_samples, temp=temp, with_lagrange=with_lagrange, with_dr3=with_dr3,
                         dr3_coefficient=dr3_coefficient, with_vib=with_vib, vib_coefficient=vib_coefficient,
                         real_world=real_world, obs_modalities=obs_modalities, goal_modalities=goal_modalities,
                         action_dim=action_dim)

    # ------------------------------------------------------------------------------ build (tacorl.py:44-126)
    def build_networks(self):
        # (the TACO-RL fields are set here: CQL_Offline.__init__ calls build_networks, and attributes cannot be
        # assigned on an nn.Module before its own __init__ ran)
        for k, v in self.__dict__.pop("_pre_init").items():
            setattr(self, k, v)
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
        ops.mark_frozen(self.perceptual_encoder)      # bf16 operand copies of frozen weights are cast once, not per step
        ops.mark_frozen(self.plan_recognition)

    # ------------------------------------------------------------------------------ optimisers
    def configure_optimizers(self):                              # cql…py:553-574 + tacorl.py:289-300
        opts = super().configure_optimizers()
        if self.finetune_action_decoder:
            opts.append(FlatAdam([p for p in self.action_decoder.parameters() if p.requires_grad],
                                 lr=self.action_decoder_lr))
        return opts

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
