"""TEST INFRASTRUCTURE — hand-composed equivalents of the reference's Hydra configs
(config/module/{play_lmp_for_rl,tacorl}.yaml, config/networks/**).  Pure dict builders with the
reference's `_target_` class paths; importable without the reference present."""

# --------------------------------------------------------------------------- configs
# Hand-composed from config/module/play_lmp_for_rl.yaml, config/networks/**,
# config/experiment/play_lmp_for_rl.yaml (Hydra composition is not available here).

def encoder_cfg(latent_dim=32, hidden_dim=256):
    # config/networks/encoder/lmp_vision_encoder.yaml
    return {
        "_target_": "tacorl.networks.visual_encoders.encoder.LMPVisionEncoder",
        "latent_dim": latent_dim,
        "hidden_dim": hidden_dim,
        "normalize_output": False,
    }


def representation_cfg(latent_dim=32):
    # config/networks/representation/lmp_encoder.yaml
    return {
        "_target_": "tacorl.networks.representation.representation_network.LateFusion",
        "_recursive_": False,
        "networks": {
            "rgb_static": encoder_cfg(latent_dim),
            "rgb_gripper": encoder_cfg(latent_dim),
        },
    }


def plan_recognition_cfg(kind="tanh_net", latent_plan_dim=16, hidden_dim=2048, max_window=16,
                         dropout_p=0.1):
    if kind == "tanh_net":  # config/networks/plan_recognition/tanh_net.yaml
        return {
            "_target_": "tacorl.networks.plan_encoders.plan_recognition_tanh_net."
                        "PlanRecognitionTanhNetwork",
            "state_dim": None, "latent_plan_dim": latent_plan_dim,
            "birnn_dropout_p": 0.0, "min_std": 0.0001, "hidden_dim": hidden_dim,
        }
    if kind == "default":  # config/networks/plan_recognition/default.yaml
        return {
            "_target_": "tacorl.networks.plan_encoders.plan_recognition_net.PlanRecognitionNetwork",
            "state_dim": None, "latent_plan_dim": latent_plan_dim,
            "birnn_dropout_p": 0.0, "min_std": 0.0001, "hidden_dim": hidden_dim,
        }
    if kind == "transformer":  # config/networks/plan_recognition/transformer.yaml
        return {
            "_target_": "tacorl.networks.plan_encoders.plan_recognition_transformer."
                        "PlanRecognitionTransformersNetwork",
            "num_heads": 8, "num_layers": 2, "encoder_hidden_size": 2048,
            "fc_hidden_size": 4096, "state_dim": None, "latent_plan_dim": latent_plan_dim,
            "min_std": 0.0001, "dropout_p": dropout_p, "encoder_normalize": False,
            "positional_normalize": False, "position_embedding": True,
            "max_position_embeddings": max_window,
        }
    raise ValueError(kind)


def actor_cfg():
    # config/networks/actor_critic/actor/default.yaml + policy/default.yaml
    return {
        "_target_": "tacorl.networks.actor_critic.actor.Actor",
        "_recursive_": False,
        "policy": {
            "_target_": "tacorl.networks.actor_critic.actor.MLPPolicy",
            "num_layers": 3, "hidden_dim": 256,
        },
    }


def critic_cfg():
    # config/networks/actor_critic/critic/default.yaml + q_network/default.yaml
    return {
        "_target_": "tacorl.networks.actor_critic.critic.Critic",
        "_recursive_": False,
        "q_network": {
            "_target_": "tacorl.networks.actor_critic.critic.MLPQNetwork",
            "num_layers": 3, "hidden_dim": 256, "last_layer_activation": "Identity",
        },
    }


def goal_encoder_cfg():
    # config/networks/goal_encoder/default.yaml
    return {
        "_target_": "tacorl.networks.visual_encoders.goal_encoder.VisualGoalEncoder",
        "in_features": None, "out_features": None, "activation_function": "ReLU",
        "last_layer_activation": "Identity", "hidden_size": 256,
    }


def action_decoder_cfg(latent_plan_dim=16, hidden_size=2048):
    # config/networks/action_decoder/logistic.yaml
    return {
        "_target_": "tacorl.networks.action_decoders.action_decoder_logistic.ActionDecoderLogistic",
        "n_mixtures": 10, "num_layers": 2, "hidden_size": hidden_size, "out_features": 7,
        "act_max_bound": [1.0] * 7, "act_min_bound": [-1.0] * 7,
        "policy_rnn_dropout_p": 0.0, "num_classes": 10, "latent_plan_dim": latent_plan_dim,
        "rnn_model": "rnn_decoder", "include_goal": False,
    }


def play_lmp_cfg(pr_kind="tanh_net", modalities=("rgb_static",), latent_plan_dim=16,
                 rnn_hidden=2048, max_window=16, dropout_p=0.1, goal_modalities=None):
    mods = list(modalities)
    goal_mods = list(goal_modalities) if goal_modalities is not None else mods[:1]
    return {
        "plan_proposal": actor_cfg(),
        "plan_recognition": plan_recognition_cfg(pr_kind, latent_plan_dim, rnn_hidden, max_window,
                                                 dropout_p),
        "goal_encoder": goal_encoder_cfg(),
        "perceptual_encoder": representation_cfg(),
        "action_decoder": action_decoder_cfg(latent_plan_dim, rnn_hidden),
        "transform_manager": {},
        "lr": 1e-4,
        "kl_beta": 1e-3,
        "plan_proposal_obs_modalities": mods,
        "plan_proposal_goal_modalities": goal_mods,
        "plan_recognition_modalities": mods,
        "action_decoder_modalities": mods,
        "real_world": True,
    }


def tacorl_cfg():
    # config/module/tacorl.yaml
    return {
        "critic": critic_cfg(),
        "critic_encoder": representation_cfg(),
        "transform_manager": {},
        "finetune_action_decoder": True,
        "action_decoder_lr": 3e-4,
        "play_lmp_dir": "/nonexistent",
        "actor_lr": 1e-4,
        "critic_lr": 3e-4,
        "discount": 0.95,
        "conservative_weight": 1.0,
        "reward_scale": 10.0,
        "n_action_samples": 4,
        "with_lagrange": True,
        "deterministic_backup": True,
        "bc_epochs": 5,
        "with_dr3": False,
        "with_vib": False,
        "real_world": True,
    }




def actor_discrete_gripper_cfg():
    # config/networks/actor_critic/actor/discrete_gripper.yaml + policy/discrete_gripper.yaml
    cfg = actor_cfg()
    cfg["discrete_gripper"] = True
    cfg["policy"]["discrete_gripper"] = True
    return cfg


def cql_offline_cfg():
    # config/module/cql_offline_goal_cond.yaml + experiment/cql_real_world.yaml (real_world: no simulator)
    return {
        "actor": actor_discrete_gripper_cfg(),
        "critic": critic_cfg(),
        "actor_encoder": representation_cfg(),
        "critic_encoder": representation_cfg(),
        "goal_encoder": goal_encoder_cfg(),
        "transform_manager": {},
        "discount": 0.99,
        "actor_lr": 1e-4,
        "critic_lr": 3e-4,
        "conservative_weight": 1.0,
        "n_action_samples": 4,
        "with_lagrange": True,
        "reward_scale": 10.0,
        "deterministic_backup": False,
        "bc_epochs": 5,
        "with_dr3": False,
        "with_vib": False,
        "real_world": True,
        "obs_modalities": ["rgb_static"],
        "goal_modalities": ["rgb_static"],
        "action_dim": 7,
    }
