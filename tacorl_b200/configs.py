"""Python equivalents of the reference's Hydra config tree for the hot path
(/root/reference/config/module/{play_lmp_for_rl,tacorl}.yaml and config/networks/**), with `_target_`
pointing at the tacorl_b200 mirrors.  The reference's own YAML/DictConfigs work unchanged too:
`tacorl_b200.utils.config.instantiate` remaps `tacorl.*` class paths."""

_N = "tacorl_b200.networks."


def lmp_vision_encoder(latent_dim=32, hidden_dim=256):            # networks/encoder/lmp_vision_encoder.yaml
    return {"_target_": _N + "visual_encoders.encoder.LMPVisionEncoder", "latent_dim": latent_dim,
            "hidden_dim": hidden_dim, "normalize_output": False}


def lmp_encoder(latent_dim=32):                                   # networks/representation/lmp_encoder.yaml
    return {"_target_": _N + "representation.representation_network.LateFusion", "_recursive_": False,
            "networks": {"rgb_static": lmp_vision_encoder(latent_dim), "rgb_gripper": lmp_vision_encoder(latent_dim)}}


def plan_recognition(kind="tanh_net", latent_plan_dim=16, hidden_dim=2048, max_window_size=16):
    if kind == "tanh_net":                                        # networks/plan_recognition/tanh_net.yaml
        return {"_target_": _N + "plan_encoders.plan_recognition_tanh_net.PlanRecognitionTanhNetwork",
                "state_dim": None, "latent_plan_dim": latent_plan_dim, "birnn_dropout_p": 0.0, "min_std": 0.0001,
                "hidden_dim": hidden_dim}
    if kind == "default":                                         # networks/plan_recognition/default.yaml
        return {"_target_": _N + "plan_encoders.plan_recognition_net.PlanRecognitionNetwork",
                "state_dim": None, "latent_plan_dim": latent_plan_dim, "birnn_dropout_p": 0.0, "min_std": 0.0001,
                "hidden_dim": hidden_dim}
    if kind == "transformer":                                     # networks/plan_recognition/transformer.yaml
        return {"_target_": _N + "plan_encoders.plan_recognition_transformer.PlanRecognitionTransformersNetwork",
                "num_heads": 8, "num_layers": 2, "encoder_hidden_size": 2048, "fc_hidden_size": 4096,
                "state_dim": None, "latent_plan_dim": latent_plan_dim, "min_std": 0.0001, "dropout_p": 0.1,
                "encoder_normalize": False, "positional_normalize": False, "position_embedding": True,
                "max_position_embeddings": max_window_size}
    raise ValueError(kind)


def actor():                                                      # networks/actor_critic/actor/default.yaml
    return {"_target_": _N + "actor_critic.actor.Actor", "_recursive_": False,
            "policy": {"_target_": _N + "actor_critic.actor.MLPPolicy", "num_layers": 3, "hidden_dim": 256}}


def critic():                                                     # networks/actor_critic/critic/default.yaml
    return {"_target_": _N + "actor_critic.critic.Critic", "_recursive_": False,
            "q_network": {"_target_": _N + "actor_critic.critic.MLPQNetwork", "num_layers": 3, "hidden_dim": 256,
                          "last_layer_activation": "Identity"}}


def goal_encoder():                                               # networks/goal_encoder/default.yaml
    return {"_target_": _N + "visual_encoders.goal_encoder.VisualGoalEncoder", "in_features": None,
            "out_features": None, "activation_function": "ReLU", "last_layer_activation": "Identity",
            "hidden_size": 256}


def action_decoder(latent_plan_dim=16, hidden_size=2048):         # networks/action_decoder/logistic.yaml
    return {"_target_": _N + "action_decoders.action_decoder_logistic.ActionDecoderLogistic", "n_mixtures": 10,
            "num_layers": 2, "hidden_size": hidden_size, "out_features": 7, "act_max_bound": [1.0] * 7,
            "act_min_bound": [-1.0] * 7, "policy_rnn_dropout_p": 0.0, "num_classes": 10,
            "latent_plan_dim": latent_plan_dim, "rnn_model": "rnn_decoder", "include_goal": False}


def play_lmp_for_rl(pr_kind="tanh_net", modalities=("rgb_static",), latent_plan_dim=16, rnn_hidden=2048,
                    max_window_size=16, goal_modalities=None):
    """config/module/play_lmp_for_rl.yaml + experiment/play_lmp_for_rl.yaml (static camera);
    modalities=(rgb_static, rgb_gripper), goal_modalities=(rgb_static, rgb_gripper), latent_plan_dim=32 =
    experiment/play_lmp_gripper_real_world.yaml:8-15 + play_lmp_real_world.yaml:10."""
    mods = list(modalities)
    goal_mods = list(goal_modalities) if goal_modalities is not None else mods[:1]
    return {"_target_": "tacorl_b200.modules.play_lmp.play_lmp_for_rl.PlayLMP", "_recursive_": False,
            "plan_proposal": actor(), "plan_recognition": plan_recognition(pr_kind, latent_plan_dim, rnn_hidden,
                                                                           max_window_size),
            "goal_encoder": goal_encoder(), "perceptual_encoder": lmp_encoder(),
            "action_decoder": action_decoder(latent_plan_dim, rnn_hidden), "lr": 1e-4, "kl_beta": 1e-3,
            "plan_proposal_obs_modalities": mods, "plan_proposal_goal_modalities": goal_mods,
            "plan_recognition_modalities": mods, "action_decoder_modalities": mods, "real_world": True}


def tacorl(play_lmp_dir="~/tacorl/models/play_lmp"):
    """config/module/tacorl.yaml."""
    return {"_target_": "tacorl_b200.modules.tacorl.tacorl.TACORL", "_recursive_": False,
            "critic": critic(), "critic_encoder": lmp_encoder(), "finetune_action_decoder": True,
            "action_decoder_lr": 3e-4, "play_lmp_dir": play_lmp_dir, "actor_lr": 1e-4, "critic_lr": 3e-4,
            "discount": 0.95, "conservative_weight": 1.0, "reward_scale": 10.0, "n_action_samples": 4,
            "with_lagrange": True, "deterministic_backup": True, "bc_epochs": 5, "with_dr3": False,
            "with_vib": False, "real_world": True}


def actor_discrete_gripper():                                     # networks/actor_critic/actor/discrete_gripper.yaml
    cfg = actor()
    cfg["discrete_gripper"] = True
    cfg["policy"]["discrete_gripper"] = True
    return cfg


def cql_offline_goal_cond(obs_modalities=("rgb_static",), goal_modalities=("rgb_static",)):
    """config/module/cql_offline_goal_cond.yaml + experiment/cql_real_world.yaml: the flat-CQL baseline."""
    return {"_target_": "tacorl_b200.modules.cql.cql_offline_lightning.CQL_Offline", "_recursive_": False,
            "actor": actor_discrete_gripper(), "critic": critic(), "actor_encoder": lmp_encoder(),
            "critic_encoder": lmp_encoder(), "goal_encoder": goal_encoder(), "discount": 0.99, "actor_lr": 1e-4,
            "critic_lr": 3e-4, "conservative_weight": 1.0, "n_action_samples": 4, "with_lagrange": True,
            "reward_scale": 10.0, "deterministic_backup": False, "bc_epochs": 5, "with_dr3": False, "with_vib": False,
            "real_world": True, "obs_modalities": list(obs_modalities), "goal_modalities": list(goal_modalities),
            "action_dim": 7}
