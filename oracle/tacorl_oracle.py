"""TEST INFRASTRUCTURE — CPU restatement (plain torch, fp32/fp64) of the reference's
PlayLMP / TACO-RL training hot path (ErickRosete/tacorl, read-only at /root/reference).

This is the ORACLE the CUDA path is checked against.  It is never imported by the
product package `tacorl_b200/`; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Pinning: the reference ships no tests or golden vectors.  This
restatement is therefore pinned against the reference *itself*, imported unmodified in
the build container through `oracle/ref_loader.py`; `oracle/make_golden.py` writes the
fixtures in `tests/golden/` and `tests/test_oracle_golden.py` re-checks them without the
reference being present.  The arithmetic primitives the reference delegates to PyTorch
(`nn.Conv2d`, `nn.RNN`, `nn.Linear`, `F.softplus/silu/log_softmax`, `optim.Adam`,
`clip_grad_norm_`; setup.cfg:16 pins torch>=1.7,<1.14, un-vendored) are restated from
their documented definitions and checked against the torch in this image.

All functions are pure: parameters come in as a dict name->tensor using the
reference's `state_dict` names; every random draw is an explicit
`noise` argument (Appendix C gives the reference's draw order).

Citations are `path:line` relative to /root/reference/src/tacorl/.
"""
import math

import torch
import torch.nn.functional as F

LOG_SIG_MAX = 2.0      # networks/actor_critic/actor.py:12
LOG_SIG_MIN = -5.0     # networks/actor_critic/actor.py:13, action_decoder_logistic.py:18
MEAN_MIN = -9.0        # networks/actor_critic/actor.py:14
MEAN_MAX = 9.0         # networks/actor_critic/actor.py:15


# ----------------------------------------------------------------------------- encoder
def spatial_softargmax(y, temperature):
    """networks/visual_encoders/utils.py:39-76 (normalize=False): softmax over H*W of
    y/temperature, expected (x=column, y=row) pixel coordinates, interleaved per channel."""
    n, c, h, w = y.shape
    p = F.softmax(y.reshape(n * c, h * w) / temperature, dim=1).reshape(n, c, h, w)
    xs = torch.arange(w, dtype=y.dtype, device=y.device)
    ys = torch.arange(h, dtype=y.dtype, device=y.device)
    ex = (p.sum(dim=2) * xs).sum(dim=2)          # sum_{i,j} p[i,j] * j
    ey = (p.sum(dim=3) * ys).sum(dim=2)          # sum_{i,j} p[i,j] * i
    return torch.stack([ex, ey], dim=-1).reshape(n, 2 * c)


def lmp_encoder_convs(P, pre, x):
    """networks/visual_encoders/encoder.py:369-390 — three valid convs + ReLU."""
    y = F.relu(F.conv2d(x, P[pre + "model.0.weight"], P[pre + "model.0.bias"], stride=4))
    y = F.relu(F.conv2d(y, P[pre + "model.2.weight"], P[pre + "model.2.bias"], stride=2))
    y = F.relu(F.conv2d(y, P[pre + "model.4.weight"], P[pre + "model.4.bias"], stride=1))
    return y


def lmp_encoder(P, pre, x):
    """LMPVisionEncoder.forward, encoder.py:410-419 (vib=False, normalize_output=False,
    Dropout(p=0) is the identity)."""
    y = lmp_encoder_convs(P, pre, x)
    f = spatial_softargmax(y, P[pre + "model.6.temperature"])
    h = F.relu(F.linear(f, P[pre + "fc_layers.0.weight"], P[pre + "fc_layers.0.bias"]))
    return F.linear(h, P[pre + "fc_layers.3.weight"], P[pre + "fc_layers.3.bias"])


def goal_encoder(P, pre, x):
    """VisualGoalEncoder.forward, networks/visual_encoders/goal_encoder.py:18-33."""
    h = F.relu(F.linear(x, P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"]))
    h = F.relu(F.linear(h, P[pre + "mlp.2.weight"], P[pre + "mlp.2.bias"]))
    return F.linear(h, P[pre + "mlp.4.weight"], P[pre + "mlp.4.bias"])


# ----------------------------------------------------------------------------- MLPs
def _num_fc_layers(P, pre):
    n = 0
    while (pre + f"fc_layers.{n}.weight") in P:
        n += 1
    return n


def mlp_policy(P, pre, x):
    """MLPPolicy.forward, networks/actor_critic/actor.py:252-270 (discrete_gripper=False)."""
    for i in range(_num_fc_layers(P, pre)):
        x = F.silu(F.linear(x, P[pre + f"fc_layers.{i}.weight"], P[pre + f"fc_layers.{i}.bias"]))
    mean = torch.clamp(F.linear(x, P[pre + "fc_mean.weight"], P[pre + "fc_mean.bias"]),
                       MEAN_MIN, MEAN_MAX)
    log_std = torch.clamp(F.linear(x, P[pre + "fc_log_std.weight"], P[pre + "fc_log_std.bias"]),
                          LOG_SIG_MIN, LOG_SIG_MAX)
    return mean, log_std.exp()


def mlp_q(P, pre, q_input):
    """MLPQNetwork.forward, networks/actor_critic/critic.py:92-97 (Identity last act)."""
    x = q_input
    for i in range(_num_fc_layers(P, pre)):
        x = F.silu(F.linear(x, P[pre + f"fc_layers.{i}.weight"], P[pre + f"fc_layers.{i}.bias"]))
    return F.linear(x, P[pre + "out.weight"], P[pre + "out.bias"])


# ----------------------------------------------------------------------------- RNN
def relu_rnn_layer(x, w_ih, w_hh, b_ih, b_hh, h0=None, reverse=False):
    """One direction of one layer of nn.RNN(nonlinearity='relu', batch_first=True):
    h_t = relu(W_ih x_t + b_ih + W_hh h_{t-1} + b_hh), h_init = 0 (SURVEY Appendix A;
    used at networks/action_decoders/rnn_models.py:8-16 and
    networks/plan_encoders/plan_recognition_tanh_net.py:23-31)."""
    B, T, _ = x.shape
    h = x.new_zeros(B, w_hh.shape[0]) if h0 is None else h0
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        h = F.relu(F.linear(x[:, t], w_ih, b_ih) + F.linear(h, w_hh, b_hh))
        outs[t] = h
    return torch.stack(outs, dim=1), h


def rnn_stack(P, pre, x, num_layers, bidirectional, h0=None):
    """nn.RNN multi-layer (dropout 0).  Returns (out, h_n) with torch's h_n ordering."""
    hn = []
    for l in range(num_layers):
        dirs = ["", "_reverse"] if bidirectional else [""]
        outs = []
        for d, suf in enumerate(dirs):
            idx = l * len(dirs) + d
            o, h = relu_rnn_layer(
                x, P[pre + f"weight_ih_l{l}{suf}"], P[pre + f"weight_hh_l{l}{suf}"],
                P[pre + f"bias_ih_l{l}{suf}"], P[pre + f"bias_hh_l{l}{suf}"],
                h0=None if h0 is None else h0[idx], reverse=(d == 1))
            outs.append(o)
            hn.append(h)
        x = torch.cat(outs, dim=-1)
    return x, torch.stack(hn, dim=0)


def plan_recognition_birnn(P, pre, emb, min_std=1e-4):
    """PlanRecognitionTanhNetwork.forward / PlanRecognitionNetwork.forward,
    networks/plan_encoders/plan_recognition_tanh_net.py:39-46 (same math in
    plan_recognition_net.py:43-50): 2-layer BiRNN, last time index, mean / softplus-std."""
    out, _ = rnn_stack(P, pre + "birnn_model.", emb, 2, True)
    x = out[:, -1]
    mean = F.linear(x, P[pre + "mean_fc.weight"], P[pre + "mean_fc.bias"])
    std = F.softplus(F.linear(x, P[pre + "variance_fc.weight"], P[pre + "variance_fc.bias"])) + min_std
    return mean, std


# ----------------------------------------------------------------------------- transformer PR
def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _drop(x, mask):
    """Inverted dropout with an explicit pre-scaled keep mask (mask = keep/(1-p)); None = eval."""
    return x if mask is None else x * mask


def plan_recognition_transformer(P, pre, emb, num_heads=8, num_layers=2, min_std=1e-4,
                                 masks=None):
    """PlanRecognitionTransformersNetwork.forward,
    networks/plan_encoders/plan_recognition_transformer.py:70-105, with torch's
    nn.TransformerEncoderLayer defaults (post-norm, relu, eps 1e-5, batch_first=False).
    `masks`: None (eval / p=0) or dict of pre-scaled dropout keep-masks:
      'input' (T,B,D); per layer l: f'attn{l}' (B*heads,T,T), f'drop1_{l}' (T,B,D),
      f'ff{l}' (T,B,FF), f'drop2_{l}' (T,B,D)."""
    masks = masks or {}
    B, T, D0 = emb.shape
    D = P[pre + "position_embeddings.weight"].shape[1]
    if D != D0:  # :36-41, :72-83 zero padding up to a multiple of num_heads
        emb = torch.cat([emb, emb.new_zeros(B, T, D - D0)], dim=-1)
    x = emb + P[pre + "position_embeddings.weight"][:T].unsqueeze(0)   # :85-90
    x = x.permute(1, 0, 2)                                              # (T,B,D)
    x = _drop(x, masks.get("input"))                                    # :97
    hd = D // num_heads
    for l in range(num_layers):
        lp = pre + f"transformer_encoder.layers.{l}."
        qkv = F.linear(x, P[lp + "self_attn.in_proj_weight"], P[lp + "self_attn.in_proj_bias"])
        q, k, v = qkv.chunk(3, dim=-1)

        def heads(t):  # (T,B,D) -> (B*heads, T, hd)
            return t.reshape(T, B * num_heads, hd).transpose(0, 1)

        q, k, v = heads(q), heads(k), heads(v)
        att = torch.softmax((q / math.sqrt(hd)) @ k.transpose(1, 2), dim=-1)
        att = _drop(att, masks.get(f"attn{l}"))
        o = (att @ v).transpose(0, 1).reshape(T, B, D)
        o = F.linear(o, P[lp + "self_attn.out_proj.weight"], P[lp + "self_attn.out_proj.bias"])
        x = layer_norm(x + _drop(o, masks.get(f"drop1_{l}")), P[lp + "norm1.weight"], P[lp + "norm1.bias"])
        ff = F.relu(F.linear(x, P[lp + "linear1.weight"], P[lp + "linear1.bias"]))
        ff = _drop(ff, masks.get(f"ff{l}"))
        ff = F.linear(ff, P[lp + "linear2.weight"], P[lp + "linear2.bias"])
        x = layer_norm(x + _drop(ff, masks.get(f"drop2_{l}")), P[lp + "norm2.weight"], P[lp + "norm2.bias"])
    x = F.linear(x.permute(1, 0, 2), P[pre + "fc.weight"], P[pre + "fc.bias"])   # :99
    x = x.mean(dim=1)                                                              # :100
    mean = F.linear(x, P[pre + "mean_fc.weight"], P[pre + "mean_fc.bias"])
    std = F.softplus(F.linear(x, P[pre + "variance_fc.weight"], P[pre + "variance_fc.bias"])) + min_std
    return mean, std


# ----------------------------------------------------------------------------- distributions
def normal_log_prob(z, mean, std):
    """torch.distributions.Normal.log_prob summed over the last dim (Independent(...,1))."""
    return (-((z - mean) ** 2) / (2 * std ** 2) - std.log() - 0.5 * math.log(2 * math.pi)).sum(-1)


def normal_kl(mu_a, std_a, mu_b, std_b):
    """KL(N_a || N_b) summed over the last dim = torch.distributions.kl_divergence on
    Independent(Normal,1) — called at modules/play_lmp/play_lmp_for_rl.py:280-285."""
    var_ratio = (std_a / std_b) ** 2
    t1 = ((mu_a - mu_b) / std_b) ** 2
    return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(-1)


def balanced_kl(mu_q, std_q, mu_p, std_p, kl_alpha=0.8, kl_balancing=True):
    """PlayLMP.compute_kl_loss, play_lmp_for_rl.py:259-285 (q = posterior/recognition,
    p = prior/proposal, both the *underlying Normals* of the TanhNormals)."""
    if not kl_balancing:
        return normal_kl(mu_q, std_q, mu_p, std_p).mean()
    return (kl_alpha * normal_kl(mu_q.detach(), std_q.detach(), mu_p, std_p).mean()
            + (1 - kl_alpha) * normal_kl(mu_q, std_q, mu_p.detach(), std_p.detach()).mean())


def atanh_clamped(x):
    """utils/misc.py:297-300."""
    return 0.5 * torch.log((1 + x).clamp(min=1e-6) / (1 - x).clamp(min=1e-6))


def tanh_normal_log_prob(mean, std, value=None, pre_tanh=None):
    """TanhNormal.log_prob, utils/distributions.py:86-108.  Returns shape (..., 1)."""
    if pre_tanh is None:
        pre_tanh = atanh_clamped(torch.clamp(value, -0.999, 0.999))
    lp = normal_log_prob(pre_tanh, mean, std)
    corr = -2.0 * (math.log(2.0) - pre_tanh - F.softplus(-2.0 * pre_tanh)).sum(-1)
    return (lp + corr).unsqueeze(-1)


# ----------------------------------------------------------------------------- action decoder
def action_decoder_forward(P, pre, latent_plan, emb, h0=None, n_dist=10):
    """ActionDecoderLogistic.forward, action_decoder_logistic.py:268-300
    (include_goal=False, discrete_gripper=True, rnn_model=rnn_decoder)."""
    B, T = emb.shape[:2]
    x = torch.cat([latent_plan.unsqueeze(1).expand(-1, T, -1), emb], dim=-1)
    r, h_n = rnn_stack(P, pre + "rnn.", x, 2, False, h0)
    probs = F.linear(r, P[pre + "prob_fc.weight"], P[pre + "prob_fc.bias"])
    means = F.linear(r, P[pre + "mean_fc.weight"], P[pre + "mean_fc.bias"])
    log_scales = torch.clamp(F.linear(r, P[pre + "log_scale_fc.weight"], P[pre + "log_scale_fc.bias"]),
                             min=LOG_SIG_MIN)
    grip = F.linear(r, P[pre + "gripper_fc.weight"], P[pre + "gripper_fc.bias"])
    A = probs.shape[-1] // n_dist
    return (probs.view(B, T, A, n_dist), log_scales.view(B, T, A, n_dist),
            means.view(B, T, A, n_dist), grip, h_n)


def log_sum_exp(x):
    """utils/misc.py:289-294."""
    m = x.max(dim=-1).values
    return m + torch.log(torch.exp(x - m.unsqueeze(-1)).sum(dim=-1))


def dlm_logistic_loss(logit_probs, log_scales, means, actions, num_classes=10,
                      act_min=-1.0, act_max=1.0):
    """ActionDecoderLogistic._logistic_loss, action_decoder_logistic.py:184-235."""
    log_scales = torch.clamp(log_scales, min=LOG_SIG_MIN)
    a = actions.unsqueeze(-1).expand_as(means)
    c = a - means
    inv = torch.exp(-log_scales)
    half = (act_max - act_min) / 2.0 / (num_classes - 1)
    plus_in = inv * (c + half)
    min_in = inv * (c - half)
    cdf_delta = torch.sigmoid(plus_in) - torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    mid_in = inv * c
    log_pdf_mid = mid_in - log_scales - 2.0 * F.softplus(mid_in)
    lp = torch.where(
        a < act_min + 1e-3, log_cdf_plus,
        torch.where(a > act_max - 1e-3, log_one_minus_cdf_min,
                    torch.where(cdf_delta > 1e-5, torch.log(torch.clamp(cdf_delta, min=1e-12)),
                                log_pdf_mid - math.log((num_classes - 1) / 2))))
    lp = lp + F.log_softmax(logit_probs, dim=-1)
    return -log_sum_exp(lp).sum(dim=-1).mean()


def dlm_loss(logit_probs, log_scales, means, gripper_act, actions, gripper_alpha=1.0):
    """ActionDecoderLogistic._loss, action_decoder_logistic.py:114-133 (discrete gripper)."""
    ll = dlm_logistic_loss(logit_probs, log_scales, means, actions[:, :, :-1])
    tgt = (actions[:, :, -1] != -1).long().reshape(-1)       # -1 -> class 0, else class 1
    ce = F.cross_entropy(gripper_act.reshape(-1, 2), tgt)
    return ll + gripper_alpha * ce


def dlm_sample(logit_probs, log_scales, means, gripper_act, u1, u2,
               gripper_bounds=(-1.0, 1.0)):
    """ActionDecoderLogistic._sample, action_decoder_logistic.py:238-266.
    u1, u2 ~ U[0,1) of shape (B,T,A,n_dist) and (B,T,A) — the two torch.rand draws."""
    r1, r2 = 1e-5, 1.0 - 1e-5
    t1 = (r1 - r2) * u1 + r2
    idx = torch.argmax(logit_probs - torch.log(-torch.log(t1)), dim=-1, keepdim=True)
    ls = log_scales.gather(-1, idx).squeeze(-1)
    mu = means.gather(-1, idx).squeeze(-1)
    u = (r1 - r2) * u2 + r2
    act = mu + torch.exp(ls) * (torch.log(u) - torch.log(1.0 - u))
    gb = torch.tensor(gripper_bounds, dtype=act.dtype, device=act.device)
    grip = gb[gripper_act.argmax(dim=-1)]
    return torch.cat([act, grip.unsqueeze(-1)], dim=2)


def gripper_accuracy(pred_actions, gt_actions):
    """play_lmp_for_rl.py:165-176."""
    pg = torch.where(pred_actions[..., -1] > 0, 1.0, -1.0)
    return (gt_actions[..., -1] == pg).float().mean()


# ----------------------------------------------------------------------------- PlayLMP
def play_lmp_forward(P, batch, noise, cfg=None):
    """PlayLMP.compute_loss / training_step, modules/play_lmp/play_lmp_for_rl.py:200-257,
    307-317.  `batch` = {'states': {mod: (B,T,3,H,W)}, 'actions': (B,T,7)}.
    `noise` = {'eps_pr': (B,L) N(0,1), 'u1','u2': U[0,1) (B,T-1,6,10)/(B,T-1,6),
               'random_plan': (B,L) U(-1,1), 'u1_rp','u2_rp'} (Appendix C order;
               'dropout_masks' for the transformer).
    Returns dict of the logged scalars (+ 'total_loss' carrying the autograd graph)."""
    cfg = cfg or {}
    mods = cfg.get("modalities", ["rgb_static"])
    goal_mods = cfg.get("goal_modalities", mods[:1])
    pr_kind = cfg.get("pr_kind", "tanh_net")
    kl_beta = cfg.get("kl_beta", 1e-3)
    kl_alpha = cfg.get("kl_alpha", 0.8)
    actions = batch["actions"]
    emb = {}
    for m in mods:                                  # get_emb_states :187-198
        x = batch["states"][m]
        B, T = x.shape[:2]
        e = lmp_encoder(P, f"perceptual_encoder.networks.{m}.", x.reshape(B * T, *x.shape[2:]))
        emb[m] = e.view(B, T, -1)
    cat = torch.cat([emb[m] for m in mods], dim=-1)
    pp_state = cat[:, 0]                            # :205-207
    pp_goal = goal_encoder(P, "goal_encoder.",
                           torch.cat([emb[m][:, -1] for m in goal_mods], dim=-1))   # :208-212
    mu_p, std_p = mlp_policy(P, "plan_proposal.policy.", torch.cat([pp_state, pp_goal], dim=-1))
    if pr_kind == "transformer":
        mu_q, std_q = plan_recognition_transformer(P, "plan_recognition.", cat,
                                                   masks=noise.get("dropout_masks"))
    else:
        mu_q, std_q = plan_recognition_birnn(P, "plan_recognition.", cat)
    kl = balanced_kl(mu_q, std_q, mu_p, std_p, kl_alpha)            # :259-285
    kl_scaled = kl * kl_beta
    plan = torch.tanh(mu_q + std_q * noise["eps_pr"])               # rsample, distributions.py:110-123
    lp, ls, mu, grip, _ = action_decoder_forward(P, "action_decoder.", plan, cat[:, :-1])
    acts = actions[:, :-1]                                          # :149-155
    pred = dlm_sample(lp, ls, mu, grip, noise["u1"], noise["u2"])   # loss_and_act :76-88
    action_loss = dlm_loss(lp, ls, mu, grip, acts)
    out = {
        "kl_loss": kl, "kl_loss_scaled": kl_scaled, "action_loss": action_loss,
        "gripper_accuracy": gripper_accuracy(pred, acts),
        "total_loss": kl_scaled + action_loss,
        "mu_q": mu_q, "std_q": std_q, "mu_p": mu_p, "std_p": std_p, "emb": cat,
        "pred_actions": pred,
    }
    if "random_plan" in noise:                                      # :243-252 (logging only)
        with torch.no_grad():
            lp2, ls2, mu2, grip2, _ = action_decoder_forward(P, "action_decoder.",
                                                             noise["random_plan"], cat[:, :-1])
            pred2 = dlm_sample(lp2, ls2, mu2, grip2, noise["u1_rp"], noise["u2_rp"])
            out["random_plan_action_loss"] = dlm_loss(lp2, ls2, mu2, grip2, acts)
            out["random_plan_gripper_accuracy"] = gripper_accuracy(pred2, acts)
    return out


def draw_play_lmp_noise(B, T, latent=16, n_act=6, n_dist=10, goal_dim=32, generator=None,
                        device="cpu"):
    """The reference's RNG draws for one PlayLMP.training_step with a BiRNN recogniser,
    in its order: rsample randn → _sample rand×2 → uniform_(B,L) →
    uniform_(B,goal) [unused value] → _sample rand×2.  With `generator=None` and the same
    torch.manual_seed this reproduces the reference's stream on CPU."""
    from torch.distributions.utils import _standard_normal
    kw = dict(device=device)
    n = {}
    if generator is None:
        n["eps_pr"] = _standard_normal((B, latent), dtype=torch.float32, device=torch.device(device))
    else:
        n["eps_pr"] = torch.randn(B, latent, generator=generator, **kw)
    n["u1"] = torch.rand(B, T - 1, n_act, n_dist, generator=generator, **kw)
    n["u2"] = torch.rand(B, T - 1, n_act, generator=generator, **kw)
    n["random_plan"] = torch.empty(B, latent, **kw).uniform_(-1.0, 1.0, generator=generator)
    torch.empty(B, goal_dim, **kw).uniform_(-1.0, 1.0, generator=generator)
    n["u1_rp"] = torch.rand(B, T - 1, n_act, n_dist, generator=generator, **kw)
    n["u2_rp"] = torch.rand(B, T - 1, n_act, generator=generator, **kw)
    return n


# ----------------------------------------------------------------------------- optimiser pieces
def adam_step(params, grads, state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) as called at
    play_lmp_for_rl.py:362-368 and cql_offline_lightning.py:553-574.
    In place on `params`; `state` = {'step': int, 'm': [...], 'v': [...]}."""
    state["step"] += 1
    t = state["step"]
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    for p, g, m, v in zip(params, grads, state["m"], state["v"]):
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def new_adam_state(params):
    return {"step": 0, "m": [torch.zeros_like(p) for p in params],
            "v": [torch.zeros_like(p) for p in params]}


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (L2), called at cql_offline_lightning.py:522-537."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def polyak_update(target, source, tau):
    """CQL_Offline.soft_update_from_to, cql_offline_lightning.py:229-232."""
    for t, s in zip(target, source):
        t.copy_(t * (1.0 - tau) + s * tau)


# ----------------------------------------------------------------------------- CQL / TACO-RL
def visual_emb(P, pre, obs, goal, obs_mods=("rgb_static",), goal_mods=("rgb_static",)):
    """Visual{Actor,Critic}Wrapper.get_emb_representation,
    networks/actor_critic/visual_actor_wrapper.py:41-62 / visual_critic_wrapper.py:50-71:
    LateFusion concatenates the per-modality embeddings in the order of the modality list
    (representation_network.py:36-65); the goal embedding goes through the goal encoder.
    obs / goal: dicts modality -> image batch."""
    e = torch.cat([lmp_encoder(P, pre + f"encoder.networks.{m}.", obs[m]) for m in obs_mods], dim=-1)
    g = torch.cat([lmp_encoder(P, pre + f"encoder.networks.{m}.", goal[m]) for m in goal_mods], dim=-1)
    return torch.cat([e, goal_encoder(P, pre + "goal_encoder.", g)], dim=-1)


def q_value(P, pre, emb, action):
    """Critic.forward, networks/actor_critic/critic.py:24-30."""
    return mlp_q(P, pre + "critic.Q.", torch.cat([emb, action], dim=-1))


def tacorl_losses(P, batch, noise, cfg, epoch=0, alpha_override=None):
    """Forward values of one TACORL.training_step (modules/tacorl/tacorl.py:254-273 →
    cql_offline_lightning.py:470-516) for the CURRENT parameters P, with the α used in the
    actor loss given by `alpha_override` (the reference steps log_alpha *before* reading α,
    :451-456).  Returns the individual losses as autograd-carrying scalars plus logged
    values; the update ordering is restated in `tacorl_training_step`.

    batch: states[mod] (B,T,3,H,W), goal[mod] (B,3,H,W), actions (B,T,7), disp (B,) int64;
           cfg['modalities'] / cfg['goal_modalities'] name the views (default: rgb_static only).
    noise: plan_noise (B,L) [z = mu + std*noise for pr_dist.sample()],
           eps_actor (B,L), eps_next (B,L), rand_actions (n*B,L) in (-1,1),
           eps_curr (n,B,L), eps_nextn (n,B,L).
    """
    n = cfg.get("n_action_samples", 4)
    discount = cfg.get("discount", 0.95)
    reward_scale = cfg.get("reward_scale", 10.0)
    target_entropy = cfg.get("target_entropy", -7.0)
    gap = cfg.get("lagrange_thresh", 5.0)
    bc_epochs = cfg.get("bc_epochs", 5)
    # modality lists inherited from the PlayLMP module (tacorl.py:56-75): all four lists of the shipped experiment
    # configs are equal (experiment/play_lmp_for_rl.yaml, play_lmp_gripper_real_world.yaml:8-15) except that the
    # goal may use fewer views
    mods = list(cfg.get("modalities", ["rgb_static"]))
    goal_mods = list(cfg.get("goal_modalities", mods[:1]))
    x0 = batch["states"][mods[0]]
    B, T = x0.shape[:2]
    out = {}
    # --- get_pr_latent_plan, tacorl.py:235-252 (frozen, eval, no_grad)
    with torch.no_grad():
        emb = torch.cat([lmp_encoder(P, f"perceptual_encoder.networks.{m}.",
                                     batch["states"][m].reshape(B * T, *batch["states"][m].shape[2:])).view(B, T, -1)
                         for m in mods], dim=-1)
        mu_q, std_q = plan_recognition_birnn(P, "plan_recognition.", emb) \
            if cfg.get("pr_kind", "default") != "transformer" else \
            plan_recognition_transformer(P, "plan_recognition.", emb)
        z_q = mu_q + std_q * noise["plan_noise"]
        # `default` PR returns Independent(Normal) (plan_recognition_net.py:55): raw sample;
        # tanh_net / transformer return TanhNormal: tanh(sample) (distributions.py:125-128)
        plan = z_q if cfg.get("pr_kind", "default") == "default" else torch.tanh(z_q)
    # --- compute_action_decoder_update, tacorl.py:206-233
    lp, ls, mu, grip, _ = action_decoder_forward(P, "action_decoder.", plan, emb[:, :-1])
    out["action_loss"] = dlm_loss(lp, ls, mu, grip, batch["actions"][:, :-1])
    # --- get_rl_batch, tacorl.py:142-179
    obs = {m: batch["states"][m][:, 0] for m in mods}
    nxt = {m: batch["states"][m][:, -1] for m in mods}
    goal = batch["goal"]
    rew = (batch["disp"] == 1).to(x0.dtype).unsqueeze(-1)
    done = rew
    _ve = visual_emb
    visual_emb_ = lambda P_, pre_, o_, g_: _ve(P_, pre_, o_, g_, mods, goal_mods)
    # --- actor & alpha, cql_offline_lightning.py:439-468
    a_emb = visual_emb_(P, "actor.", obs, goal)
    mu_a, std_a = mlp_policy(P, "actor.actor.policy.", a_emb)
    z = mu_a + std_a * noise["eps_actor"]
    curr_actions = torch.tanh(z)
    curr_log_pi = tanh_normal_log_prob(mu_a, std_a, pre_tanh=z)
    out["alpha_loss"] = -(P["log_alpha"][0] * (curr_log_pi + target_entropy).detach()).mean()
    alpha = P["log_alpha"][0].exp() if alpha_override is None else alpha_override
    out["alpha"] = alpha
    q1_emb = visual_emb_(P, "q1.", obs, goal)
    q2_emb = visual_emb_(P, "q2.", obs, goal)
    if epoch < bc_epochs:
        plp = tanh_normal_log_prob(mu_a, std_a, value=plan)
        out["actor_loss"] = (alpha * curr_log_pi - plp).mean()
    else:
        qv = torch.min(q_value(P, "q1.", q1_emb, curr_actions), q_value(P, "q2.", q2_emb, curr_actions))
        out["actor_loss"] = (alpha * curr_log_pi - qv).mean()
    # --- Bellman, :284-314 (deterministic_backup=True)
    with torch.no_grad():
        an_emb = visual_emb_(P, "actor.", nxt, goal)
        mu_n, std_n = mlp_policy(P, "actor.actor.policy.", an_emb)
        next_actions = torch.tanh(mu_n + std_n * noise["eps_next"])
        tq = torch.min(q_value(P, "target_q1.", visual_emb_(P, "target_q1.", nxt, goal), next_actions),
                       q_value(P, "target_q2.", visual_emb_(P, "target_q2.", nxt, goal), next_actions))
        q_target = reward_scale * rew + (1 - done) * discount * tq
    q1_data = q_value(P, "q1.", q1_emb, plan)
    q2_data = q_value(P, "q2.", q2_emb, plan)
    out["bellman_q1_loss"] = F.mse_loss(q1_data, q_target)
    out["bellman_q2_loss"] = F.mse_loss(q2_data, q_target)
    # --- conservative, :316-406
    L = plan.shape[-1]
    rand_a = noise["rand_actions"]                                  # (n*B, L)
    rep = lambda e: e.unsqueeze(0).expand(n, *e.shape).reshape(n * B, -1)   # expand_obs, misc.py:132-153
    with torch.no_grad():
        zc = mu_a.detach() + std_a.detach() * noise["eps_curr"]             # (n,B,L)
        ac, lpc = torch.tanh(zc), tanh_normal_log_prob(mu_a.detach(), std_a.detach(), pre_tanh=zc)
        zn = mu_n + std_n * noise["eps_nextn"]
        an, lpn = torch.tanh(zn), tanh_normal_log_prob(mu_n, std_n, pre_tanh=zn)
    rand_density = math.log(0.5 ** L)
    for i, (qe, name) in enumerate(((q1_emb, "q1."), (q2_emb, "q2."))):
        qr = q_value(P, name, rep(qe), rand_a).view(n, B).t()
        qc = q_value(P, name, rep(qe), ac.reshape(n * B, L)).view(n, B).t()
        qn = q_value(P, name, rep(qe), an.reshape(n * B, L)).view(n, B).t()
        cat = torch.cat([qr - rand_density, qc - lpc.squeeze(-1).t(), qn - lpn.squeeze(-1).t()], dim=1)
        qd = q1_data if i == 0 else q2_data
        cons = torch.logsumexp(cat, dim=1).mean() - qd.mean()
        out[f"q{i+1}_data"], out[f"q{i+1}_random"], out[f"q{i+1}_policy"] = qd.mean(), qr.mean(), qc.mean()
        out[f"cons_raw_q{i+1}"] = cons
    alpha_prime = torch.clamp(P["log_alpha_prime"][0].exp(), min=0.0, max=1e6)
    out["alpha_prime"] = alpha_prime
    out["conservative_q1_loss"] = alpha_prime * (out["cons_raw_q1"] - gap)
    out["conservative_q2_loss"] = alpha_prime * (out["cons_raw_q2"] - gap)
    out["alpha_prime_loss"] = (-out["conservative_q1_loss"] - out["conservative_q2_loss"]) * 0.5
    out["q1_loss"] = out["bellman_q1_loss"] + out["conservative_q1_loss"]
    out["q2_loss"] = out["bellman_q2_loss"] + out["conservative_q2_loss"]
    out["plan"] = plan
    return out


def draw_tacorl_noise(B, latent=16, n=4, generator=None, device="cpu"):
    """The reference's RNG draws for one TACORL.training_step:
    pr_dist.sample() normal (B,L) → rsample randn (B,L) → sample normal (B,L) →
    uniform_(n*B,L) → sample_n normal (n,B,L) ×2."""
    from torch.distributions.utils import _standard_normal
    kw = dict(device=device)

    def nrm(*shape):
        return torch.empty(*shape, **kw).normal_(generator=generator)

    d = {"plan_noise": nrm(B, latent)}
    if generator is None:
        d["eps_actor"] = _standard_normal((B, latent), dtype=torch.float32, device=torch.device(device))
    else:
        d["eps_actor"] = torch.randn(B, latent, generator=generator, **kw)
    d["eps_next"] = nrm(B, latent)
    d["rand_actions"] = torch.zeros(n * B, latent, **kw).uniform_(-1.0, 1.0, generator=generator)
    d["eps_curr"] = nrm(n, B, latent)
    d["eps_nextn"] = nrm(n, B, latent)
    return d


TACORL_GROUPS = ("alpha", "actor", "q1", "q2", "alpha_prime", "decoder")


def tacorl_param_groups(P):
    """Parameter lists of the six Adams in `optimizers()` order,
    cql_offline_lightning.py:553-574 + tacorl.py:289-300 (filter requires_grad, module
    registration order = state_dict order)."""
    def grp(prefix):
        return [k for k in P if k.startswith(prefix) and P[k].dtype.is_floating_point
                and not _is_buffer(k)]
    return {
        "alpha": ["log_alpha"], "actor": grp("actor."), "q1": grp("q1."), "q2": grp("q2."),
        "alpha_prime": ["log_alpha_prime"], "decoder": grp("action_decoder."),
    }


_BUFFER_SUFFIXES = ("one_hot_embedding_eye", "ones", "gripper_bounds", "action_max_bound",
                    "action_min_bound")


def _is_buffer(name):
    return name.split(".")[-1] in _BUFFER_SUFFIXES


def new_tacorl_opt_state(P):
    groups = tacorl_param_groups(P)
    return {g: new_adam_state([P[k] for k in ks]) for g, ks in groups.items()}


def tacorl_training_step(P, opt, batch, noise, cfg=None, epoch=0):
    """One TACORL.training_step with optimisation, in the reference's order
    (tacorl.py:254-273, cql_offline_lightning.py:470-542):
      decoder Adam → α Adam → [losses with the NEW α, OLD α′] → α′ Adam →
      actor clip+Adam → q1 clip+Adam → q2 clip+Adam → Polyak(q1→target_q1, q2→target_q2).
    P: dict of leaf tensors (requires_grad for trainables); updated in place.
    Returns the logged scalars (detached)."""
    cfg = cfg or {}
    lrs = {"alpha": cfg.get("actor_lr", 1e-4), "actor": cfg.get("actor_lr", 1e-4),
           "q1": cfg.get("critic_lr", 3e-4), "q2": cfg.get("critic_lr", 3e-4),
           "alpha_prime": cfg.get("critic_lr", 3e-4), "decoder": cfg.get("action_decoder_lr", 3e-4)}
    tau = cfg.get("tau", 0.005)
    clip = cfg.get("clip_grad_val", 1.0)
    groups = tacorl_param_groups(P)

    def step(group, loss, do_clip=False):
        ps = [P[k] for k in groups[group]]
        gs = list(torch.autograd.grad(loss, ps, retain_graph=True, allow_unused=True))
        gs = [torch.zeros_like(p) if g is None else g.clone() for p, g in zip(ps, gs)]
        if do_clip:
            clip_grad_norm(gs, clip)
        with torch.no_grad():
            adam_step(ps, gs, opt[group], lrs[group])
        return gs

    out = tacorl_losses(P, batch, noise, cfg, epoch)
    logged = {"alpha_loss": out["alpha_loss"].detach().clone(),
              "action_loss": out["action_loss"].detach().clone()}
    grads = {}
    if cfg.get("finetune_action_decoder", True):
        grads["decoder"] = step("decoder", out["action_loss"])
    grads["alpha"] = step("alpha", out["alpha_loss"])
    # everything below is rebuilt with the stepped α (cql…py:451-456); the other parameters the
    # losses read are unchanged until their own step, so one more forward gives the same values
    # the reference's retained graph holds.
    out = tacorl_losses(P, batch, noise, cfg, epoch)
    for k in ("alpha", "actor_loss", "bellman_q1_loss", "bellman_q2_loss", "conservative_q1_loss",
              "conservative_q2_loss", "q1_loss", "q2_loss", "alpha_prime", "alpha_prime_loss",
              "q1_data", "q1_random", "q1_policy", "q2_data", "q2_random", "q2_policy"):
        logged[k] = out[k].detach().clone()
    grads["alpha_prime"] = step("alpha_prime", out["alpha_prime_loss"])
    grads["actor"] = step("actor", out["actor_loss"], do_clip=cfg.get("clip_grad", True))
    grads["q1"] = step("q1", out["q1_loss"], do_clip=cfg.get("clip_grad", True))
    grads["q2"] = step("q2", out["q2_loss"], do_clip=cfg.get("clip_grad", True))
    with torch.no_grad():
        for q in ("q1", "q2"):
            src = [P[k] for k in groups[q]]
            tgt = [P["target_" + k] for k in groups[q]]
            polyak_update(tgt, src, tau)
    logged["plan"] = out["plan"]
    return logged, grads


# ----------------------------------------------------------------------------- flat CQL baseline
def mlp_policy_gripper(P, pre, x):
    """MLPPolicy.forward with discrete_gripper=True, networks/actor_critic/actor.py:252-270:
    (mean, std) over the continuous dims + 2 open/close logits from the same trunk."""
    h = x
    for i in range(_num_fc_layers(P, pre)):
        h = F.silu(F.linear(h, P[pre + f"fc_layers.{i}.weight"], P[pre + f"fc_layers.{i}.bias"]))
    mean = torch.clamp(F.linear(h, P[pre + "fc_mean.weight"], P[pre + "fc_mean.bias"]), MEAN_MIN, MEAN_MAX)
    log_std = torch.clamp(F.linear(h, P[pre + "fc_log_std.weight"], P[pre + "fc_log_std.bias"]),
                          LOG_SIG_MIN, LOG_SIG_MAX)
    logits = F.linear(h, P[pre + "gripper_action.weight"], P[pre + "gripper_action.bias"])
    return mean, log_std.exp(), logits


def gumbel_argmax(logits, u, clamp):
    """Index drawn by GumbelSoftmax, utils/distributions.py:28-48 (temperature 0.5 does not move the argmax):
    sample():            argmax(normalised_logits - log(-log(u))), u ~ uniform_(0, 1)                     (clamp False)
    rsample(hard=True):  argmax of the relaxed sample = argmax(normalised_logits + gumbel(clamp_probs(u))) (clamp True;
                         torch ExpRelaxedCategorical.rsample, clamp_probs eps = float32 eps)
    `logits` of a torch Categorical-family distribution are normalised: logits - logsumexp(logits)."""
    norm = logits - logits.logsumexp(dim=-1, keepdim=True)
    if clamp:
        eps = torch.finfo(u.dtype).eps
        u = u.clamp(min=eps, max=1 - eps)
    return torch.argmax(norm - torch.log(-torch.log(u)), dim=-1)


def gripper_log_prob(logits, index):
    """GumbelSoftmax.log_prob of a class index, utils/distributions.py:50-58: log_softmax(logits)[index], keepdim."""
    onehot = F.one_hot(index.long(), logits.shape[-1]).to(logits.dtype)
    return (onehot * F.log_softmax(logits, dim=-1)).sum(dim=-1, keepdim=True)


def cql_losses(P, batch, noise, cfg, epoch=0, alpha_override=None):
    """Forward values of one CQL_Offline.training_step of the flat baseline (modules/cql/cql_offline_lightning.py:
    470-516 with config/module/cql_offline_goal_cond.yaml: discrete-gripper actor, entropy-regularised backup,
    Lagrange) for the CURRENT parameters P; alpha_override as in tacorl_losses.

    batch: observations / next_observations = {"observation": {mod: (B,3,H,W)}, "goal": {mod: (B,3,H,W)}},
           actions (B,7) with the last channel in {-1,+1}, rewards (B,), terminals (B,)
           (datamodule/dataset/goal_cond_replay_buffer_dataset.py:277-296).
    noise: eps_actor (B,6), u_actor (B,2), eps_next (B,6), u_next (B,2), rand_actions (n*B,7),
           eps_curr (n,B,6), u_curr (n,B,2), eps_nextn (n,B,6), u_nextn (n,B,2)   (draw order: draw_cql_noise)."""
    n = cfg.get("n_action_samples", 4)
    discount = cfg.get("discount", 0.99)
    reward_scale = cfg.get("reward_scale", 10.0)
    target_entropy = cfg.get("target_entropy", -7.0)
    gap = cfg.get("lagrange_thresh", 5.0)
    bc_epochs = cfg.get("bc_epochs", 5)
    det_backup = cfg.get("deterministic_backup", False)
    mods = list(cfg.get("modalities", ["rgb_static"]))
    goal_mods = list(cfg.get("goal_modalities", mods[:1]))
    obs, goal = batch["observations"]["observation"], batch["observations"]["goal"]
    nxt = batch["next_observations"]["observation"]
    actions = batch["actions"].float()
    B = actions.shape[0]
    rew = batch["rewards"].float().reshape(B, 1)
    done = batch["terminals"].int().reshape(B, 1)           # overwrite_batch, :119-148
    emb = lambda pre, o: visual_emb(P, pre, o, goal, mods, goal_mods)
    pol = "actor.actor.policy."
    out = {}
    # --- actor & alpha, :439-468 + Actor.get_actions(reparameterize=True), actor.py:66-98
    mu_a, std_a, lg_a = mlp_policy_gripper(P, pol, emb("actor.", obs))
    z = mu_a + std_a * noise["eps_actor"]
    grip_idx = gumbel_argmax(lg_a, noise["u_actor"], clamp=True)
    curr_log_pi = tanh_normal_log_prob(mu_a, std_a, pre_tanh=z) + gripper_log_prob(lg_a, grip_idx)
    curr_actions = torch.cat([torch.tanh(z), grip_idx.unsqueeze(-1).to(z.dtype) * 2.0 - 1], dim=-1)
    out["alpha_loss"] = -(P["log_alpha"][0] * (curr_log_pi + target_entropy).detach()).mean()
    alpha = P["log_alpha"][0].exp() if alpha_override is None else alpha_override
    out["alpha"] = alpha
    q1_emb, q2_emb = emb("q1.", obs), emb("q2.", obs)
    if epoch < bc_epochs:
        # Actor.log_prob, actor.py:143-156
        plp = tanh_normal_log_prob(mu_a, std_a, value=actions[..., :-1]) \
            + gripper_log_prob(lg_a, actions[..., -1] / 2 + 0.5)
        out["actor_loss"] = (alpha * curr_log_pi - plp).mean()
    else:
        qv = torch.min(q_value(P, "q1.", q1_emb, curr_actions), q_value(P, "q2.", q2_emb, curr_actions))
        out["actor_loss"] = (alpha * curr_log_pi - qv).mean()
    # --- Bellman, :284-314
    with torch.no_grad():
        mu_n, std_n, lg_n = mlp_policy_gripper(P, pol, emb("actor.", nxt))
        zn1 = mu_n + std_n * noise["eps_next"]
        gi_n = gumbel_argmax(lg_n, noise["u_next"], clamp=False)
        next_log_pi = tanh_normal_log_prob(mu_n, std_n, pre_tanh=zn1) + gripper_log_prob(lg_n, gi_n)
        next_actions = torch.cat([torch.tanh(zn1), gi_n.unsqueeze(-1).to(zn1.dtype) * 2.0 - 1], dim=-1)
        tq = torch.min(q_value(P, "target_q1.", emb("target_q1.", nxt), next_actions),
                       q_value(P, "target_q2.", emb("target_q2.", nxt), next_actions))
        if not det_backup:
            tq = tq - alpha.detach() * next_log_pi
        q_target = reward_scale * rew + (1 - done) * discount * tq
    q1_data = q_value(P, "q1.", q1_emb, actions)
    q2_data = q_value(P, "q2.", q2_emb, actions)
    out["bellman_q1_loss"] = F.mse_loss(q1_data, q_target)
    out["bellman_q2_loss"] = F.mse_loss(q2_data, q_target)
    # --- conservative, :238-282, 316-406
    A = actions.shape[-1]
    rand_a = noise["rand_actions"].clone()
    rand_a[..., -1] = torch.where(rand_a[..., -1] >= 0, 1.0, -1.0)
    rep = lambda e: e.unsqueeze(0).expand(n, *e.shape).reshape(n * B, -1)

    def sample_n(mu, std, lg, eps, u):          # Actor.sample_n_with_log_prob, actor.py:117-141
        zz = mu + std * eps
        gi = gumbel_argmax(lg.unsqueeze(0).expand(n, *lg.shape), u, clamp=False)
        lp = tanh_normal_log_prob(mu, std, pre_tanh=zz) + gripper_log_prob(lg, gi)
        return torch.cat([torch.tanh(zz), gi.unsqueeze(-1).to(zz.dtype) * 2 - 1], dim=-1), lp

    with torch.no_grad():
        ac, lpc = sample_n(mu_a.detach(), std_a.detach(), lg_a.detach(), noise["eps_curr"], noise["u_curr"])
        an, lpn = sample_n(mu_n, std_n, lg_n, noise["eps_nextn"], noise["u_nextn"])
    rand_density = math.log(0.5 ** A)
    for i, (qe, name) in enumerate(((q1_emb, "q1."), (q2_emb, "q2."))):
        qr = q_value(P, name, rep(qe), rand_a).view(n, B).t()
        qc = q_value(P, name, rep(qe), ac.reshape(n * B, A)).view(n, B).t()
        qn = q_value(P, name, rep(qe), an.reshape(n * B, A)).view(n, B).t()
        cat = torch.cat([qr - rand_density, qc - lpc.squeeze(-1).t(), qn - lpn.squeeze(-1).t()], dim=1)
        qd = q1_data if i == 0 else q2_data
        cons = torch.logsumexp(cat, dim=1).mean() - qd.mean()
        out[f"q{i+1}_data"], out[f"q{i+1}_random"], out[f"q{i+1}_policy"] = qd.mean(), qr.mean(), qc.mean()
        out[f"cons_raw_q{i+1}"] = cons
    alpha_prime = torch.clamp(P["log_alpha_prime"][0].exp(), min=0.0, max=1e6)
    out["alpha_prime"] = alpha_prime
    out["conservative_q1_loss"] = alpha_prime * (out["cons_raw_q1"] - gap)
    out["conservative_q2_loss"] = alpha_prime * (out["cons_raw_q2"] - gap)
    out["alpha_prime_loss"] = (-out["conservative_q1_loss"] - out["conservative_q2_loss"]) * 0.5
    out["q1_loss"] = out["bellman_q1_loss"] + out["conservative_q1_loss"]
    out["q2_loss"] = out["bellman_q2_loss"] + out["conservative_q2_loss"]
    return out


def draw_cql_noise(B, action_dim=7, n=4, generator=None, device="cpu"):
    """The reference's RNG draws for one flat CQL_Offline.training_step, in order:
    rsample randn (B,A-1) -> RelaxedOneHotCategorical.rsample torch.rand (B,2) -> sample normal (B,A-1) ->
    GumbelSoftmax.sample uniform_ (B,2) -> uniform_(-1,1) (n*B,A) -> [normal (n,B,A-1), uniform_ (n,B,2)] x 2."""
    from torch.distributions.utils import _standard_normal
    kw = dict(device=device)
    C = action_dim - 1

    def nrm(*shape):
        return torch.empty(*shape, **kw).normal_(generator=generator)

    def uni(*shape):
        return torch.empty(*shape, **kw).uniform_(0, 1, generator=generator)

    d = {}
    if generator is None:
        d["eps_actor"] = _standard_normal((B, C), dtype=torch.float32, device=torch.device(device))
    else:
        d["eps_actor"] = torch.randn(B, C, generator=generator, **kw)
    d["u_actor"] = torch.rand(B, 2, generator=generator, **kw)
    d["eps_next"] = nrm(B, C)
    d["u_next"] = uni(B, 2)
    d["rand_actions"] = torch.zeros(n * B, action_dim, **kw).uniform_(-1.0, 1.0, generator=generator)
    d["eps_curr"] = nrm(n, B, C)
    d["u_curr"] = uni(n, B, 2)
    d["eps_nextn"] = nrm(n, B, C)
    d["u_nextn"] = uni(n, B, 2)
    return d


CQL_NOISE_ORDER = ("eps_actor", "u_actor", "eps_next", "u_next", "rand_actions", "eps_curr", "u_curr", "eps_nextn",
                   "u_nextn")


def cql_training_step(P, opt, batch, noise, cfg=None, epoch=0):
    """One flat CQL_Offline.training_step with optimisation in the reference's order (:470-542):
    alpha Adam -> [losses with the NEW alpha, OLD alpha'] -> alpha' Adam -> actor clip+Adam -> q1 clip+Adam ->
    q2 clip+Adam -> Polyak.  Same conventions as tacorl_training_step."""
    cfg = cfg or {}
    lrs = {"alpha": cfg.get("actor_lr", 1e-4), "actor": cfg.get("actor_lr", 1e-4),
           "q1": cfg.get("critic_lr", 3e-4), "q2": cfg.get("critic_lr", 3e-4),
           "alpha_prime": cfg.get("critic_lr", 3e-4)}
    tau = cfg.get("tau", 0.005)
    clip = cfg.get("clip_grad_val", 1.0)
    groups = tacorl_param_groups(P)

    def step(group, loss, do_clip=False):
        ps = [P[k] for k in groups[group]]
        gs = list(torch.autograd.grad(loss, ps, retain_graph=True, allow_unused=True))
        gs = [torch.zeros_like(p) if g is None else g.clone() for p, g in zip(ps, gs)]
        if do_clip:
            clip_grad_norm(gs, clip)
        with torch.no_grad():
            adam_step(ps, gs, opt[group], lrs[group])
        return gs

    out = cql_losses(P, batch, noise, cfg, epoch)
    logged = {"alpha_loss": out["alpha_loss"].detach().clone()}
    grads = {"alpha": step("alpha", out["alpha_loss"])}
    out = cql_losses(P, batch, noise, cfg, epoch)
    for k in ("alpha", "actor_loss", "bellman_q1_loss", "bellman_q2_loss", "conservative_q1_loss",
              "conservative_q2_loss", "q1_loss", "q2_loss", "alpha_prime", "alpha_prime_loss",
              "q1_data", "q1_random", "q1_policy", "q2_data", "q2_random", "q2_policy"):
        logged[k] = out[k].detach().clone()
    grads["alpha_prime"] = step("alpha_prime", out["alpha_prime_loss"])
    grads["actor"] = step("actor", out["actor_loss"], do_clip=cfg.get("clip_grad", True))
    grads["q1"] = step("q1", out["q1_loss"], do_clip=cfg.get("clip_grad", True))
    grads["q2"] = step("q2", out["q2_loss"], do_clip=cfg.get("clip_grad", True))
    with torch.no_grad():
        for q in ("q1", "q2"):
            polyak_update([P["target_" + k] for k in groups[q]], [P[k] for k in groups[q]], tau)
    return logged, grads


def trainable_names(P):
    return [k for k, v in P.items() if v.dtype.is_floating_point and not _is_buffer(k)]


def play_lmp_training_step(P, opt, batch, noise, cfg=None, lr=1e-4):
    """PlayLMP.training_step + backward + the single Adam of configure_optimizers
    (play_lmp_for_rl.py:307-317, 362-368).  P leaves need requires_grad; updated in place.
    Parameters whose grad is None (e.g. the transformer's unused `layernorm`) are skipped, as
    torch.optim.Adam does.  Returns (out dict, {name: grad})."""
    names = trainable_names(P)
    out = play_lmp_forward(P, batch, noise, cfg)
    gs = torch.autograd.grad(out["total_loss"], [P[k] for k in names], allow_unused=True)
    grads = {k: g for k, g in zip(names, gs) if g is not None}
    if opt is not None:
        if "names" not in opt:
            opt["names"] = list(grads)
            opt.update(new_adam_state([P[k] for k in opt["names"]]))
        with torch.no_grad():
            adam_step([P[k] for k in opt["names"]], [grads[k] for k in opt["names"]], opt, lr)
    return out, grads


def params_from(sd, frozen_prefixes=()):
    """Clone a state_dict into oracle leaves (requires_grad on trainable float tensors)."""
    P = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if t.dtype.is_floating_point and not _is_buffer(k) and not k.startswith(tuple(frozen_prefixes)):
            t.requires_grad_(True)
        P[k] = t
    return P


TACORL_FROZEN = ("perceptual_encoder.", "plan_recognition.", "target_q1.", "target_q2.")
