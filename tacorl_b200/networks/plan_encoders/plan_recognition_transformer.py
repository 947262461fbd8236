"""Mirror of PlanRecognitionTransformersNetwork,
/root/reference/src/tacorl/networks/plan_encoders/plan_recognition_transformer.py:10-105
(config/networks/plan_recognition/transformer.yaml — the plan recogniser config/module/play_lmp_for_rl.yaml
ships with).  Same ctor kwargs and state_dict keys as the reference (torch nn.TransformerEncoder naming)."""
import torch
import torch.nn as nn

from ... import ops
from ...utils import rng
from ...utils.distributions import TanhNormal
from ..layers import Linear


class _LayerNormParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class _SelfAttnParams(nn.Module):
    """nn.MultiheadAttention parameter names: in_proj_weight (3D,D), in_proj_bias, out_proj.{weight,bias}."""

    def __init__(self, dim):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = Linear(dim, dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class _EncoderLayerParams(nn.Module):
    def __init__(self, dim, ff):
        super().__init__()
        self.self_attn = _SelfAttnParams(dim)
        self.linear1 = Linear(dim, ff)
        self.linear2 = Linear(ff, dim)
        self.norm1 = _LayerNormParams(dim)
        self.norm2 = _LayerNormParams(dim)


class _EncoderParams(nn.Module):
    def __init__(self, dim, ff, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayerParams(dim, ff) for _ in range(num_layers)])


class PlanRecognitionTransformersNetwork(nn.Module):
    def __init__(self, state_dim: int, latent_plan_dim: int, num_heads: int = 8, num_layers: int = 2,
                 encoder_hidden_size: int = 2048, fc_hidden_size: int = 4096, encoder_normalize: bool = False,
                 positional_normalize: bool = False, position_embedding: bool = True,
                 max_position_embeddings: int = 16, dropout_p: float = 0.01, min_std: float = 0.0001):
        super().__init__()
        if encoder_normalize or positional_normalize or not position_embedding:
            raise NotImplementedError("kernels cover transformer.yaml: learned position embedding, no extra norms")
        self.state_dim = state_dim
        self.latent_plan_dim = latent_plan_dim
        self.padding = False
        self.hidden_size = fc_hidden_size
        self.position_embedding = position_embedding
        self.encoder_normalize = encoder_normalize
        self.positional_normalize = positional_normalize
        self.min_std = min_std
        self.num_heads = num_heads
        self.dropout_p = dropout_p
        mod = self.state_dim % num_heads
        if mod != 0:                                   # :36-41
            self.padding = True
            self.pad = num_heads - mod
            self.state_dim += self.pad
        self.position_embeddings = nn.Embedding(max_position_embeddings, self.state_dim)   # parameter container
        self.layernorm = _LayerNormParams(self.state_dim)       # unused unless positional_normalize (kept for the state_dict)
        self.transformer_encoder = _EncoderParams(self.state_dim, encoder_hidden_size, num_layers)
        self.fc = Linear(self.state_dim, fc_hidden_size)
        self.mean_fc = Linear(fc_hidden_size, latent_plan_dim)
        self.variance_fc = Linear(fc_hidden_size, latent_plan_dim)

    def forward(self, perceptual_emb: torch.Tensor) -> TanhNormal:
        B, T, _ = perceptual_emb.shape
        D, H, p, dev = self.state_dim, self.num_heads, self.dropout_p, perceptual_emb.device
        tr = self.training
        # + learned position embedding, (B,T,D) -> (T,B,D), input dropout (:85-97)
        x = ops.posemb(perceptual_emb, self.position_embeddings.weight, rng.dropout_mask((T, B, D), p, dev, tr))
        for layer in self.transformer_encoder.layers:
            sa = layer.self_attn
            qkv = ops.linear(x, sa.in_proj_weight, sa.in_proj_bias)
            o = ops.attention(qkv, H, rng.dropout_mask((B * H, T, T), p, dev, tr))
            o = sa.out_proj(o)
            x = ops.add_layernorm(x, o, layer.norm1.weight, layer.norm1.bias, rng.dropout_mask((T, B, D), p, dev, tr))
            ff = layer.linear1(x, act="relu")
            ff = ops.mask_mul(ff, rng.dropout_mask(tuple(ff.shape), p, dev, tr))
            ff = layer.linear2(ff)
            x = ops.add_layernorm(x, ff, layer.norm2.weight, layer.norm2.bias, rng.dropout_mask((T, B, D), p, dev, tr))
        y = self.fc(x.transpose(0, 1))                 # (B,T,fc_hidden)   :99
        y = ops.mean_time(y)                           # :100
        w = torch.cat([self.mean_fc.weight, self.variance_fc.weight], dim=0)
        b = torch.cat([self.mean_fc.bias, self.variance_fc.bias], dim=0)
        mean, std = ops.softplus_head(ops.linear(y, w, b), self.min_std)
        return TanhNormal(mean, std)
