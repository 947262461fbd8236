"""Mirror of /root/reference/src/tacorl/utils/distributions.py:15-58 (GumbelSoftmax), :61-153 (TanhNormal) on the
tacorl_b200 kernels, plus the diagonal Normal the reference gets from torch.distributions."""
import torch

from .. import ops
from . import rng


class DiagNormal:
    """Independent(Normal(mean, std), 1) as used by the reference (mean / stddev / sample / rsample)."""

    def __init__(self, mean, std):
        self.mean = mean
        self.stddev = std

    @property
    def loc(self):
        return self.mean

    @property
    def scale(self):
        return self.stddev

    def rsample(self):
        eps = rng.standard_normal(self.mean.shape, self.mean.device)
        z, _ = ops.tanh_rsample(self.mean, self.stddev, eps, apply_tanh=False)
        return z

    def sample(self, sample_shape=()):
        shape = tuple(sample_shape) + tuple(self.mean.shape)
        eps = rng.normal_noise(shape, self.mean.device)
        with torch.no_grad():
            z, _ = ops.tanh_rsample(self.mean.detach(), self.stddev.detach(), eps, apply_tanh=False)
        return z


class TanhNormal:
    """X = tanh(Z), Z ~ N(mean, std)."""

    def __init__(self, normal_mean, normal_std):
        self.normal_mean = normal_mean
        self.normal_std = normal_std
        self.normal = DiagNormal(normal_mean, normal_std)

    def sample_n(self, n, return_pre_tanh_value=False):          # distributions.py:78-84
        eps = rng.normal_noise((n,) + tuple(self.normal_mean.shape), self.normal_mean.device)
        with torch.no_grad():
            a, z = ops.tanh_rsample(self.normal_mean.detach(), self.normal_std.detach(), eps, True)
        return (a, z) if return_pre_tanh_value else a

    def log_prob(self, value, pre_tanh_value=None):              # distributions.py:86-108
        if pre_tanh_value is None:
            return ops.tanh_logprob(self.normal_mean, self.normal_std, value, from_value=True)
        return ops.tanh_logprob(self.normal_mean, self.normal_std, pre_tanh_value, from_value=False)

    def rsample_with_pretanh(self):                              # distributions.py:110-116
        eps = rng.standard_normal(self.normal_mean.shape, self.normal_mean.device)
        return ops.tanh_rsample(self.normal_mean, self.normal_std, eps, True)

    def rsample(self):
        return self.rsample_with_pretanh()[0]

    def sample(self):                                            # distributions.py:125-128
        eps = rng.normal_noise(self.normal_mean.shape, self.normal_mean.device)
        with torch.no_grad():
            a, _ = ops.tanh_rsample(self.normal_mean.detach(), self.normal_std.detach(), eps, True)
        return a

    def sample_and_logprob(self):                                # distributions.py:130-135
        eps = rng.normal_noise(self.normal_mean.shape, self.normal_mean.device)
        with torch.no_grad():
            a, z = ops.tanh_rsample(self.normal_mean.detach(), self.normal_std.detach(), eps, True)
        return a, self.log_prob(a, z)

    def rsample_and_logprob(self):                               # distributions.py:137-140
        a, z = self.rsample_with_pretanh()
        return a, self.log_prob(a, z)

    def rsample_logprob_and_pretanh(self):
        a, z = self.rsample_with_pretanh()
        return a, self.log_prob(a, z), z

    @property
    def mean(self):
        return torch.tanh(self.normal_mean)

    @property
    def stddev(self):
        return self.normal_std


class GumbelSoftmax:
    """The open/close gripper distribution of the discrete-gripper actor (distributions.py:15-58).  Every call site of
    the reference reduces a draw to its class index (actor.py:84-91, 128-129), so draws return the index (float 0 / 1,
    trailing dim 1); `hard` relaxed samples are not materialised."""

    def __init__(self, temperature: float = 0.5, logits=None):
        self.temperature = temperature            # (does not move the argmax of a draw)
        self.logits = logits

    def sample(self, sample_shape=()):                           # distributions.py:28-38: uniform_(0, 1), no clamp
        shape = tuple(sample_shape) + tuple(self.logits.shape)
        u = rng.uniform(shape, 0.0, 1.0, self.logits.device)
        return ops.gripper_gumbel(self.logits, u, clamp=False)

    def rsample_index(self):
        """argmax(rsample(hard=True)) (distributions.py:40-48 + actor.py:84-85): torch.rand clamped by clamp_probs."""
        u = rng.rand(tuple(self.logits.shape), self.logits.device)
        return ops.gripper_gumbel(self.logits, u, clamp=True)

    def log_prob(self, index):                                   # distributions.py:50-58
        return ops.gripper_logprob(self.logits, index)
