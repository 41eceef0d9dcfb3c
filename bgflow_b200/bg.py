"""Generator-level host orchestration (stays PyTorch, as in the reference).

Mirrors bgflow/bg.py:13-165 (loss functionals + BoltzmannGenerator) and the pieces of the prior
the hot path touches: ``NormalDistribution.sample/energy`` (bgflow/distribution/normal.py:31-46,
75-92).  Only the flow call inside runs the hand-written kernels.
"""

import math
from collections.abc import Iterable

import torch

__all__ = ["BoltzmannGenerator", "NormalDistribution", "UniformDistribution", "unnormalized_kl_div",
           "unormalized_nll", "log_weights", "log_weights_given_latent", "effective_sample_size",
           "sampling_efficiency"]


def _pack(seq):
    if isinstance(seq, torch.Tensor):
        return (seq,)
    if isinstance(seq, Iterable):
        return tuple(seq)
    return seq


class NormalDistribution(torch.nn.Module):
    """Isotropic Gaussian prior (normal.py:17-92 without the covariance option)."""

    def __init__(self, dim, mean=None):
        super().__init__()
        self.dim = dim
        self.event_shapes = [torch.Size([dim])]
        self._has_mean = mean is not None
        if self._has_mean:
            assert len(mean.shape) == 1 and mean.shape[-1] == dim, "`mean` must be a vector of size `dim`"
            self.register_buffer("_mean", mean)
        else:
            self.register_buffer("_mean", torch.zeros(dim))

    def energy(self, x, temperature=1.0):
        if self._has_mean:
            x = x - self._mean
        x = x / (temperature ** 0.5)
        log_z = self.dim / 2 * math.log(2 * math.pi * temperature)
        return 0.5 * x.pow(2).sum(dim=-1, keepdim=True) + log_z

    def sample(self, n_samples, temperature=1.0):
        x = torch.randn(n_samples, self.dim, dtype=self._mean.dtype, device=self._mean.device)
        x = x * (temperature ** 0.5)
        if self._has_mean:
            x = x + self._mean
        return x


class UniformDistribution(torch.nn.Module):
    """Uniform prior on a box (the builder's default prior, bgflow/distribution/distributions.py:71-117)."""

    def __init__(self, low, high):
        super().__init__()
        self.register_buffer("_low", torch.as_tensor(low, dtype=torch.get_default_dtype()))
        self.register_buffer("_high", torch.as_tensor(high, dtype=torch.get_default_dtype()))
        self.dim = self._low.shape[-1]
        self.event_shapes = [torch.Size([self.dim])]

    def energy(self, x, temperature=1.0):
        inside = ((x >= self._low) & (x <= self._high)).all(dim=-1, keepdim=True)
        log_vol = torch.log(self._high - self._low).sum()
        e = torch.full_like(x[..., :1], float("inf"))
        return torch.where(inside, log_vol.expand_as(e), e)

    def sample(self, n_samples, temperature=1.0):
        u = torch.rand(n_samples, self.dim, dtype=self._low.dtype, device=self._low.device)
        return self._low + (self._high - self._low) * u


def unnormalized_kl_div(prior, flow, target, n_samples, temperature=1.0):
    """bg.py:13-17"""
    z = _pack(prior.sample(n_samples, temperature=temperature))
    *x, dlogp = flow(*z, temperature=temperature)
    return target.energy(*x, temperature=temperature) - dlogp


def unormalized_nll(prior, flow, *x, temperature=1.0):
    """bg.py:20-22"""
    *z, neg_dlogp = flow(*x, inverse=True, temperature=temperature)
    return prior.energy(*z, temperature=temperature) - neg_dlogp


def log_weights_given_latent(x, z, dlogp, prior, target, temperature=1.0, normalize=True):
    """bg.py:54-64"""
    x, z = _pack(x), _pack(z)
    logw = prior.energy(*z, temperature=temperature) + dlogp - target.energy(*x, temperature=temperature)
    if normalize:
        logw = logw - torch.logsumexp(logw, dim=0)
    return logw.view(-1)


def log_weights(*x, prior, flow, target, temperature=1.0, normalize=True):
    """bg.py:25-29"""
    *z, neg_dlogp = flow(*x, inverse=True, temperature=temperature)
    return log_weights_given_latent(x, z, -neg_dlogp, prior, target, temperature=temperature,
                                    normalize=normalize)


def effective_sample_size(log_weights):
    return torch.exp(2 * torch.logsumexp(log_weights, dim=0) - torch.logsumexp(2 * log_weights, dim=0))


def sampling_efficiency(log_weights):
    return effective_sample_size(log_weights) / len(log_weights)


class BoltzmannGenerator(torch.nn.Module):
    """prior -> flow -> target (bg.py:77-165): ``sample``, ``energy`` (NLL), ``kldiv``,
    ``log_weights``."""

    def __init__(self, prior, flow, target):
        super().__init__()
        self._prior = prior
        self._flow = flow
        self._target = target
        src = target if target is not None else prior
        self.event_shapes = getattr(src, "event_shapes", None)

    @property
    def flow(self):
        return self._flow

    @property
    def prior(self):
        return self._prior

    def sample(self, n_samples, temperature=1.0, with_latent=False, with_dlogp=False, with_energy=False,
               with_log_weights=False, with_weights=False):
        z = _pack(self._prior.sample(n_samples, temperature=temperature))
        *x, dlogp = self._flow(*z, temperature=temperature)
        results = list(x)
        if with_latent:
            results.extend(z)
        if with_dlogp:
            results.append(dlogp)
        if with_energy or with_log_weights or with_weights:
            bg_energy = self._prior.energy(*z, temperature=temperature) + dlogp
            if with_energy:
                results.append(bg_energy)
            if with_log_weights or with_weights:
                logw = bg_energy - self._target.energy(*x, temperature=temperature)
                if with_log_weights:
                    results.append(logw)
                if with_weights:
                    results.append(torch.softmax(logw, dim=0).view(-1))
        return (*results,) if len(results) > 1 else results[0]

    def energy(self, *x, temperature=1.0):
        return unormalized_nll(self._prior, self._flow, *x, temperature=temperature)

    def kldiv(self, n_samples, temperature=1.0):
        return unnormalized_kl_div(self._prior, self._flow, self._target, n_samples, temperature=temperature)

    def log_weights(self, *x, temperature=1.0, normalize=True):
        return log_weights(*x, prior=self._prior, flow=self._flow, target=self._target,
                           temperature=temperature, normalize=normalize)

    def log_weights_given_latent(self, x, z, dlogp, temperature=1.0, normalize=True):
        return log_weights_given_latent(x, z, dlogp, self._prior, self._target, temperature=temperature,
                                        normalize=normalize)

    def trigger(self, function_name):
        return self.flow.trigger(function_name)
