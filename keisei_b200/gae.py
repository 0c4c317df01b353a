"""Generalised Advantage Estimation — drop-in for `keisei.training.gae` (reference gae.py:8-296).

Four entry points with the reference's names, argument meaning and error behaviour:
`compute_gae` (1-D or 2-D), `compute_gae_padded`, `compute_gae_gpu`, `compute_gae_padded_gpu`.
CUDA tensors run the single-launch reverse-scan kernel `kb_gae_scan` (csrc/gae.cu) instead of the
reference's 2·T launches; CPU tensors (the reference's rollout buffer lives on the host) run the
same recurrence as a vectorised host loop. Outputs never carry a graph (training targets only).
"""
from __future__ import annotations

import torch

from . import _lib


def _scan_cuda(rewards, values, terminated, next_value, gamma, lam, override, lengths):
    """rewards/values (T,N) on CUDA, fp32 or fp64; returns (T,N) of values.dtype."""
    T, N = rewards.shape
    dt = values.dtype
    if dt not in (torch.float32, torch.float64):
        dt = torch.float32
    dev = values.device
    v = values.detach().to(dt).contiguous()
    r = rewards.detach().to(device=dev, dtype=dt).contiguous()
    nv = next_value.detach().to(device=dev, dtype=dt).reshape(-1)
    if nv.numel() == 1 and N != 1:
        nv = nv.expand(N)
    nv = nv.contiguous()
    if nv.numel() != N:
        raise ValueError(f"next_value has {nv.numel()} elements, expected {N}")
    term = terminated.detach().to(dev)
    if term.dtype == torch.bool:
        term_kind, term = 0, term.contiguous()
    else:
        term_kind, term = 1, term.to(dt).contiguous()
    ov = None
    if override is not None:
        ov = override.detach().to(device=dev, dtype=dt).contiguous()
    ln = None
    if lengths is not None:
        ln = lengths.detach().to(device=dev, dtype=torch.int32).contiguous()
    adv = torch.empty_like(v)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_gae_scan(
            _lib.ptr(r), _lib.ptr(v), _lib.ptr(term), term_kind, _lib.ptr(nv), _lib.ptr(ov), _lib.ptr(ln),
            _lib.ptr(adv), T, N, float(gamma), float(lam), 1 if dt == torch.float64 else 0,
            _lib.stream_ptr(dev))
    _lib.check(rc, "kb_gae_scan")
    return adv.to(values.dtype)


def _scan_host(rewards, values, terminated, next_value, gamma, lam, override, lengths):
    """Host recurrence for CPU tensors, (T,N) layout. Same rounding order as the reference."""
    T, N = rewards.shape
    dt = values.dtype
    rewards = rewards.to(dt)
    nxt = torch.empty_like(values)
    if T > 1:
        nxt[:-1] = values[1:]
    nxt[-1] = next_value
    if lengths is not None:
        last = (lengths.to(torch.long) - 1).clamp(min=0)
        nxt[last, torch.arange(N)] = next_value.to(dt)
    if override is not None:
        ov = override.to(dt)
        nxt = torch.where(torch.isnan(ov), nxt, ov)
    nd = 1.0 - terminated.float()
    delta = rewards + gamma * nxt * nd - values
    decay = gamma * lam * nd
    out = torch.empty_like(values)
    carry = torch.zeros(N, dtype=delta.dtype)
    for t in range(T - 1, -1, -1):
        carry = delta[t] + decay[t] * carry
        out[t] = carry
    return out.to(dt)


def _dispatch(rewards, values, terminated, next_value, gamma, lam, override, lengths):
    if rewards.numel() == 0:
        return torch.zeros_like(rewards, dtype=values.dtype)
    if values.is_cuda:
        return _scan_cuda(rewards, values, terminated, next_value, gamma, lam, override, lengths)
    return _scan_host(rewards, values, terminated, next_value, gamma, lam, override, lengths)


@torch.no_grad()
def compute_gae(rewards, values, terminated, next_value, gamma, lam, next_value_override=None):
    """1-D (T,) with scalar bootstrap, or 2-D (T, N) with (N,) bootstrap. Reference gae.py:8-73."""
    if rewards.ndim == 1:
        ov = None if next_value_override is None else next_value_override.reshape(-1, 1)
        nv = next_value.reshape(-1)[:1] if isinstance(next_value, torch.Tensor) else torch.tensor([next_value])
        out = _dispatch(rewards.reshape(-1, 1), values.reshape(-1, 1), terminated.reshape(-1, 1),
                        nv.to(values.device), gamma, lam, ov, None)
        return out.reshape(-1)
    if rewards.ndim != 2:
        raise ValueError(f"compute_gae supports 1D (T,) or 2D (T, N) input, got shape {tuple(rewards.shape)}")
    return _dispatch(rewards, values, terminated, next_value, gamma, lam, next_value_override, None)


@torch.no_grad()
def compute_gae_padded(rewards, values, terminated, next_values, lengths, gamma, lam,
                       next_value_override=None):
    """Padded (T_max, N) variant: bootstrap stamped at lengths[i]-1. Reference gae.py:76-148."""
    return _dispatch(rewards, values, terminated, next_values, gamma, lam, next_value_override, lengths)


@torch.no_grad()
def compute_gae_gpu(rewards, values, terminated, next_value, gamma, lam, next_value_override=None):
    """(T, N) only; 1-D input is rejected like the reference (gae.py:187-190)."""
    if rewards.ndim != 2:
        raise ValueError(f"compute_gae_gpu only supports 2D (T, N) input, got shape {rewards.shape}")
    return _dispatch(rewards, values, terminated, next_value, gamma, lam, next_value_override, None)


@torch.no_grad()
def compute_gae_padded_gpu(rewards, values, terminated, next_values, lengths, gamma, lam,
                           next_value_override=None):
    """Reference gae.py:221-296."""
    if rewards.ndim != 2:
        raise ValueError(
            f"compute_gae_padded_gpu only supports 2D (T_max, N) input, got shape {rewards.shape}")
    return _dispatch(rewards, values, terminated, next_values, gamma, lam, next_value_override, lengths)


@torch.no_grad()
def normalize_advantages_(adv: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """(A - mean) / (std_unbiased + eps) over the whole buffer, in place (katago_ppo.py:797-798)."""
    if adv.numel() <= 1:
        return adv
    if adv.is_cuda:
        if adv.dtype != torch.float32 or not adv.is_contiguous():
            raise ValueError("normalize_advantages_ needs a contiguous float32 CUDA tensor")
        with torch.cuda.device(adv.device):
            rc = _lib.load().kb_advantage_normalize(_lib.ptr(adv), adv.numel(), float(eps),
                                                    _lib.stream_ptr(adv.device))
        _lib.check(rc, "kb_advantage_normalize")
        return adv
    adv.copy_((adv - adv.mean()) / (adv.std() + eps))
    return adv
