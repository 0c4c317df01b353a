"""CPU oracle for the keisei hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import this module, and only as the checker / CPU baseline. Nothing under `keisei_b200/`
imports it.

What it is: a CPU restatement of the reference's algorithm for the path named by
BASELINE.json `north_star`, each function citing the reference file:line it follows. The
arithmetic of the reference lives in PyTorch (third-party, not vendored; `uv.lock` pins
torch 2.11.0, the same version installed here), so the floating-point pieces are restated with
plain fp32 `torch.nn.functional` calls on CPU (a torch fp32 reference is the sanctioned oracle
for a floating-point kernel); the integer / scan pieces (GAE, masking, action indexing) are
restated in numpy with explicit loops.

Pinning: `oracle/make_golden.py` imports the REAL reference from /root/reference in the build
container, runs it on seeded inputs and stores input/output vectors under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against every one of them, and against the
reference's own known-answer vectors (tests/test_gae.py:10-96, tests/test_katago_ppo.py:449-489,
tests/test_se_resnet.py:173-219). Parity is therefore PINNED for: model forward (train/eval),
BN running-stat updates, loss values, parameter gradients, GAE (all four variants),
advantage normalisation, rollout log-prob semantics, value adapters.
"""
from __future__ import annotations

import math
from typing import Mapping

import numpy as np
import torch
import torch.nn.functional as F

def _f(x: torch.Tensor) -> torch.Tensor:
    """The reference's `.float()` (bf16 / fp16 autocast outputs -> fp32 for the loss math); float64 inputs — the
    rounding-free yardstick some tests compute — stay float64."""
    return x if x.dtype == torch.float64 else x.float()


SPATIAL_MOVE_TYPES = 139  # reference models/katago_base.py:41
ACTION_SPACE = 81 * 139   # reference models/katago_base.py:43


# ---------------------------------------------------------------------------------------------
# GAE — reference keisei/training/gae.py
# ---------------------------------------------------------------------------------------------
def gae_numpy(rewards, values, terminated, next_value, gamma, lam, override=None, lengths=None):
    """Pure-loop restatement of compute_gae / compute_gae_padded (gae.py:54-73, :117-148).

    (T, N) arrays. fp32 arithmetic with every operation rounded separately, python-float
    scalars rounded to fp32 first (PyTorch's scalar-times-tensor rule).
    """
    r = np.asarray(rewards, dtype=np.float32)
    v = np.asarray(values, dtype=np.float32)
    term = np.asarray(terminated).astype(np.float32)
    nvb = np.broadcast_to(np.asarray(next_value, dtype=np.float32).reshape(-1), (r.shape[1],))
    T, N = r.shape
    g = np.float32(gamma)
    gl = np.float32(gamma * lam)
    adv = np.zeros((T, N), dtype=np.float32)
    for n in range(N):
        last_step = T - 1
        if lengths is not None:
            last_step = max(int(lengths[n]) - 1, 0)
        last = np.float32(0.0)
        for t in range(T - 1, -1, -1):
            nv = nvb[n] if t == T - 1 else v[t + 1, n]
            if lengths is not None and t == last_step:
                nv = nvb[n]
            if override is not None and not np.isnan(override[t, n]):
                nv = np.float32(override[t, n])
            nd = np.float32(1.0) - term[t, n]
            delta = np.float32(np.float32(r[t, n] + np.float32(np.float32(g * nv) * nd)) - v[t, n])
            last = np.float32(delta + np.float32(np.float32(gl * nd) * last))
            adv[t, n] = last
    return adv


def normalize_advantages(adv: np.ndarray) -> np.ndarray:
    """katago_ppo.py:797-798 — unbiased std, eps 1e-8, skipped when numel <= 1."""
    a = np.asarray(adv, dtype=np.float64)
    if a.size <= 1:
        return np.asarray(adv, dtype=np.float32)
    return ((a - a.mean()) / (a.std(ddof=1) + 1e-8)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# SE-ResNet — reference keisei/training/models/se_resnet.py
# ---------------------------------------------------------------------------------------------
def _bn(x, sd, prefix, training, new_stats, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d semantics (se_resnet.py:51,53,111,121): batch statistics with biased variance
    for normalisation and unbiased variance for the running buffer in training mode."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.numel() / x.shape[1]
        if new_stats is not None:
            new_stats[prefix + ".running_mean"] = (1 - momentum) * rm + momentum * mean.detach()
            new_stats[prefix + ".running_var"] = (1 - momentum) * rv + momentum * var.detach() * n / max(n - 1, 1)
            new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = rm, rv
    xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def global_pool(x):
    """se_resnet.py:93-98 — mean, amax, population std over 9x9, concatenated."""
    return torch.cat([x.mean(dim=(-2, -1)), x.amax(dim=(-2, -1)), x.std(dim=(-2, -1), correction=0)], dim=-1)


def seresnet_forward(sd: Mapping[str, torch.Tensor], obs: torch.Tensor, num_blocks: int,
                     training: bool, new_stats: dict | None = None):
    """Functional restatement of SEResNetModel._forward_impl (se_resnet.py:132-159) and
    GlobalPoolBiasBlock.forward (se_resnet.py:68-90) over a state_dict. Returns
    (policy (B,9,9,139), value_logits (B,3), score (B,1))."""
    x = F.relu(_bn(F.conv2d(obs, sd["input_conv.weight"], padding=1), sd, "input_bn", training, new_stats))
    for i in range(num_blocks):
        p = f"blocks.{i}."
        out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], padding=1), sd, p + "bn1", training, new_stats))
        g = global_pool(x)
        g = F.linear(F.relu(F.linear(g, sd[p + "global_fc.0.weight"], sd[p + "global_fc.0.bias"])),
                     sd[p + "global_fc.2.weight"], sd[p + "global_fc.2.bias"])
        out = out + g[:, :, None, None]
        out = _bn(F.conv2d(out, sd[p + "conv2.weight"], padding=1), sd, p + "bn2", training, new_stats)
        se = out.mean(dim=(-2, -1))
        se = F.linear(F.relu(F.linear(se, sd[p + "se_fc1.weight"], sd[p + "se_fc1.bias"])),
                      sd[p + "se_fc2.weight"], sd[p + "se_fc2.bias"])
        c = out.shape[1]
        scale, shift = se[:, :c], se[:, c:]
        out = out * torch.sigmoid(scale)[:, :, None, None] + shift[:, :, None, None]
        x = F.relu(out + x)
    pol = F.relu(_bn(F.conv2d(x, sd["policy_conv1.weight"]), sd, "policy_bn1", training, new_stats))
    pol = F.conv2d(pol, sd["policy_conv2.weight"], sd["policy_conv2.bias"]).permute(0, 2, 3, 1)
    pool = global_pool(x)
    v = F.linear(F.relu(F.linear(pool, sd["value_fc1.weight"], sd["value_fc1.bias"])),
                 sd["value_fc2.weight"], sd["value_fc2.bias"])
    s = F.linear(F.relu(F.linear(pool, sd["score_fc1.weight"], sd["score_fc1.bias"])),
                 sd["score_fc2.weight"], sd["score_fc2.bias"])
    return pol, v, s


# ---------------------------------------------------------------------------------------------
# Plain ResNet baseline — reference keisei/training/models/resnet.py (BASELINE.json configs[3])
# ---------------------------------------------------------------------------------------------
def resnet_forward(sd: Mapping[str, torch.Tensor], obs: torch.Tensor, num_layers: int, training: bool,
                   new_stats: dict | None = None):
    """Functional restatement of ResNetModel.forward (resnet.py:64-84) and ResidualBlock.forward
    (resnet.py:33-37) over a state_dict. Returns (policy_logits (B,11259), value (B,1) tanh)."""
    x = F.relu(_bn(F.conv2d(obs, sd["input_conv.weight"], padding=1), sd, "input_bn", training, new_stats))
    for i in range(num_layers):
        p = f"blocks.{i}."
        out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], padding=1), sd, p + "bn1", training, new_stats))
        out = _bn(F.conv2d(out, sd[p + "conv2.weight"], padding=1), sd, p + "bn2", training, new_stats)
        x = F.relu(out + x)
    pol = F.relu(_bn(F.conv2d(x, sd["policy_conv.weight"]), sd, "policy_bn", training, new_stats)).flatten(1)
    logits = F.linear(pol, sd["policy_fc.weight"], sd["policy_fc.bias"])
    v = F.relu(_bn(F.conv2d(x, sd["value_conv.weight"]), sd, "value_bn", training, new_stats)).flatten(1)
    v = F.relu(F.linear(v, sd["value_fc1.weight"], sd["value_fc1.bias"]))
    return logits, torch.tanh(F.linear(v, sd["value_fc2.weight"], sd["value_fc2.bias"]))


def scalar_ppo_losses(policy_logits, value, legal_mask, actions, old_log_probs, advantages, returns,
                      clip_epsilon=0.2, value_loss_coeff=0.5, entropy_coeff=0.01):
    """Standard PPO for the scalar contract. COMPOSITION UNPINNED (the reference trainer was deleted,
    CHANGELOG.md:250-254): ppo_clip_loss (katago_ppo.py:33-43) + value_loss_coeff * MSE(value.squeeze(-1),
    returns) (ScalarValueAdapter.compute_value_loss, value_adapter.py:49-59) - entropy_coeff * masked
    entropy (katago_ppo.py:880-888), coefficients from PPOParams (algorithm_registry.py:11-19)."""
    logp_all = masked_log_softmax(policy_logits, legal_mask)
    new_logp = logp_all.gather(1, actions.unsqueeze(1)).squeeze(1)
    ratio = (new_logp - old_log_probs).exp()
    policy_loss = -torch.min(ratio * advantages, ratio.clamp(1 - clip_epsilon, 1 + clip_epsilon) * advantages).mean()
    entropy = -(logp_all.exp() * logp_all.masked_fill(~legal_mask, 0.0)).sum(dim=-1).mean()
    value_loss = F.mse_loss(_f(value).squeeze(-1), returns)
    loss = policy_loss + value_loss_coeff * value_loss - entropy_coeff * entropy
    return {"loss": loss, "policy_loss": policy_loss, "value_loss": value_loss, "entropy": entropy,
            "new_log_probs": new_logp}


# ---------------------------------------------------------------------------------------------
# KataGo-PPO losses — reference keisei/training/katago_ppo.py
# ---------------------------------------------------------------------------------------------
def masked_log_softmax(flat_logits, legal_mask):
    """katago_ppo.py:873-874."""
    return F.log_softmax(_f(flat_logits).masked_fill(~legal_mask, float("-inf")), dim=-1)


def ppo_losses(policy_logits, value_logits, score_lead, legal_mask, actions, old_log_probs, advantages,
               value_cats, score_targets, clip_epsilon=0.2, lambda_policy=1.0, lambda_value=1.5,
               lambda_score=0.02, entropy_coeff=0.01):
    """Restates katago_ppo.py:858-924 (non-adapter path; the adapter path value_adapter.py:98-126
    is the same arithmetic). Returns dict of scalar tensors incl. 'loss' (differentiable)."""
    B = policy_logits.shape[0]
    logp_all = masked_log_softmax(policy_logits.reshape(B, -1), legal_mask)
    new_logp = logp_all.gather(1, actions.unsqueeze(1)).squeeze(1)
    ratio = (new_logp - old_log_probs).exp()                                   # katago_ppo.py:40
    surr1 = ratio * advantages
    surr2 = ratio.clamp(1 - clip_epsilon, 1 + clip_epsilon) * advantages
    policy_loss = -torch.min(surr1, surr2).mean()                              # :43
    probs = logp_all.exp()
    entropy = -(probs * logp_all.masked_fill(~legal_mask, 0.0)).sum(dim=-1).mean()  # :886-888
    if (value_cats >= 0).any():                                                # :51-57
        value_loss = F.cross_entropy(_f(value_logits), value_cats, ignore_index=-1)
    else:
        value_loss = _f(value_logits).sum() * 0.0
    score_loss = F.mse_loss(_f(score_lead).squeeze(-1), score_targets)     # :910-912
    loss = lambda_policy * policy_loss + lambda_value * value_loss + lambda_score * score_loss \
        - entropy_coeff * entropy                                              # :914-924
    return {"loss": loss, "policy_loss": policy_loss, "value_loss": value_loss,
            "score_loss": score_loss, "entropy": entropy, "new_log_probs": new_logp}


def scalar_value(value_logits, score_lead=None, alpha=0.0):
    """katago_ppo.py:533-541 and value_adapter.py:79-96."""
    p = F.softmax(_f(value_logits), dim=-1)
    v = p[:, 0] - p[:, 2]
    if alpha == 0.0 or score_lead is None:
        return v
    return (1 - alpha) * v + alpha * _f(score_lead).squeeze(-1).clamp(-1, 1)


def rollout_log_prob(flat_logits, legal_mask, actions):
    """katago_ppo.py:599-605 for a GIVEN action: softmax over masked logits in the logits' dtype,
    Categorical's renormalise + eps-clamp (eps = finfo(probs.dtype).eps), log, gather."""
    masked = flat_logits.masked_fill(~legal_mask, float("-inf"))
    probs = F.softmax(masked, dim=-1)
    dist = torch.distributions.Categorical(probs, validate_args=False)
    return dist.log_prob(actions)


def action_index(row: int, col: int, move_type: int) -> int:
    """spatial_action_mapper.rs:5,26-28 / se_resnet.py:145-146: flat = (row*9+col)*139 + move_type."""
    return (row * 9 + col) * SPATIAL_MOVE_TYPES + move_type


def amax_tie_grad(x: np.ndarray) -> np.ndarray:
    """d amax / dx with ties split evenly (probed: a k-way tie gives 1/k each)."""
    m = x.max()
    tie = (x == m).astype(np.float64)
    return tie / tie.sum()
