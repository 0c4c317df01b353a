"""Split-merge environment step: the learner and its league opponents each move their own environments (SURVEY 8(f)
rank 3; reference keisei/training/katago_loop.py:284-431 `split_merge_step`, result type :64-73).

Same signature, same result fields, same error texts as the reference function, so `dropin.install_into_reference()` can
put it behind `keisei.training.katago_loop.split_merge_step`. What differs is how the step is executed:

* The partition of the environments is pure host arithmetic on the two numpy arrays the caller already holds
  (`current_players`, `env_opponent_ids`): the index sets are built with numpy and shipped to the device as ONE int64
  tensor. The reference derives them on the device (`learner_mask.nonzero()`, `opponent_mask.cpu()`), which costs two
  host synchronisations per step before the first forward can be enqueued.
* All sub-batches are gathered with one `index_select` per input tensor (the groups are contiguous slices of the
  result), and all actions go back with one `index_copy_`.
* When every model is a keisei_b200 SE-ResNet on the observations' CUDA device, the forwards run as parallel branches
  of ONE replayed CUDA graph (`models.rollout_forward_many`, each branch sized for its share of the SMs) and every
  sub-batch is sampled by the packed-mask kernel (`policy_ops.policy_sample`); the zero-legal guard reads one flag
  word per group in a single host read after everything has been enqueued.
* Any other model (the reference's tests pass `MagicMock`s; an opponent may live on another GPU, katago_loop.py:252-281)
  takes the plain path: one forward per model under `no_grad`, `Categorical` sampling on CPU tensors / the sampling
  kernel on CUDA tensors, cross-device opponents evaluated where they live.

Module modes follow the reference: the learner is put in eval mode for its forward and left there (the trainer's
`update()` switches back to train mode, katago_loop.py:330-333); opponents are never toggled (they are loaded in eval mode).
The kernel path does not touch the `training` flags at all (the mode is an argument of the C call).
"""
from __future__ import annotations

import sys
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch
import torch.nn.functional as F

from . import policy_ops
from .models.se_resnet import SEResNetModel, rollout_forward_many


@dataclass
class SplitMergeResult:
    """Field for field the reference's result type (katago_loop.py:64-73)."""
    actions: torch.Tensor             # (num_envs,) int64, merged over learner and opponents
    learner_mask: torch.Tensor        # (num_envs,) bool
    opponent_mask: torch.Tensor       # (num_envs,) bool
    learner_log_probs: torch.Tensor   # (n_learner,)
    learner_values: torch.Tensor      # (n_learner,)
    learner_indices: torch.Tensor     # (n_learner,) int64 indices into the env array


def _result_type() -> type:
    """The reference's own dataclass when its loop module is loaded (callers may isinstance-check), else ours."""
    loop = sys.modules.get("keisei.training.katago_loop")
    return getattr(loop, "SplitMergeResult", SplitMergeResult) if loop is not None else SplitMergeResult


def _scalar_value(value_logits: torch.Tensor) -> torch.Tensor:
    p = F.softmax(value_logits, dim=-1)
    return p[:, 0] - p[:, 2]


def _model_device(model: Any, default: torch.device) -> torch.device:
    try:
        return next(model.parameters()).device
    except (StopIteration, AttributeError, TypeError):
        return default


def _same_device(a: torch.device, b: torch.device) -> bool:
    if a.type != b.type:
        return False
    if a.type != "cuda":
        return True
    ia = a.index if a.index is not None else torch.cuda.current_device()
    ib = b.index if b.index is not None else torch.cuda.current_device()
    return ia == ib


class _Group:
    """One (model, environments) pair of the step."""
    __slots__ = ("who", "model", "idx_np", "lo", "hi", "device", "obs", "masks", "out", "actions", "log_probs", "values",
                 "legal", "flags")

    def __init__(self, who: str, model: Any, idx_np: np.ndarray, lo: int, device: torch.device | None) -> None:
        self.who, self.model, self.idx_np, self.lo, self.hi, self.device = who, model, idx_np, lo, lo + len(idx_np), device
        self.obs = self.masks = self.out = self.actions = self.log_probs = self.values = self.legal = self.flags = None


def _zero_legal_error(g: _Group, zero_rows: list[int]) -> RuntimeError:
    envs = g.idx_np[zero_rows].tolist()
    who = "Learner" if g.who == "learner" else "Opponent"
    return RuntimeError(f"{who} envs {envs} have zero legal actions — all-False legal mask would produce NaN")


def _sample(g: _Group, value_adapter: Any | None, seed: int | None) -> None:
    """Masked sample (+ log-prob and scalar value for the learner) of one group from its network output."""
    out, masks, n = g.out, g.masks, g.obs.shape[0]
    flat = out.policy_logits.reshape(n, -1)
    learner = g.who == "learner"
    if flat.is_cuda:
        alpha = float(getattr(value_adapter, "score_blend_alpha", 0.0)) if (learner and value_adapter is not None) else 0.0
        fused_value = learner and (value_adapter is None or hasattr(value_adapter, "score_blend_alpha"))
        g.actions, g.log_probs, g.values, g.legal, g.flags = policy_ops.policy_sample(
            flat, masks, out.value_logits if fused_value else None, out.score_lead if fused_value else None, alpha, seed=seed)
        if learner and not fused_value:
            g.values = value_adapter.scalar_value_blended(out.value_logits, out.score_lead)
        return
    counts = masks.sum(dim=-1)
    if bool((counts == 0).any()):
        raise _zero_legal_error(g, (counts == 0).nonzero(as_tuple=True)[0].tolist())
    dist = torch.distributions.Categorical(F.softmax(flat.masked_fill(~masks, float("-inf")), dim=-1), validate_args=False)
    g.actions = dist.sample()
    if learner:
        g.log_probs = dist.log_prob(g.actions)
        g.values = (value_adapter.scalar_value_blended(out.value_logits, out.score_lead) if value_adapter is not None
                    else _scalar_value(out.value_logits))


@torch.no_grad()
def split_merge_step(
    obs: torch.Tensor,
    legal_masks: torch.Tensor,
    current_players: np.ndarray,
    learner_model: torch.nn.Module,
    opponent_model: torch.nn.Module | None = None,
    opponent_models: dict[int, torch.nn.Module] | None = None,
    env_opponent_ids: np.ndarray | None = None,
    learner_side: int | np.ndarray = 0,
    value_adapter: Any | None = None,
    opponent_devices: dict[int, torch.device | None] | None = None,
    *,
    seed: int | None = None,
    strict_guards: bool = True,
):
    """One environment step with the learner moving the environments where it is to move and each opponent moving its
    own (reference katago_loop.py:284-431). Legacy mode: `opponent_model=`; cohort mode: `opponent_models={id: model}` +
    `env_opponent_ids`. Returns learner-side data only (log-probs, values, indices) plus the merged actions.

    `seed` (keyword, ours): Philox seed of the sampling kernel (default: `torch.initial_seed()`); `strict_guards=False`
    skips the zero-legal host read on the kernel path (the reference always checks)."""
    if opponent_models is None and opponent_model is not None:
        opponents: dict[int, Any] = {0: opponent_model}
        env_ids = None
    elif opponent_models is not None:
        opponents, env_ids = opponent_models, env_opponent_ids
    else:
        raise ValueError("Must provide either opponent_model or opponent_models")

    num_envs, device = obs.shape[0], obs.device
    # ---- partition on the host: (model, env indices) groups, learner first, opponents in dict order ----
    is_learner = np.ascontiguousarray(np.asarray(current_players) == learner_side)
    if is_learner.shape != (num_envs,):
        is_learner = np.broadcast_to(is_learner, (num_envs,)).copy()
    groups: list[_Group] = []
    learner_idx = np.flatnonzero(is_learner)
    cursor = 0
    if learner_idx.size:
        groups.append(_Group("learner", learner_model, learner_idx, 0, None))
        cursor = learner_idx.size
    not_learner = ~is_learner
    for opp_id, model in opponents.items():
        sel = not_learner if env_ids is None else ((np.asarray(env_ids) == opp_id) & not_learner)
        idx = np.flatnonzero(sel)
        if idx.size == 0:
            continue
        if opponent_devices is not None:
            od = opponent_devices.get(opp_id)
        else:
            md = _model_device(model, device)
            od = md if isinstance(md, torch.device) and not _same_device(md, device) else None
        groups.append(_Group("opponent", model, idx, cursor, od))
        cursor += idx.size

    learner_mask = torch.from_numpy(is_learner).to(device=device, dtype=torch.bool)
    opponent_mask = ~learner_mask
    actions = torch.zeros(num_envs, dtype=torch.long, device=device)
    empty = torch.zeros(0, device=device)
    if not groups:
        return _result_type()(actions=actions, learner_mask=learner_mask, opponent_mask=opponent_mask, learner_log_probs=empty,
                              learner_values=empty, learner_indices=torch.zeros(0, dtype=torch.long, device=device))

    # ---- one index upload, one gather per input tensor ----
    idx_all = torch.from_numpy(np.concatenate([g.idx_np for g in groups]).astype(np.int64, copy=False)).to(device)
    obs_all = obs.index_select(0, idx_all)
    masks_all = legal_masks.index_select(0, idx_all)
    for g in groups:
        g.obs, g.masks = obs_all[g.lo:g.hi], masks_all[g.lo:g.hi]
    learner_indices = idx_all[:learner_idx.size]

    # ---- forwards ----
    grouped = (device.type == "cuda" and len(groups) > 1
               and all(g.device is None and isinstance(g.model, SEResNetModel) and g.model.kernel_supported()
                       and _same_device(_model_device(g.model, device), device) for g in groups))
    if grouped:
        outs = rollout_forward_many([(g.model, g.obs) for g in groups], eval_mode=True)
        for g, out in zip(groups, outs):
            g.out = out
    else:
        for g in groups:
            if g.device is not None:                       # opponent on another GPU: evaluate it where it lives
                g.obs, g.masks = g.obs.to(g.device), g.masks.to(g.device)
            if isinstance(g.model, SEResNetModel) and g.obs.is_cuda:
                g.out = g.model.rollout_forward(g.obs, eval_mode=True)
            else:
                if g.who == "learner":
                    g.model.eval()
                g.out = g.model(g.obs)
    # ---- sampling; zero-legal guard ----
    for g in groups:
        _sample(g, value_adapter, seed)
    flagged = [g for g in groups if g.flags is not None]
    if strict_guards and flagged:
        host = torch.stack([g.flags[0].to(device) for g in flagged]).tolist()   # one host read for every kernel-sampled group
        for g, bad in zip(flagged, host):
            if int(bad) != 0:
                raise _zero_legal_error(g, (g.legal == 0).nonzero(as_tuple=True)[0].tolist())
    # ---- merge ----
    actions.index_copy_(0, idx_all, torch.cat([g.actions.to(device) for g in groups]))
    learner = groups[0] if groups[0].who == "learner" else None
    return _result_type()(
        actions=actions, learner_mask=learner_mask, opponent_mask=opponent_mask,
        learner_log_probs=learner.log_probs if learner is not None else empty,
        learner_values=learner.values if learner is not None else empty,
        learner_indices=learner_indices)
