"""Supervised-learning loss and step on the same CUDA forward / backward (reference
keisei/sl/trainer.py:131-166: `SLTrainer.train_epoch` — policy cross-entropy over the 11,259 flat actions, W/D/L
cross-entropy, score MSE, weighted sum, GradScaler + clip_grad_norm_ + optimiser step).

`sl_losses` is the loss block of that loop; `SLStep` is its per-batch body for a `keisei_b200` model (the dataset,
DataLoader, scheduler and checkpointing of `SLTrainer` are not on the hot path and stay the reference's). On CUDA the
policy cross-entropy is the masked-policy kernel with NO mask (`KB_MASK_NONE`: one read of the logits, the row's
log-sum-exp and the target's log-prob), the value cross-entropy and the score MSE are `kb_value_losses_*`; CPU tensors run
the reference's PyTorch expressions.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch.amp import GradScaler

from . import policy_ops


@dataclass(frozen=True)
class SLLossWeights:
    """Reference SLConfig defaults (sl/trainer.py:45-47)."""
    lambda_policy: float = 1.0
    lambda_value: float = 1.5
    lambda_score: float = 0.02


def sl_losses(policy_logits: torch.Tensor, value_logits: torch.Tensor, score_lead: torch.Tensor, policy_targets: torch.Tensor,
              value_targets: torch.Tensor, score_targets: torch.Tensor, weights: SLLossWeights = SLLossWeights()):
    """(loss, policy_loss, value_loss, score_loss) — reference sl/trainer.py:147-158. `policy_logits` is (B, 9, 9, 139) or
    the flat / padded (B, >= 11259) buffer; targets are int64 (B,), int64 (B,) in {0, 1, 2}, float (B,)."""
    B = value_logits.shape[0]
    flat = policy_logits.reshape(B, -1) if policy_logits.ndim != 2 else policy_logits
    if flat.is_cuda:
        zeros = torch.zeros(B, device=flat.device)
        _, new_logp, _, _, _, flags = policy_ops.ppo_policy_loss(flat, None, policy_targets, zeros, zeros, 0.0)
        policy_loss = -new_logp.mean()
        out3 = policy_ops.value_losses(value_logits, value_targets, score_lead, score_targets)
        value_loss, score_loss = out3[0], out3[1]
    else:
        policy_loss = F.cross_entropy(flat.float(), policy_targets)
        value_loss = F.cross_entropy(value_logits.float(), value_targets)
        score_loss = F.mse_loss(score_lead.float().squeeze(-1), score_targets)
    loss = weights.lambda_policy * policy_loss + weights.lambda_value * value_loss + weights.lambda_score * score_loss
    return loss, policy_loss, value_loss, score_loss


class SLStep:
    """The body of `SLTrainer.train_epoch`'s batch loop (sl/trainer.py:131-166) for one batch dict with the dataset's keys
    (`observation`, `policy_target`, `value_target`, `score_target`): forward, losses, scaled backward, unscale, clip,
    optimiser step, scaler update. Returns the three loss values as tensors (no host sync)."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, weights: SLLossWeights = SLLossWeights(),
                 grad_clip: float = 1.0, use_amp: bool = False) -> None:
        self.model, self.optimizer, self.weights, self.grad_clip = model, optimizer, weights, grad_clip
        device = next(model.parameters()).device
        self.device = device
        self.scaler = GradScaler(enabled=use_amp and device.type == "cuda")
        if hasattr(model, "configure_amp"):
            model.configure_amp(enabled=use_amp, dtype=torch.bfloat16 if use_amp else torch.float16, device_type=device.type)

    def __call__(self, batch: dict) -> dict[str, torch.Tensor]:
        dev = self.device
        obs = batch["observation"].to(dev, non_blocking=True)
        pt, vt, st = (batch[k].to(dev, non_blocking=True) for k in ("policy_target", "value_target", "score_target"))
        self.model.train()
        out = self.model(obs)
        # the kernel model keeps the padded (B, 11264) logits buffer of its last CUDA forward: consume it in place
        flat = getattr(self.model, "last_policy_buffer", None)
        in_place = flat is not None and flat.is_cuda and flat.requires_grad and flat.shape[0] == obs.shape[0]
        flat = flat[:, :11259] if in_place else out.policy_logits
        loss, pl, vl, sl = sl_losses(flat, out.value_logits, out.score_lead, pt, vt, st, self.weights)
        self.optimizer.zero_grad(set_to_none=True)
        self.scaler.scale(loss).backward()
        self.scaler.unscale_(self.optimizer)
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clip)
        self.scaler.step(self.optimizer)
        self.scaler.update()
        if hasattr(self.model, "invalidate_packed_weights"):
            self.model.invalidate_packed_weights()
        return {"policy_loss": pl.detach(), "value_loss": vl.detach(), "score_loss": sl.detach()}
