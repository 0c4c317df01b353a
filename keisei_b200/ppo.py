"""Standard (scalar-value) PPO for the `resnet` baseline — BASELINE.json configs[3].

PARITY UNPINNED for the loss COMPOSITION: the reference deleted its scalar `PPOAlgorithm`
(CHANGELOG.md:250-254), so nothing in /root/reference pins how the terms were combined. This is the
restatement SURVEY.md §8(c) gives, assembled only from pieces that DO survive in the reference:

  loss = ppo_clip_loss(new_logp, old_logp, adv, clip_epsilon)             katago_ppo.py:33-43
       + value_loss_coeff * ScalarValueAdapter.compute_value_loss(...)    value_adapter.py:49-59 (MSE vs returns)
       - entropy_coeff * H(masked policy)                                 katago_ppo.py:880-888
  returns = advantages (GAE, un-normalised) + values                      gae.py:8-73

with the surviving `PPOParams` defaults (algorithm_registry.py:11-19). The model side (ResNetModel
forward / backward) IS pinned against the reference class (tests/golden/resnet_tiny.npz).

Everything else — rollout buffer, GAE dispatch, advantage normalisation, minibatch loop, guards,
optimiser / clip / GradScaler, data-parallel gradient sync — is inherited from KataGoPPOAlgorithm; the
CUDA path runs `kb_resnet_forward` / `kb_ppo_policy_fwd,bwd` / `kb_resnet_backward`.
"""
from __future__ import annotations

from typing import Any

import torch
import torch.nn.functional as F
from torch.amp import autocast

from . import policy_ops, resnet_ops
from .algorithm_registry import PPOParams
from .katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from .model_ops import POLICY_A
from .models.base import BaseModel
from .models.resnet import ResNetModel
from .value_adapter import ScalarValueAdapter


class PPOAlgorithm(KataGoPPOAlgorithm):
    """`select_actions(obs, legal_masks)` / `update(buffer, next_values)` with the trainer interface of
    katago_ppo (same buffer type, metrics keys; `score_loss` is always 0 for the scalar contract)."""

    def __init__(self, params: PPOParams, model: BaseModel, *, gae_lambda: float = 0.95, grad_clip: float = 1.0,
                 use_amp: bool = False) -> None:
        kp = KataGoPPOParams(learning_rate=params.learning_rate, gamma=params.gamma, gae_lambda=gae_lambda,
                             clip_epsilon=params.clip_epsilon, epochs_per_batch=params.epochs_per_batch,
                             batch_size=params.batch_size, lambda_policy=1.0, lambda_value=params.value_loss_coeff,
                             lambda_score=0.0, lambda_entropy=params.entropy_coeff, grad_clip=grad_clip, use_amp=use_amp)
        super().__init__(kp, model)
        self.ppo_params = params
        self._scalar_adapter = ScalarValueAdapter()

    def _kernel_model(self, device: torch.device) -> ResNetModel | None:
        base = self._base()
        if device.type == "cuda" and isinstance(base, ResNetModel) and base.kernel_supported():
            return base
        return None

    # ---- rollout ---------------------------------------------------------------------------------
    @torch.no_grad()
    def select_actions(self, obs: torch.Tensor, legal_masks: torch.Tensor, value_adapter: Any | None = None):
        """Eval-mode forward, zero-legal guard, masked sample + log-prob (katago_ppo.py:589-606 semantics),
        scalar value = value.squeeze(-1) (value_adapter.py:46-47). Leaves the model in train mode."""
        device = next(self.model.parameters()).device
        model = self.forward_model
        fast = isinstance(model, ResNetModel) and obs.is_cuda   # mode passed to the kernels: no eval() / train() walks
        if not fast:
            model.eval()
        try:
            tok = self._events(device, "select_actions_forward_ms")
            with autocast(device_type=device.type, dtype=torch.bfloat16, enabled=self.params.use_amp):
                logits, value = model.eval_forward(obs) if fast else model(obs)
            self._events_end(tok)
            values = self._scalar_adapter.scalar_value_from_output(value).float()
            if device.type == "cuda":
                actions, log_probs, _, legal, flags = policy_ops.policy_sample(logits, legal_masks, seed=self._sample_seed)
                if self.strict_guards and int(flags[0].item()) != 0:
                    zero_envs = (legal == 0).nonzero(as_tuple=True)[0].tolist()
                    raise RuntimeError(f"Environments {zero_envs} have zero legal actions — "
                                       f"all-False legal mask would produce NaN")
                return actions, log_probs, values
            legal_counts = legal_masks.sum(dim=-1)
            if (legal_counts == 0).any():
                zero_envs = (legal_counts == 0).nonzero(as_tuple=True)[0].tolist()
                raise RuntimeError(f"Environments {zero_envs} have zero legal actions — "
                                   f"all-False legal mask would produce NaN")
            masked = logits.masked_fill(~legal_masks, float("-inf"))
            dist = torch.distributions.Categorical(F.softmax(masked, dim=-1), validate_args=False)
            actions = dist.sample()
            return actions, dist.log_prob(actions), values
        finally:
            if not fast or not self.forward_model.training:
                self.forward_model.train()

    # ---- one optimisation step -----------------------------------------------------------------------
    def _losses(self, flat_logits, value, _score, mb, value_adapter):
        p = self.params
        masks, actions, old_lp, adv, returns = mb[0], mb[1], mb[2], mb[3], mb[6]
        policy_loss, entropy, flags = self._policy_terms(flat_logits, masks, actions, old_lp, adv)
        adapter = value_adapter if value_adapter is not None else self._scalar_adapter
        value_loss = adapter.compute_value_loss(value.float(), returns, None, None)
        loss = p.lambda_policy * policy_loss + p.lambda_value * value_loss - self.current_entropy_coeff * entropy
        return loss, policy_loss, value_loss, torch.zeros((), device=flat_logits.device), entropy, flags

    def _step_fused(self, km: ResNetModel, obs, mb, value_adapter):
        tables = km._ptr_tables()
        params, buffers = tables.params, tables.buffers
        dtype = torch.bfloat16 if self.params.use_amp else km._act_dtype(obs.device)
        code = 0 if dtype == torch.float32 else 1
        wpack = km._packed(params, buffers, dtype)
        with torch.no_grad():
            policy_buf, value, ws, new_stats = resnet_ops.resnet_forward_raw(obs, tables, wpack, True, code,
                                                                             bool(km.use_tensor_cores))
            km._store_running_stats(buffers, new_stats)
        policy_buf.requires_grad_(True); value.requires_grad_(True)
        loss, pl, vl, sl, ent, flags = self._losses(policy_buf[:, :POLICY_A], value, None, mb, value_adapter)
        self._check_flags(flags)
        self.optimizer.zero_grad(set_to_none=True)
        self.scaler.scale(loss).backward()
        with torch.no_grad():
            flat = resnet_ops.resnet_backward_raw(tables, wpack, ws, policy_buf.grad, value.grad, code,
                                                  bool(km.use_tensor_cores), km._grad_sizes)
            if self.grad_sync is not None:
                self.grad_sync.all_reduce_flat(flat)
            self._flat_grad = flat   # consumed by KataGoPPOAlgorithm._optimizer_tail
            off = 0
            for prm in params:
                n = prm.numel()
                prm.grad = flat[off:off + n].view(prm.shape)
                off += n
        return pl, vl, sl, ent, value.detach()

    def _step_autograd(self, obs, mb, value_adapter, amp_dtype, amp_dev):
        with autocast(device_type=amp_dev, dtype=amp_dtype, enabled=self.params.use_amp):
            logits, value = self.forward_model(obs)
            loss, pl, vl, sl, ent, flags = self._losses(logits, value, None, mb, value_adapter)
        self._check_flags(flags)
        self.optimizer.zero_grad(set_to_none=True)
        self.scaler.scale(loss).backward()
        if self.grad_sync is not None:
            self.grad_sync.all_reduce_params(self.model.parameters())
        return pl, vl, sl, ent, value.detach()
