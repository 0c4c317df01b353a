"""The two model contracts of the hot path, in one place.

* scalar contract (reference keisei/training/models/base.py:11-27): `forward(obs (B,50,9,9)) -> (policy_logits
  (B, 11259) raw, value (B, 1) in [-1, 1])` — the plain ResNet baseline;
* multi-head contract (reference keisei/training/models/katago_base.py:14-78): `forward(obs) -> KataGoOutput` with a
  spatial policy `(B, 9, 9, 139)`, W/D/L logits `(B, 3)` and a score lead `(B, 1)`, plus the AMP switch the trainer
  flips (`configure_amp`, frozen after `torch.compile` in the reference; on CUDA bf16 AMP selects the tcgen05 kernels).

Names, attributes and error behaviour follow the reference so the trainer, the loop, the league and the showcase
sidecar can hold either implementation; `base.py` / `katago_base.py` re-export from here under the reference's module
names."""
from __future__ import annotations

import abc
import dataclasses

import torch
from torch import nn

_BOARD = 9
_MOVE_TYPES = 139            # per-square move encodings (spatial_action_mapper.rs)
_ACTIONS = _BOARD * _BOARD * _MOVE_TYPES   # 11,259 flat actions: (row * 9 + col) * 139 + move_type


class BaseModel(abc.ABC, nn.Module):
    """Scalar-value contract."""

    OBS_CHANNELS, BOARD_SIZE, ACTION_SPACE = 50, _BOARD, _ACTIONS

    @abc.abstractmethod
    def forward(self, obs: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError


@dataclasses.dataclass
class KataGoOutput:
    """What a multi-head model returns; all three tensors are raw (no mask, no softmax, no tanh)."""

    policy_logits: torch.Tensor   # (B, 9, 9, 139)
    value_logits: torch.Tensor    # (B, 3): win / draw / loss
    score_lead: torch.Tensor      # (B, 1)


class KataGoBaseModel(abc.ABC, nn.Module):
    """Multi-head contract. Subclasses implement `_forward_impl`; `forward` is the public entry."""

    BOARD_SIZE, SPATIAL_MOVE_TYPES, SPATIAL_ACTION_SPACE = _BOARD, _MOVE_TYPES, _ACTIONS

    def __init__(self) -> None:
        super().__init__()
        # the trainer's AMP decision, recorded on the model (read by the CUDA path to pick bf16 vs fp32 kernels)
        self._amp_enabled, self._amp_dtype, self._amp_device_type = False, torch.float16, "cpu"
        self._amp_frozen = False   # the reference sets this after torch.compile; configure_amp then refuses

    def configure_amp(self, enabled: bool, dtype: torch.dtype = torch.float16, device_type: str = "cuda") -> None:
        if self._amp_frozen:
            raise RuntimeError("configure_amp() must not be called after torch.compile() — "
                               "changing AMP attributes would trigger silent recompilation")
        self._amp_enabled, self._amp_dtype, self._amp_device_type = enabled, dtype, device_type

    @abc.abstractmethod
    def _forward_impl(self, obs: torch.Tensor) -> KataGoOutput:
        raise NotImplementedError

    def forward(self, obs: torch.Tensor) -> KataGoOutput:
        return self._forward_impl(obs)
