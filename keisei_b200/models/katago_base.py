"""KataGo-style multi-head model contract — mirror of keisei/training/models/katago_base.py:14-78.

Same names, attributes and error behaviour as the reference so the trainer, the loop, the league
and the showcase sidecar can hold either implementation.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import torch
import torch.nn as nn


@dataclass
class KataGoOutput:
    """policy_logits (B, 9, 9, 139) raw/unmasked; value_logits (B, 3) W/D/L; score_lead (B, 1)."""

    policy_logits: torch.Tensor
    value_logits: torch.Tensor
    score_lead: torch.Tensor


class KataGoBaseModel(ABC, nn.Module):
    BOARD_SIZE = 9
    SPATIAL_MOVE_TYPES = 139
    SPATIAL_ACTION_SPACE = 81 * 139  # 11,259

    def __init__(self) -> None:
        super().__init__()
        self._amp_enabled: bool = False
        self._amp_dtype: torch.dtype = torch.float16
        self._amp_device_type: str = "cpu"
        self._amp_frozen: bool = False

    def configure_amp(self, enabled: bool, dtype: torch.dtype = torch.float16, device_type: str = "cuda") -> None:
        """Reference katago_base.py:52-66. On CUDA, bf16 AMP selects the bf16 tcgen05 kernels."""
        if self._amp_frozen:
            raise RuntimeError(
                "configure_amp() must not be called after torch.compile() — "
                "changing AMP attributes would trigger silent recompilation")
        self._amp_enabled = enabled
        self._amp_dtype = dtype
        self._amp_device_type = device_type

    def forward(self, obs: torch.Tensor) -> KataGoOutput:
        return self._forward_impl(obs)

    @abstractmethod
    def _forward_impl(self, obs: torch.Tensor) -> KataGoOutput: ...
