"""Scalar-value model contract — mirror of keisei/training/models/base.py:11-27.

Input: observation (batch, 50, 9, 9). Output: (policy_logits (batch, 11259) raw/unmasked,
value (batch, 1) tanh-activated)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
import torch.nn as nn


class BaseModel(ABC, nn.Module):
    OBS_CHANNELS = 50
    BOARD_SIZE = 9
    ACTION_SPACE = 11259

    @abstractmethod
    def forward(self, obs: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]: ...
