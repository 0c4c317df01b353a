"""Algorithm registry — mirror of keisei/training/algorithm_registry.py:22-40."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from .katago_ppo import KataGoPPOParams


@dataclass(frozen=True)
class PPOParams:
    """Reference algorithm_registry.py:11-19 — the only surviving piece of the deleted scalar PPO trainer
    (not registered in `_PARAM_SCHEMAS` there either); consumed by keisei_b200.ppo.PPOAlgorithm."""
    learning_rate: float = 3e-4
    gamma: float = 0.99
    clip_epsilon: float = 0.2
    epochs_per_batch: int = 4
    batch_size: int = 256
    entropy_coeff: float = 0.01
    value_loss_coeff: float = 0.5

_PARAM_SCHEMAS: dict[str, type] = {"katago_ppo": KataGoPPOParams}
VALID_ALGORITHMS = set(_PARAM_SCHEMAS.keys())


def validate_algorithm_params(algorithm: str, params: dict[str, Any]) -> object:
    if algorithm not in _PARAM_SCHEMAS:
        raise ValueError(f"Unknown algorithm '{algorithm}'. Valid: {sorted(VALID_ALGORITHMS)}")
    try:
        return _PARAM_SCHEMAS[algorithm](**params)
    except TypeError as e:
        raise TypeError(f"Invalid params for '{algorithm}': {e}") from e
