"""Algorithm-name -> parameter-schema lookup (reference keisei/training/algorithm_registry.py:22-40: `katago_ppo` is
the only registered trainer; the loop type-checks the validated object with isinstance, katago_loop.py:538-541)."""
from __future__ import annotations

import dataclasses
from typing import Any

from .katago_ppo import KataGoPPOParams

_PARAM_SCHEMAS: dict[str, type] = {"katago_ppo": KataGoPPOParams}
VALID_ALGORITHMS = set(_PARAM_SCHEMAS)


def validate_algorithm_params(algorithm: str, params: dict[str, Any]) -> object:
    """Instantiate the schema of `algorithm`; ValueError for an unknown name, TypeError for unknown / missing fields."""
    schema = _PARAM_SCHEMAS.get(algorithm)
    if schema is None:
        raise ValueError(f"Unknown algorithm '{algorithm}'. Valid: {sorted(VALID_ALGORITHMS)}")
    try:
        return schema(**params)
    except TypeError as e:
        raise TypeError(f"Invalid params for '{algorithm}': {e}") from e


@dataclasses.dataclass(frozen=True)
class PPOParams:
    """Hyper-parameters of the scalar-contract PPO (BASELINE config 4). The reference deleted that trainer
    (CHANGELOG.md:250-254) and kept only this record (algorithm_registry.py:11-19, unregistered there too);
    keisei_b200.ppo.PPOAlgorithm consumes it."""

    learning_rate: float = 3e-4
    gamma: float = 0.99
    clip_epsilon: float = 0.2
    epochs_per_batch: int = 4
    batch_size: int = 256
    entropy_coeff: float = 0.01
    value_loss_coeff: float = 0.5
