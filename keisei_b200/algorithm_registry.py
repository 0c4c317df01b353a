"""Algorithm registry — mirror of keisei/training/algorithm_registry.py:22-40."""
from __future__ import annotations

from typing import Any

from .katago_ppo import KataGoPPOParams

_PARAM_SCHEMAS: dict[str, type] = {"katago_ppo": KataGoPPOParams}
VALID_ALGORITHMS = set(_PARAM_SCHEMAS.keys())


def validate_algorithm_params(algorithm: str, params: dict[str, Any]) -> object:
    if algorithm not in _PARAM_SCHEMAS:
        raise ValueError(f"Unknown algorithm '{algorithm}'. Valid: {sorted(VALID_ALGORITHMS)}")
    try:
        return _PARAM_SCHEMAS[algorithm](**params)
    except TypeError as e:
        raise TypeError(f"Invalid params for '{algorithm}': {e}") from e
