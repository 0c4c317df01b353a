"""Model registry — mirror of keisei/training/model_registry.py:17-100 for the architectures on
the hot path. `install_into_reference()` swaps these classes into the reference's own registry so
`keisei.training.model_registry.build_model("se_resnet", ...)` returns the B200 implementation."""
from __future__ import annotations

from typing import Any, NamedTuple

import torch.nn as nn

from .models.resnet import ResNetModel, ResNetParams
from .models.se_resnet import SEResNetModel, SEResNetParams


class ArchitectureSpec(NamedTuple):
    model_cls: type[nn.Module]
    params_cls: type
    contract: str  # "scalar" or "multi_head"
    obs_channels: int


_REGISTRY: dict[str, ArchitectureSpec] = {
    "resnet": ArchitectureSpec(ResNetModel, ResNetParams, "scalar", 50),
    "se_resnet": ArchitectureSpec(SEResNetModel, SEResNetParams, "multi_head", 50),
}

VALID_ARCHITECTURES = set(_REGISTRY.keys())


def _get_spec(architecture: str) -> ArchitectureSpec:
    if architecture not in _REGISTRY:
        raise ValueError(f"Unknown architecture '{architecture}'. Valid: {sorted(VALID_ARCHITECTURES)}")
    return _REGISTRY[architecture]


def validate_model_params(architecture: str, params: dict[str, Any]) -> object:
    spec = _get_spec(architecture)
    try:
        validated = spec.params_cls(**params)
    except TypeError as e:
        raise TypeError(f"Invalid params for '{architecture}': {e}") from e
    if architecture == "se_resnet":
        if validated.channels <= 0:
            raise ValueError(f"se_resnet: channels must be > 0, got {validated.channels}")
        if validated.se_reduction <= 0:
            raise ValueError(f"se_resnet: se_reduction must be > 0, got {validated.se_reduction}")
        if validated.channels // validated.se_reduction < 1:
            raise ValueError(f"se_resnet: channels ({validated.channels}) // se_reduction "
                             f"({validated.se_reduction}) must be >= 1")
    elif architecture == "resnet":
        if validated.hidden_size <= 0:
            raise ValueError(f"resnet: hidden_size must be > 0, got {validated.hidden_size}")
        if validated.num_layers < 0:
            raise ValueError(f"resnet: num_layers must be >= 0, got {validated.num_layers}")
    return validated


def build_model(architecture: str, params: dict[str, Any]) -> nn.Module:
    validated = validate_model_params(architecture, params)
    return _get_spec(architecture).model_cls(validated)


def get_model_contract(architecture: str) -> str:
    return _get_spec(architecture).contract


def get_obs_channels(architecture: str) -> int:
    return _get_spec(architecture).obs_channels


def install_into_reference() -> None:
    """Drop-in: replace the `se_resnet` entry of an importable `keisei` with this implementation
    (see INTEGRATION.md). The reference's params dataclass is kept so its isinstance checks hold."""
    import keisei.training.model_registry as ref  # noqa: PLC0415 — optional dependency

    old = ref._REGISTRY["se_resnet"]

    class _Adapter(SEResNetModel):
        def __init__(self, params):  # accepts the reference's SEResNetParams
            super().__init__(SEResNetParams(**{f: getattr(params, f) for f in SEResNetParams.__dataclass_fields__}))
            self.params = params

    _Adapter.__name__ = "SEResNetModel"
    ref._REGISTRY["se_resnet"] = ref.ArchitectureSpec(_Adapter, old.params_cls, old.contract, old.obs_channels)
