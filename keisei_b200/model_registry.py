"""Model registry — mirror of keisei/training/model_registry.py:17-100 for the architectures on
the hot path. `install_into_reference()` swaps these classes into the reference's own registry so
`keisei.training.model_registry.build_model("se_resnet", ...)` returns the B200 implementation."""
from __future__ import annotations

from typing import Any, Callable, NamedTuple

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
VALID_ARCHITECTURES = set(_REGISTRY)

# value checks per architecture: (field, predicate, requirement text) — the dataclass only checks names / arity
_POSITIVE = (lambda v: v > 0, "must be > 0")
_NON_NEGATIVE = (lambda v: v >= 0, "must be >= 0")
_FIELD_RULES: dict[str, tuple[tuple[str, tuple[Callable[[Any], bool], str]], ...]] = {
    "se_resnet": (("channels", _POSITIVE), ("se_reduction", _POSITIVE)),
    "resnet": (("hidden_size", _POSITIVE), ("num_layers", _NON_NEGATIVE)),
}


def _spec_of(architecture: str) -> ArchitectureSpec:
    try:
        return _REGISTRY[architecture]
    except KeyError:
        raise ValueError(f"Unknown architecture '{architecture}'. Valid: {sorted(VALID_ARCHITECTURES)}") from None


def validate_model_params(architecture: str, params: dict[str, Any]) -> object:
    """`params_cls(**params)` plus the value checks: TypeError for bad keyword names, ValueError for bad values or an
    unknown architecture (reference model_registry.py:34-72)."""
    spec = _spec_of(architecture)
    try:
        obj = spec.params_cls(**params)
    except TypeError as e:
        raise TypeError(f"Invalid params for '{architecture}': {e}") from e
    for field, (ok, text) in _FIELD_RULES.get(architecture, ()):
        value = getattr(obj, field)
        if not ok(value):
            raise ValueError(f"{architecture}: {field} {text}, got {value}")
    if architecture == "se_resnet" and obj.channels // obj.se_reduction < 1:   # the SE hidden width would be zero
        raise ValueError(f"se_resnet: channels ({obj.channels}) // se_reduction ({obj.se_reduction}) must be >= 1")
    return obj


def build_model(architecture: str, params: dict[str, Any]) -> nn.Module:
    return _spec_of(architecture).model_cls(validate_model_params(architecture, params))


def get_model_contract(architecture: str) -> str:
    return _spec_of(architecture).contract


def get_obs_channels(architecture: str) -> int:
    return _spec_of(architecture).obs_channels


def install_into_reference() -> None:
    """Drop-in: see `keisei_b200.dropin.install_into_reference` (registry entries, trainer, buffer and GAE functions of an
    importable `keisei`)."""
    from .dropin import install_into_reference as _install  # noqa: PLC0415
    _install()
