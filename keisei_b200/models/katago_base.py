"""Reference module name for the multi-head contract (keisei/training/models/katago_base.py); see contracts.py."""
from .contracts import KataGoBaseModel, KataGoOutput  # noqa: F401
