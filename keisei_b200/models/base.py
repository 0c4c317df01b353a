"""Reference module name for the scalar-value contract (keisei/training/models/base.py); see contracts.py."""
from .contracts import BaseModel  # noqa: F401
