"""Host-side mirrors of the reference model classes on the hot path (keisei/training/models/)."""
from .katago_base import KataGoBaseModel, KataGoOutput  # noqa: F401
from .se_resnet import GlobalPoolBiasBlock, SEResNetModel, SEResNetParams, rollout_forward_many  # noqa: F401
from .base import BaseModel  # noqa: F401
from .resnet import ResidualBlock, ResNetModel, ResNetParams  # noqa: F401
