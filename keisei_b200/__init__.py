"""keisei_b200 — B200-native (sm_100a) hot path for Keisei's SE-ResNet + KataGo-PPO.

Host code is Python/PyTorch; every device op on the path is a hand-written CUDA kernel in
`libkeisei_b200.so`, reached through the C-ABI declared in `include/keisei_b200.h`.
"""
from . import _lib  # noqa: F401
from ._lib import KeiseiB200Error  # noqa: F401

__version__ = "0.1.0"
