"""Plain ResNet policy+value network — drop-in for the reference's `resnet` registry entry
(keisei/training/models/resnet.py:13-84; BASELINE.json configs[3]).

Parameters and buffers live in standard `nn.Conv2d` / `nn.BatchNorm2d` / `nn.Linear` containers with
the reference's attribute names, so `state_dict()` keys, shapes and registration order are identical.
A CUDA observation runs the whole network through one C-ABI call (`keisei_b200::resnet_forward`,
csrc/resnet.cu): tcgen05 implicit-GEMM convolutions in bf16 (under bf16 autocast or after
`configure_amp(True, torch.bfloat16)`), fp32-accurate SIMT kernels otherwise. CPU tensors run the
same graph with plain PyTorch ops (host logic for the reference's CPU-only tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from .. import resnet_ops
from .._lib import KeiseiB200Error
from ..model_ops import POLICY_A
from .base import BaseModel


@dataclass(frozen=True)
class ResNetParams:
    hidden_size: int
    num_layers: int

    def __post_init__(self) -> None:
        if self.hidden_size <= 0:
            raise ValueError(f"hidden_size must be > 0, got {self.hidden_size}")
        if self.num_layers < 0:
            raise ValueError(f"num_layers must be >= 0, got {self.num_layers}")


class ResidualBlock(nn.Module):
    """relu(bn2(conv2(relu(bn1(conv1(x))))) + x) (reference resnet.py:25-37). Parameter container;
    `forward` is the CPU path."""

    def __init__(self, channels: int) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = torch.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return torch.relu(out + x)


class ResNetModel(BaseModel):
    """`forward(obs) -> (policy_logits (B, 11259), value (B, 1))`; ValueError on a bad observation shape."""

    def __init__(self, params: ResNetParams) -> None:
        super().__init__()
        self.params = params
        ch = params.hidden_size
        self.input_conv = nn.Conv2d(self.OBS_CHANNELS, ch, 3, padding=1, bias=False)
        self.input_bn = nn.BatchNorm2d(ch)
        self.blocks = nn.Sequential(*[ResidualBlock(ch) for _ in range(params.num_layers)])
        policy_channels = 2
        self.policy_conv = nn.Conv2d(ch, policy_channels, 1, bias=False)
        self.policy_bn = nn.BatchNorm2d(policy_channels)
        self.policy_fc = nn.Linear(policy_channels * self.BOARD_SIZE * self.BOARD_SIZE, self.ACTION_SPACE)
        value_channels = 1
        self.value_conv = nn.Conv2d(ch, value_channels, 1, bias=False)
        self.value_bn = nn.BatchNorm2d(value_channels)
        self.value_fc1 = nn.Linear(value_channels * self.BOARD_SIZE * self.BOARD_SIZE, ch)
        self.value_fc2 = nn.Linear(ch, 1)
        # kernel-side state (not part of state_dict)
        self._amp_enabled = False
        self._amp_dtype = torch.float16
        self._wpack: torch.Tensor | None = None
        self._wpack_key: tuple | None = None
        self.use_tensor_cores: bool = True
        self.last_policy_buffer: torch.Tensor | None = None
        self._tables_cache = None
        self._grad_sizes: list[int] = []

    # ---- kernel plumbing (same shape as SEResNetModel's) -------------------------------------------
    def configure_amp(self, enabled: bool, dtype: torch.dtype = torch.float16, device_type: str = "cuda") -> None:
        """keisei_b200 extension (the reference BaseModel relies on an outer autocast): bf16 selects the
        tcgen05 kernels without an autocast context."""
        self._amp_enabled, self._amp_dtype = enabled, dtype

    def _desc(self) -> list[int]:
        return [self.params.num_layers, self.params.hidden_size, self.OBS_CHANNELS]

    def kernel_supported(self) -> bool:
        return self.params.hidden_size % 4 == 0 and self.params.hidden_size <= 1024

    def _ptr_tables(self) -> "resnet_ops.PointerTables":
        t = self._tables_cache
        if t is None or not t.valid():
            params, buffers = list(self.parameters()), list(self.buffers())
            t = self._tables_cache = resnet_ops.PointerTables(params, buffers, self._desc())
            self._grad_sizes = [p.numel() for p in params]
        return t

    def _apply(self, fn, *args, **kwargs):
        self._tables_cache = None
        self._wpack_key = None
        return super()._apply(fn, *args, **kwargs)

    def _act_dtype(self, device: torch.device) -> torch.dtype:
        if self._amp_enabled and self._amp_dtype == torch.bfloat16:
            return torch.bfloat16
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def _packed(self, params, buffers, dtype: torch.dtype) -> torch.Tensor:
        dev = params[0].device
        key = (dtype, dev, sum(p._version for p in params) + sum(b._version for b in buffers), params[0].data_ptr())
        if self._wpack is None or self._wpack_key != key:
            code = 0 if dtype == torch.float32 else 1
            nbytes = resnet_ops.wpack_bytes(self._desc(), code)
            if self._wpack is None or self._wpack.numel() != nbytes or self._wpack.device != dev:
                self._wpack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            resnet_ops.pack_weights(params, buffers, self._desc(), code, self._wpack)
            self._wpack_key = key
        return self._wpack

    @torch.no_grad()
    def _store_running_stats(self, buffers: list[torch.Tensor], new_stats: torch.Tensor) -> None:
        dst, src, nbt = [], [], []
        for layer in range(len(buffers) // 3):
            rm, rv, n = buffers[3 * layer], buffers[3 * layer + 1], buffers[3 * layer + 2]
            c = rm.numel()
            dst += [rm, rv]
            src += [new_stats[layer, 0, :c], new_stats[layer, 1, :c]]
            nbt.append(n)
        torch._foreach_copy_(dst, src)
        torch._foreach_add_(nbt, 1)

    def eval_forward(self, obs: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Inference-mode forward on CUDA whatever the modules' `training` flags say, without flipping them (the kernels
        take the mode as an argument; `model.eval()` ... `model.train()` per rollout step is pure Python overhead).
        Anything the kernels do not cover gets the ordinary eval() ... train() bracket."""
        if obs.is_cuda and obs.ndim == 4 and self.kernel_supported():
            return self._forward_cuda(obs, training=False)
        was = self.training
        self.eval()
        try:
            return self.forward(obs)
        finally:
            self.train(was)

    def _forward_cuda(self, obs: torch.Tensor, training: bool | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        if not self.kernel_supported():
            raise KeiseiB200Error(f"ResNetParams {self.params} not supported by the CUDA kernels "
                                  "(hidden_size must be a multiple of 4, <= 1024)")
        tables = self._ptr_tables()
        params, buffers = tables.params, tables.buffers
        dtype = self._act_dtype(obs.device)
        code = 0 if dtype == torch.float32 else 1
        training = self.training if training is None else training
        wpack = self._packed(params, buffers, dtype)
        if torch.is_grad_enabled() and training:
            policy_buf, value, _ws, new_stats = resnet_ops.resnet_forward(
                obs, params, buffers, wpack, self._desc(), training, code, bool(self.use_tensor_cores))
        else:
            policy_buf, value, _ws, new_stats = resnet_ops.resnet_forward_raw(
                obs, tables, wpack, training, code, bool(self.use_tensor_cores))
        if training:
            self._store_running_stats(buffers, new_stats)
        self.last_policy_buffer = policy_buf
        return policy_buf[:, :POLICY_A], value

    # ---- CPU path: plain PyTorch on the same parameters (reference resnet.py:64-84) -------------------
    def _forward_host(self, obs: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        x = torch.relu(self.input_bn(self.input_conv(obs)))
        x = self.blocks(x)
        p = torch.relu(self.policy_bn(self.policy_conv(x))).flatten(1)
        policy_logits = self.policy_fc(p)
        v = torch.relu(self.value_bn(self.value_conv(x))).flatten(1)
        v = torch.relu(self.value_fc1(v))
        return policy_logits, torch.tanh(self.value_fc2(v))

    def forward(self, obs: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        if obs.ndim != 4 or obs.shape[1] != self.OBS_CHANNELS or obs.shape[2] != self.BOARD_SIZE or obs.shape[3] != self.BOARD_SIZE:
            hint = ""
            if obs.ndim == 4 and obs.shape[-1] == self.OBS_CHANNELS:
                hint = " (input appears to be NHWC — expected NCHW)"
            raise ValueError(
                f"Expected obs shape (batch, {self.OBS_CHANNELS}, {self.BOARD_SIZE}, {self.BOARD_SIZE}), "
                f"got {tuple(obs.shape)}{hint}")
        if obs.is_cuda:
            return self._forward_cuda(obs)
        return self._forward_host(obs)
