"""Data-parallel plumbing — mirror of keisei/training/distributed.py:39-157 plus the gradient
exchange that replaces the reference's DistributedDataParallel wrap (katago_loop.py:494-508).

One process per GPU (torchrun env: RANK / LOCAL_RANK / WORLD_SIZE), NCCL over NVLink 5 / NVSwitch
for CUDA, gloo for CPU tests. The PPO update's only exchange step is the gradient average: the
keisei_b200 backward emits every parameter gradient in ONE flat fp32 buffer (213.7 MB for the
40x256 model), so the exchange is a single `all_reduce` — no bucketing, NVLS-eligible. Rollout
inference shards by process with no collective. Per-rank semantics are the reference's:
`batch_size`, GAE and advantage normalisation are per rank; gradients are averaged; parameters are
broadcast from rank 0 at construction.
"""
from __future__ import annotations

import logging
import os
import random
from dataclasses import dataclass, field
from typing import Iterable

import numpy as np
import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)


def _resolve_device(is_distributed: bool, local_rank: int) -> torch.device:
    if is_distributed and torch.cuda.is_available():
        return torch.device(f"cuda:{local_rank}")
    if torch.cuda.is_available():
        return torch.device("cuda")
    return torch.device("cpu")


@dataclass(frozen=True, slots=True)
class DistributedContext:
    rank: int
    local_rank: int
    world_size: int
    is_distributed: bool
    device: torch.device = field(init=False)

    def __post_init__(self) -> None:
        object.__setattr__(self, "device", _resolve_device(self.is_distributed, self.local_rank))

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def _require_env(key: str) -> str:
    val = os.environ.get(key)
    if val is None:
        raise RuntimeError(f"torchrun env var {key!r} is missing. Ensure RANK, LOCAL_RANK, and WORLD_SIZE are all set. "
                           f"Launch with: torchrun --nproc_per_node=N your_script.py")
    return val


def get_distributed_context() -> DistributedContext:
    rank = os.environ.get("RANK")
    if rank is None:
        return DistributedContext(rank=0, local_rank=0, world_size=1, is_distributed=False)
    return DistributedContext(rank=int(rank), local_rank=int(_require_env("LOCAL_RANK")),
                              world_size=int(_require_env("WORLD_SIZE")), is_distributed=True)


def setup_distributed(ctx: DistributedContext, backend: str | None = None) -> None:
    if not ctx.is_distributed:
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    elif backend == "nccl" and not torch.cuda.is_available():
        raise RuntimeError("backend='nccl' requires CUDA but torch.cuda.is_available() is False. "
                           "Use backend='gloo' for CPU-only distributed training, or set backend=None to auto-select.")
    try:
        if torch.cuda.is_available():
            torch.cuda.set_device(ctx.local_rank)
        dist.init_process_group(backend=backend)
        logger.info("DP initialized: rank=%d, local_rank=%d, world_size=%d, backend=%s", ctx.rank, ctx.local_rank,
                    ctx.world_size, backend)
    except Exception:
        logger.error("DP init failed: rank=%d, local_rank=%d, world_size=%d, MASTER_ADDR=%s, MASTER_PORT=%s", ctx.rank,
                     ctx.local_rank, ctx.world_size, os.environ.get("MASTER_ADDR", "<unset>"),
                     os.environ.get("MASTER_PORT", "<unset>"))
        raise


def cleanup_distributed(ctx: DistributedContext) -> None:
    if ctx.is_distributed and dist.is_initialized():
        dist.destroy_process_group()


def seed_all_ranks(seed: int) -> None:
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


class BatchNormSync:
    """SyncBatchNorm exchange for `SEResNetModel.convert_sync_batchnorm` (reference katago_loop.py:494-497:
    `torch.nn.SyncBatchNorm.convert_sync_batchnorm` before the DDP wrap, `sync_batchnorm = true` by default).

    The C schedule hands over one (2*C,) float64 device slice per BatchNorm layer — (sum x, sum x^2) in the forward,
    (sum dz, sum dz*z) in the backward — and this object sums it over the ranks on the current stream (NCCL orders
    itself after the producing kernel and before the consuming one). 4 KB per call, 2*(2*blocks+2) calls per step."""

    def __init__(self, process_group=None) -> None:
        if not dist.is_initialized():
            raise RuntimeError("BatchNormSync needs an initialised process group (setup_distributed)")
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)

    @torch.no_grad()
    def all_reduce_(self, sums: torch.Tensor) -> torch.Tensor:
        if self.world_size > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        return sums


class GradSync:
    """Gradient averaging across ranks for `KataGoPPOAlgorithm.grad_sync`.

    `all_reduce_flat` — the fused path: one collective over the flat gradient buffer.
    `all_reduce_params` — generic path (CPU / wrapped models): gradients are flattened into buckets
    of `bucket_bytes`, reduced, and copied back.
    """

    def __init__(self, process_group=None, bucket_bytes: int = 256 << 20) -> None:
        if not dist.is_initialized():
            raise RuntimeError("GradSync needs an initialised process group (setup_distributed)")
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.bucket_bytes = bucket_bytes

    @torch.no_grad()
    def broadcast_parameters(self, module: torch.nn.Module, src: int = 0) -> None:
        """Reference DDP ctor semantics: every rank starts from rank 0's parameters and buffers."""
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=self.group)

    @torch.no_grad()
    def all_reduce_flat(self, flat: torch.Tensor) -> torch.Tensor:
        if self.world_size == 1:
            return flat
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world_size)
        return flat

    @torch.no_grad()
    def all_reduce_params(self, params: Iterable[torch.nn.Parameter]) -> None:
        if self.world_size == 1:
            return
        grads = [p.grad for p in params if p.grad is not None]
        bucket: list[torch.Tensor] = []
        size = 0
        def flush():
            nonlocal bucket, size
            if not bucket:
                return
            flat = torch.cat([g.reshape(-1) for g in bucket])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world_size)
            off = 0
            for g in bucket:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            bucket, size = [], 0
        for g in grads:
            bucket.append(g)
            size += g.numel() * g.element_size()
            if size >= self.bucket_bytes:
                flush()
        flush()
