"""Data-parallel plumbing: the gradient exchange and the SyncBatchNorm statistic exchange that replace the reference's
DistributedDataParallel + SyncBatchNorm wrap (katago_loop.py:494-508).

One process per GPU (torchrun env: RANK / LOCAL_RANK / WORLD_SIZE), NCCL over NVLink 5 / NVSwitch
for CUDA, gloo for CPU tests. The PPO update's only exchange step is the gradient average: the
keisei_b200 backward emits every parameter gradient in ONE flat fp32 buffer (213.7 MB for the
40x256 model); its residual blocks are contiguous runs of that buffer, finished from the last block to
the first, so the exchange is a handful of ~25 MB `all_reduce` calls on a communication stream, each
launched as soon as the backward schedule has enqueued its bucket's kernels (`kb_bucket_hook`) —
DDP's bucketed overlap without DDP. Rollout
inference shards by process with no collective. Per-rank semantics are the reference's:
`batch_size`, GAE and advantage normalisation are per rank; gradients are averaged; parameters are
broadcast from rank 0 at construction.
"""
from __future__ import annotations

import ctypes
import logging
import os
from typing import Iterable, NamedTuple

import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)


# The process-group bootstrap (torchrun env discovery, init / destroy, per-rank seeding) is NOT part of the hot path:
# under the reference's loop its own `keisei.training.distributed` keeps doing that job (katago_loop.py:1962-1985) and
# this package only plugs GradSync / BatchNormSync into the trainer. The names are re-exported when the reference is
# importable so `from keisei_b200.distributed import setup_distributed` keeps working in a drop-in install.
try:  # pragma: no cover - depends on the environment
    from keisei.training.distributed import (DistributedContext, cleanup_distributed, get_distributed_context,  # noqa: F401
                                             seed_all_ranks, setup_distributed)
except Exception:  # noqa: BLE001  (standalone use: bench.py, tests)
    pass


class RankEnv(NamedTuple):
    """What the launcher (torchrun) told this process. `launched` is False for a plain single process."""
    rank: int
    local_rank: int
    world_size: int
    launched: bool


def rank_env() -> RankEnv:
    """Standalone launcher-env reader for bench.py / tests: RANK, LOCAL_RANK and WORLD_SIZE come as a set."""
    have = {k: os.environ.get(k) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    if have["RANK"] is None:
        return RankEnv(0, 0, 1, False)
    missing = [k for k, v in have.items() if v is None]
    if missing:
        raise RuntimeError(f"launcher environment incomplete: RANK is set but {missing} missing (use torchrun)")
    return RankEnv(int(have["RANK"]), int(have["LOCAL_RANK"]), int(have["WORLD_SIZE"]), True)


def init_from_env(backend: str | None = None) -> RankEnv:
    """One process per GPU: bind the local device and join the process group (NCCL on CUDA, gloo on CPU)."""
    env = rank_env()
    if env.launched and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        cuda = torch.cuda.is_available()
        backend = backend or ("nccl" if cuda else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(env.local_rank)
            dist.init_process_group(backend, device_id=torch.device(f"cuda:{env.local_rank}"))
        else:
            dist.init_process_group(backend)
    return env


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


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

    def begin_update(self, n_samples: int, batch_size: int, device) -> None:
        """Called by `KataGoPPOAlgorithm.update()` before the first minibatch (collective)."""
        _check_equal_shards(self.group, self.world_size, n_samples, batch_size, device)

    def check(self) -> None:
        """NCCL reports a lost peer itself (watchdog / timeout): nothing to poll."""


class KbPeerCtx(ctypes.Structure):
    """include/keisei_b200.h: kb_peer_ctx"""
    _fields_ = [("peers", ctypes.c_void_p * 16), ("rank", ctypes.c_int), ("world", ctypes.c_int), ("n_slots", ctypes.c_int),
                ("reserved", ctypes.c_int), ("slot_doubles", ctypes.c_longlong), ("seq", ctypes.c_ulonglong),
                ("status", ctypes.POINTER(ctypes.c_ulonglong)), ("timeout_ms", ctypes.c_longlong)]


class PeerLostError(RuntimeError):
    """A SyncBatchNorm exchange gave up waiting for a peer rank (the step's statistics and gradients are NaN)."""


def _check_equal_shards(group, world: int, n_samples: int, batch_size: int, device) -> None:
    """SyncBatchNorm normalises with count = local_batch * world and every rank must make the same number of exchanges:
    all ranks have to run the same minibatch size and the same number of minibatches (the reference's DDP +
    SyncBatchNorm deadlocks or mis-normalises in the same situation). Checked once per update(), one tiny all-gather."""
    if world <= 1:
        return
    mine = torch.tensor([int(n_samples), int(batch_size)], dtype=torch.int64, device=device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    seen = {tuple(int(v) for v in p.tolist()) for p in parts}
    if len(seen) != 1:
        raise RuntimeError(f"SyncBatchNorm needs equal shards on every rank: (samples, batch_size) per rank = "
                           f"{[tuple(int(v) for v in p.tolist()) for p in parts]}")


class PeerBatchNormSync:
    """SyncBatchNorm exchange as ONE kernel over NVLink peer memory (csrc/peer_sync.cu) instead of an NCCL collective:
    every rank of the node owns a CUDA-IPC-shared buffer; an exchange stores this rank's sums into all peers' buffers,
    raises a flag there, waits for the peers' flags locally and adds the rows in rank order (bit-identical statistics
    on every rank). ~5 us instead of ~21 us per exchange, 164 exchanges per 40-block step, and no Python in the loop:
    the C schedule calls `kb_peer_allreduce_hook` directly. Single node only (all ranks must be IPC peers); use
    `BatchNormSync` (NCCL) otherwise. Call `close()` on every rank when done (collective; garbage collection only
    unmaps and frees without the barrier if the process group is already gone)."""

    def __init__(self, process_group=None, max_channels: int = 1024, n_slots: int = 4, timeout_s: float | None = None) -> None:
        from . import _lib
        self._status = None
        if not dist.is_initialized():
            raise RuntimeError("PeerBatchNormSync needs an initialised process group (setup_distributed)")
        self._lib = _lib
        lib = _lib.load()
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        if self.world_size > 16:
            raise ValueError("PeerBatchNormSync supports up to 16 ranks of one node")
        slot = 2 * int(max_channels)
        nbytes = lib.kb_peer_buffer_bytes(self.world_size, n_slots, slot)
        own, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _lib.check(lib.kb_peer_buffer_create(nbytes, ctypes.byref(own), handle), "kb_peer_buffer_create")
        self._own = own.value
        self._collective = self.world_size > 1   # close() meets the other ranks at a barrier before freeing
        handles: list = [None] * self.world_size
        dist.all_gather_object(handles, bytes(handle), group=process_group)
        self.ctx = KbPeerCtx()
        self._opened: list[int] = []
        failure: Exception | None = None
        try:
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.ctx.peers[r] = self._own
                    continue
                p = ctypes.c_void_p()
                hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(lib.kb_peer_buffer_open(hb, ctypes.byref(p)), f"kb_peer_buffer_open(rank {r})")
                self.ctx.peers[r] = p.value
                self._opened.append(p.value)
        except Exception as e:  # noqa: BLE001
            failure = e
        # every rank learns whether EVERY rank mapped every buffer; on any failure all ranks back out together (a rank
        # that raised alone would leave the others waiting in the next collective)
        oks: list = [None] * self.world_size
        dist.all_gather_object(oks, failure is None, group=process_group)
        if not all(oks):
            self.close()   # collective: unmap, barrier, free
            raise RuntimeError(f"PeerBatchNormSync: CUDA IPC mapping failed on rank(s) {[r for r, o in enumerate(oks) if not o]}"
                               + (f": {failure}" if failure is not None else ""))
        self.ctx.rank, self.ctx.world, self.ctx.n_slots, self.ctx.slot_doubles, self.ctx.seq = \
            self.rank, self.world_size, n_slots, slot, 0
        self._arm_status(timeout_s)
        self.c_hook = ctypes.cast(lib.kb_peer_allreduce_hook, ctypes.c_void_p)
        self.c_user = ctypes.cast(ctypes.pointer(self.ctx), ctypes.c_void_p)
        dist.barrier(group=process_group)  # every rank has mapped every buffer before the first exchange

    @classmethod
    def from_local_buffers(cls, ptrs: list[int], rank: int, slot_doubles: int, n_slots: int = 4,
                           timeout_s: float | None = None) -> "PeerBatchNormSync":
        """Ranks emulated inside ONE process (tests): the 'peers' are plain device buffers of this process."""
        from . import _lib
        self = cls.__new__(cls)
        self._lib, self.group, self.world_size, self.rank = _lib, None, len(ptrs), rank
        self._own, self._opened, self._collective = None, [], False
        self.ctx = KbPeerCtx()
        for r, p in enumerate(ptrs):
            self.ctx.peers[r] = p
        self.ctx.rank, self.ctx.world, self.ctx.n_slots, self.ctx.slot_doubles, self.ctx.seq = rank, len(ptrs), n_slots, slot_doubles, 0
        self._status = None
        self._arm_status(timeout_s)
        self.c_hook = ctypes.cast(_lib.load().kb_peer_allreduce_hook, ctypes.c_void_p)
        self.c_user = ctypes.cast(ctypes.pointer(self.ctx), ctypes.c_void_p)
        return self

    def _arm_status(self, timeout_s: float | None) -> None:
        """Host-visible failure word (pinned mapped memory) + how long an exchange waits for the slowest rank. Ranks skew
        by seconds in normal operation (checkpointing / evaluation on rank 0, uneven CPU rollouts, first-call builds), so
        the default is two minutes (KB_PEER_TIMEOUT_S overrides)."""
        if timeout_s is None:
            timeout_s = float(os.environ.get("KB_PEER_TIMEOUT_S", "120"))
        word = ctypes.POINTER(ctypes.c_ulonglong)()
        self._lib.check(self._lib.load().kb_peer_status_create(ctypes.byref(word)), "kb_peer_status_create")
        self._status = word
        self.ctx.status = word
        self.ctx.timeout_ms = max(1, int(timeout_s * 1000))

    def begin_update(self, n_samples: int, batch_size: int, device) -> None:
        """Called by `KataGoPPOAlgorithm.update()` before the first minibatch (collective): checks the equal-shard
        assumption of the exchange and lines the ranks up on the host, so the first exchange of the update does not
        start with whatever skew the rollout phase left."""
        if self.group is None and not (dist.is_available() and dist.is_initialized()):
            return   # thread-emulated ranks (tests)
        _check_equal_shards(self.group, self.world_size, n_samples, batch_size, device)

    def check(self) -> None:
        """Raise `PeerLostError` if any exchange since the last check timed out. Reads one pinned host word; meaningful
        after the stream has been synchronised (the trainer calls it right after its per-step host reads)."""
        if self._status and self._status[0] != 0:
            word = int(self._status[0])
            self._status[0] = 0
            raise PeerLostError(f"SyncBatchNorm exchange #{(word & 0xffffffffffff) - 1} on rank {self.rank} timed out waiting for rank "
                                f"{(word >> 48) - 1} (after {self.ctx.timeout_ms / 1000:.0f} s): peer rank lost. The step's statistics were "
                                f"poisoned with NaN and NOT committed to the BatchNorm running buffers.")

    @torch.no_grad()
    def all_reduce_(self, sums: torch.Tensor) -> torch.Tensor:
        """The same exchange from Python (float64 tensor on this rank's device), e.g. for tests."""
        if self.world_size > 1:
            with torch.cuda.device(sums.device):
                rc = self._lib.load().kb_peer_allreduce_hook(self.c_user, sums.data_ptr(), sums.numel(),
                                                             torch.cuda.current_stream(sums.device).cuda_stream)
            self._lib.check(rc, "kb_peer_allreduce_hook")
        return sums

    def close(self, collective: bool = True) -> None:
        """Unmap the peers' buffers, then free this rank's. COLLECTIVE (unless `collective=False`) when the object was built
        over a process group:
        a buffer must not be freed while another rank still has it mapped (CUDA IPC), so every rank unmaps first and
        the ranks meet at a barrier before any of them frees. Idempotent."""
        lib = self._lib.load()
        for p in self._opened:
            lib.kb_peer_buffer_close(p)
        self._opened = []
        if getattr(self, "_status", None):
            self.ctx.status = ctypes.POINTER(ctypes.c_ulonglong)()
            lib.kb_peer_status_destroy(self._status)
            self._status = None
        if self._own:
            if collective and getattr(self, "_collective", False) and dist.is_available() and dist.is_initialized():
                try:
                    dist.barrier(group=self.group)
                except Exception:  # noqa: BLE001  (process group already torn down: nothing left to wait for)
                    pass
            lib.kb_peer_buffer_destroy(self._own)
            self._own = None

    def __del__(self) -> None:  # pragma: no cover - interpreter shutdown order
        try:
            self.close(collective=False)
        except Exception:
            pass


class GradSync:
    """Gradient averaging across ranks for `KataGoPPOAlgorithm.grad_sync`.

    `all_reduce_flat` — the fused path: one collective over the flat gradient buffer.
    `all_reduce_params` — generic path (CPU / wrapped models): gradients are flattened into buckets
    of `bucket_bytes`, reduced, and copied back.
    """

    def __init__(self, process_group=None, bucket_bytes: int = 48 << 20, overlap: bool = True) -> None:
        if not dist.is_initialized():
            raise RuntimeError("GradSync needs an initialised process group (setup_distributed)")
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        # DDP's default is 25 MB buckets (reference katago_loop.py:498-504 keeps it); each bucket costs a host-side NCCL
        # enqueue (~50 us) on a path that is launch-bound at 1024 samples per GPU, and NVSwitch moves 48 MB in ~0.15 ms,
        # so fewer, larger buckets measure better here (5 per 213.7 MB gradient)
        self.bucket_bytes = bucket_bytes
        # fused CUDA path: all-reduce each bucket on a communication stream as soon as the backward schedule has enqueued
        # the kernels that produce it (kb_bucket_hook), so only the last bucket (the stem) is exposed
        self.overlap = overlap
        self.last_overlap_buckets = 0
        self._comm_streams: dict = {}

    def comm_stream(self, device: torch.device) -> "torch.cuda.Stream":
        s = self._comm_streams.get(device)
        if s is None:
            s = self._comm_streams[device] = torch.cuda.Stream(device)
        return s

    @torch.no_grad()
    def broadcast_parameters(self, module: torch.nn.Module, src: int = 0) -> None:
        """Reference DDP ctor semantics: every rank starts from rank 0's parameters and buffers."""
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=self.group)
        for m in module.modules():   # writes through .data bypass the autograd version counters the weight cache keys on
            if hasattr(m, "invalidate_packed_weights"):
                m.invalidate_packed_weights()

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
