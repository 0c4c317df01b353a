"""Observation / legal-mask ingest from the VecEnv into HBM (reference katago_loop.py:1529-1530:
`torch.from_numpy(np.asarray(step_result.observations)).to(self.device)` and the same for the masks — two pageable,
synchronous host->device copies per environment step, 27.5 KB per sample).

`PinnedIngest` keeps `depth` (default 2) slots of page-locked host staging buffers and device buffers and a dedicated copy
stream: `submit()` copies the step's arrays into the next pinned slot and enqueues the host->device copies (and, on
request, the 8x bit-packing of the masks, `kb_pack_mask_bits`) on the copy stream; `get()` hands out the device tensors
after making the caller's stream wait on the slot's event. With two slots the transfer of step k+1 overlaps the network
forward of step k whenever the caller has the next observations early (split-merge sub-batches, several env groups,
evaluation / SL streams); in the strictly serial rollout loop it still replaces pageable copies by pinned asynchronous
ones. The fp32 NCHW -> bf16 NHWC cast is not a separate pass: the network's first kernel (`pack_obs_kernel`, csrc/blocks.cu)
reads the fp32 NCHW observation as it arrives and writes the padded bf16 NHWC tile the stem convolution's TMA loads.
"""
from __future__ import annotations

import numpy as np
import torch

from . import policy_ops


class PinnedIngest:
    def __init__(self, device: torch.device | str, num_envs: int, obs_shape: tuple[int, ...] = (50, 9, 9),
                 action_space: int = 11259, depth: int = 2, pack_masks: bool = False) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("PinnedIngest stages into CUDA memory; on CPU use the tensors directly")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.num_envs, self.obs_shape, self.action_space, self.depth, self.pack_masks = num_envs, tuple(obs_shape), action_space, depth, pack_masks
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()   # noqa: E731
        self._h_obs = [pin(num_envs, *obs_shape, dtype=torch.float32) for _ in range(depth)]
        self._h_mask = [pin(num_envs, action_space, dtype=torch.bool) for _ in range(depth)]
        self._d_obs = [torch.empty((num_envs, *obs_shape), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self._d_mask = [torch.empty((num_envs, action_space), dtype=torch.bool, device=self.device) for _ in range(depth)]
        self._d_bits: list = [None] * depth
        self._count = [0] * depth
        self._ready = [torch.cuda.Event() for _ in range(depth)]       # H2D of the slot finished (copy stream)
        self._consumed = [torch.cuda.Event() for _ in range(depth)]    # the consumer's work on the slot was enqueued
        self._used = [False] * depth
        self._stream = torch.cuda.Stream(self.device)
        self._next = 0
        self.h2d_bytes_per_sample = int(np.prod(obs_shape)) * 4 + action_space

    def submit(self, obs, legal_masks) -> int:
        """Stage one step (numpy arrays or CPU tensors, n <= num_envs rows) and start its transfer. Returns the slot."""
        slot = self._next
        self._next = (slot + 1) % self.depth
        obs_t = torch.as_tensor(np.asarray(obs) if not isinstance(obs, torch.Tensor) else obs)
        mask_t = torch.as_tensor(np.asarray(legal_masks) if not isinstance(legal_masks, torch.Tensor) else legal_masks)
        n = obs_t.shape[0]
        if n > self.num_envs or tuple(obs_t.shape[1:]) != self.obs_shape or tuple(mask_t.shape) != (n, self.action_space):
            raise ValueError(f"ingest: unexpected shapes {tuple(obs_t.shape)} / {tuple(mask_t.shape)}")
        if self._used[slot]:
            self._consumed[slot].synchronize()      # the GPU is done reading this slot's device buffers ...
            self._ready[slot].synchronize()         # ... and its previous transfer no longer reads the pinned staging
        # sources that already live in page-locked memory (a VecEnv writing into pinned arrays) are copied from directly —
        # the caller keeps them unchanged until the slot's transfer has finished (`get()` orders consumers after it);
        # pageable sources (plain numpy arrays) go through the slot's pinned staging buffers first
        direct = (obs_t.is_pinned() and mask_t.is_pinned() and obs_t.dtype == torch.float32 and mask_t.dtype == torch.bool
                  and obs_t.is_contiguous() and mask_t.is_contiguous())
        if direct:
            src_obs, src_mask = obs_t, mask_t
        else:
            self._h_obs[slot][:n].copy_(obs_t)      # host memcpy into page-locked memory (also converts dtype if needed)
            self._h_mask[slot][:n].copy_(mask_t)
            src_obs, src_mask = self._h_obs[slot][:n], self._h_mask[slot][:n]
        with torch.cuda.stream(self._stream):
            self._d_obs[slot][:n].copy_(src_obs, non_blocking=True)
            self._d_mask[slot][:n].copy_(src_mask, non_blocking=True)
            if self.pack_masks:
                self._d_bits[slot] = policy_ops.pack_mask_bits(self._d_mask[slot][:n])
            self._ready[slot].record(self._stream)
        self._count[slot] = n
        self._used[slot] = True
        return slot

    def get(self, slot: int) -> tuple[torch.Tensor, torch.Tensor]:
        """(observations, legal masks) of a submitted slot on the device, ordered after their transfer on the CURRENT
        stream. The masks are the bit-packed int32 rows when `pack_masks` was requested, else the bool tensor. The tensors
        stay valid until the slot is submitted again (`depth` submissions later)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[slot])
        n = self._count[slot]
        masks = self._d_bits[slot] if self.pack_masks else self._d_mask[slot][:n]
        return self._d_obs[slot][:n], masks

    def release(self, slot: int) -> None:
        """Mark the consumer's work on the slot as enqueued (call after the forward that reads it was launched)."""
        self._consumed[slot].record(torch.cuda.current_stream(self.device))
