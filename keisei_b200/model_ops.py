"""torch.library custom ops for the SE-ResNet forward / backward (csrc/model.cu).

`keisei_b200::seresnet_forward`   one C call that enqueues the whole network on the current stream
`keisei_b200::seresnet_backward`  the matching backward (training-mode workspaces only)

Autograd is registered on the forward op: gradients flow from (policy, value_logits, score_lead)
to every parameter. The observation never receives a gradient (reference: obs is data).
"""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_longlong, c_void_p
from typing import List

import torch

from . import _lib

POLICY_A = 81 * 139          # 11,259
POLICY_PITCH = 11264         # padded logits row (16-byte aligned in bf16 and fp32)


class SeResnetDesc(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("num_blocks", "channels", "se_hidden", "gpool_channels", "policy_channels",
                                      "value_fc", "score_fc", "obs_channels")]


_P = c_void_p
_lib.register_signature("kb_seresnet_num_params", c_longlong, [_P])
_lib.register_signature("kb_seresnet_num_buffers", c_longlong, [_P])
_lib.register_signature("kb_seresnet_wpack_bytes", c_longlong, [_P, c_int])
_lib.register_signature("kb_seresnet_workspace_bytes", c_longlong, [_P, c_int, c_int, c_int])
_lib.register_signature("kb_seresnet_pack_weights", c_int, [_P, _P, _P, c_int, _P, c_longlong, _P])
_lib.register_signature("kb_seresnet_forward", c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_longlong, _P,
                                                        c_longlong, _P, _P, c_int, c_int, _P])
_lib.register_signature("kb_seresnet_backward", c_int, [_P, _P, _P, c_int, c_int, _P, c_longlong, _P, c_longlong, _P,
                                                         _P, _P, c_int, c_int, _P])
_lib.register_signature("kb_seresnet_forward_sync", c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_longlong,
                                                             _P, c_longlong, _P, _P, c_int, c_int, _P, _P, c_int, _P])
_lib.register_signature("kb_seresnet_backward_sync", c_int, [_P, _P, _P, c_int, c_int, _P, c_longlong, _P, c_longlong,
                                                              _P, _P, _P, c_int, c_int, _P, _P, c_int, _P])
_lib.register_signature("kb_conv3x3_forward", c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P,
                                                       _P, _P, _P, c_int, _P])
_lib.register_signature("kb_conv3x3_se_tail", c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, c_int, _P])
_lib.register_signature("kb_conv3x3_wgrad", c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_longlong, c_int, _P])
_lib.register_signature("kb_conv3x3_wgrad_ws_bytes", c_longlong, [c_int, c_int, c_int])
_lib.register_signature("kb_pack_conv_weight", c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P])

_lib.register_signature("kb_pack_linear_weight", c_int, [_P, _P, c_int, c_int, c_int, c_int, _P])
_lib.register_signature("kb_linear_tc", c_int, [_P, c_longlong, c_int, _P, c_int, c_int, _P, _P, c_int, _P, c_longlong, _P,
                                                 c_longlong, c_int, c_int, c_longlong, c_int, _P])

_lib.register_signature("kb_se_block_tail", c_int, [_P] * 13 + [c_int] + [_P] * 3 + [c_int, c_int, c_int, c_int, _P])
_lib.register_signature("kb_se_block_tail_variant", c_int, [_P] * 13 + [c_int] + [_P] * 3 + [c_int, c_int, c_int, c_int, c_int, _P])

_lib.register_signature("kb_seresnet_set_bucket_hook", c_int, [_P, _P])

_DT = {torch.float32: 0, torch.bfloat16: 1}
_DT_INV = {0: torch.float32, 1: torch.bfloat16}
_sm_count_cache: dict[int, int] = {}


# ---- SyncBatchNorm plumbing (include/keisei_b200.h: kb_allreduce_hook) ------------------------
_HOOK_T = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_longlong, c_void_p)
_BN_SYNCS: dict[int, object] = {}     # handle -> sync object, for the torch.library ops (schemas carry ints only)


def register_bn_sync(sync) -> int:
    """Handle for a BatchNorm sync object (`.world_size`, `.all_reduce_(float64 tensor)`); 0 means none."""
    if sync is None:
        return 0
    h = id(sync)
    _BN_SYNCS[h] = sync
    return h


class _BnHook:
    """C callback that sums a (2*C,) float64 slice of the workspace over the ranks, in stream order.
    The schedule calls it between each convolution and its BatchNorm finalize (forward) and before each
    BatchNorm-backward finalize. Exceptions cannot cross the C frame: they are parked and re-raised after."""

    def __init__(self, ws: torch.Tensor, sync) -> None:
        self.error: BaseException | None = None
        base, nbytes = ws.data_ptr(), ws.numel()

        def cb(_user, buf, n, _stream):
            try:
                off = buf - base
                if off < 0 or off + 8 * n > nbytes:
                    raise _lib.KeiseiB200Error("BatchNorm sync buffer lies outside the workspace")
                sync.all_reduce_(ws[off:off + 8 * n].view(torch.float64))
                return 0
            except BaseException as e:  # noqa: BLE001
                self.error = e
                return 1

        self.fn = _HOOK_T(cb)
        self.ptr = ctypes.cast(self.fn, c_void_p)

    def check(self) -> None:
        if self.error is not None:
            raise self.error


class _NativeHook:
    """A sync object that brings its own C hook (`c_hook` function pointer, `c_user` context pointer): the schedule calls
    straight into the library (distributed.PeerBatchNormSync -> kb_peer_allreduce_hook), no Python in the loop."""

    def __init__(self, sync) -> None:
        self.ptr, self.user = sync.c_hook, sync.c_user

    def check(self) -> None:
        return None   # the kernels are only enqueued here; a lost peer is reported by sync.check() after the step's host read


_BUCKET_HOOK_T = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p)


class _BucketHook:
    """Overlapped, bucketed gradient all-reduce (reference katago_loop.py:498-504: DDP's 25 MB buckets). The backward
    schedule calls back as each run of parameter gradients has been enqueued (heads, then blocks last to first); blocks are
    contiguous in the flat gradient, so consecutive callbacks extend one contiguous range, which is all-reduced on a
    communication stream — behind events recorded on the schedule's two streams — as soon as it reaches `bucket_bytes`.
    `finish()` reduces what is left (the stem and the last partial bucket) and joins everything back into the caller's
    stream. Exceptions cannot cross the C frame: they are parked and re-raised by `finish()`."""

    def __init__(self, flat: torch.Tensor, offsets: list[int], grad_sync) -> None:
        import torch.distributed as dist
        self.flat, self.offsets, self.sync, self.dist = flat, offsets, grad_sync, dist
        self.dev = flat.device
        self.comm = grad_sync.comm_stream(self.dev)
        self.works: list = []
        self.error: BaseException | None = None
        self.lo = self.hi = None            # pending contiguous range [lo, hi) of flat elements
        self.fired: list[tuple[int, int]] = []   # ranges already handed to a collective
        self.bucket_elems = max(1, int(grad_sync.bucket_bytes) // 4)
        self.launched = 0

        def cb(_user, _bucket, first_param, n_params, main_stream, side_stream):
            try:
                lo, hi = self.offsets[first_param], self.offsets[first_param + n_params]
                if self.lo is None:
                    self.lo, self.hi = lo, hi
                elif hi == self.lo:
                    self.lo = lo
                elif lo == self.hi:
                    self.hi = hi
                else:                       # not contiguous with the pending range: flush it first
                    self._fire(main_stream, side_stream)
                    self.lo, self.hi = lo, hi
                if self.hi - self.lo >= self.bucket_elems:
                    self._fire(main_stream, side_stream)
                return 0
            except BaseException as e:  # noqa: BLE001
                self.error = e
                return 1

        self.fn = _BUCKET_HOOK_T(cb)
        self.ptr = ctypes.cast(self.fn, c_void_p)

    def _stream(self, handle):
        """cudaStream_t handle from the C side -> torch stream (NULL is the legacy default stream); wrappers are cached."""
        if not handle:
            return torch.cuda.default_stream(self.dev)
        cache = self.sync.__dict__.setdefault("_ext_streams", {})
        s = cache.get((self.dev, int(handle)))
        if s is None:
            s = cache[(self.dev, int(handle))] = torch.cuda.ExternalStream(int(handle), device=self.dev)
        return s

    def _fire(self, main_stream, side_stream) -> None:
        if self.lo is None or self.hi <= self.lo:
            return
        ev_main, ev_side = torch.cuda.Event(), torch.cuda.Event()
        ev_main.record(self._stream(main_stream))
        self.comm.wait_event(ev_main)
        if side_stream and side_stream != main_stream:
            ev_side.record(self._stream(side_stream))
            self.comm.wait_event(ev_side)
        with torch.cuda.stream(self.comm):
            self.works.append(self.dist.all_reduce(self.flat[self.lo:self.hi], op=self.dist.ReduceOp.SUM, group=self.sync.group,
                                                   async_op=True))
        self.launched += 1
        self.fired.append((self.lo, self.hi))
        self.lo = self.hi = None

    def finish(self) -> None:
        """After the C call returned: reduce every range not yet handed to a collective and join the caller's stream."""
        if self.error is not None:
            raise self.error
        cur = torch.cuda.current_stream(self.dev)
        # the flat buffer is [stem | blocks 0..nb-1 | heads]; callbacks came heads first, then blocks in descending order:
        # what is still missing is the pending range and whatever no callback covered (the stem) — the gaps of `fired`
        if self.lo is not None:
            self._fire(cur.cuda_stream, 0)
        gaps, pos = [], 0
        for lo, hi in sorted(self.fired):
            if lo > pos:
                gaps.append((pos, lo))
            pos = max(pos, hi)
        if pos < self.flat.numel():
            gaps.append((pos, self.flat.numel()))
        for lo, hi in gaps:
            self.lo, self.hi = lo, hi
            self._fire(cur.cuda_stream, 0)
        for w in self.works:
            w.wait()                      # the caller's current stream waits for the collective
        self.works.clear()


def _sync_args(ws: torch.Tensor, sync):
    """(hook object | None, hook pointer | None, hook user pointer | None, world) for the *_sync entry points."""
    if sync is None or int(sync.world_size) <= 1:
        return None, None, None, 1
    if getattr(sync, "c_hook", None) is not None:
        h = _NativeHook(sync)
        return h, h.ptr, h.user, int(sync.world_size)
    h = _BnHook(ws, sync)
    return h, h.ptr, None, int(sync.world_size)


def sm_count(device: torch.device) -> int:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _sm_count_cache:
        _sm_count_cache[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _sm_count_cache[idx]


def _desc(desc: List[int]) -> SeResnetDesc:
    return SeResnetDesc(*[int(v) for v in desc])


def _ptr_table(tensors) -> ctypes.Array:
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def wpack_bytes(desc: List[int], dtype_code: int) -> int:
    d = _desc(desc)
    n = _lib.load().kb_seresnet_wpack_bytes(ctypes.byref(d), dtype_code)
    if n < 0:
        _lib.check(-1, "kb_seresnet_wpack_bytes")
    return int(n)


def workspace_bytes(desc: List[int], B: int, training: bool, dtype_code: int) -> int:
    d = _desc(desc)
    n = _lib.load().kb_seresnet_workspace_bytes(ctypes.byref(d), B, 1 if training else 0, dtype_code)
    if n < 0:
        _lib.check(-1, "kb_seresnet_workspace_bytes")
    return int(n)


def _check_tables(params, buffers, d: SeResnetDesc):
    n_p, n_b = 16 + 14 * d.num_blocks, 6 + 6 * d.num_blocks
    if len(params) != n_p or len(buffers) != n_b:
        raise ValueError(f"expected {n_p} params / {n_b} buffers, got {len(params)} / {len(buffers)}")
    for t in params:
        if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
            raise ValueError("model parameters must be contiguous float32 CUDA tensors")


@torch.no_grad()
def pack_weights(params, buffers, desc: List[int], dtype_code: int, wpack: torch.Tensor) -> None:
    d = _desc(desc)
    _check_tables(params, buffers, d)
    dev = wpack.device
    pt, bt = _ptr_table(params), _ptr_table(buffers)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_seresnet_pack_weights(ctypes.byref(d), pt, bt, dtype_code, wpack.data_ptr(),
                                                  wpack.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "kb_seresnet_pack_weights")


class PointerTables:
    """Host arrays of device pointers for a model's parameters / buffers, built once and reused while
    the tensors keep their storage (building them costs ~1 ms for the 822 tensors of a 40-block net)."""

    def __init__(self, params, buffers, desc: List[int]) -> None:
        self.desc = _desc(desc)
        _check_tables(params, buffers, self.desc)
        self.params, self.buffers = params, buffers
        self.pt, self.bt = _ptr_table(params), _ptr_table(buffers)
        self._probe = (params[0].data_ptr(), params[-1].data_ptr(), buffers[0].data_ptr(), buffers[-1].data_ptr())

    def valid(self) -> bool:
        p, b = self.params, self.buffers
        return self._probe == (p[0].data_ptr(), p[-1].data_ptr(), b[0].data_ptr(), b[-1].data_ptr())


@torch.no_grad()
def seresnet_forward_raw(obs: torch.Tensor, tables: PointerTables, wpack: torch.Tensor, training: bool, dtype_code: int,
                         use_tc: bool, bn_sync=None, out=None, num_sms: int | None = None):
    """The C call without the torch.library dispatcher (no-grad callers: rollout, the fused trainer step).
    Returns (policy_buf, value_logits, score_lead, workspace, new_stats). `out` = (policy_buf, value, score) row
    slices of caller-owned buffers to write into (the two-stream rollout runs half batches side by side).
    `num_sms`: the SM budget the persistent kernels size their grids for (default: the whole device); the grouped
    multi-model rollout gives every branch its share so that the branches are co-resident."""
    if not obs.is_cuda:
        raise _lib.KeiseiB200Error("keisei_b200 seresnet_forward needs CUDA tensors")
    d = tables.desc
    dev = obs.device
    B = obs.shape[0]
    obs_c = obs.detach().to(torch.float32).contiguous()
    ws = torch.empty(int(_lib.load().kb_seresnet_workspace_bytes(ctypes.byref(d), B, 1 if training else 0, dtype_code)),
                     dtype=torch.uint8, device=dev)
    if out is None:
        policy = torch.empty((B, POLICY_PITCH), dtype=_DT_INV[dtype_code], device=dev)
        value = torch.empty((B, 3), dtype=torch.float32, device=dev)
        score = torch.empty((B, 1), dtype=torch.float32, device=dev)
    else:
        policy, value, score = out
        if (policy.shape != (B, POLICY_PITCH) or policy.dtype != _DT_INV[dtype_code] or not policy.is_contiguous()
                or value.shape != (B, 3) or not value.is_contiguous() or score.shape != (B, 1) or not score.is_contiguous()):
            raise ValueError("seresnet_forward_raw: `out` buffers have the wrong shape / dtype / layout")
    policy[:, POLICY_A:].zero_()
    cmax = max(d.channels, d.policy_channels)
    new_stats = torch.empty((2 * d.num_blocks + 2, 2, cmax) if training else (0,), dtype=torch.float32, device=dev)
    hook, hook_ptr, hook_user, world = _sync_args(ws, bn_sync if training else None)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_seresnet_forward_sync(
            ctypes.byref(d), tables.pt, tables.bt, new_stats.data_ptr() if training else None, wpack.data_ptr(),
            obs_c.data_ptr(), B, 1 if training else 0, dtype_code, ws.data_ptr(), ws.numel(), policy.data_ptr(),
            POLICY_PITCH, value.data_ptr(), score.data_ptr(), 1 if use_tc else 0,
            sm_count(dev) if num_sms is None else max(2, min(int(num_sms), sm_count(dev))), hook_ptr, hook_user, world,
            _lib.stream_ptr(dev))
    if hook is not None:
        hook.check()
    _lib.check(rc, "kb_seresnet_forward")
    return policy, value, score, ws, new_stats


@torch.no_grad()
def seresnet_backward_raw(tables: PointerTables, wpack: torch.Tensor, ws: torch.Tensor, dpolicy: torch.Tensor,
                          dvalue: torch.Tensor, dscore: torch.Tensor, dtype_code: int, use_tc: bool,
                          sizes: List[int] | None = None, bn_sync=None, grad_sync=None) -> torch.Tensor:
    """The C backward without the dispatcher. Returns the flat fp32 gradient (parameter-table order). With `grad_sync`
    (a `distributed.GradSync` of more than one rank, overlap enabled) the gradient comes back all-reduced (SUMMED over
    the ranks; the caller divides): buckets are reduced on a communication stream while the backward is still running."""
    d = tables.desc
    dev = ws.device
    B = dvalue.shape[0]
    dpol = dpolicy
    if dpol.dtype != _DT_INV[dtype_code] or dpol.stride(1) != 1 or dpol.stride(0) < POLICY_A:
        dpol = dpol.to(_DT_INV[dtype_code]).contiguous()
    dv = dvalue.to(torch.float32).contiguous()
    ds = dscore.to(torch.float32).reshape(B).contiguous()
    if sizes is None:
        sizes = [p.numel() for p in tables.params]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    gt = (c_void_p * len(sizes))()
    base, off = flat.data_ptr(), 0
    for i, n in enumerate(sizes):
        gt[i] = base + 4 * off
        off += n
    hook, hook_ptr, hook_user, world = _sync_args(ws, bn_sync)
    bucket = None
    if grad_sync is not None and int(grad_sync.world_size) > 1 and getattr(grad_sync, "overlap", False):
        offsets = [0]
        for n in sizes:
            offsets.append(offsets[-1] + n)
        bucket = _BucketHook(flat, offsets, grad_sync)
    lib = _lib.load()
    with torch.cuda.device(dev):
        if bucket is not None:
            lib.kb_seresnet_set_bucket_hook(bucket.ptr, None)
        try:
            rc = lib.kb_seresnet_backward_sync(
                ctypes.byref(d), tables.pt, wpack.data_ptr(), B, dtype_code, ws.data_ptr(), ws.numel(), dpol.data_ptr(),
                dpol.stride(0), dv.data_ptr(), ds.data_ptr(), gt, 1 if use_tc else 0, sm_count(dev), hook_ptr, hook_user, world,
                _lib.stream_ptr(dev))
        finally:
            if bucket is not None:
                lib.kb_seresnet_set_bucket_hook(None, None)
    if hook is not None:
        hook.check()
    if bucket is not None and bucket.error is not None:
        raise bucket.error
    _lib.check(rc, "kb_seresnet_backward")
    if bucket is not None:
        bucket.finish()
        grad_sync.last_overlap_buckets = bucket.launched
    return flat


@torch.library.custom_op("keisei_b200::seresnet_forward", mutates_args=())
def seresnet_forward(obs: torch.Tensor, params: List[torch.Tensor], buffers: List[torch.Tensor], wpack: torch.Tensor,
                     desc: List[int], training: bool, dtype_code: int, use_tc: bool,
                     bn_sync: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (policy_buf (B, 11264) act-dtype, value_logits (B,3) f32, score_lead (B,1) f32, workspace u8,
    new_stats f32 [(2*nb+2), 2, Cmax]: the updated BatchNorm running mean/var rows in training mode).
    Functional: `buffers` is read-only here; the module copies new_stats back (see SEResNetModel)."""
    if not obs.is_cuda:
        raise _lib.KeiseiB200Error("keisei_b200::seresnet_forward needs CUDA tensors")
    d = _desc(desc)
    _check_tables(params, buffers, d)
    dev = obs.device
    B = obs.shape[0]
    obs_c = obs.detach().to(torch.float32).contiguous()
    ws = torch.empty(workspace_bytes(desc, B, training, dtype_code), dtype=torch.uint8, device=dev)
    policy = torch.empty((B, POLICY_PITCH), dtype=_DT_INV[dtype_code], device=dev)
    policy[:, POLICY_A:].zero_()
    value = torch.empty((B, 3), dtype=torch.float32, device=dev)
    score = torch.empty((B, 1), dtype=torch.float32, device=dev)
    cmax = max(d.channels, d.policy_channels)
    new_stats = torch.empty((2 * d.num_blocks + 2, 2, cmax) if training else (0,), dtype=torch.float32, device=dev)
    pt, bt = _ptr_table(params), _ptr_table(buffers)
    hook, hook_ptr, hook_user, world = _sync_args(ws, _BN_SYNCS[bn_sync] if (bn_sync and training) else None)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_seresnet_forward_sync(
            ctypes.byref(d), pt, bt, new_stats.data_ptr() if training else None, wpack.data_ptr(), obs_c.data_ptr(), B, 1 if training else 0, dtype_code,
            ws.data_ptr(), ws.numel(), policy.data_ptr(), POLICY_PITCH, value.data_ptr(), score.data_ptr(),
            1 if use_tc else 0, sm_count(dev), hook_ptr, hook_user, world, _lib.stream_ptr(dev))
    if hook is not None:
        hook.check()
    _lib.check(rc, "kb_seresnet_forward")
    return policy, value, score, ws, new_stats


@seresnet_forward.register_fake
def _(obs, params, buffers, wpack, desc, training, dtype_code, use_tc, bn_sync):
    B = obs.shape[0]
    return (obs.new_empty((B, POLICY_PITCH), dtype=_DT_INV[dtype_code]), obs.new_empty((B, 3), dtype=torch.float32),
            obs.new_empty((B, 1), dtype=torch.float32), obs.new_empty((1,), dtype=torch.uint8),
            obs.new_empty((1,), dtype=torch.float32))


@torch.library.custom_op("keisei_b200::seresnet_backward", mutates_args=())
def seresnet_backward(params: List[torch.Tensor], wpack: torch.Tensor, ws: torch.Tensor, dpolicy: torch.Tensor,
                      dvalue: torch.Tensor, dscore: torch.Tensor, desc: List[int], dtype_code: int,
                      use_tc: bool, bn_sync: int) -> torch.Tensor:
    """Returns ONE flat fp32 gradient buffer (parameters concatenated in table order)."""
    d = _desc(desc)
    dev = ws.device
    B = dvalue.shape[0]
    dpol = dpolicy
    if dpol.dtype != _DT_INV[dtype_code] or dpol.stride(1) != 1 or dpol.stride(0) < POLICY_A:
        dpol = dpol.to(_DT_INV[dtype_code]).contiguous()
    dv = dvalue.to(torch.float32).contiguous()
    ds = dscore.to(torch.float32).reshape(B).contiguous()
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    grads = [g.view(p.shape) for g, p in zip(flat.split(sizes), params)]
    pt, gt = _ptr_table(params), _ptr_table(grads)
    hook, hook_ptr, hook_user, world = _sync_args(ws, _BN_SYNCS[bn_sync] if bn_sync else None)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_seresnet_backward_sync(
            ctypes.byref(d), pt, wpack.data_ptr(), B, dtype_code, ws.data_ptr(), ws.numel(), dpol.data_ptr(),
            dpol.stride(0), dv.data_ptr(), ds.data_ptr(), gt, 1 if use_tc else 0, sm_count(dev), hook_ptr, hook_user, world,
            _lib.stream_ptr(dev))
    if hook is not None:
        hook.check()
    _lib.check(rc, "kb_seresnet_backward")
    return flat


@seresnet_backward.register_fake
def _(params, wpack, ws, dpolicy, dvalue, dscore, desc, dtype_code, use_tc, bn_sync):
    return params[0].new_empty((sum(p.numel() for p in params),), dtype=torch.float32)


def _fwd_setup(ctx, inputs, output):
    obs, params, buffers, wpack, desc, training, dtype_code, use_tc, bn_sync = inputs
    policy, value, score, ws, _new_stats = output
    ctx.desc, ctx.dtype_code, ctx.use_tc, ctx.training, ctx.bn_sync = desc, dtype_code, use_tc, training, bn_sync
    ctx.n_params = len(params)
    ctx.n_buffers = len(buffers)
    ctx.save_for_backward(wpack, ws, *params)
    ctx.policy_meta = (policy.shape, policy.dtype, policy.device)
    ctx.B = obs.shape[0]


def _fwd_backward(ctx, g_policy, g_value, g_score, g_ws, g_stats):
    if not ctx.training:
        raise RuntimeError("keisei_b200::seresnet_forward was run in eval mode; backward needs training=True "
                           "(batch-statistics BatchNorm and saved activations)")
    wpack, ws, *params = ctx.saved_tensors
    shape, dtype, dev = ctx.policy_meta
    if g_policy is None:
        g_policy = torch.zeros(shape, dtype=dtype, device=dev)
    if g_value is None:
        g_value = torch.zeros((ctx.B, 3), dtype=torch.float32, device=dev)
    if g_score is None:
        g_score = torch.zeros((ctx.B, 1), dtype=torch.float32, device=dev)
    flat = seresnet_backward(list(params), wpack, ws, g_policy, g_value, g_score, ctx.desc, ctx.dtype_code, ctx.use_tc,
                             ctx.bn_sync)
    grads = [g.view(p.shape) for g, p in zip(flat.split([p.numel() for p in params]), params)]
    return None, grads, [None] * ctx.n_buffers, None, None, None, None, None, None


seresnet_forward.register_autograd(_fwd_backward, setup_context=_fwd_setup)


# ---- single-conv helpers (tests / profiling) -------------------------------------------------
@torch.no_grad()
def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype, cin_pad: int | None = None, with_dgrad: bool = False):
    """(Cout,Cin,3,3) fp32 -> wf (Cout,9,Cinp) [, wd (Cinp,9,Cout)] in `dtype`."""
    Cout, Cin = w.shape[0], w.shape[1]
    Cinp = cin_pad or Cin
    wf = torch.empty((Cout, 9, Cinp), dtype=dtype, device=w.device)
    wd = torch.empty((Cinp, 9, Cout), dtype=dtype, device=w.device) if with_dgrad else None
    with torch.cuda.device(w.device):
        rc = _lib.load().kb_pack_conv_weight(w.contiguous().data_ptr(), wf.data_ptr(), _lib.ptr(wd), Cout, Cin, Cinp,
                                             _DT[dtype], _lib.stream_ptr(w.device))
    _lib.check(rc, "kb_pack_conv_weight")
    return (wf, wd) if with_dgrad else wf


@torch.no_grad()
def conv3x3(x: torch.Tensor, wf: torch.Tensor, backend: int = 0, scale=None, shift=None, relu: bool = False,
            gbias=None, want_sums: bool = False, want_board_mean: bool = False, want_pool: bool = False):
    """x (B,81,Cin) NHWC, wf (Cout,9,Cin). Returns (out, ch_sums|None, board_mean|None, pool|None)."""
    B, _, Cin = x.shape
    Cout = wf.shape[0]
    dev = x.device
    out = torch.empty((B, 81, Cout), dtype=x.dtype, device=dev)
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device=dev) if want_sums else None
    bm = torch.empty((B, Cout), dtype=torch.float32, device=dev) if want_board_mean else None
    pool = torch.empty((B, 3 * Cout), dtype=torch.float32, device=dev) if want_pool else None
    with torch.cuda.device(dev):
        rc = _lib.load().kb_conv3x3_forward(
            x.contiguous().data_ptr(), wf.contiguous().data_ptr(), out.data_ptr(), B, Cin, Cout, _DT[x.dtype], backend,
            _lib.ptr(scale), _lib.ptr(shift), 1 if relu else 0, _lib.ptr(gbias), _lib.ptr(sums), _lib.ptr(bm),
            _lib.ptr(pool), sm_count(dev), _lib.stream_ptr(dev))
    _lib.check(rc, "kb_conv3x3_forward")
    return out, sums, bm, pool


@torch.no_grad()
def conv3x3_se_tail(x: torch.Tensor, wf: torch.Tensor, scale, shift, res: torch.Tensor, w1, b1, w2, b2):
    """Evaluation-mode conv2 + block tail in ONE kernel (csrc/conv_tc.cu, CTA-pair kernel with the fused SE epilogue).
    x, res (B,81,C) bf16; wf (256,9,C) bf16; scale/shift (256,) folded BatchNorm; SE weights fp32.
    Returns (out (B,81,256) bf16, pool (B,768) fp32, pool_bf16 (B,768))."""
    B, _, Cin = x.shape
    Cout, S = wf.shape[0], w1.shape[0]
    dev = x.device
    out = torch.empty((B, 81, Cout), dtype=torch.bfloat16, device=dev)
    pool = torch.empty((B, 3 * Cout), dtype=torch.float32, device=dev)
    pool_bf = torch.empty((B, 3 * Cout), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_conv3x3_se_tail(
            x.contiguous().data_ptr(), wf.contiguous().data_ptr(), out.data_ptr(), B, Cin, Cout, scale.contiguous().data_ptr(),
            shift.contiguous().data_ptr(), res.contiguous().data_ptr(), w1.contiguous().data_ptr(), b1.contiguous().data_ptr(),
            w2.contiguous().data_ptr(), b2.contiguous().data_ptr(), S, pool.data_ptr(), pool_bf.data_ptr(), sm_count(dev),
            _lib.stream_ptr(dev))
    _lib.check(rc, "kb_conv3x3_se_tail")
    return out, pool, pool_bf


@torch.no_grad()
def conv3x3_wgrad(x: torch.Tensor, dy: torch.Tensor, cin_true: int | None = None, backend: int = 0) -> torch.Tensor:
    """dW (Cout, Cin_true, 3, 3) fp32 from x (B,81,Cin) and dy (B,81,Cout)."""
    B, _, Cin = x.shape
    Cout = dy.shape[2]
    ct = cin_true or Cin
    dw = torch.zeros((Cout, ct, 3, 3), dtype=torch.float32, device=x.device)
    ws = None
    if backend == 1:
        ws = torch.empty(int(_lib.load().kb_conv3x3_wgrad_ws_bytes(Cin, Cout, sm_count(x.device))), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().kb_conv3x3_wgrad(x.contiguous().data_ptr(), dy.contiguous().data_ptr(), dw.data_ptr(), B, Cin,
                                          Cout, ct, _DT[x.dtype], backend, _lib.ptr(ws), 0 if ws is None else ws.numel(),
                                          sm_count(x.device), _lib.stream_ptr(x.device))
    _lib.check(rc, "kb_conv3x3_wgrad")
    return dw


@torch.no_grad()
def linear_tc(x: torch.Tensor, w: torch.Tensor, bias=None, scale=None, relu: bool = False, want_f32: bool = True,
              want_bf16: bool = False):
    """Y = act((x @ w.T) * scale + bias) on the tcgen05 path. x (M,K) any float dtype (cast to bf16, K padded to 64),
    w (N,K) fp32. Returns (y_f32 (M,N) | None, y_bf16 (M, ceil64(N)) | None)."""
    M, K = x.shape
    N = w.shape[0]
    Kp, Np, Nb = (K + 63) // 64 * 64, (N + 127) // 128 * 128, (N + 63) // 64 * 64
    dev = x.device
    xb = torch.zeros((M, Kp), dtype=torch.bfloat16, device=dev)
    xb[:, :K] = x
    wp = torch.empty((Np, Kp), dtype=torch.bfloat16, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.kb_pack_linear_weight(w.float().contiguous().data_ptr(), wp.data_ptr(), N, K, Np, Kp, _lib.stream_ptr(dev)),
                   "kb_pack_linear_weight")
        yf = torch.empty((M, N), dtype=torch.float32, device=dev) if want_f32 else None
        yb = torch.empty((M, Nb), dtype=torch.bfloat16, device=dev) if want_bf16 else None
        _lib.check(lib.kb_linear_tc(xb.data_ptr(), M, Kp, wp.data_ptr(), N, Np, _lib.ptr(scale), _lib.ptr(bias), 1 if relu else 0,
                                    _lib.ptr(yf), N, _lib.ptr(yb), Nb, Nb, 0, 0, sm_count(dev), _lib.stream_ptr(dev)),
                   "kb_linear_tc")
    return yf, yb


@torch.no_grad()
def se_block_tail(z: torch.Tensor, res: torch.Tensor, board_mean: torch.Tensor, w1, b1, w2, b2, bn_a=None, bn_b=None,
                  want_ties: bool = False, se_raw: bool = True, variant: int = 0):
    """Fused SE MLP + scale/shift + residual + ReLU + pool statistics (csrc/se_apply.cu) on bf16 (B,81,C) tiles.
    Returns (out, pool (B,3C), ties|None, se_in (B,C), seh (B,S), se (B,2C))."""
    B, _, C = z.shape
    S = w1.shape[0]
    dev = z.device
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)  # noqa: E731
    out = torch.empty_like(z)
    pool, se_in, seh, se = f(B, 3 * C), f(B, C), f(B, S), f(B, 2 * C)
    ties = f(B, C) if want_ties else None
    with torch.cuda.device(dev):
        rc = _lib.load().kb_se_block_tail_variant(
            z.contiguous().data_ptr(), res.contiguous().data_ptr(), out.data_ptr(), _lib.ptr(bn_a), _lib.ptr(bn_b),
            board_mean.contiguous().data_ptr(), w1.contiguous().data_ptr(), b1.contiguous().data_ptr(),
            w2.contiguous().data_ptr(), b2.contiguous().data_ptr(), se_in.data_ptr(), seh.data_ptr(), se.data_ptr(),
            1 if se_raw else 0, pool.data_ptr(), None, _lib.ptr(ties), B, C, S, int(variant), sm_count(dev),
            _lib.stream_ptr(dev))
    _lib.check(rc, "kb_se_block_tail")
    return out, pool, ties, se_in, seh, se
