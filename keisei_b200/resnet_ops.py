"""torch.library custom ops for the plain ResNet baseline (csrc/resnet.cu; reference
keisei/training/models/resnet.py:25-84).

`keisei_b200::resnet_forward`   one C call that enqueues the whole network on the current stream
`keisei_b200::resnet_backward`  the matching backward (training-mode workspaces only)

Same conventions as `model_ops` (SE-ResNet): autograd is registered on the forward op, the
gradients come back as ONE flat fp32 buffer in parameter-table order, the observation never
receives a gradient.
"""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_longlong, c_void_p
from typing import List

import torch

from . import _lib
from .model_ops import POLICY_A, POLICY_PITCH, _DT_INV, _ptr_table, sm_count


class ResnetDesc(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("num_layers", "hidden_size", "obs_channels")]


_P = c_void_p
_lib.register_signature("kb_resnet_num_params", c_longlong, [_P])
_lib.register_signature("kb_resnet_num_buffers", c_longlong, [_P])
_lib.register_signature("kb_resnet_wpack_bytes", c_longlong, [_P, c_int])
_lib.register_signature("kb_resnet_workspace_bytes", c_longlong, [_P, c_int, c_int, c_int])
_lib.register_signature("kb_resnet_pack_weights", c_int, [_P, _P, _P, c_int, _P, c_longlong, _P])
_lib.register_signature("kb_resnet_forward", c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_longlong, _P,
                                                      c_longlong, _P, c_int, c_int, _P])
_lib.register_signature("kb_resnet_backward", c_int, [_P, _P, _P, c_int, c_int, _P, c_longlong, _P, c_longlong, _P, _P,
                                                       c_int, c_int, _P])


def _desc(desc: List[int]) -> ResnetDesc:
    return ResnetDesc(*[int(v) for v in desc])


def _check_tables(params, buffers, d: ResnetDesc) -> None:
    n_p, n_b = 15 + 6 * d.num_layers, 9 + 6 * d.num_layers
    if len(params) != n_p or len(buffers) != n_b:
        raise ValueError(f"expected {n_p} params / {n_b} buffers, got {len(params)} / {len(buffers)}")
    for t in params:
        if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
            raise ValueError("model parameters must be contiguous float32 CUDA tensors")


def wpack_bytes(desc: List[int], dtype_code: int) -> int:
    d = _desc(desc)
    n = _lib.load().kb_resnet_wpack_bytes(ctypes.byref(d), dtype_code)
    if n < 0:
        _lib.check(-1, "kb_resnet_wpack_bytes")
    return int(n)


def workspace_bytes(desc: List[int], B: int, training: bool, dtype_code: int) -> int:
    d = _desc(desc)
    n = _lib.load().kb_resnet_workspace_bytes(ctypes.byref(d), B, 1 if training else 0, dtype_code)
    if n < 0:
        _lib.check(-1, "kb_resnet_workspace_bytes")
    return int(n)


@torch.no_grad()
def pack_weights(params, buffers, desc: List[int], dtype_code: int, wpack: torch.Tensor) -> None:
    d = _desc(desc)
    _check_tables(params, buffers, d)
    dev = wpack.device
    pt, bt = _ptr_table(params), _ptr_table(buffers)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_resnet_pack_weights(ctypes.byref(d), pt, bt, dtype_code, wpack.data_ptr(), wpack.numel(),
                                                _lib.stream_ptr(dev))
    _lib.check(rc, "kb_resnet_pack_weights")


class PointerTables:
    """Host arrays of device pointers for the model's parameters / buffers (see model_ops.PointerTables)."""

    def __init__(self, params, buffers, desc: List[int]) -> None:
        self.desc = _desc(desc)
        _check_tables(params, buffers, self.desc)
        self.params, self.buffers = params, buffers
        self.pt, self.bt = _ptr_table(params), _ptr_table(buffers)
        self._probe = (params[0].data_ptr(), params[-1].data_ptr(), buffers[0].data_ptr(), buffers[-1].data_ptr())

    def valid(self) -> bool:
        p, b = self.params, self.buffers
        return self._probe == (p[0].data_ptr(), p[-1].data_ptr(), b[0].data_ptr(), b[-1].data_ptr())


def _forward_call(d: ResnetDesc, pt, bt, obs: torch.Tensor, wpack: torch.Tensor, training: bool, dtype_code: int, use_tc: bool):
    dev = obs.device
    B = obs.shape[0]
    obs_c = obs.detach().to(torch.float32).contiguous()
    ws = torch.empty(int(_lib.load().kb_resnet_workspace_bytes(ctypes.byref(d), B, 1 if training else 0, dtype_code)),
                     dtype=torch.uint8, device=dev)
    policy = torch.empty((B, POLICY_PITCH), dtype=_DT_INV[dtype_code], device=dev)
    policy[:, POLICY_A:].zero_()
    value = torch.empty((B, 1), dtype=torch.float32, device=dev)
    new_stats = torch.empty((2 * d.num_layers + 3, 2, d.hidden_size) if training else (0,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_resnet_forward(
            ctypes.byref(d), pt, bt, new_stats.data_ptr() if training else None, wpack.data_ptr(), obs_c.data_ptr(), B,
            1 if training else 0, dtype_code, ws.data_ptr(), ws.numel(), policy.data_ptr(), POLICY_PITCH, value.data_ptr(),
            1 if use_tc else 0, sm_count(dev), _lib.stream_ptr(dev))
    _lib.check(rc, "kb_resnet_forward")
    return policy, value, ws, new_stats


@torch.no_grad()
def resnet_forward_raw(obs: torch.Tensor, tables: PointerTables, wpack: torch.Tensor, training: bool, dtype_code: int,
                       use_tc: bool):
    """The C call without the dispatcher. Returns (policy_buf (B,11264), value (B,1) f32, workspace, new_stats)."""
    if not obs.is_cuda:
        raise _lib.KeiseiB200Error("keisei_b200 resnet_forward needs CUDA tensors")
    return _forward_call(tables.desc, tables.pt, tables.bt, obs, wpack, training, dtype_code, use_tc)


def _backward_call(d: ResnetDesc, pt, sizes: List[int], wpack, ws, dpolicy, dvalue, dtype_code: int, use_tc: bool) -> torch.Tensor:
    dev = ws.device
    B = dvalue.shape[0]
    dpol = dpolicy
    if dpol.dtype != _DT_INV[dtype_code] or dpol.stride(1) != 1 or dpol.stride(0) < POLICY_A:
        dpol = dpol.to(_DT_INV[dtype_code]).contiguous()
    dv = dvalue.to(torch.float32).reshape(B).contiguous()
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    gt = (c_void_p * len(sizes))()
    base, off = flat.data_ptr(), 0
    for i, n in enumerate(sizes):
        gt[i] = base + 4 * off
        off += n
    with torch.cuda.device(dev):
        rc = _lib.load().kb_resnet_backward(
            ctypes.byref(d), pt, wpack.data_ptr(), B, dtype_code, ws.data_ptr(), ws.numel(), dpol.data_ptr(), dpol.stride(0),
            dv.data_ptr(), gt, 1 if use_tc else 0, sm_count(dev), _lib.stream_ptr(dev))
    _lib.check(rc, "kb_resnet_backward")
    return flat


@torch.no_grad()
def resnet_backward_raw(tables: PointerTables, wpack, ws, dpolicy, dvalue, dtype_code: int, use_tc: bool,
                        sizes: List[int] | None = None) -> torch.Tensor:
    if sizes is None:
        sizes = [p.numel() for p in tables.params]
    return _backward_call(tables.desc, tables.pt, sizes, wpack, ws, dpolicy, dvalue, dtype_code, use_tc)


@torch.library.custom_op("keisei_b200::resnet_forward", mutates_args=())
def resnet_forward(obs: torch.Tensor, params: List[torch.Tensor], buffers: List[torch.Tensor], wpack: torch.Tensor,
                   desc: List[int], training: bool, dtype_code: int,
                   use_tc: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (policy_buf (B, 11264) act-dtype, value (B,1) f32 tanh, workspace u8, new_stats f32)."""
    if not obs.is_cuda:
        raise _lib.KeiseiB200Error("keisei_b200::resnet_forward needs CUDA tensors")
    d = _desc(desc)
    _check_tables(params, buffers, d)
    return _forward_call(d, _ptr_table(params), _ptr_table(buffers), obs, wpack, training, dtype_code, use_tc)


@resnet_forward.register_fake
def _(obs, params, buffers, wpack, desc, training, dtype_code, use_tc):
    B = obs.shape[0]
    return (obs.new_empty((B, POLICY_PITCH), dtype=_DT_INV[dtype_code]), obs.new_empty((B, 1), dtype=torch.float32),
            obs.new_empty((1,), dtype=torch.uint8), obs.new_empty((1,), dtype=torch.float32))


@torch.library.custom_op("keisei_b200::resnet_backward", mutates_args=())
def resnet_backward(params: List[torch.Tensor], wpack: torch.Tensor, ws: torch.Tensor, dpolicy: torch.Tensor,
                    dvalue: torch.Tensor, desc: List[int], dtype_code: int, use_tc: bool) -> torch.Tensor:
    """Returns ONE flat fp32 gradient buffer (parameters concatenated in table order)."""
    return _backward_call(_desc(desc), _ptr_table(params), [p.numel() for p in params], wpack, ws, dpolicy, dvalue,
                          dtype_code, use_tc)


@resnet_backward.register_fake
def _(params, wpack, ws, dpolicy, dvalue, desc, dtype_code, use_tc):
    return params[0].new_empty((sum(p.numel() for p in params),), dtype=torch.float32)


def _fwd_setup(ctx, inputs, output):
    obs, params, buffers, wpack, desc, training, dtype_code, use_tc = inputs
    policy, _value, ws, _new_stats = output
    ctx.desc, ctx.dtype_code, ctx.use_tc, ctx.training = desc, dtype_code, use_tc, training
    ctx.n_buffers = len(buffers)
    ctx.save_for_backward(wpack, ws, *params)
    ctx.policy_meta = (policy.shape, policy.dtype, policy.device)
    ctx.B = obs.shape[0]


def _fwd_backward(ctx, g_policy, g_value, g_ws, g_stats):
    if not ctx.training:
        raise RuntimeError("keisei_b200::resnet_forward was run in eval mode; backward needs training=True "
                           "(batch-statistics BatchNorm and saved activations)")
    wpack, ws, *params = ctx.saved_tensors
    shape, dtype, dev = ctx.policy_meta
    if g_policy is None:
        g_policy = torch.zeros(shape, dtype=dtype, device=dev)
    if g_value is None:
        g_value = torch.zeros((ctx.B, 1), dtype=torch.float32, device=dev)
    flat = resnet_backward(list(params), wpack, ws, g_policy, g_value, ctx.desc, ctx.dtype_code, ctx.use_tc)
    grads = [g.view(p.shape) for g, p in zip(flat.split([p.numel() for p in params]), params)]
    return None, grads, [None] * ctx.n_buffers, None, None, None, None, None


resnet_forward.register_autograd(_fwd_backward, setup_context=_fwd_setup)
