"""ctypes loader for libkeisei_b200.so — the C-ABI boundary (include/keisei_b200.h).

The library is loaded lazily and exactly once. There is no fallback: if a CUDA tensor reaches
one of the ops and the library is missing or fails to load, `KeiseiB200Error` is raised.
"""
from __future__ import annotations

import ctypes
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_ulonglong, c_void_p
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libkeisei_b200.so"
_lock = threading.Lock()
_lib: ctypes.CDLL | None = None


class KeiseiB200Error(RuntimeError):
    """The CUDA library is missing, failed to load, or an entry point returned an error."""


_P = c_void_p
_SIGS: dict[str, tuple[object, list[object]]] = {
    "kb_abi_version": (c_int, []),
    "kb_compiled_sm": (c_int, []),
    "kb_last_error": (c_char_p, []),
    "kb_launch_count": (c_ulonglong, []),
    "kb_device_sm_count": (c_int, [c_int]),
    "kb_gae_scan": (c_int, [_P, _P, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_double, c_double, c_int, _P]),
    "kb_advantage_normalize": (c_int, [_P, c_longlong, c_float, _P]),
    "kb_policy_sample": (c_int, [_P, c_int, c_longlong, _P, _P, _P, c_float, c_int, c_int, c_ulonglong,
                                 c_ulonglong, c_int, _P, _P, _P, _P, _P, _P, c_int, c_longlong, _P]),
    "kb_ppo_policy_fwd": (c_int, [_P, c_int, c_longlong, _P, _P, _P, _P, c_int, c_int, c_float,
                                  _P, _P, _P, _P, _P, _P, c_int, c_longlong, _P]),
    "kb_ppo_policy_bwd": (c_int, [_P, c_int, c_longlong, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P,
                                  c_longlong, c_int, c_longlong, _P]),
    "kb_flat_grad_stats": (c_int, [_P, c_longlong, _P, c_int, _P]),
    "kb_adam_step_flat": (c_int, [_P] * 9 + [c_int, c_int, c_int, _P, _P] + [c_float] * 6 + [_P, _P, _P]),
    "kb_pack_mask_bits": (c_int, [_P, _P, c_longlong, c_int, c_int, _P]),
    "kb_gather_minibatch": (c_int, [_P] * 9 + [c_longlong, c_int, c_int, c_int] + [_P] * 8 + [_P]),
    "kb_value_losses_fwd": (c_int, [_P, _P, _P, _P, c_int, _P, _P]),
    "kb_value_losses_bwd": (c_int, [_P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "kb_peer_buffer_bytes": (c_longlong, [c_int, c_int, c_longlong]),
    "kb_peer_buffer_create": (c_int, [c_longlong, _P, _P]),
    "kb_peer_buffer_open": (c_int, [_P, _P]),
    "kb_peer_buffer_close": (c_int, [_P]),
    "kb_peer_buffer_destroy": (c_int, [_P]),
    "kb_peer_status_create": (c_int, [_P]),
    "kb_peer_status_destroy": (c_int, [_P]),
    "kb_peer_allreduce_f64": (c_int, [_P, c_longlong, _P, c_ulonglong, _P]),
    "kb_peer_allreduce_emulate": (c_int, [_P, c_longlong, _P, c_int, _P]),
    "kb_peer_allreduce_hook": (c_int, [_P, _P, c_longlong, _P]),
}


def register_signature(name: str, restype: object, argtypes: list[object]) -> None:
    """Used by sibling modules that own further entry points (model, conv)."""
    _SIGS[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib_path() -> Path:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not _LIB_PATH.exists():
            raise KeiseiB200Error(
                f"{_LIB_PATH} not found — build it with `python -m keisei_b200.build` "
                "(or __graft_entry__.build()); there is no fallback for CUDA tensors")
        try:
            lib = ctypes.CDLL(str(_LIB_PATH))
        except OSError as e:  # pragma: no cover - environment specific
            raise KeiseiB200Error(f"failed to load {_LIB_PATH}: {e}") from e
        for name, (restype, argtypes) in _SIGS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise KeiseiB200Error(f"{_LIB_PATH} does not export {name}; stale build?") from e
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().kb_last_error()
        raise KeiseiB200Error(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().kb_launch_count())


def ptr(t) -> int | None:
    """data_ptr of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
