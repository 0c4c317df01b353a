"""The optimiser tail of a fused KataGo-PPO step in two kernel launches (csrc/optim.cu): GradScaler inf-check / unscale,
global-norm clip and the Adam update over the ONE flat fp32 gradient buffer the backward produces (reference
keisei/training/katago_ppo.py:494-495, :926-933).

The optimizer object stays `torch.optim.Adam` — the loop replaces `.optimizer` at seat rotation (katago_loop.py:1859) and
checkpoints save / restore `optimizer.state_dict()` positionally (checkpoint.py:123) — and its state stays in ordinary
per-parameter tensors (`exp_avg`, `exp_avg_sq`, `step`): fresh state is created here as views of three flat buffers, state
loaded from a checkpoint is used where it is. Only the arithmetic moves: instead of ~50 foreach / multi-tensor launches,
`kb_flat_grad_stats` (one read of the gradient) and `kb_adam_step_flat` (one pass over p, g, m, v).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_CHUNK = 1 << 16


class FlatAdamTail:
    """Per-trainer cache of the device tables (parameter / moment pointers, offsets, chunk list) + the launch."""

    def __init__(self) -> None:
        self._key = None
        self._tables: dict | None = None

    @staticmethod
    def supports(optimizer: torch.optim.Optimizer, params: list[torch.Tensor], flat: torch.Tensor) -> bool:
        if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1 or not flat.is_cuda:
            return False
        g = optimizer.param_groups[0]
        if g.get("amsgrad") or g.get("maximize") or g.get("weight_decay", 0) != 0 or g.get("differentiable") or g.get("capturable"):
            return False
        if isinstance(g["lr"], torch.Tensor) or flat.dtype != torch.float32 or flat.data_ptr() % 16 != 0:
            return False
        return all(p.dtype == torch.float32 and p.is_contiguous() and p.device == flat.device for p in params)

    def _build(self, optimizer, params, dev) -> dict:
        state = optimizer.state
        sizes = [p.numel() for p in params]
        total = sum(sizes)
        fresh = all(len(state[p]) == 0 for p in params) if all(p in state for p in params) else not any(p in state and len(state[p]) for p in params)
        steps = torch.zeros(len(params), dtype=torch.float32, device=dev)
        if fresh:
            m_flat, v_flat = torch.zeros(total, device=dev), torch.zeros(total, device=dev)
            off = 0
            for i, (p, n) in enumerate(zip(params, sizes)):
                state[p]["step"] = steps[i]
                state[p]["exp_avg"] = m_flat[off:off + n].view_as(p)
                state[p]["exp_avg_sq"] = v_flat[off:off + n].view_as(p)
                off += n
        else:   # state restored from a checkpoint (or created by a stock step): use the tensors where they are
            for i, p in enumerate(params):
                st = state[p]
                if len(st) == 0:
                    st["step"] = steps[i]
                    st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
                    continue
                for k in ("exp_avg", "exp_avg_sq"):
                    if st[k].device != dev or st[k].dtype != torch.float32 or not st[k].is_contiguous():
                        st[k] = st[k].to(device=dev, dtype=torch.float32).contiguous()
                steps[i] = float(st["step"])
                st["step"] = steps[i]
        ptr = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)   # noqa: E731
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += n
        chunk_tensor, chunk_start = [], []
        for i, n in enumerate(sizes):
            for s in range(0, n, _CHUNK):
                chunk_tensor.append(i)
                chunk_start.append(s)
        return {"p": ptr(params), "m": ptr([state[p]["exp_avg"] for p in params]), "v": ptr([state[p]["exp_avg_sq"] for p in params]),
                "steps": steps, "g_off": torch.tensor(offs, dtype=torch.int64, device=dev),
                "sizes": torch.tensor(sizes, dtype=torch.int64, device=dev),
                "chunk_tensor": torch.tensor(chunk_tensor, dtype=torch.int32, device=dev),
                "chunk_start": torch.tensor(chunk_start, dtype=torch.int64, device=dev), "n_chunks": len(chunk_tensor),
                "stats": torch.zeros(2, dtype=torch.float64, device=dev), "out": torch.zeros(2, dtype=torch.float32, device=dev),
                "total": total}

    def _key_of(self, optimizer, params) -> tuple:
        st = optimizer.state
        p0, pl = params[0], params[-1]
        s0, sl = st.get(p0, {}), st.get(pl, {})
        return (id(optimizer), len(params), p0.data_ptr(), pl.data_ptr(),
                s0["exp_avg"].data_ptr() if "exp_avg" in s0 else 0, sl["exp_avg_sq"].data_ptr() if "exp_avg_sq" in sl else 0,
                s0["step"].data_ptr() if isinstance(s0.get("step"), torch.Tensor) and s0["step"].is_cuda else 0)

    @torch.no_grad()
    def step(self, flat: torch.Tensor, optimizer: torch.optim.Adam, params: list[torch.Tensor], max_norm: float,
             inv_scale: torch.Tensor | None = None, grad_div: float = 1.0, num_sms: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
        """Returns (gradient norm of the unscaled gradient, found_inf) as 0-dim device tensors; no host sync."""
        dev = flat.device
        key = self._key_of(optimizer, params)
        if self._tables is None or key != self._key:
            self._tables = self._build(optimizer, params, dev)
            self._key = self._key_of(optimizer, params)
        t = self._tables
        if t["total"] != flat.numel():
            raise ValueError("flat gradient size does not match the optimizer's parameters")
        g = optimizer.param_groups[0]
        lib = _lib.load()
        stream = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.kb_flat_grad_stats(flat.data_ptr(), flat.numel(), t["stats"].data_ptr(), num_sms, stream), "kb_flat_grad_stats")
            _lib.check(lib.kb_adam_step_flat(
                flat.data_ptr(), t["p"].data_ptr(), t["m"].data_ptr(), t["v"].data_ptr(), t["steps"].data_ptr(), t["g_off"].data_ptr(),
                t["sizes"].data_ptr(), t["chunk_tensor"].data_ptr(), t["chunk_start"].data_ptr(), len(params), t["n_chunks"], _CHUNK,
                t["stats"].data_ptr(), _lib.ptr(inv_scale), float(grad_div), float(max_norm), float(g["lr"]), float(g["betas"][0]),
                float(g["betas"][1]), float(g["eps"]), t["out"].data_ptr(), t["out"].data_ptr() + 4, stream), "kb_adam_step_flat")
        return t["out"][0], t["out"][1]
