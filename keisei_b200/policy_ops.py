"""torch.library custom ops over the masked-policy / loss kernels (csrc/policy.cu).

`keisei_b200::ppo_policy_loss`  fused masked log-softmax + gather + entropy + clipped surrogate
                                (reference katago_ppo.py:33-43, :858-888), autograd registered.
`keisei_b200::value_losses`     W/D/L cross-entropy (ignore_index=-1) + score MSE
                                (reference katago_ppo.py:46-57, :910-912), autograd registered.
`policy_sample`                 rollout mask->softmax->sample->log-prob (+ scalar value)
                                (reference katago_ppo.py:589-613), no grad.

Logit rows may be padded: a (B, A) tensor whose row stride is >= A is consumed in place.
All ops require CUDA tensors; there is no CPU implementation behind them.
"""
from __future__ import annotations

import itertools

import torch

from . import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1}


def _check_logits(logits: torch.Tensor) -> tuple[int, int, int]:
    if not logits.is_cuda:
        raise _lib.KeiseiB200Error("keisei_b200 policy ops need CUDA tensors (no CPU fallback on this path)")
    if logits.ndim != 2 or logits.stride(1) != 1:
        raise ValueError(f"logits must be (B, A) with unit inner stride, got {tuple(logits.shape)} / {logits.stride()}")
    if logits.dtype not in _DT:
        raise ValueError(f"logits dtype must be float32 or bfloat16, got {logits.dtype}")
    return logits.shape[0], logits.shape[1], logits.stride(0)


MASK_BYTES, MASK_BITS, MASK_NONE = 0, 1, 2      # include/keisei_b200.h: KB_MASK_*


def mask_words(num_actions: int) -> int:
    """32-bit words per bit-packed mask row (352 for the 11,259-action space: 1,408 B instead of 11,259 B)."""
    return (int(num_actions) + 31) // 32


def mask_row_bytes(mask: torch.Tensor | None) -> int:
    """Bytes of legal-mask storage the kernels read per row (algorithmic-traffic accounting in bench.py)."""
    return 0 if mask is None else int(mask.shape[1]) * mask.element_size()


def pack_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """(rows, A) bool / uint8 CUDA tensor -> (rows, ceil(A/32)) int32 bit-packed rows (action i = bit i & 31 of word i >> 5).
    The packed form is what `KataGoRolloutBuffer(device=cuda)` stores and what the policy kernels read 8x cheaper."""
    if not mask.is_cuda:
        raise _lib.KeiseiB200Error("pack_mask_bits needs a CUDA tensor")
    if mask.dtype != torch.bool and mask.dtype != torch.uint8:
        mask = mask != 0
    mask = mask.contiguous()
    rows, A = mask.shape
    words = mask_words(A)
    bits = torch.empty((rows, words), dtype=torch.int32, device=mask.device)
    with torch.cuda.device(mask.device):
        rc = _lib.load().kb_pack_mask_bits(_lib.ptr(mask), _lib.ptr(bits), rows, A, words, _lib.stream_ptr(mask.device))
    _lib.check(rc, "kb_pack_mask_bits")
    return bits


def unpack_mask_bits(bits: torch.Tensor, num_actions: int) -> torch.Tensor:
    """(rows, words) int32 -> (rows, A) bool (host-side convenience: flatten() consumers that want the reference layout)."""
    idx = torch.arange(num_actions, device=bits.device)
    return ((bits[:, idx >> 5] >> (idx & 31)) & 1).bool()


def _mask_args(mask: torch.Tensor | None, B: int, A: int) -> tuple[torch.Tensor | None, int, int]:
    """(tensor, kind, pitch) for the C entry points: None -> no mask; int32 (B, words) -> bit-packed; else (B, A) bytes."""
    if mask is None:
        return None, MASK_NONE, 0
    if mask.dtype == torch.int32:
        if mask.ndim != 2 or mask.shape[0] != B or mask.shape[1] * 32 < A:
            raise ValueError(f"bit-packed legal mask shape {tuple(mask.shape)} does not cover {(B, A)}")
        return mask.contiguous(), MASK_BITS, int(mask.shape[1])
    if mask.shape != (B, A):
        raise ValueError(f"legal mask shape {tuple(mask.shape)} != {(B, A)}")
    if mask.dtype != torch.bool and mask.dtype != torch.uint8:
        mask = mask != 0
    return mask.contiguous(), MASK_BYTES, A


# ---------------------------------------------------------------------------------------------
# ppo_policy_loss
# ---------------------------------------------------------------------------------------------
@torch.library.custom_op("keisei_b200::ppo_policy_loss", mutates_args=())
def ppo_policy_loss(logits: torch.Tensor, mask: torch.Tensor | None, actions: torch.Tensor,
                    old_log_probs: torch.Tensor, advantages: torch.Tensor,
                    clip_epsilon: float) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (out2=[policy_loss, entropy], new_log_probs, row_entropy, row_lse, dlogp, flags).
    flags = [rows with zero legal actions, rows with NaN raw logits].
    `mask`: (B, A) bool / uint8, (B, ceil(A/32)) int32 bit-packed (`pack_mask_bits`), or None (all legal)."""
    B, A, stride = _check_logits(logits)
    dev = logits.device
    mask, mkind, mpitch = _mask_args(mask, B, A)
    actions = actions.to(torch.int64).contiguous()
    old = old_log_probs.to(torch.float32).contiguous()
    adv = advantages.to(torch.float32).contiguous()
    out2 = torch.empty(2, device=dev, dtype=torch.float32)
    new_logp = torch.empty(B, device=dev, dtype=torch.float32)
    row_ent = torch.empty(B, device=dev, dtype=torch.float32)
    row_lse = torch.empty(B, device=dev, dtype=torch.float32)
    dlogp = torch.empty(B, device=dev, dtype=torch.float32)
    flags = torch.zeros(2, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_ppo_policy_fwd(
            _lib.ptr(logits), _DT[logits.dtype], stride, _lib.ptr(mask), _lib.ptr(actions), _lib.ptr(old),
            _lib.ptr(adv), B, A, float(clip_epsilon), _lib.ptr(new_logp), _lib.ptr(row_ent), _lib.ptr(row_lse),
            _lib.ptr(dlogp), _lib.ptr(out2), _lib.ptr(flags), mkind, mpitch, _lib.stream_ptr(dev))
    _lib.check(rc, "kb_ppo_policy_fwd")
    return out2, new_logp, row_ent, row_lse, dlogp, flags


@ppo_policy_loss.register_fake
def _(logits, mask, actions, old_log_probs, advantages, clip_epsilon):
    B = logits.shape[0]
    f = lambda *s: logits.new_empty(s, dtype=torch.float32)
    return f(2), f(B), f(B), f(B), f(B), logits.new_empty((2,), dtype=torch.int32)


@torch.library.custom_op("keisei_b200::ppo_policy_loss_backward", mutates_args=())
def ppo_policy_loss_backward(logits: torch.Tensor, mask: torch.Tensor | None, actions: torch.Tensor,
                             row_lse: torch.Tensor, row_entropy: torch.Tensor, dlogp: torch.Tensor,
                             g_out2: torch.Tensor) -> torch.Tensor:
    B, A, stride = _check_logits(logits)
    dev = logits.device
    mask, mkind, mpitch = _mask_args(mask, B, A)
    actions = actions.to(torch.int64).contiguous()
    g = g_out2.to(torch.float32).contiguous()
    # same padded pitch as the logits so the producer's backward can consume it in place
    dbuf = torch.empty((B, stride), device=dev, dtype=logits.dtype)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_ppo_policy_bwd(
            _lib.ptr(logits), _DT[logits.dtype], stride, _lib.ptr(mask), _lib.ptr(actions), B, A,
            _lib.ptr(row_lse), _lib.ptr(row_entropy), _lib.ptr(dlogp), g.data_ptr(), g.data_ptr() + 4,
            _lib.ptr(dbuf), stride, mkind, mpitch, _lib.stream_ptr(dev))
    _lib.check(rc, "kb_ppo_policy_bwd")
    return dbuf[:, :A]


@ppo_policy_loss_backward.register_fake
def _(logits, mask, actions, row_lse, row_entropy, dlogp, g_out2):
    return torch.empty_strided(logits.shape, logits.stride(), dtype=logits.dtype, device=logits.device)


def _ppl_setup(ctx, inputs, output):
    logits, mask, actions, _, _, _ = inputs
    _, _, row_ent, row_lse, dlogp, _ = output
    ctx.save_for_backward(logits, mask, actions, row_lse, row_ent, dlogp)


def _ppl_backward(ctx, g_out2, g_new_logp, g_row_ent, g_row_lse, g_dlogp, g_flags):
    logits, mask, actions, row_lse, row_ent, dlogp = ctx.saved_tensors
    dlogits = None
    if g_out2 is not None or g_new_logp is None:
        if g_out2 is None:
            g_out2 = torch.zeros(2, device=logits.device)
        dlogits = ppo_policy_loss_backward(logits, mask, actions, row_lse, row_ent, dlogp, g_out2)
    if g_new_logp is not None:
        # gradient arriving directly at the per-row log-probs (supervised cross-entropy = -mean(new_log_probs)):
        # d new_logp[r] / d logits[r, i] = onehot - p  — the same kernel with the incoming row weights in place of dlogp
        one = torch.tensor([1.0, 0.0], device=logits.device)
        extra = ppo_policy_loss_backward(logits, mask, actions, row_lse, row_ent, g_new_logp.to(torch.float32).contiguous(), one)
        dlogits = extra if dlogits is None else dlogits + extra
    return dlogits, None, None, None, None, None


ppo_policy_loss.register_autograd(_ppl_backward, setup_context=_ppl_setup)


# ---------------------------------------------------------------------------------------------
# value_losses
# ---------------------------------------------------------------------------------------------
@torch.library.custom_op("keisei_b200::value_losses", mutates_args=())
def value_losses(value_logits: torch.Tensor, value_cats: torch.Tensor, score_pred: torch.Tensor,
                 score_targets: torch.Tensor) -> torch.Tensor:
    """Returns out3 = [wdl_cross_entropy, score_mse, n_valid]."""
    if not value_logits.is_cuda:
        raise _lib.KeiseiB200Error("keisei_b200::value_losses needs CUDA tensors")
    B = value_logits.shape[0]
    dev = value_logits.device
    vl = value_logits.to(torch.float32).contiguous()
    sp = score_pred.to(torch.float32).reshape(B).contiguous()
    st = score_targets.to(torch.float32).reshape(B).contiguous()
    cats = value_cats.to(torch.int64).contiguous()
    out3 = torch.empty(3, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_value_losses_fwd(_lib.ptr(vl), _lib.ptr(cats), _lib.ptr(sp), _lib.ptr(st), B,
                                             _lib.ptr(out3), _lib.stream_ptr(dev))
    _lib.check(rc, "kb_value_losses_fwd")
    return out3


@value_losses.register_fake
def _(value_logits, value_cats, score_pred, score_targets):
    return value_logits.new_empty((3,), dtype=torch.float32)


@torch.library.custom_op("keisei_b200::value_losses_backward", mutates_args=())
def value_losses_backward(value_logits: torch.Tensor, value_cats: torch.Tensor, score_pred: torch.Tensor,
                          score_targets: torch.Tensor, out3: torch.Tensor,
                          g_out3: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    B = value_logits.shape[0]
    dev = value_logits.device
    vl = value_logits.to(torch.float32).contiguous()
    sp = score_pred.to(torch.float32).reshape(B).contiguous()
    st = score_targets.to(torch.float32).reshape(B).contiguous()
    cats = value_cats.to(torch.int64).contiguous()
    g = g_out3.to(torch.float32).contiguous()
    dv = torch.empty((B, 3), device=dev, dtype=torch.float32)
    ds = torch.empty(B, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.load().kb_value_losses_bwd(_lib.ptr(vl), _lib.ptr(cats), _lib.ptr(sp), _lib.ptr(st), B,
                                             _lib.ptr(out3), g.data_ptr(), g.data_ptr() + 4, _lib.ptr(dv),
                                             _lib.ptr(ds), _lib.stream_ptr(dev))
    _lib.check(rc, "kb_value_losses_bwd")
    return dv.to(value_logits.dtype), ds.reshape(score_pred.shape).to(score_pred.dtype)


@value_losses_backward.register_fake
def _(value_logits, value_cats, score_pred, score_targets, out3, g_out3):
    return torch.empty_like(value_logits), torch.empty_like(score_pred)


def _vl_setup(ctx, inputs, output):
    value_logits, value_cats, score_pred, score_targets = inputs
    ctx.save_for_backward(value_logits, value_cats, score_pred, score_targets, output)


def _vl_backward(ctx, g_out3):
    value_logits, value_cats, score_pred, score_targets, out3 = ctx.saved_tensors
    dv, ds = value_losses_backward(value_logits, value_cats, score_pred, score_targets, out3, g_out3)
    return dv, None, ds, None


value_losses.register_autograd(_vl_backward, setup_context=_vl_setup)


# ---------------------------------------------------------------------------------------------
# rollout sampling
# ---------------------------------------------------------------------------------------------
_sample_calls = itertools.count()


@torch.no_grad()
def policy_sample(logits: torch.Tensor, mask: torch.Tensor | None, value_logits: torch.Tensor | None = None,
                  score_lead: torch.Tensor | None = None, alpha: float = 0.0, *, seed: int | None = None,
                  offset: int | None = None, logprob_mode: int | None = None,
                  forced_actions: torch.Tensor | None = None, dense: bool | None = None):
    """One launch: actions (int64), log_probs (fp32), scalar values (fp32 or None), legal_count, flags.

    Sampling is Gumbel-max over the legal entries with a Philox4x32-10 stream keyed by
    (seed, row, offset, action index); illegal actions have probability exactly 0.
    logprob_mode 0 = fp32 Categorical semantics, 1 = bf16-autocast semantics (eps = 2^-7 clamp);
    default follows the logits dtype like the reference does.
    `forced_actions` (int64, (B,)) skips the draw and reports the log-prob of the given actions.
    Bit-packed masks go to the warp-per-row kernel that touches only the legal logits. Bool / uint8 masks: `dense=False`
    packs them first (`pack_mask_bits`) for that kernel, `dense=True` keeps the byte mask and the one-CTA-per-row kernel
    that stages the whole row (same draw, log-probs equal up to summation order); the default is to pack
    (tools/bench_sample.py has the timings of both).
    """
    B, A, stride = _check_logits(logits)
    dev = logits.device
    if mask is not None and mask.dtype != torch.int32 and not dense:
        if mask.shape != (B, A):
            raise ValueError(f"legal mask shape {tuple(mask.shape)} != {(B, A)}")
        mask = pack_mask_bits(mask)
    mask, mkind, mpitch = _mask_args(mask, B, A)
    if seed is None:
        seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    if offset is None:
        offset = next(_sample_calls)
    if logprob_mode is None:
        logprob_mode = 1 if logits.dtype == torch.bfloat16 else 0
    actions = torch.empty(B, device=dev, dtype=torch.int64)
    logp = torch.empty(B, device=dev, dtype=torch.float32)
    legal = torch.empty(B, device=dev, dtype=torch.int32)
    flags = torch.zeros(2, device=dev, dtype=torch.int32)
    vl = sc = values = None
    if value_logits is not None:
        vl = value_logits.to(torch.float32).contiguous()
        values = torch.empty(B, device=dev, dtype=torch.float32)
        if score_lead is not None and alpha != 0.0:
            sc = score_lead.to(torch.float32).reshape(B).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.load().kb_policy_sample(
            _lib.ptr(logits), _DT[logits.dtype], stride, _lib.ptr(mask), _lib.ptr(vl), _lib.ptr(sc), float(alpha),
            B, A, int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset) & 0xFFFFFFFF, int(logprob_mode),
            _lib.ptr(None if forced_actions is None else forced_actions.to(torch.int64).contiguous()),
            _lib.ptr(actions),
            _lib.ptr(logp), _lib.ptr(values), _lib.ptr(legal), _lib.ptr(flags), mkind, mpitch, _lib.stream_ptr(dev))
    _lib.check(rc, "kb_policy_sample")
    return actions, logp, values, legal, flags


@torch.no_grad()
def gather_minibatch(obs: torch.Tensor, mask_bits: torch.Tensor, actions: torch.Tensor, old_log_probs: torch.Tensor,
                     advantages: torch.Tensor, value_cats: torch.Tensor, score_targets: torch.Tensor, returns: torch.Tensor,
                     idx: torch.Tensor):
    """ONE launch for the eight index-gathers of a shuffled minibatch (reference katago_ppo.py:829-841) out of the
    device-resident rollout storage. obs (N, ...) fp32 contiguous, mask_bits (N, words) int32; the rest (N,).
    Returns (obs[idx], mask_bits[idx], actions[idx], old_log_probs[idx], advantages[idx], value_cats[idx],
    score_targets[idx], returns[idx])."""
    if not obs.is_cuda:
        raise _lib.KeiseiB200Error("gather_minibatch needs CUDA tensors")
    N, M, dev = obs.shape[0], idx.numel(), obs.device
    obs_floats = obs[0].numel()
    words = mask_bits.shape[1]
    if (obs.dtype != torch.float32 or not obs.is_contiguous() or mask_bits.dtype != torch.int32 or not mask_bits.is_contiguous()
            or obs_floats % 2 != 0 or mask_bits.shape[0] != N):
        raise ValueError("gather_minibatch: obs must be contiguous float32 with an even row size, mask_bits contiguous int32 (N, words)")
    f32 = lambda t: t.to(torch.float32).contiguous()    # noqa: E731
    i64 = lambda t: t.to(torch.int64).contiguous()      # noqa: E731
    actions, value_cats, idx = i64(actions), i64(value_cats), i64(idx)
    old_log_probs, advantages, score_targets, returns = f32(old_log_probs), f32(advantages), f32(score_targets), f32(returns)
    o_obs = torch.empty((M,) + tuple(obs.shape[1:]), dtype=torch.float32, device=dev)
    o_bits = torch.empty((M, words), dtype=torch.int32, device=dev)
    o_act, o_cats = torch.empty(M, dtype=torch.int64, device=dev), torch.empty(M, dtype=torch.int64, device=dev)
    o_old, o_adv, o_score, o_ret = (torch.empty(M, dtype=torch.float32, device=dev) for _ in range(4))
    with torch.cuda.device(dev):
        rc = _lib.load().kb_gather_minibatch(
            _lib.ptr(obs), _lib.ptr(mask_bits), _lib.ptr(actions), _lib.ptr(old_log_probs), _lib.ptr(advantages), _lib.ptr(value_cats),
            _lib.ptr(score_targets), _lib.ptr(returns), _lib.ptr(idx), N, M, obs_floats, words, _lib.ptr(o_obs), _lib.ptr(o_bits),
            _lib.ptr(o_act), _lib.ptr(o_old), _lib.ptr(o_adv), _lib.ptr(o_cats), _lib.ptr(o_score), _lib.ptr(o_ret), _lib.stream_ptr(dev))
    _lib.check(rc, "kb_gather_minibatch")
    return o_obs, o_bits, o_act, o_old, o_adv, o_cats, o_score, o_ret
