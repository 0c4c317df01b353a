"""KataGo-style multi-head PPO — drop-in for keisei/training/katago_ppo.py (reference lines cited
per symbol). Same constructor, methods, attributes, metrics keys and exception types as the
reference trainer; the device work on CUDA goes through the sm_100a kernels:

  select_actions  -> one C call for the network + `kb_policy_sample` (mask/softmax/sample/log-prob/value)
  update          -> `kb_gae_scan` + `kb_advantage_normalize`, then per minibatch
                     `kb_seresnet_forward` / fused PPO losses / `kb_seresnet_backward`, with the
                     gradients produced as ONE flat fp32 buffer (a single NCCL all-reduce in data-parallel
                     runs, see keisei_b200/distributed.py). Optimiser, GradScaler and clipping stay PyTorch.

Neither Triton nor torch.compile runs on this path: `compile_mode` is accepted for config
compatibility and ignored (a warning is logged).
"""
from __future__ import annotations

import logging
import sys
from dataclasses import dataclass
from typing import Any, Callable

import torch
import torch.nn.functional as F
from torch.amp import GradScaler, autocast

from . import gae as gae_mod
from . import model_ops, policy_ops
from .models.katago_base import KataGoBaseModel
from .models.se_resnet import SEResNetModel

SCORE_NORMALIZATION = 76.0  # reference keisei/sl/dataset.py:32

_log = logging.getLogger(__name__)


def _gae_fn(name: str):
    """`compute_gae*` resolved by MODULE ATTRIBUTE at call time: the reference's `update()` does a local
    `from keisei.training.gae import ...`, so its tests (tests/test_split_merge_gae_opt.py:336-372) and callers patch
    `keisei.training.gae.<name>` and expect the trainer to see it. When that module is loaded and the attribute is NOT
    the reference's own pristine function (a spy, or this package's kernel-backed function put there by
    `install_into_reference()`), it wins; otherwise this package's `gae` module is used (also looked up per call)."""
    ref = sys.modules.get("keisei.training.gae")
    if ref is not None:
        fn = getattr(ref, name, None)
        if fn is not None and getattr(fn, "__module__", None) != "keisei.training.gae":
            return fn
    return getattr(gae_mod, name)


def _amp_dtype_and_device(use_amp: bool, device: torch.device) -> tuple[torch.dtype, str]:
    """Reference katago_ppo.py:19-30."""
    if not use_amp:
        dtype = torch.float16
    elif device.type == "cpu":
        dtype = torch.bfloat16
    elif torch.cuda.is_bf16_supported():
        dtype = torch.bfloat16
    else:
        dtype = torch.float16
    return dtype, device.type


def ppo_clip_loss(new_log_probs, old_log_probs, advantages, clip_epsilon: float) -> torch.Tensor:
    """Clipped surrogate (reference katago_ppo.py:33-43). Host-side form; the CUDA update path fuses
    this into `keisei_b200::ppo_policy_loss`."""
    ratio = (new_log_probs - old_log_probs).exp()
    clipped = ratio.clamp(1 - clip_epsilon, 1 + clip_epsilon)
    return -torch.min(ratio * advantages, clipped * advantages).mean()


def wdl_cross_entropy_loss(value_logits: torch.Tensor, value_cats: torch.Tensor) -> torch.Tensor:
    """W/D/L cross-entropy, ignore_index=-1, graph-connected zero when nothing is valid
    (reference katago_ppo.py:46-57)."""
    if not (value_cats >= 0).any():
        return value_logits.sum() * 0.0
    return F.cross_entropy(value_logits, value_cats, ignore_index=-1)


def compute_value_metrics(value_logits: torch.Tensor, value_targets: torch.Tensor) -> dict[str, float]:
    """Reference katago_ppo.py:60-78."""
    pred = value_logits.argmax(dim=-1)
    return {
        "value_accuracy": (pred == value_targets).float().mean().item(),
        "frac_predicted_win": (pred == 0).float().mean().item(),
        "frac_predicted_draw": (pred == 1).float().mean().item(),
        "frac_predicted_loss": (pred == 2).float().mean().item(),
    }


@dataclass(frozen=True)
class KataGoPPOParams:
    """Reference katago_ppo.py:81-116 (same fields, defaults and validation)."""
    learning_rate: float = 2e-4
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_epsilon: float = 0.2
    epochs_per_batch: int = 4
    batch_size: int = 256
    lambda_policy: float = 1.0
    lambda_value: float = 1.5
    lambda_score: float = 0.02
    lambda_entropy: float = 0.01
    score_normalization: float = SCORE_NORMALIZATION
    grad_clip: float = 1.0
    use_amp: bool = False
    compile_mode: str | None = None
    compile_dynamic: bool = True
    entropy_decay_epochs: int = 0
    score_blend_alpha: float = 0.0
    use_terminated_for_gae: bool = True

    def __post_init__(self) -> None:
        if self.batch_size <= 0:
            raise ValueError(f"batch_size must be > 0, got {self.batch_size}")
        if self.epochs_per_batch <= 0:
            raise ValueError(f"epochs_per_batch must be > 0, got {self.epochs_per_batch}")
        if not 0.0 <= self.gamma <= 1.0:
            raise ValueError(f"gamma must be in [0, 1], got {self.gamma}")
        if not 0.0 <= self.gae_lambda <= 1.0:
            raise ValueError(f"gae_lambda must be in [0, 1], got {self.gae_lambda}")
        if self.clip_epsilon < 0.0:
            raise ValueError(f"clip_epsilon must be >= 0, got {self.clip_epsilon}")
        if self.learning_rate <= 0.0:
            raise ValueError(f"learning_rate must be > 0, got {self.learning_rate}")
        if self.grad_clip <= 0.0:
            raise ValueError(f"grad_clip must be > 0, got {self.grad_clip}")


class _DeviceFlat(dict):
    """`flatten()` result of a device-resident buffer: everything the reference's dict has, with `legal_masks` (the
    (T*N, A) bool tensor) unpacked from the stored bit-packed rows only if somebody asks for it — `update()` consumes
    `legal_masks_packed` directly and never does."""

    def __init__(self, base: dict, bits: torch.Tensor, num_actions: int) -> None:
        super().__init__(base)
        self._bits, self._num_actions = bits, num_actions
        dict.__setitem__(self, "legal_masks_packed", bits)

    def __missing__(self, key):
        if key == "legal_masks":
            value = policy_ops.unpack_mask_bits(self._bits, self._num_actions)
            dict.__setitem__(self, key, value)
            return value
        raise KeyError(key)

    def __contains__(self, key) -> bool:
        return key == "legal_masks" or dict.__contains__(self, key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def _materialise(self) -> None:
        self["legal_masks"]   # noqa: B018  (side effect: unpack + cache)

    def keys(self):
        self._materialise()
        return dict.keys(self)

    def items(self):
        self._materialise()
        return dict.items(self)

    def values(self):
        self._materialise()
        return dict.values(self)

    def __iter__(self):
        self._materialise()
        return dict.__iter__(self)


class KataGoRolloutBuffer:
    """Rollout storage with the reference's interface and guards (katago_ppo.py:128-388): `add`, `flatten`, `clear`,
    `size`, `fill_alternating_perspective_overrides`.

    `device=None` (default) is the reference's host-resident buffer: every `add` copies the step to the CPU and
    `update()` ships observations + masks back (225 MB for T=128 x N=64). `device="cuda:k"` keeps the same storage in
    HBM: `add` is device-to-device, the guards run as one fused flag reduction with a single
    host read per step, `flatten()` returns device views and `update()` performs no host<->device copy of the
    per-sample data at all. The legal masks of a CUDA buffer are stored BIT-PACKED (1,408 B instead of 11,259 B per
    sample; `kb_pack_mask_bits` in `add`): `flatten()["legal_masks_packed"]` is what the update's gather and loss kernels
    read, `flatten()["legal_masks"]` unpacks the reference's bool layout on demand."""

    _FIELDS = ("observations", "actions", "log_probs", "values", "rewards", "dones", "terminated", "legal_masks",
               "value_categories", "score_targets")

    def __init__(self, num_envs: int, obs_shape: tuple[int, ...], action_space: int,
                 device: torch.device | str | None = None) -> None:
        self.num_envs = num_envs
        self.obs_shape = obs_shape
        self.action_space = action_space
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self._alloc_samples = 0
        self._write_offset = 0
        self._step_count = 0
        self._storage: dict[str, torch.Tensor] = {}
        self._has_env_ids = False
        self._has_next_value_override = False

    def _new_storage(self, cap: int) -> dict[str, torch.Tensor]:
        dev = self.device
        st = {
            "observations": torch.empty(cap, *self.obs_shape, device=dev),
            "actions": torch.empty(cap, dtype=torch.long, device=dev),
            "log_probs": torch.empty(cap, device=dev),
            "values": torch.empty(cap, device=dev),
            "rewards": torch.empty(cap, device=dev),
            "dones": torch.empty(cap, dtype=torch.bool, device=dev),
            "terminated": torch.empty(cap, dtype=torch.bool, device=dev),
            "legal_masks": (torch.empty(cap, policy_ops.mask_words(self.action_space), dtype=torch.int32, device=dev)
                            if dev.type == "cuda" else torch.empty(cap, self.action_space, dtype=torch.bool, device=dev)),
            "value_categories": torch.empty(cap, dtype=torch.long, device=dev),
            "score_targets": torch.empty(cap, device=dev),
        }
        if self._has_env_ids:
            st["env_ids"] = torch.empty(cap, dtype=torch.long, device=dev)
        if self._has_next_value_override:
            st["next_value_override"] = torch.full((cap,), float("nan"), device=dev)
        return st

    def _ensure_capacity(self, n_samples: int) -> None:
        needed = self._write_offset + n_samples
        if needed <= self._alloc_samples:
            return
        cap = max(needed * 2, 512 * self.num_envs)
        if self.device.type == "cuda":
            # HBM is not host RAM: no speculative 512-step floor (that alone would be 0.9 GB for 64 envs)
            cap = max(needed, min(cap, 2 * max(needed, 128 * self.num_envs)))
        fresh = self._new_storage(cap)
        off = self._write_offset
        if off > 0:
            for key, t in fresh.items():
                if key in self._storage:
                    t[:off] = self._storage[key][:off]
        self._storage = fresh
        self._alloc_samples = cap

    @property
    def size(self) -> int:
        return self._step_count

    def clear(self) -> None:
        self._write_offset = 0
        self._step_count = 0

    def add(self, obs, actions, log_probs, values, rewards, dones, terminated, legal_masks, value_categories,
            score_targets, env_ids=None, next_value_override=None) -> None:
        dev = self.device
        host = lambda t: t.detach().to(dev)  # noqa: E731  (CPU buffer: the reference's .cpu(); CUDA buffer: stays in HBM)
        obs_c, act_c, lp_c, val_c, rew_c = host(obs), host(actions), host(log_probs), host(values), host(rewards)
        done_c, term_c = host(dones), host(terminated)
        mask_c, cats_c, score_c = host(legal_masks), host(value_categories), host(score_targets)
        if dev.type == "cuda":
            # the four guards as one device-side flag vector and a single host read
            abs_max = score_c.abs().max()
            bad_term, bad_cats, nan_score, big_score = torch.stack([
                (term_c.bool() & ~done_c.bool()).any(), ((cats_c < -1) | (cats_c > 2)).any(), score_c.isnan().any(),
                abs_max > 3.5]).tolist()
        else:
            bad_term = bool((term_c.bool() & ~done_c.bool()).any())
            bad_cats = nan_score = big_score = None
        if bad_term:
            raise AssertionError(
                "terminated must be a subset of dones: every terminated position must also be done. "
                "Got terminated=True where dones=False — likely a call site passing the merged signal.")
        if bad_cats is None or bad_cats:
            invalid = set(cats_c.unique().tolist()) - {-1, 0, 1, 2}
            if invalid:
                raise ValueError(f"value_categories contains invalid values {invalid}. "
                                 f"Expected only {{-1=ignore, 0=W, 1=D, 2=L}}.")
        if nan_score is None:
            nan_score = bool(score_c.isnan().any())
        if nan_score:
            raise ValueError("score_targets contains NaN. With per-step material balance, "
                             "all targets should be real-valued.")
        if big_score is None:
            abs_max = score_c.abs().max()
            big_score = bool(abs_max > 3.5)
        if big_score:
            raise ValueError(f"score_targets appear unnormalized: max abs value = {abs_max.item():.1f}. "
                             f"Expected in [-1.7, +1.7] typical, theoretical max 2.58 (guard 3.5).")
        n = obs_c.shape[0]
        if self._step_count == 0 and env_ids is not None:
            self._has_env_ids = True
        if self._step_count == 0 and next_value_override is not None:
            self._has_next_value_override = True
        self._ensure_capacity(n)
        sl = slice(self._write_offset, self._write_offset + n)
        st = self._storage
        if dev.type == "cuda":
            if tuple(mask_c.shape) != (n, self.action_space):
                raise ValueError(f"legal_masks shape {tuple(mask_c.shape)} != {(n, self.action_space)}")
            mask_c = policy_ops.pack_mask_bits(mask_c)   # stored bit-packed in HBM
        for key, val in zip(self._FIELDS, (obs_c, act_c, lp_c, val_c, rew_c, done_c, term_c, mask_c, cats_c, score_c)):
            st[key][sl] = val
        if env_ids is not None:
            if "env_ids" not in st:
                st["env_ids"] = torch.empty(self._alloc_samples, dtype=torch.long, device=dev)
            st["env_ids"][sl] = host(env_ids)
        if next_value_override is not None:
            if "next_value_override" not in st:
                st["next_value_override"] = torch.full((self._alloc_samples,), float("nan"), device=dev)
                self._has_next_value_override = True
            st["next_value_override"][sl] = host(next_value_override).to(torch.float32)
        elif self._has_next_value_override and "next_value_override" in st:
            st["next_value_override"][sl] = float("nan")  # no stale cells from a previous epoch
        self._write_offset += n
        self._step_count += 1

    def fill_alternating_perspective_overrides(self) -> None:
        """Reference katago_ppo.py:320-362: for non-terminal cells without a caller override,
        next_value_override[t] = -values[t+1] (the next ply is the opponent's frame)."""
        if self._has_env_ids:
            return
        T, N = self._step_count, self.num_envs
        if T <= 1 or self._write_offset != T * N:
            return
        if "next_value_override" not in self._storage:
            self._storage["next_value_override"] = torch.full((self._alloc_samples,), float("nan"), device=self.device)
            self._has_next_value_override = True
        ov = self._storage["next_value_override"][:T * N].view(T, N)
        values = self._storage["values"][:T * N].view(T, N)
        term = self._storage["terminated"][:T * N].view(T, N).bool()
        target = torch.isnan(ov[:-1]) & ~term[:-1]
        ov[:-1][target] = -values[1:][target]

    def flatten(self) -> dict[str, torch.Tensor]:
        if self._step_count == 0:
            raise ValueError("Cannot flatten an empty buffer. Call add() at least once before flatten().")
        off = self._write_offset
        st = self._storage
        packed = self.device.type == "cuda"
        out = {"observations": st["observations"][:off].reshape(-1, *self.obs_shape)}
        if not packed:
            out["legal_masks"] = st["legal_masks"][:off].reshape(-1, self.action_space)
        for key in ("actions", "log_probs", "values", "rewards", "dones", "terminated", "value_categories", "score_targets"):
            out[key] = st[key][:off].reshape(-1)
        if self._has_env_ids and "env_ids" in st:
            out["env_ids"] = st["env_ids"][:off].reshape(-1)
        if self._has_next_value_override and "next_value_override" in st:
            out["next_value_override"] = st["next_value_override"][:off].reshape(-1)
        if packed:
            return _DeviceFlat(out, st["legal_masks"][:off], self.action_space)
        return out


class KataGoPPOAlgorithm:
    """Reference katago_ppo.py:391-991."""

    def __init__(self, params: KataGoPPOParams, model: KataGoBaseModel, forward_model: torch.nn.Module | None = None,
                 warmup_epochs: int = 0, warmup_entropy: float = 0.05) -> None:
        self.params = params
        self.model = model
        self.forward_model = forward_model or model
        fm_base = self.forward_model.module if hasattr(self.forward_model, "module") else self.forward_model
        m_base = self.model.module if hasattr(self.model, "module") else self.model
        assert fm_base is m_base, (
            "forward_model and model must share parameters — compile + grad clipping requires this")
        self.compiled_train: Callable[..., Any] | None = None
        self.compiled_eval: Callable[..., Any] | None = None
        if params.compile_mode is not None:
            _log.warning("compile_mode=%r ignored: the keisei_b200 hot path runs hand-written sm_100a kernels, "
                         "not torch.compile", params.compile_mode)
        self._timing_events: dict[str, list] = {"select_actions_forward_ms": [], "update_forward_backward_ms": [], "gae_ms": []}
        self.timings: dict[str, list[float]] = {"select_actions_forward_ms": [], "update_forward_backward_ms": [], "gae_ms": []}
        device = next(model.parameters()).device
        if hasattr(model, "configure_amp"):
            amp_dtype, amp_dev = _amp_dtype_and_device(params.use_amp, device)
            model.configure_amp(enabled=params.use_amp, dtype=amp_dtype, device_type=amp_dev)
        # Same optimiser and hyper-parameters as the reference (katago_ppo.py:494: Adam defaults); on CUDA PyTorch's fused
        # multi-tensor implementation (one launch per ~64 tensors for the whole moment/parameter update instead of five
        # foreach passes) — identical `state_dict()` layout, so checkpoints interchange (checkpoint.py:123).
        self.optimizer = torch.optim.Adam(model.parameters(), lr=params.learning_rate, fused=device.type == "cuda")
        self.scaler = GradScaler(enabled=params.use_amp and device.type == "cuda")
        self.warmup_epochs = warmup_epochs
        self.warmup_entropy = warmup_entropy
        self.current_entropy_coeff = params.lambda_entropy
        # keisei_b200 extensions (not in the reference): data-parallel gradient sync and guard strictness
        self.grad_sync = None          # keisei_b200.distributed.GradSync or None
        self.strict_guards = True      # check NaN / zero-legal flags every minibatch (1 host sync, reference: 2)
        self._sample_seed: int | None = None
        self._flat_grad: torch.Tensor | None = None   # flat gradient of the last fused step (consumed by _optimizer_tail)
        self.fused_optimizer_tail = True   # norm + clip + Adam as two launches over the flat gradient (keisei_b200/optim.py)

    # ---- small helpers ---------------------------------------------------------------------------
    def get_entropy_coeff(self, epoch: int) -> float:
        """Reference katago_ppo.py:500-516."""
        if epoch < self.warmup_epochs:
            return self.warmup_entropy
        decay = self.params.entropy_decay_epochs
        elapsed = epoch - self.warmup_epochs
        if decay <= 0 or elapsed >= decay:
            return self.params.lambda_entropy
        return self.warmup_entropy + (elapsed / decay) * (self.params.lambda_entropy - self.warmup_entropy)

    def flush_timings(self) -> None:
        """Reference katago_ppo.py:518-531: the only synchronisation point of the event timers."""
        for key, pairs in self._timing_events.items():
            self.timings[key] = [s.elapsed_time(e) for s, e in pairs]
            pairs.clear()

    @staticmethod
    def scalar_value(value_logits: torch.Tensor) -> torch.Tensor:
        """P(W) - P(L) (reference katago_ppo.py:533-541)."""
        p = F.softmax(value_logits, dim=-1)
        return p[:, 0] - p[:, 2]

    def _base(self) -> torch.nn.Module:
        return self.model.module if hasattr(self.model, "module") else self.model

    def _kernel_model(self, device: torch.device) -> SEResNetModel | None:
        """The fused C-ABI path applies when the (unwrapped) model is the keisei_b200 SE-ResNet on CUDA."""
        base = self._base()
        if device.type == "cuda" and isinstance(base, SEResNetModel) and base.kernel_supported():
            return base
        return None

    def _events(self, device, key):
        if device.type != "cuda":
            return None
        stream = torch.cuda.current_stream(device)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        return key, s, e, stream

    def _events_end(self, tok) -> None:
        if tok is not None:
            key, s, e, stream = tok
            e.record(stream)
            self._timing_events[key].append((s, e))

    # ---- rollout ---------------------------------------------------------------------------------
    @torch.no_grad()
    def select_actions(self, obs: torch.Tensor, legal_masks: torch.Tensor, value_adapter: Any | None = None):
        """Reference katago_ppo.py:543-617: eval-mode forward, zero-legal guard (RuntimeError), masked
        sample + log-prob, scalar value (blended through the adapter when given). Leaves
        `forward_model` in train mode."""
        device = next(self.model.parameters()).device
        model = self.forward_model
        # the kernel path takes the mode as an argument: no eval() / train() walk over ~600 submodules per step (1.5 ms of
        # Python with the GPU idle); every other path gets the reference's eval() ... train() bracket
        fast = isinstance(model, SEResNetModel) and obs.is_cuda
        if not fast:
            model.eval()
        try:
            tok = self._events(device, "select_actions_forward_ms")
            # small batches replay a captured CUDA graph of the network (launch-bound otherwise)
            output = model.rollout_forward(obs, eval_mode=True) if isinstance(model, SEResNetModel) else model(obs)
            self._events_end(tok)
            B = obs.shape[0]
            if device.type == "cuda":
                flat = output.policy_logits.reshape(B, -1)
                alpha = float(getattr(value_adapter, "score_blend_alpha", 0.0)) if value_adapter is not None else 0.0
                fused_value = value_adapter is None or hasattr(value_adapter, "score_blend_alpha")
                actions, log_probs, values, legal, flags = policy_ops.policy_sample(
                    flat, legal_masks, output.value_logits if fused_value else None, output.score_lead, alpha,
                    seed=self._sample_seed)
                if self.strict_guards and int(flags[0].item()) != 0:
                    zero_envs = (legal == 0).nonzero(as_tuple=True)[0].tolist()
                    raise RuntimeError(f"Environments {zero_envs} have zero legal actions — "
                                       f"all-False legal mask would produce NaN")
                if not fused_value:
                    values = value_adapter.scalar_value_blended(output.value_logits, output.score_lead)
                return actions, log_probs, values
            legal_counts = legal_masks.sum(dim=-1)
            if (legal_counts == 0).any():
                zero_envs = (legal_counts == 0).nonzero(as_tuple=True)[0].tolist()
                raise RuntimeError(f"Environments {zero_envs} have zero legal actions — "
                                   f"all-False legal mask would produce NaN")
            masked = output.policy_logits.reshape(B, -1).masked_fill(~legal_masks, float("-inf"))
            dist = torch.distributions.Categorical(F.softmax(masked, dim=-1), validate_args=False)
            actions = dist.sample()
            log_probs = dist.log_prob(actions)
            if value_adapter is not None:
                values = value_adapter.scalar_value_blended(output.value_logits, output.score_lead)
            else:
                values = self.scalar_value(output.value_logits)
            return actions, log_probs, values
        finally:
            if not fast or not self.forward_model.training:
                self.forward_model.train()

    @torch.no_grad()
    def select_actions_many(self, batches, models=None, value_adapter: Any | None = None):
        """Action selection for several independent sub-batches in one go — the learner's and the league opponents'
        sub-batches of a split-merge step (reference katago_loop.py:284-431 runs one model after the other).

        `batches`: list of (obs, legal_masks); `models`: one model per batch (default: this algorithm's forward model for
        all). On CUDA the networks run as parallel branches of ONE replayed CUDA graph (`rollout_forward_many`), then each
        sub-batch is sampled exactly as in `select_actions`. Returns a list of (actions, log_probs, values)."""
        if models is None:
            models = [self.forward_model] * len(batches)
        if len(models) != len(batches):
            raise ValueError("select_actions_many: one model per batch expected")
        device = next(self.model.parameters()).device
        grouped = device.type == "cuda" and all(isinstance(m, SEResNetModel) for m in models)
        if not grouped:
            return [self.select_actions(o, k, value_adapter) for o, k in batches]
        try:
            from .models import rollout_forward_many
            outs = rollout_forward_many([(m, o) for m, (o, _) in zip(models, batches)], eval_mode=True)
            alpha = float(getattr(value_adapter, "score_blend_alpha", 0.0)) if value_adapter is not None else 0.0
            fused_value = value_adapter is None or hasattr(value_adapter, "score_blend_alpha")
            results, checks = [], []
            for out, (obs, masks) in zip(outs, batches):
                flat = out.policy_logits.reshape(obs.shape[0], -1)
                actions, log_probs, values, legal, flags = policy_ops.policy_sample(
                    flat, masks, out.value_logits if fused_value else None, out.score_lead, alpha, seed=self._sample_seed)
                if not fused_value:
                    values = value_adapter.scalar_value_blended(out.value_logits, out.score_lead)
                results.append((actions, log_probs, values))
                checks.append((flags, legal))
            if self.strict_guards:   # one host read for all sub-batches
                bad = torch.stack([f[0] for f, _ in checks]).tolist()
                for i, b in enumerate(bad):
                    if int(b) != 0:
                        zero_envs = (checks[i][1] == 0).nonzero(as_tuple=True)[0].tolist()
                        raise RuntimeError(f"Environments {zero_envs} have zero legal actions — "
                                           f"all-False legal mask would produce NaN")
            return results
        finally:
            if not self.forward_model.training:
                self.forward_model.train()

    # ---- advantages ------------------------------------------------------------------------------
    def _advantages(self, data, T: int, N: int, next_values: torch.Tensor, device: torch.device) -> torch.Tensor:
        """GAE over the buffer on `device` (reference katago_ppo.py:649-773): (T,N) grid, per-env padded
        (env_ids) or flat fallback; fp32 forced; returns a flat fp32 tensor on `device`, un-normalised."""
        p = self.params
        key = "terminated" if p.use_terminated_for_gae else "dones"
        total = data["rewards"].numel()
        nv = next_values.detach().float()
        if total == T * N:
            ov = data.get("next_value_override")
            tok = self._events(device, "gae_ms")
            adv = _gae_fn("compute_gae_gpu")(
                data["rewards"].reshape(T, N).float().to(device), data["values"].reshape(T, N).float().to(device),
                data[key].reshape(T, N).to(device), nv.to(device), gamma=p.gamma, lam=p.gae_lambda,
                next_value_override=None if ov is None else ov.reshape(T, N).float().to(device)) \
                if device.type == "cuda" else _gae_fn("compute_gae")(
                    data["rewards"].reshape(T, N).float(), data["values"].reshape(T, N).float(), data[key].reshape(T, N),
                    nv.cpu(), gamma=p.gamma, lam=p.gae_lambda,
                    next_value_override=None if ov is None else ov.reshape(T, N).float())
            self._events_end(tok)
            return adv.reshape(-1)
        nv_cpu = nv.cpu()
        # ragged buffers (split-merge rollouts): the index bookkeeping below is host-side, on the small 1-D fields only
        data = {k: (v.cpu() if v.is_cuda and k not in ("observations", "legal_masks", "legal_masks_packed") else v)
                for k, v in dict.items(data)}   # dict.items: a device buffer's lazy `legal_masks` is not materialised
        if "env_ids" in data:
            env_ids = data["env_ids"]
            order = torch.argsort(env_ids, stable=True)
            uniq, counts = env_ids[order].unique_consecutive(return_counts=True)
            if uniq.max() >= nv_cpu.shape[0]:
                raise IndexError(f"env_id {uniq.max().item()} >= next_values size {nv_cpu.shape[0]}")
            lengths = counts
            n_env, max_t = len(uniq), int(counts.max())
            # (t, column) coordinates of every sample in the padded grid
            col = torch.repeat_interleave(torch.arange(n_env), counts)
            starts = torch.cumsum(counts, 0) - counts
            row = torch.arange(total) - torch.repeat_interleave(starts, counts)
            def pad(src, fill):
                out = torch.full((max_t, n_env), fill, dtype=torch.float32)
                out[row, col] = src[order].float()
                return out
            rewards_p, values_p = pad(data["rewards"], 0.0), pad(data["values"], 0.0)
            term_p = pad(data[key], 1.0)  # padding = terminated so nothing propagates through it
            ov_p = pad(data["next_value_override"], float("nan")) if "next_value_override" in data else None
            nv_cols = nv_cpu[uniq]
            fn = _gae_fn("compute_gae_padded_gpu") if device.type == "cuda" else _gae_fn("compute_gae_padded")
            mv = (lambda t: t.to(device)) if device.type == "cuda" else (lambda t: t)
            padded = fn(mv(rewards_p), mv(values_p), mv(term_p), mv(nv_cols), lengths, gamma=p.gamma, lam=p.gae_lambda,
                        next_value_override=None if ov_p is None else mv(ov_p))
            adv = torch.zeros(total, device=padded.device)
            adv[order.to(padded.device)] = padded[row.to(padded.device), col.to(padded.device)]
            return adv
        adv = _gae_fn("compute_gae")(data["rewards"].float(), data["values"].float(), data[key], nv_cpu.mean(),
                                  gamma=p.gamma, lam=p.gae_lambda)
        return adv.to(device)

    # ---- one optimisation step -----------------------------------------------------------------------
    def _policy_terms(self, flat_logits, masks, actions, old_lp, adv):
        """(policy_loss, entropy, flags): masked log-softmax, gather, entropy and the clipped surrogate
        (reference katago_ppo.py:858-888, :33-43) — one fused kernel on CUDA."""
        p = self.params
        if flat_logits.is_cuda:
            if masks is not None and masks.dtype != torch.int32:
                # byte masks handed in directly (the device-resident buffer already stores packed rows): pack once, the
                # forward and backward kernels over packed masks read 8x less mask and skip every illegal logit
                masks = policy_ops.pack_mask_bits(masks)
            out2, _, _, _, _, flags = policy_ops.ppo_policy_loss(flat_logits, masks, actions, old_lp, adv, p.clip_epsilon)
            return out2[0], out2[1], flags
        if flat_logits.isnan().any():
            raise RuntimeError("NaN in raw policy logits from model forward pass")
        if (masks.sum(dim=-1) == 0).any():
            raise RuntimeError("Batch contains samples with zero legal actions in update(). "
                               "Check that terminal-state masks are not stored in the buffer.")
        logp_all = F.log_softmax(flat_logits.float().masked_fill(~masks, float("-inf")), dim=-1)
        new_lp = logp_all.gather(1, actions.unsqueeze(1)).squeeze(1)
        policy_loss = ppo_clip_loss(new_lp, old_lp, adv, p.clip_epsilon)
        entropy = -(logp_all.exp() * logp_all.masked_fill(~masks, 0.0)).sum(dim=-1).mean()
        return policy_loss, entropy, None

    def _losses(self, flat_logits, value_logits, score_lead, mb, value_adapter):
        """(loss, policy_loss, value_loss, score_loss, entropy, flags) for CUDA or CPU tensors
        (reference katago_ppo.py:858-924)."""
        p = self.params
        masks, actions, old_lp, adv, cats, score_t = mb[:6]
        policy_loss, entropy, flags = self._policy_terms(flat_logits, masks, actions, old_lp, adv)
        if value_adapter is not None:
            value_score = value_adapter.compute_value_loss(value_logits, returns=None, value_cats=cats,
                                                           score_targets=score_t, score_pred=score_lead)
            value_loss, score_loss = value_score, torch.zeros((), device=flat_logits.device)
        elif flat_logits.is_cuda:
            out3 = policy_ops.value_losses(value_logits, cats, score_lead, score_t)
            value_loss, score_loss = out3[0], out3[1]
            value_score = p.lambda_value * value_loss + p.lambda_score * score_loss
        else:
            value_loss = wdl_cross_entropy_loss(value_logits.float(), cats)
            score_loss = F.mse_loss(score_lead.float().squeeze(-1), score_t)
            value_score = p.lambda_value * value_loss + p.lambda_score * score_loss
        loss = p.lambda_policy * policy_loss + value_score - self.current_entropy_coeff * entropy
        return loss, policy_loss, value_loss, score_loss, entropy, flags

    def _check_flags(self, flags) -> None:
        if flags is None or not self.strict_guards:
            return
        zero_legal, nan_rows = flags.tolist()
        if nan_rows:
            raise RuntimeError("NaN in raw policy logits from model forward pass")
        if zero_legal:
            raise RuntimeError("Batch contains samples with zero legal actions in update(). "
                               "Check that terminal-state masks are not stored in the buffer.")

    def _step_fused(self, km: SEResNetModel, obs, mb, value_adapter):
        """Forward, losses, backward through the C-ABI with ONE flat gradient buffer."""
        tables = km._ptr_tables()
        params, buffers = tables.params, tables.buffers
        dtype = km._act_dtype(obs.device)
        code = 0 if dtype == torch.float32 else 1
        wpack = km._packed(params, buffers, dtype)
        with torch.no_grad():
            policy_buf, value, score, ws, new_stats = model_ops.seresnet_forward_raw(
                obs, tables, wpack, True, code, bool(km.use_tensor_cores), km.bn_sync)
            km._store_running_stats(buffers, new_stats)
        policy_buf.requires_grad_(True); value.requires_grad_(True); score.requires_grad_(True)
        loss, pl, vl, sl, ent, flags = self._losses(policy_buf[:, :model_ops.POLICY_A], value, score, mb, value_adapter)
        self._check_flags(flags)
        if self.strict_guards and km.bn_sync is not None and hasattr(km.bn_sync, "check"):
            km.bn_sync.check()   # after the host read above: a forward exchange that lost a peer raises here, not as NaN later
        self.optimizer.zero_grad(set_to_none=True)
        self.scaler.scale(loss).backward()
        with torch.no_grad():
            gs = self.grad_sync
            overlapped = gs is not None and int(gs.world_size) > 1 and getattr(gs, "overlap", False)
            flat = model_ops.seresnet_backward_raw(tables, wpack, ws, policy_buf.grad, value.grad, score.grad, code,
                                                   bool(km.use_tensor_cores), km._grad_sizes, km.bn_sync,
                                                   grad_sync=gs if overlapped else None)
            self._grad_div = 1.0
            if overlapped:
                self._grad_div = float(gs.world_size)   # the buckets came back SUMMED over the ranks: divided in the optimiser tail
            elif gs is not None:
                gs.all_reduce_flat(flat)
            self._flat_grad = flat   # every p.grad below is a view of it: the optimiser tail works on this one buffer
            off = 0
            for prm in params:
                n = prm.numel()
                prm.grad = flat[off:off + n].view(prm.shape)
                off += n
        return pl, vl, sl, ent, value.detach()

    def _optimizer_tail(self) -> torch.Tensor:
        """unscale -> clip -> optimiser step -> scaler update (reference katago_ppo.py:926-933), returns the gradient norm.

        After a fused step all gradients are views of ONE flat fp32 buffer, so the GradScaler's inf check / unscale,
        `clip_grad_norm_` and — for a plain `torch.optim.Adam` — the Adam update itself run as TWO launches on that buffer
        (`FlatAdamTail`: one read of the gradient, one pass over p / g / m / v) instead of ~50 foreach / multi-tensor
        launches over 576 tensors; the optimizer's `state_dict()` is unchanged. Any other optimizer keeps the stock
        `scaler.step()`, and any step whose gradients did not come from `_step_fused` goes through the stock PyTorch
        calls end to end."""
        p = self.params
        flat = getattr(self, "_flat_grad", None)
        self._flat_grad = None
        grad_div = getattr(self, "_grad_div", 1.0)
        self._grad_div = 1.0
        base = self._base()
        if hasattr(base, "invalidate_packed_weights"):
            base.invalidate_packed_weights()   # the fused optimiser does not bump Tensor._version: re-pack explicitly
        params = self.optimizer.param_groups[0]["params"] if len(self.optimizer.param_groups) == 1 else None
        usable = (flat is not None and params is not None and len(params) > 0 and params[0].grad is not None
                  and params[0].grad.data_ptr() == flat.data_ptr()
                  and sum(q.numel() for q in params) == flat.numel())
        from .optim import FlatAdamTail
        fused_adam = usable and self.fused_optimizer_tail and FlatAdamTail.supports(self.optimizer, params, flat)
        if flat is not None and grad_div != 1.0 and not fused_adam:
            flat.div_(grad_div)     # the stock paths below expect the rank-averaged gradient
            grad_div = 1.0
        if not usable:
            self.scaler.unscale_(self.optimizer)
            grad_norm = torch.nn.utils.clip_grad_norm_(self.model.parameters(), p.grad_clip)
            self.scaler.step(self.optimizer)
            self.scaler.update()
            return grad_norm
        tail = getattr(self, "_flat_adam", None)
        if tail is None:
            tail = self._flat_adam = FlatAdamTail()
        inv_scale = None
        st = None
        if self.scaler.is_enabled():
            # GradScaler.unscale_ on the flat buffer; the scaler's bookkeeping is filled in exactly as unscale_ does, so
            # scaler.update() adapts the scale as usual (and scaler.step() sees found_inf on the stock path)
            from torch.amp.grad_scaler import OptState
            st = self.scaler._per_optimizer_states[id(self.optimizer)]
            if st["stage"] is OptState.UNSCALED:
                raise RuntimeError("unscale_() has already been called on this optimizer since the last update().")
            scale = self.scaler._scale
            if scale is None:
                raise RuntimeError("GradScaler has no scale yet: scaler.scale(loss) must run before the optimiser tail")
            inv_scale = scale.double().reciprocal().float()
        if fused_adam:
            # two launches: one read of the gradient (norm + non-finite flag), one pass over (p, g, m, v) that unscales,
            # clips and applies Adam in place on PyTorch's own parameter / state tensors (csrc/optim.cu)
            grad_norm, found_inf = tail.step(flat, self.optimizer, params, p.grad_clip, inv_scale, grad_div,
                                             model_ops.sm_count(flat.device))
            if st is not None:
                from torch.amp.grad_scaler import OptState
                st["found_inf_per_device"] = {flat.device: found_inf.reshape(1)}
                st["stage"] = OptState.STEPPED
                self.scaler.update()
            return grad_norm
        if st is not None:
            from torch.amp.grad_scaler import OptState
            found_inf = torch.zeros((), dtype=torch.float32, device=flat.device)
            torch._amp_foreach_non_finite_check_and_unscale_([flat], found_inf, inv_scale)
            st["found_inf_per_device"] = {flat.device: found_inf}
            st["stage"] = OptState.UNSCALED
        grad_norm = torch.linalg.vector_norm(flat)
        flat.mul_(torch.clamp(p.grad_clip / (grad_norm + 1e-6), max=1.0))   # clip_grad_norm_: always multiplies
        self.scaler.step(self.optimizer)
        self.scaler.update()
        return grad_norm

    def _step_autograd(self, obs, mb, value_adapter, amp_dtype, amp_dev):
        with autocast(device_type=amp_dev, dtype=amp_dtype, enabled=self.params.use_amp):
            output = self.forward_model(obs)
            flat_logits = output.policy_logits.reshape(obs.shape[0], -1)
            loss, pl, vl, sl, ent, flags = self._losses(flat_logits, output.value_logits, output.score_lead, mb, value_adapter)
        self._check_flags(flags)
        self.optimizer.zero_grad(set_to_none=True)
        self.scaler.scale(loss).backward()
        if self.grad_sync is not None:
            self.grad_sync.all_reduce_params(self.model.parameters())
        return pl, vl, sl, ent, output.value_logits.detach()

    # ---- update --------------------------------------------------------------------------------------
    def update(self, buffer: KataGoRolloutBuffer, next_values: torch.Tensor, value_adapter: Any | None = None,
               heartbeat_fn: Any | None = None) -> dict[str, float]:
        """Reference katago_ppo.py:619-991. Returns the same metrics dict; clears the buffer; leaves
        `forward_model` in train mode."""
        self.forward_model.train()
        assert self.forward_model.training
        self._timing_events["update_forward_backward_ms"].clear()
        self._timing_events["gae_ms"].clear()
        data = buffer.flatten()
        T, N = buffer.size, buffer.num_envs
        total = data["rewards"].numel()
        device = next(self.model.parameters()).device
        p = self.params

        if device.type == "cuda" and data["observations"].device == device:
            side = None   # device-resident buffer (KataGoRolloutBuffer(device=...)): nothing to ship, masks already bit-packed
            gpu_obs = data["observations"]
            gpu_masks = data["legal_masks_packed"] if dict.__contains__(data, "legal_masks_packed") else data["legal_masks"]
        elif device.type == "cuda":
            side = torch.cuda.Stream(device)
            with torch.cuda.stream(side):
                gpu_obs = data["observations"].pin_memory().to(device, non_blocking=True)
                gpu_masks = data["legal_masks"].pin_memory().to(device, non_blocking=True)
                gpu_masks = policy_ops.pack_mask_bits(gpu_masks)   # 8x fewer mask bytes for every gather and loss kernel below
        else:
            side = None
            gpu_obs, gpu_masks = data["observations"], data["legal_masks"]

        adv = self._advantages(data, T, N, next_values, device).float().contiguous()
        returns = adv + data["values"].reshape(-1).float().to(device)  # value targets of the scalar contract (ppo.py)
        if adv.numel() > 1:
            gae_mod.normalize_advantages_(adv)
        mv = lambda t: t.to(device, non_blocking=True)  # noqa: E731
        g_actions, g_old, g_cats, g_score = mv(data["actions"]), mv(data["log_probs"]), mv(data["value_categories"]), mv(data["score_targets"])
        if side is not None:
            torch.cuda.current_stream(device).wait_stream(side)

        batch_size = min(p.batch_size, total)
        amp_dtype, amp_dev = _amp_dtype_and_device(p.use_amp, device)
        km = self._kernel_model(device) if self.forward_model is self._base() else None
        bn_sync = getattr(km, "bn_sync", None) if km is not None else None
        if bn_sync is not None and hasattr(bn_sync, "begin_update"):
            # SyncBatchNorm exchanges assume equal shards and the same number of minibatches on every rank (count =
            # local batch x world): verified here, and the ranks are lined up before the first exchange
            bn_sync.begin_update(total, batch_size, device)
        zero = lambda: torch.zeros((), device=device)  # noqa: E731
        acc = {k: zero() for k in ("policy_loss", "value_loss", "score_loss", "entropy", "gradient_norm")}
        n_updates = 0
        last_value_logits = last_cats = None
        # one gather kernel per minibatch (observations, bit-packed masks, six scalars) instead of eight index ops
        fused_gather = (device.type == "cuda" and gpu_masks.dtype == torch.int32 and gpu_obs.dtype == torch.float32
                        and gpu_obs.is_contiguous() and gpu_obs[0].numel() % 2 == 0)
        for _ in range(p.epochs_per_batch):
            perm = torch.randperm(total, device=device)
            for start in range(0, total, batch_size):
                idx = perm[start:start + batch_size]
                if fused_gather:
                    obs_b, *mb = policy_ops.gather_minibatch(gpu_obs, gpu_masks, g_actions, g_old, adv, g_cats, g_score, returns, idx)
                    mb = tuple(mb)
                else:
                    obs_b = gpu_obs[idx]
                    mb = (gpu_masks[idx], g_actions[idx], g_old[idx], adv[idx], g_cats[idx], g_score[idx], returns[idx])
                tok = self._events(device, "update_forward_backward_ms")
                if km is not None:
                    pl, vl, sl, ent, v_logits = self._step_fused(km, obs_b, mb, value_adapter)
                else:
                    pl, vl, sl, ent, v_logits = self._step_autograd(obs_b, mb, value_adapter, amp_dtype, amp_dev)
                grad_norm = self._optimizer_tail()
                self._events_end(tok)
                acc["policy_loss"] += pl.detach(); acc["value_loss"] += vl.detach(); acc["score_loss"] += sl.detach()
                acc["entropy"] += ent.detach()
                acc["gradient_norm"] += grad_norm.detach() if isinstance(grad_norm, torch.Tensor) else float(grad_norm)
                n_updates += 1
                last_value_logits, last_cats = v_logits, mb[4]
                if heartbeat_fn is not None:
                    heartbeat_fn()
        del gpu_obs, gpu_masks
        buffer.clear()
        denom = max(n_updates, 1)
        metrics = {k: (v / denom).item() for k, v in acc.items()}
        if bn_sync is not None and hasattr(bn_sync, "check"):
            bn_sync.check()   # the .item() reads above synchronised the stream: report a lost peer instead of NaN metrics
        if last_value_logits is not None and last_value_logits.shape[-1] == 3:
            valid = last_cats >= 0
            if valid.any():
                metrics.update(compute_value_metrics(last_value_logits[valid], last_cats[valid]))
        self.forward_model.train()
        return metrics
