"""SE-ResNet with KataGo-style global-pool bias — drop-in for the reference's `se_resnet` registry
entry (keisei/training/models/se_resnet.py:15-159).

Parameters and buffers are ordinary fp32 `nn.Parameter`s held by standard `nn.Conv2d` /
`nn.BatchNorm2d` / `nn.Linear` containers with the reference's attribute names, so `state_dict()`
keys, shapes and registration order are identical (strict loads, positional Adam state).
The containers are never *called* on CUDA: a CUDA observation runs the whole network through one
C-ABI call (`keisei_b200::seresnet_forward`, csrc/model.cu) — tcgen05 implicit-GEMM convolutions
in bf16 when AMP is configured, fp32-accurate SIMT kernels otherwise. CPU tensors (the reference's
showcase sidecar and unit tests are CPU-only) run the same graph with plain PyTorch CPU ops.
"""
from __future__ import annotations

import os
import threading
from collections import OrderedDict
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import model_ops
from .katago_base import KataGoBaseModel, KataGoOutput


@dataclass(frozen=True)
class SEResNetParams:
    num_blocks: int = 40
    channels: int = 256
    se_reduction: int = 16
    global_pool_channels: int = 128
    policy_channels: int = 32
    value_fc_size: int = 256
    score_fc_size: int = 128
    obs_channels: int = 50

    def __post_init__(self) -> None:
        for name in ("num_blocks", "channels", "se_reduction", "global_pool_channels", "policy_channels",
                     "value_fc_size", "score_fc_size", "obs_channels"):
            if getattr(self, name) < 1:
                raise ValueError(f"{name} must be >= 1, got {getattr(self, name)}")
        if self.channels // self.se_reduction < 1:
            raise ValueError(
                f"channels ({self.channels}) // se_reduction ({self.se_reduction}) must be >= 1")


def rollout_bucket(batch: int) -> int:
    """Graph-replayed rollout batches are padded up to a bucket so that the learner / opponent sub-batches of a
    split-merge step — whose sizes change on almost every step (`obs[learner_mask]`, katago_loop.py:337-344) — hit a
    handful of captured graphs instead of capturing a new one per size. Eval-mode boards are independent, so the padding
    rows only cost their (launch-bound) share of compute and the outputs are sliced back."""
    step = 8 if batch <= 64 else 32 if batch <= 512 else 128 if batch <= 2048 else 256
    return (batch + step - 1) // step * step


class _GraphCache:
    """Captured rollout graphs of one model. Keys carry the calling THREAD (a tournament thread or DynamicTrainer using the
    same model concurrently gets its own graphs and static output buffers, dynamic_trainer.py:44-50), the bucketed batch
    size(s), dtype and device. LRU, bounded by bytes of static buffers and by entry count. A key is captured on its
    SECOND sighting: a size that shows up once never pays a warm-up + capture + instantiate."""

    def __init__(self, max_bytes: int, max_entries: int = 32) -> None:
        self.max_bytes, self.max_entries = max_bytes, max_entries
        self.entries: "OrderedDict[tuple, dict]" = OrderedDict()
        self.sightings: dict = {}
        self.bytes = 0
        self.captures = self.hits = self.misses = 0
        self.lock = threading.Lock()

    def __len__(self) -> int:
        return len(self.entries)

    # copy.deepcopy(model) / pickling (opponent snapshots, torch.save(model)): captured graphs never travel with a copy
    def __deepcopy__(self, memo) -> "_GraphCache":
        return _GraphCache(self.max_bytes, self.max_entries)

    def __reduce__(self):
        return (_GraphCache, (self.max_bytes, self.max_entries))

    def clear(self) -> None:
        with self.lock:
            self.entries.clear(); self.sightings.clear(); self.bytes = 0

    def lookup(self, key):
        with self.lock:
            ent = self.entries.get(key)
            if ent is not None:
                self.entries.move_to_end(key)
                self.hits += 1
            else:
                self.misses += 1
            return ent

    def drop(self, key) -> None:
        with self.lock:
            ent = self.entries.pop(key, None)
            if ent is not None:
                self.bytes -= ent["nbytes"]

    def should_capture(self, key) -> bool:
        with self.lock:
            n = self.sightings.get(key, 0) + 1
            if len(self.sightings) > 4096:
                self.sightings.clear()
            self.sightings[key] = n
            return n >= 2

    def store(self, key, ent: dict) -> None:
        with self.lock:
            self.entries[key] = ent
            self.bytes += ent["nbytes"]
            self.captures += 1
            while len(self.entries) > 1 and (self.bytes > self.max_bytes or len(self.entries) > self.max_entries):
                _, old = self.entries.popitem(last=False)
                self.bytes -= old["nbytes"]


def _tensor_bytes(obj) -> int:
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    if isinstance(obj, (tuple, list)):
        return sum(_tensor_bytes(o) for o in obj)
    return 0


_capture_lock = threading.Lock()   # stream capture is process-wide state: one capture at a time


def _global_pool(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) -> (B,3C): mean, max, population std (reference se_resnet.py:93-98)."""
    return torch.cat([x.mean(dim=(-2, -1)), x.amax(dim=(-2, -1)), x.std(dim=(-2, -1), correction=0)], dim=-1)


class GlobalPoolBiasBlock(nn.Module):
    """conv1 -> BN -> ReLU -> + global_fc(pool(block input)) -> conv2 -> BN -> SE(scale, shift)
    -> + residual -> ReLU  (reference se_resnet.py:40-90). Parameter container; `forward` is the
    CPU path."""

    def __init__(self, channels: int, se_reduction: int, global_pool_channels: int) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(channels)
        self.global_fc = nn.Sequential(
            nn.Linear(channels * 3, global_pool_channels), nn.ReLU(), nn.Linear(global_pool_channels, channels))
        se_hidden = channels // se_reduction
        self.se_fc1 = nn.Linear(channels, se_hidden)
        self.se_fc2 = nn.Linear(se_hidden, channels * 2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = F.relu(self.bn1(self.conv1(x)))
        out = out + self.global_fc(_global_pool(x))[:, :, None, None]
        out = self.bn2(self.conv2(out))
        se = self.se_fc2(F.relu(self.se_fc1(out.mean(dim=(-2, -1)))))
        scale, shift = se.chunk(2, dim=-1)
        out = out * torch.sigmoid(scale)[:, :, None, None] + shift[:, :, None, None]
        return F.relu(out + x)


class SEResNetModel(KataGoBaseModel):
    """3-head SE-ResNet. `forward(obs) -> KataGoOutput`; ValueError on a bad observation shape."""

    def __init__(self, params: SEResNetParams) -> None:
        super().__init__()
        self.params = params
        ch = params.channels
        self.input_conv = nn.Conv2d(params.obs_channels, ch, 3, padding=1, bias=False)
        self.input_bn = nn.BatchNorm2d(ch)
        self.blocks = nn.Sequential(*[
            GlobalPoolBiasBlock(ch, params.se_reduction, params.global_pool_channels)
            for _ in range(params.num_blocks)])
        self.policy_conv1 = nn.Conv2d(ch, params.policy_channels, 1, bias=False)
        self.policy_bn1 = nn.BatchNorm2d(params.policy_channels)
        self.policy_conv2 = nn.Conv2d(params.policy_channels, self.SPATIAL_MOVE_TYPES, 1)
        self.value_fc1 = nn.Linear(ch * 3, params.value_fc_size)
        self.value_fc2 = nn.Linear(params.value_fc_size, 3)
        self.score_fc1 = nn.Linear(ch * 3, params.score_fc_size)
        self.score_fc2 = nn.Linear(params.score_fc_size, 1)
        # kernel-side state (not part of state_dict)
        self._wpack: torch.Tensor | None = None
        self._wpack_key: tuple | None = None
        self.use_tensor_cores: bool = True   # tcgen05 convolutions whenever the bf16 path is selected
        self.last_policy_buffer: torch.Tensor | None = None  # padded (B, 11264) logits of the last CUDA forward
        self._tables_cache = None
        self._grad_sizes: list[int] = []
        # captured rollout forwards: (thread, bucketed batch, dtype, device, use_tc) -> graph (see _GraphCache)
        self._graphs = _GraphCache(int(float(os.environ.get("KB_GRAPH_CACHE_GB", "24")) * (1 << 30)))
        self._group_graphs = _GraphCache(int(float(os.environ.get("KB_GRAPH_CACHE_GB", "24")) * (1 << 30)))
        # rollout batches up to this size replay a CUDA graph (0 disables); KB_GRAPH_MAX_BATCH overrides the default
        self.graph_max_batch: int = int(os.environ.get("KB_GRAPH_MAX_BATCH", "4096"))
        self.graph_replayed_kernels: int = 0  # library kernels launched through graph replays (kb_launch_count sees captures only)
        self.bn_sync = None                  # distributed.BatchNormSync: global-batch BatchNorm statistics (SyncBatchNorm)
        # graph-replayed rollout batches of at least this many boards can run as two half batches on two captured
        # branches (boards are independent in eval mode): one half's HBM-bound block tails then execute under the other
        # half's tensor-bound convolutions. Off by default (0) since the block tail is fused into conv2's epilogue for the
        # 256-channel network (csrc/conv_tc.cu, fused SE tail) and no separate tail kernels are left to overlap; it still
        # pays when KB_FUSED_TAIL=0 or for widths the fused tail does not cover (KB_ROLLOUT_SPLIT_MIN overrides)
        self.rollout_split_min: int = int(os.environ.get("KB_ROLLOUT_SPLIT_MIN", "0"))
        self.rollout_split_ways: int = int(os.environ.get("KB_ROLLOUT_SPLIT_WAYS", "2"))   # number of parallel branches

    # ---- kernel plumbing ---------------------------------------------------------------------
    def _desc(self) -> list[int]:
        p = self.params
        return [p.num_blocks, p.channels, p.channels // p.se_reduction, p.global_pool_channels, p.policy_channels,
                p.value_fc_size, p.score_fc_size, p.obs_channels]

    def kernel_supported(self) -> bool:
        p = self.params
        return p.channels % 4 == 0 and p.channels <= 1024 and p.policy_channels <= 1024 and p.obs_channels <= 128

    def convert_sync_batchnorm(self, bn_sync) -> "SEResNetModel":
        """The CUDA-path counterpart of `torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)` (reference
        katago_loop.py:494-497, on by default under DDP): training-mode BatchNorm layers normalise with the statistics
        of the batch over ALL ranks. `bn_sync` is a `keisei_b200.distributed.BatchNormSync` (anything with
        `.world_size` and `.all_reduce_(float64 tensor)`); None switches back to per-rank statistics. Modules, parameter
        names and `state_dict` are untouched. CPU tensors keep the plain PyTorch path (torch's SyncBatchNorm itself is
        GPU-only)."""
        self.bn_sync = bn_sync
        return self

    def _tables(self) -> tuple[list[torch.Tensor], list[torch.Tensor]]:
        return list(self.parameters()), list(self.buffers())

    def _ptr_tables(self) -> "model_ops.PointerTables":
        """Cached parameter/buffer pointer tables (rebuilt when the module is moved or re-materialised)."""
        t = self._tables_cache
        if t is None or not t.valid():
            params, buffers = self._tables()
            t = self._tables_cache = model_ops.PointerTables(params, buffers, self._desc())
            self._grad_sizes = [p.numel() for p in params]
        return t

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): storages change
        self._tables_cache = None
        self._wpack_key = None
        self._graphs.clear()
        self._group_graphs.clear()
        return super()._apply(fn, *args, **kwargs)

    def invalidate_packed_weights(self) -> None:
        """Force a re-pack of the kernel-side weight shadow before the next CUDA forward. The cache key also watches the
        autograd version counters, but writes that bypass them — `Adam(fused=True).step()`, `t.data` writes such as
        `GradSync.broadcast_parameters`, `p.data.copy_` — need this explicit call (the trainer makes it after every
        optimiser step; `load_state_dict` does through the hook below)."""
        self._wpack_key = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed_weights()
        return out

    def _act_dtype(self, device: torch.device) -> torch.dtype:
        if self._amp_enabled and self._amp_dtype == torch.bfloat16:
            return torch.bfloat16
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def _packed(self, params, buffers, dtype: torch.dtype) -> torch.Tensor:
        dev = params[0].device
        key = (dtype, dev, sum(p._version for p in params) + sum(b._version for b in buffers),
               params[0].data_ptr())
        if self._wpack is None or self._wpack_key != key:
            code = 0 if dtype == torch.float32 else 1
            nbytes = model_ops.wpack_bytes(self._desc(), code)
            if self._wpack is None or self._wpack.numel() != nbytes or self._wpack.device != dev:
                self._wpack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            model_ops.pack_weights(params, buffers, self._desc(), code, self._wpack)
            self._wpack_key = key
        return self._wpack

    def _forward_cuda(self, obs: torch.Tensor) -> KataGoOutput:
        if not self.kernel_supported():
            raise model_ops._lib.KeiseiB200Error(
                f"SEResNetParams {self.params} not supported by the CUDA kernels (channels must be a multiple of 4)")
        tables = self._ptr_tables()
        params, buffers = tables.params, tables.buffers
        dtype = self._act_dtype(obs.device)
        code = 0 if dtype == torch.float32 else 1
        training = self.training
        wpack = self._packed(params, buffers, dtype)
        if torch.is_grad_enabled() and training:
            # autograd path: the torch.library op (gradients flow to every parameter)
            policy_buf, value, score, _ws, new_stats = model_ops.seresnet_forward(
                obs, params, buffers, wpack, self._desc(), training, code, bool(self.use_tensor_cores),
                model_ops.register_bn_sync(self.bn_sync))
        else:
            # no graph needed (rollout / evaluation): straight to the C-ABI, no dispatcher overhead
            policy_buf, value, score, _ws, new_stats = model_ops.seresnet_forward_raw(
                obs, tables, wpack, training, code, bool(self.use_tensor_cores), self.bn_sync)
        if training:
            self._store_running_stats(buffers, new_stats)
        B = obs.shape[0]
        self.last_policy_buffer = policy_buf
        policy = policy_buf[:, :model_ops.POLICY_A].view(B, 9, 9, self.SPATIAL_MOVE_TYPES)
        return KataGoOutput(policy_logits=policy, value_logits=value, score_lead=score)

    # ---- CUDA-graph replay for small rollout batches ---------------------------------------------
    @torch.no_grad()
    def _eval_forward_impl(self, obs: torch.Tensor) -> KataGoOutput:
        """`_forward_impl` in evaluation mode whatever the module's current mode (restored afterwards)."""
        if not self.training:
            return self._forward_impl(obs)
        self.eval()
        try:
            return self._forward_impl(obs)
        finally:
            self.train()

    def rollout_forward(self, obs: torch.Tensor, eval_mode: bool = False) -> KataGoOutput:
        """Eval-mode no-grad forward for action selection (reference katago_ppo.py:575-580 / katago_loop.py:337-344,
        404-406: per-step inference on 64..512 boards, several sub-batches per step in league play).

        For batches up to `graph_max_batch` the ~290 kernel launches of the network are captured into a CUDA graph per
        (thread, batch BUCKET, dtype) and replayed: at these sizes the step is launch-bound, not GPU-bound. The batch is
        padded up to `rollout_bucket(B)` inside the graph's static input (boards are independent in eval mode) so
        varying sub-batch sizes share a few graphs; a bucket is captured the second time it is seen. The returned
        tensors are views of the graph's static output buffers — owned by the calling thread, valid until its next
        `rollout_forward` in the same bucket (`select_actions` consumes them immediately). Larger batches, CPU tensors
        and training mode go through the ordinary `forward`.

        `eval_mode=True`: evaluate in inference mode WITHOUT touching the module's `training` flags — the kernels take the
        mode as an argument, and `model.eval()` + `model.train()` walk ~600 submodules in Python (0.72 + 0.76 ms measured,
        with the GPU idle: a quarter of a 512-board step). `select_actions` uses it."""
        if ((self.training and not eval_mode) or not obs.is_cuda or obs.shape[0] > self.graph_max_batch or not self.kernel_supported()
                or obs.ndim != 4 or tuple(obs.shape[1:]) != (self.params.obs_channels, 9, 9) or obs.shape[0] < 1):
            return self._eval_forward_impl(obs) if eval_mode else self._forward_impl(obs)
        tables = self._ptr_tables()
        dtype = self._act_dtype(obs.device)
        code = 0 if dtype == torch.float32 else 1
        wpack = self._packed(tables.params, tables.buffers, dtype)   # re-packs in place when a parameter changed
        B = obs.shape[0]
        Bb = rollout_bucket(B)
        key = (threading.get_ident(), Bb, code, obs.device, bool(self.use_tensor_cores))
        ent = self._graphs.lookup(key)
        if ent is not None and (ent["wpack_ptr"] != wpack.data_ptr() or ent["tables"] is not tables):
            self._graphs.drop(key)
            ent = None
        if ent is None:
            if not self._graphs.should_capture(key):
                # first sighting of this bucket: the same C call the graph would replay, launched directly
                policy_buf, value, score, _ws, _ = self._captured_forward(obs, tables, wpack, code)
                self.last_policy_buffer = policy_buf
                policy = policy_buf[:, :model_ops.POLICY_A].view(B, 9, 9, self.SPATIAL_MOVE_TYPES)
                return KataGoOutput(policy_logits=policy, value_logits=value, score_lead=score)
            static_obs = torch.zeros((Bb, self.params.obs_channels, 9, 9), dtype=torch.float32, device=obs.device)
            static_obs[:B].copy_(obs)
            with _capture_lock:
                cur = torch.cuda.current_stream(obs.device)
                side = torch.cuda.Stream(obs.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):  # warm-up outside the capture: one-time function attributes / driver lookups
                    model_ops.seresnet_forward_raw(static_obs, tables, wpack, False, code, bool(self.use_tensor_cores))
                cur.wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                n0 = model_ops._lib.launch_count()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    out = self._captured_forward(static_obs, tables, wpack, code)
                n_kernels = model_ops._lib.launch_count() - n0   # library kernels inside the graph (each replay launches them)
            ent = {"graph": graph, "obs": static_obs, "out": out, "kernels": n_kernels, "wpack_ptr": wpack.data_ptr(),
                   "tables": tables, "nbytes": _tensor_bytes(static_obs) + _tensor_bytes(out)}
            self._graphs.store(key, ent)
        ent["obs"][:B].copy_(obs)
        ent["graph"].replay()
        self.graph_replayed_kernels += ent["kernels"]
        policy_buf, value, score, _ws, _ = ent["out"]
        policy_buf, value, score = policy_buf[:B], value[:B], score[:B]
        self.last_policy_buffer = policy_buf
        policy = policy_buf[:, :model_ops.POLICY_A].view(B, 9, 9, self.SPATIAL_MOVE_TYPES)
        return KataGoOutput(policy_logits=policy, value_logits=value, score_lead=score)

    def _captured_forward(self, obs: torch.Tensor, tables, wpack: torch.Tensor, code: int, num_sms: int | None = None):
        """Body of the rollout graph: one C call, or two half-batch calls on two branches of the capture."""
        B = obs.shape[0]
        use_tc = bool(self.use_tensor_cores)
        if num_sms is not None or self.rollout_split_min <= 0 or B < self.rollout_split_min:
            return model_ops.seresnet_forward_raw(obs, tables, wpack, False, code, use_tc, num_sms=num_sms)
        # split on 3-board tile boundaries; every part but the last gets a whole number of conv rounds (74 board groups x
        # 2 channel halves = one CTA per SM), so only the last part pays a partial round
        groups = (B + 2) // 3
        ways = max(2, min(self.rollout_split_ways, groups))
        if groups >= 2 * ways * 74:
            per = max(74, int(round(groups / ways / 74)) * 74)
        else:
            per = max(1, groups // ways)
        cuts = [min(B, k * per * 3) for k in range(1, ways)]
        bounds = sorted(set([0] + [c for c in cuts if 0 < c < B] + [B]))
        dev = obs.device
        dtype = torch.float32 if code == 0 else torch.bfloat16
        policy = torch.empty((B, model_ops.POLICY_PITCH), dtype=dtype, device=dev)
        value = torch.empty((B, 3), dtype=torch.float32, device=dev)
        score = torch.empty((B, 1), dtype=torch.float32, device=dev)
        cur = torch.cuda.current_stream(dev)
        spans = list(zip(bounds[:-1], bounds[1:]))
        # fork every side branch BEFORE anything is enqueued on the capturing stream: a branch that waited on the stream
        # after part 0 was enqueued would depend on all of part 0 and the parts would run one after the other
        sides = [torch.cuda.Stream(dev) for _ in spans[1:]]
        for side in sides:
            side.wait_stream(cur)
        workspaces = []
        for k, (lo, hi) in enumerate(spans):
            outs = (policy[lo:hi], value[lo:hi], score[lo:hi])
            if k == 0:
                workspaces.append(model_ops.seresnet_forward_raw(obs[lo:hi], tables, wpack, False, code, use_tc, out=outs)[3])
            else:
                with torch.cuda.stream(sides[k - 1]):
                    workspaces.append(model_ops.seresnet_forward_raw(obs[lo:hi], tables, wpack, False, code, use_tc, out=outs)[3])
        for side in sides:
            cur.wait_stream(side)
        return policy, value, score, tuple(workspaces), None

    @torch.no_grad()
    def _store_running_stats(self, buffers: list[torch.Tensor], new_stats: torch.Tensor) -> None:
        """Copy the kernel's updated BatchNorm running statistics back into the registered buffers
        (buffer order: running_mean, running_var, num_batches_tracked per BN layer) with batched
        foreach ops; bumps the buffers' version counters so the folded eval weights get refreshed."""
        dst, src, nbt = [], [], []
        for layer in range(len(buffers) // 3):
            rm, rv, n = buffers[3 * layer], buffers[3 * layer + 1], buffers[3 * layer + 2]
            c = rm.numel()
            dst += [rm, rv]
            src += [new_stats[layer, 0, :c], new_stats[layer, 1, :c]]
            nbt.append(n)
        torch._foreach_copy_(dst, src)
        torch._foreach_add_(nbt, 1)

    # ---- CPU path (showcase sidecar, unit tests): plain PyTorch on the same parameters ----------
    def _forward_host(self, obs: torch.Tensor) -> KataGoOutput:
        x = F.relu(self.input_bn(self.input_conv(obs)))
        x = self.blocks(x)
        p = F.relu(self.policy_bn1(self.policy_conv1(x)))
        p = self.policy_conv2(p).permute(0, 2, 3, 1)
        pool = _global_pool(x)
        v = self.value_fc2(F.relu(self.value_fc1(pool)))
        s = self.score_fc2(F.relu(self.score_fc1(pool)))
        return KataGoOutput(policy_logits=p, value_logits=v, score_lead=s)

    def _forward_impl(self, obs: torch.Tensor) -> KataGoOutput:
        p = self.params
        if obs.ndim != 4 or obs.shape[1] != p.obs_channels or obs.shape[2] != 9 or obs.shape[3] != 9:
            raise ValueError(f"Expected obs shape (batch, {p.obs_channels}, 9, 9), got {tuple(obs.shape)}")
        if obs.is_cuda:
            return self._forward_cuda(obs)
        if self._amp_enabled:
            with torch.amp.autocast(device_type=self._amp_device_type, dtype=self._amp_dtype):
                return self._forward_host(obs)
        return self._forward_host(obs)


# ---- grouped rollout: several (model, sub-batch) pairs as parallel branches of ONE CUDA graph -------------------------
@torch.no_grad()
def _sm_shares(batches: "list[int]", sms: int) -> "list[int]":
    """SM budget of every branch of a grouped rollout graph, proportional to its boards, in whole CTA pairs (>= 1 pair each).
    The convolutions are persistent one-CTA-per-SM kernels: sized for the whole device, the branches of a graph would take
    turns; sized for their share they run side by side, each CTA pair with about as many boards as if the sub-batches had
    been one batch."""
    pairs_total = max(len(batches), int(sms // 2 * float(os.environ.get("KB_GROUP_SM_FRACTION", "1.0"))))
    total = max(1, sum(batches))
    shares = [max(1, (pairs_total * b) // total) for b in batches]
    spare = pairs_total - sum(shares)
    order = sorted(range(len(batches)), key=lambda i: -(batches[i] / shares[i]))   # most boards per pair first
    k = 0
    while spare > 0 and order:
        shares[order[k % len(order)]] += 1
        spare -= 1
        k += 1
    return [2 * s for s in shares]


def rollout_forward_many(pairs: "list[tuple[SEResNetModel, torch.Tensor]]", eval_mode: bool = False) -> "list[KataGoOutput]":
    """Eval-mode forward of several independent (model, observations) pairs — the learner and its K league opponents
    on their sub-batches (reference katago_loop.py:284-431 `split_merge_step`, match_utils.py:124-293 `play_batch`: one
    sequential forward per model per step, 64..256 boards each). At these sizes a forward is a chain of ~290 kernels of
    10-20 us that fill a fraction of the GPU, so the pairs are captured as PARALLEL BRANCHES of one CUDA graph (one
    branch per pair, each on its own stream during capture) and replayed with a single launch: the GPU runs the
    branches side by side. Sub-batches are padded to `rollout_bucket` sizes inside the graph's static inputs, so the
    varying learner / opponent splits of consecutive steps reuse a few graphs (captured on the second sighting; the
    first falls back to one `rollout_forward` per pair). Results are bit-identical to `model.rollout_forward(obs)` per
    pair. Returned tensors are views of static buffers owned by the calling thread (valid until its next call with the
    same bucket signature). The cache lives on the first model of the group. Falls back to one `rollout_forward` per
    pair for CPU tensors, training-mode models or unsupported shapes. `eval_mode=True`: inference whatever the modules'
    `training` flags say, without toggling them (see `rollout_forward`)."""
    if not pairs:
        return []
    ok = all(isinstance(m, SEResNetModel) and (eval_mode or not m.training) and o.is_cuda and m.kernel_supported() and o.ndim == 4
             and tuple(o.shape[1:]) == (m.params.obs_channels, 9, 9) and 0 < o.shape[0] <= max(m.graph_max_batch, 0)
             for m, o in pairs) and len({o.device for _, o in pairs}) == 1
    if not ok or len(pairs) == 1:
        return [m.rollout_forward(o, eval_mode) if isinstance(m, SEResNetModel) else m.rollout_forward(o) for m, o in pairs]
    dev = pairs[0][1].device
    prep = []
    for m, o in pairs:
        tables = m._ptr_tables()
        dtype = m._act_dtype(dev)
        code = 0 if dtype == torch.float32 else 1
        wpack = m._packed(tables.params, tables.buffers, dtype)
        prep.append((m, o, tables, wpack, code))
    cache = pairs[0][0]._group_graphs
    key = (threading.get_ident(), dev) + tuple((id(m), rollout_bucket(o.shape[0]), code, bool(m.use_tensor_cores), wpack.data_ptr(), id(tables))
                                               for m, o, tables, wpack, code in prep)
    ent = cache.lookup(key)
    if ent is None:
        if not cache.should_capture(key):
            return [m.rollout_forward(o, eval_mode) for m, o in pairs]
        statics = []
        for _, o, *_rest in prep:
            so = torch.zeros((rollout_bucket(o.shape[0]),) + tuple(o.shape[1:]), dtype=torch.float32, device=dev)
            so[:o.shape[0]].copy_(o)
            statics.append(so)
        with _capture_lock:
            cur = torch.cuda.current_stream(dev)
            warm = torch.cuda.Stream(dev)
            warm.wait_stream(cur)
            with torch.cuda.stream(warm):  # warm-up outside the capture: one-time function attributes / driver lookups
                for (m, _, tables, wpack, code), so in zip(prep, statics):
                    model_ops.seresnet_forward_raw(so, tables, wpack, False, code, bool(m.use_tensor_cores))
            cur.wait_stream(warm)
            graph = torch.cuda.CUDAGraph()
            n0 = model_ops._lib.launch_count()
            outs = []
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                cap = torch.cuda.current_stream(dev)
                # fork every branch before anything is enqueued on the capturing stream (a later fork would depend on pair 0)
                branches = [torch.cuda.Stream(dev) for _ in prep[1:]]
                for br in branches:
                    br.wait_stream(cap)
                shares = _sm_shares([so.shape[0] for so in statics], model_ops.sm_count(dev))
                for i, ((m, _, tables, wpack, code), so) in enumerate(zip(prep, statics)):
                    if i == 0:
                        outs.append(m._captured_forward(so, tables, wpack, code, num_sms=shares[i]))
                    else:
                        with torch.cuda.stream(branches[i - 1]):
                            outs.append(m._captured_forward(so, tables, wpack, code, num_sms=shares[i]))
                for br in branches:
                    cap.wait_stream(br)
            n_kernels = model_ops._lib.launch_count() - n0
        ent = {"graph": graph, "obs": statics, "outs": outs, "kernels": n_kernels,
               "keep": [(tables, wpack) for _, _, tables, wpack, _ in prep], "nbytes": _tensor_bytes(statics) + _tensor_bytes(outs)}
        cache.store(key, ent)
    for so, (_, o, *_rest) in zip(ent["obs"], prep):
        so[:o.shape[0]].copy_(o)
    ent["graph"].replay()
    pairs[0][0].graph_replayed_kernels += ent["kernels"]
    res = []
    for (m, o, *_rest), out in zip(prep, ent["outs"]):
        n = o.shape[0]
        policy_buf, value, score = out[0][:n], out[1][:n], out[2][:n]
        m.last_policy_buffer = policy_buf
        res.append(KataGoOutput(policy_logits=policy_buf[:, :model_ops.POLICY_A].view(n, 9, 9, m.SPATIAL_MOVE_TYPES),
                                value_logits=value, score_lead=score))
    return res
