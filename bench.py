#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], config 5 shape per rank when N > 1): SE-ResNet 40x256 rollout
inference, bf16, 4096 synthetic boards per GPU. One "step" = one `select_actions`-equivalent pass:
network forward (eval BatchNorm) + legal-mask softmax + sample + log-prob + scalar value.
`value` = positions/s with inputs resident in HBM (device-timed, max over ranks); `e2e` = the same
through `KataGoPPOAlgorithm.select_actions` with HOST buffers (pinned H2D of obs + mask and D2H of
actions/log-probs/values inside the timed region). Rollout shards across ranks with no collective
(weak scaling). The line also carries `update` (BASELINE.json configs[2]: KataGo-PPO update,
8192 samples split over the ranks, fwd+loss+bwd+all-reduce+clip+Adam) as extra keys, the dominant
kernel's `roofline`, and the `cpu_baseline` (oracle port on the host cores, bounded sample).

`--impl reference` times the reference's CPU path (the oracle restatement — the reference itself is
Python/PyTorch and /root/reference does not exist on the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MODEL_CFG = dict(num_blocks=40, channels=256)           # reference SEResNetParams defaults (se_resnet.py:15-24)
ROLLOUT_B = 4096                                        # BASELINE.json configs[1]
UPDATE_GLOBAL_B = 8192                                  # BASELINE.json configs[2]
A = 11259
WORKLOAD = "SE-ResNet 40x256 rollout, 4096 boards/GPU"
CONV_TRAFFIC_BYTES = 303.8e6                            # ncu dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/ncu_r2_conv_pair.txt)
CONV_FLOP_PER_POS = 2 * 81 * 256 * 2304                 # one 256->256 3x3 conv, SURVEY.md 8(d): 95.55 MFLOP
FWD_FLOP_PER_POS = 18.66e6 + 80 * 95.55e6               # trunk convs only: 7.663 GFLOP


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clocks / throttle reasons via NVML during the timed region."""

    def __init__(self, index: int) -> None:
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def synth_boards(B: int, seed: int, device):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(B, 50, 9, 9, generator=g)
    mask = torch.zeros(B, A, dtype=torch.bool)
    idx = torch.randint(0, A, (B, 80), generator=g)      # ~80 legal moves per position (realistic)
    mask.scatter_(1, idx, True)
    return obs.to(device), mask.to(device)


def dist_setup(n: int):
    import torch.distributed as dist
    if n > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        return dist.get_rank(), dist.get_world_size(), local
    return 0, 1, 0


def timed(fn, steps: int, warmup: int, device, world: int) -> float:
    """ms per step: CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def _cpu_threads() -> int:
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return torch.get_num_threads()


def cpu_baseline(batch: int = 256, reps: int = 6, warmup: int = 1) -> dict:
    """Oracle port (fp32 CPU PyTorch restatement of the reference) on the host cores: rollout step
    (eval forward + mask/softmax/Categorical sample + log-prob + scalar value) on a bounded sample."""
    from oracle import keisei_oracle as O
    from keisei_b200.models import SEResNetModel, SEResNetParams
    cores = _cpu_threads()
    torch.manual_seed(0)
    sd = SEResNetModel(SEResNetParams(**MODEL_CFG)).state_dict()
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(batch, 50, 9, 9, generator=g)
    mask = torch.zeros(batch, A, dtype=torch.bool)
    mask.scatter_(1, torch.randint(0, A, (batch, 80), generator=g), True)

    def step():
        with torch.no_grad():
            p, v, s = O.seresnet_forward(sd, obs, MODEL_CFG["num_blocks"], training=False)
            flat = p.reshape(batch, -1)
            probs = torch.softmax(flat.masked_fill(~mask, float("-inf")), -1)
            a = torch.distributions.Categorical(probs, validate_args=False).sample()
            O.rollout_log_prob(flat, mask, a)
            O.scalar_value(v)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return {"value": batch / dt, "unit": "positions/s", "cores": cores, "kind": "port",
            "sample": f"oracle fp32 rollout, batch {batch} x {reps}"}


def _oracle_train_step(cfg: dict, batch: int, seed: int):
    """Closure running one KataGo-PPO minibatch forward + losses + backward of the oracle (fp32, CPU)."""
    from oracle import keisei_oracle as O
    from keisei_b200.models import SEResNetModel, SEResNetParams
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in SEResNetModel(SEResNetParams(**cfg)).state_dict().items()}
    for t in sd.values():
        if t.is_floating_point():
            t.requires_grad_(True)
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(batch, 50, 9, 9, generator=g)
    mask = torch.rand(batch, A, generator=g) < 0.01          # ~110 legal per row
    acts = torch.randint(0, A, (batch,), generator=g)
    mask[torch.arange(batch), acts] = True
    old, adv = -3 * torch.rand(batch, generator=g), torch.randn(batch, generator=g)
    cats = torch.randint(-1, 3, (batch,), generator=g)
    score_t = torch.randn(batch, generator=g).clamp(-1.5, 1.5)
    nb = cfg["num_blocks"]

    def step():
        for t in sd.values():
            t.grad = None
        p, v, s = O.seresnet_forward(sd, obs, nb, training=True)
        O.ppo_losses(p, v, s, mask, acts, old, adv, cats, score_t)["loss"].backward()
    return step


def cpu_config1(reps: int = 5) -> dict:
    """BASELINE.json configs[0] — the mandatory CPU point: SE-ResNet 4x64 KataGo-PPO minibatch
    fwd+bwd, batch 256, all host cores, fp32, median of `reps` after 1 warm-up (profile_hotpath.py:168-206 discipline)."""
    cores = _cpu_threads()
    step = _oracle_train_step(dict(num_blocks=4, channels=64), 256, 3)
    step()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    med = sorted(ts)[len(ts) // 2]
    return {"value": 256 / med, "unit": "samples/s", "cores": cores, "ms_per_step": med * 1e3, "kind": "port"}


def cpu_update_baseline(batch: int = 32, reps: int = 2) -> dict:
    """CPU baseline of configs[2]: oracle fwd + ppo_losses + backward of the 40x256 model on a bounded sample
    (batch 32, per-sample throughput), all host cores."""
    cores = _cpu_threads()
    step = _oracle_train_step(MODEL_CFG, batch, 4)
    step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return {"value": batch / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"oracle fp32 fwd+loss+bwd, 40x256, batch {batch}, {reps} reps"}


def run_reference(args) -> None:
    """The reference's CPU path (oracle port) for the same metric / config: every step is one rollout pass over a bounded
    sample (batch 256 of the 4096 boards), exactly `--steps` timed steps after `--warmup` warm-ups."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 256
    t0 = time.perf_counter()
    cb = cpu_baseline(batch=batch, reps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {"metric": "rollout positions/s", "value": cb["value"], "unit": "positions/s", "n_gpus": args.gpus, "steps": max(1, args.steps),
            "warmup": max(1, args.warmup), "ms_per_step": 1000.0 * batch / cb["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "timed_sample": f"batch {batch} per step on host cores, scaled per position"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1)}
    print(json.dumps(line), flush=True)


def conv_roofline(device, reps: int = 20) -> dict:
    """Dominant kernel: conv3x3_tc_kernel (256->256, B=4096, folded-BN+ReLU+gpool-bias epilogue), timed back to back
    with CUDA events on the launching stream, alone — so the denominator is the measured BURST bf16 peak."""
    from keisei_b200 import model_ops
    pk = peaks()
    B, C = ROLLOUT_B, 256
    torch.manual_seed(0)
    xs = [torch.randn(B, 81, C, device=device).bfloat16() for _ in range(2)]  # 2 x 170 MB > 126 MB L2
    wf = model_ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=device) / 48, torch.bfloat16)
    sc, sh = torch.ones(C, device=device), torch.zeros(C, device=device)
    gb = torch.zeros(B, C, device=device)
    for i in range(3):
        model_ops.conv3x3(xs[i & 1], wf, backend=1, scale=sc, shift=sh, relu=True, gbias=gb)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        model_ops.conv3x3(xs[i & 1], wf, backend=1, scale=sc, shift=sh, relu=True, gbias=gb)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / reps
    flops = CONV_FLOP_PER_POS * B
    achieved = flops / (ms * 1e-3) / 1e12
    peak = pk["bf16_tflops"]
    return {"kernel": "conv3x3_tc2 (cta_group::2) 256->256 B=4096", "bound": "tensor", "achieved": round(achieved, 1), "peak": peak,
            "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture under profiles/
            "traffic": CONV_TRAFFIC_BYTES, "ms": round(ms, 4), "peak_src": f"{pk['source']} burst"}


def hbm_kernels(device) -> dict:
    """Achieved fraction of the measured HBM copy peak for the bandwidth-bound kernels of the path, each timed alone
    with CUDA events over alternating inputs larger than L2. Algorithmic bytes per unit: SURVEY.md 8(d)."""
    from keisei_b200 import policy_ops
    pk = peaks()["hbm_gbs"]
    out = {}
    g = torch.Generator(device=device).manual_seed(5)

    def timeit(fn, reps=10):
        """ms per call. The calls are captured into ONE CUDA graph (both alternating input sets, 2 calls) and the graph is
        replayed: these kernels run 10-100 us, less than the Python + allocator work of issuing them, and in the real step
        they are enqueued behind a long-running forward — the device time is what the roofline fraction is about."""
        for i in range(4):
            fn(i)
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                fn(0)
                fn(1)
        torch.cuda.current_stream(device).wait_stream(side)
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / (2 * reps)

    def rows(B):
        sets = []
        for _ in range(2):
            lg = torch.randn(B, 11264, device=device, generator=g).bfloat16()
            mk = torch.rand(B, A, device=device, generator=g) < 0.007
            ac = torch.randint(0, A, (B,), device=device, generator=g)
            mk[torch.arange(B, device=device), ac] = True
            sets.append((lg, mk, ac))
        return sets
    B = ROLLOUT_B
    sets = rows(B)
    vl = torch.randn(B, 3, device=device)
    # rollout sampling: (i) the dense one-CTA-per-row kernel on byte masks (whole row staged: SURVEY 8(d) algorithmic bytes
    # = bf16 logits + byte mask), (ii) what select_actions runs since round 2: bit-packed masks + the warp-per-row kernel
    # that gathers ONLY the legal logits (~80 of 11,259 here) — its traffic is the packed mask + one 32-byte sector per
    # legal logit, which is what its fraction is computed on (on the dense bytes it would read > 1)
    ms = timeit(lambda i: policy_ops.policy_sample(sets[i & 1][0][:, :A], sets[i & 1][1], vl, dense=True))
    bytes_row = A * 2 + policy_ops.mask_row_bytes(sets[0][1])
    out["policy_sample_dense"] = {"ms": round(ms, 4), "gbs": round(B * bytes_row / ms / 1e6, 1), "frac": round(B * bytes_row / ms / 1e6 / pk, 3)}
    bits = [policy_ops.pack_mask_bits(sets[k][1]) for k in range(2)]
    legal = float(sets[0][1].sum(1).float().mean())
    ms = timeit(lambda i: policy_ops.policy_sample(sets[i & 1][0][:, :A], bits[i & 1], vl))
    sparse_row = policy_ops.mask_row_bytes(bits[0]) + 32 * legal
    out["policy_sample"] = {"ms": round(ms, 4), "gbs": round(B * sparse_row / ms / 1e6, 1), "frac": round(B * sparse_row / ms / 1e6 / pk, 3),
                            "legal_per_row": round(legal, 1), "bytes_per_row": round(sparse_row), "dense_bytes_per_row": A * 2 + bits[0].shape[1] * 4}
    ms = timeit(lambda i: policy_ops.pack_mask_bits(sets[i & 1][1]))
    pack_row = A + bits[0].shape[1] * 4
    out["pack_mask_bits"] = {"ms": round(ms, 4), "gbs": round(B * pack_row / ms / 1e6, 1), "frac": round(B * pack_row / ms / 1e6 / pk, 3)}
    # update-side loss kernels, on what update() feeds them: bit-packed masks out of the device-resident buffer. Forward =
    # one streaming read of the raw bf16 logits (the reference's NaN guard looks at every logit) + the packed mask;
    # backward = one write of the (B, 11264) bf16 gradient + the packed mask + a 32-byte sector per legal logit (illegal
    # logits are never read)
    B = UPDATE_GLOBAL_B
    sets = rows(B)
    bits = [policy_ops.pack_mask_bits(sets[k][1]) for k in range(2)]
    mask_row = policy_ops.mask_row_bytes(bits[0])
    old, adv = -3 * torch.rand(B, device=device), torch.randn(B, device=device)
    keep = {}

    def fwd(i):
        lg, _, ac = sets[i & 1]
        keep[i & 1] = policy_ops.ppo_policy_loss(lg[:, :A], bits[i & 1], ac, old, adv, 0.2)
    ms = timeit(fwd)
    fwd_bytes = A * 2 + mask_row
    out["ppo_policy_fwd"] = {"ms": round(ms, 4), "gbs": round(B * fwd_bytes / ms / 1e6, 1), "frac": round(B * fwd_bytes / ms / 1e6 / pk, 3),
                             "bytes_per_row": fwd_bytes, "includes": "ppo_policy_reduce (1 CTA)"}
    g2 = torch.ones(2, device=device)

    def bwd(i):
        lg, _, ac = sets[i & 1]
        o = keep[i & 1]
        policy_ops.ppo_policy_loss_backward(lg[:, :A], bits[i & 1], ac, o[3], o[2], o[4], g2)
    ms = timeit(bwd)
    bwd_bytes = 11264 * 2 + mask_row + 32 * legal
    out["ppo_policy_bwd"] = {"ms": round(ms, 4), "gbs": round(B * bwd_bytes / ms / 1e6, 1), "frac": round(B * bwd_bytes / ms / 1e6 / pk, 3),
                             "bytes_per_row": round(bwd_bytes)}
    return out


def eager_port_gpu(device, obs, mask) -> dict:
    """Informational: the oracle port (plain PyTorch ops) on the SAME B200 under bf16 autocast — what PyTorch itself
    dispatches (cuDNN convolutions) for the rollout step, timed like the product arm. Library kernels, not this repo's."""
    from oracle import keisei_oracle as O
    from keisei_b200.models import SEResNetModel, SEResNetParams
    torch.manual_seed(0)
    sd = {k: v.to(device) for k, v in SEResNetModel(SEResNetParams(**MODEL_CFG)).state_dict().items()}
    B = obs.shape[0]

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            p, v, s = O.seresnet_forward(sd, obs, MODEL_CFG["num_blocks"], training=False)
            flat = p.reshape(B, -1)
            probs = torch.softmax(flat.masked_fill(~mask, float("-inf")), -1)
            a = torch.distributions.Categorical(probs, validate_args=False).sample()
            O.rollout_log_prob(flat, mask, a)
            O.scalar_value(v)
    ms = timed(step, 5, 3, device, 1)
    return {"value": B / (ms * 1e-3), "unit": "positions/s", "ms_per_step": ms, "what": "oracle port, torch eager + cuDNN, bf16 autocast, channels-first"}


def run_ours(args) -> None:
    from keisei_b200 import _lib
    from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
    from keisei_b200.models import SEResNetModel, SEResNetParams

    rank, world, local = dist_setup(args.gpus)
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    torch.manual_seed(0)
    model = SEResNetModel(SEResNetParams(**MODEL_CFG)).to(device)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True, batch_size=UPDATE_GLOBAL_B // world), model)
    obs, mask = synth_boards(ROLLOUT_B, 100 + rank, device)
    # non-trivial running statistics: a few train-mode batches
    model.train()
    with torch.no_grad():
        for i in range(2):
            model(obs[i * 256:(i + 1) * 256])
    B = ROLLOUT_B
    detail: dict = {}

    def step_device():
        algo.select_actions(obs, mask)

    ingest = RolloutIngest(algo, obs.cpu(), mask.cpu(), device)

    with ClockSampler(local) as clk:
        # kernels of this library launched in the timed region: direct launches (kb_launch_count) + the kernels inside
        # the CUDA graph that select_actions replays (counted once at capture, added per replay by the model)
        n0 = _lib.launch_count() + model.graph_replayed_kernels
        ms = timed(step_device, args.steps, args.warmup, device, world)
        launches = (_lib.launch_count() + model.graph_replayed_kernels - n0) * args.steps // (args.steps + args.warmup)
    ms_e2e = timed(ingest.step, args.steps, args.warmup, device, world)
    value = world * B / (ms * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)
    pk = peaks()

    line = {"metric": "rollout positions/s", "value": round(value, 1), "unit": "positions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "impl": "ours",
            "config": {"workload": WORKLOAD, "parallelism": f"shard{world}-no-comm", "cache": "inputs>L2"},
            "e2e": {"value": round(e2e, 1), "unit": "positions/s", "ms_per_step": round(ms_e2e, 4),
                    "h2d_bytes_per_step": ingest.h2d_bytes, "d2h_bytes_per_step": ingest.d2h_bytes},
            "gpu_launches": int(launches), "clocks": clk.summary(),
            "model_frac": round((FWD_FLOP_PER_POS * B / (ms * 1e-3) / 1e12) / pk["bf16_tflops_sustained"], 4)}

    upd = None
    if not args.no_update:
        upd, upd_detail = bench_update(args, algo, model, device, rank, world)
        detail["update"] = upd_detail
    if args.extra:
        detail["league_rollout"] = bench_league(algo, device, rank, world)
        algo.optimizer.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        detail["resnet_update"] = bench_resnet_update(args, device, rank, world)
    if rank == 0:
        line["roofline"] = conv_roofline(device)
        hk = hbm_kernels(device)
        detail["hbm_kernels"] = hk
        line["hbm_frac"] = {k: v["frac"] for k, v in hk.items()}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
            c1 = cpu_config1()
            detail["cpu_config1"] = c1
            line["cpu_config1"] = {"value": round(c1["value"], 1), "unit": "samples/s", "cores": c1["cores"]}   # 4x64 b256 PPO fwd+bwd
            line["cpu_baseline"]["value"] = round(line["cpu_baseline"]["value"], 2)
            if upd is not None:
                cu = cpu_update_baseline()
                detail["update_cpu_baseline"] = cu
                upd["cpu"] = {"value": round(cu["value"], 2), "cores": cu["cores"], "kind": "port", "sample": "b32 fwd+bwd"}
            if args.extra:
                try:
                    detail["gpu_eager_port"] = eager_port_gpu(device, obs, mask)
                    line["eager_port"] = round(detail["gpu_eager_port"]["value"], 1)
                except Exception as e:  # noqa: BLE001  (informational leg only)
                    detail["gpu_eager_port"] = {"error": repr(e)[:200]}
        if upd is not None:
            line["update"] = upd       # LAST key: the driver keeps the tail of the line
        print("BENCH_DETAIL " + json.dumps(detail), file=sys.stderr, flush=True)
        if args.detail_file:
            Path(args.detail_file).parent.mkdir(parents=True, exist_ok=True)
            Path(args.detail_file).write_text(json.dumps({"line": line, "detail": detail}, indent=1))
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


class RolloutIngest:
    """End-to-end rollout step from HOST buffers through the public API: `keisei_b200.ingest.PinnedIngest` (pinned,
    double-buffered host -> device transfer of observations + legal masks, inside the timed region) -> `select_actions`
    -> device -> host read of actions / log-probs / values. The transfer of step k+1 is submitted before the forward of
    step k, as a caller with the next observations in hand would."""

    def __init__(self, algo, obs_cpu, mask_cpu, device) -> None:
        from keisei_b200.ingest import PinnedIngest
        self.algo, self.device = algo, device
        n = obs_cpu.shape[0]
        # two distinct host batches in page-locked memory (the contract's "pinned host memory")
        self.src = [(obs_cpu.pin_memory(), mask_cpu.pin_memory()), (obs_cpu.flip(0).contiguous().pin_memory(), mask_cpu.flip(0).contiguous().pin_memory())]
        self.ingest = PinnedIngest(device, n, tuple(obs_cpu.shape[1:]), mask_cpu.shape[1], depth=2, pack_masks=True)
        self.h_out = [torch.empty(n, dtype=torch.int64).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n).pin_memory()]
        self.h2d_bytes = obs_cpu.numel() * 4 + mask_cpu.numel()
        self.d2h_bytes = n * 16
        self.k = 0
        self.slot = self.ingest.submit(*self.src[0])

    def step(self) -> None:
        cur = self.slot
        self.k += 1
        self.slot = self.ingest.submit(*self.src[self.k & 1])      # next step's observations: staged + H2D while this step computes
        d_obs, d_mask = self.ingest.get(cur)
        a, lp, v = self.algo.select_actions(d_obs, d_mask)
        self.ingest.release(cur)
        self.h_out[0].copy_(a, non_blocking=True); self.h_out[1].copy_(lp, non_blocking=True); self.h_out[2].copy_(v, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()


def _update_batch(Bu: int, seed: int, device):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(Bu, 50, 9, 9, generator=g).to(device)
    mask = torch.zeros(Bu, A, dtype=torch.bool)
    actions = torch.randint(0, A, (Bu,), generator=g)
    mask.scatter_(1, torch.randint(0, A, (Bu, 80), generator=g), True)
    mask[torch.arange(Bu), actions] = True
    mb = (mask.to(device), actions.to(device), (-3 * torch.rand(Bu, generator=g)).to(device), torch.randn(Bu, generator=g).to(device),
          torch.randint(-1, 3, (Bu,), generator=g).to(device), torch.randn(Bu, generator=g).clamp(-1.5, 1.5).to(device))
    return obs, mb


def bench_update(args, algo, model, device, rank, world) -> tuple[dict, dict]:
    """KataGo-PPO update step (BASELINE.json configs[2]): 8192 samples / world per rank; one step =
    forward (batch-stat BN) + fused losses + backward + gradient all-reduce + unscale/clip/Adam.
    Returns (compact dict for the JSON line, detail dict)."""
    import torch.distributed as dist
    from keisei_b200.distributed import BatchNormSync, GradSync
    Bu = UPDATE_GLOBAL_B // world
    obs, mb = _update_batch(Bu, 7 + rank, device)
    if world > 1:
        algo.grad_sync = GradSync()
        algo.grad_sync.broadcast_parameters(model)
    model.train()
    km = algo._kernel_model(device)

    def step():
        algo._step_fused(km, obs, mb, None)
        algo._optimizer_tail()

    steps, warm = max(2, args.steps), max(3, args.warmup)
    detail: dict = {"steps": steps, "warmup": warm, "global_batch": UPDATE_GLOBAL_B, "per_gpu_batch": Bu}
    ms_local_bn = ms_nccl_bn = ms_no_ar = allreduce_ms = None
    peer = None
    bn_kind = "local"
    if world > 1:
        # per-rank BatchNorm statistics first (plain DDP), then the reference's default: SyncBatchNorm
        # (katago_loop.py:494-497, sync_batchnorm = true) — the headline number for N > 1. The statistic exchange is
        # timed both ways: NCCL all-reduce per layer, and this library's one-kernel exchange over NVLink peer memory.
        from keisei_b200.distributed import PeerBatchNormSync
        gs = algo.grad_sync
        # step without any gradient exchange (weights diverge: timing only) against the step with the overlapped bucketed
        # exchange, INTERLEAVED (A B A B): the two legs differ by ~1 ms on a power-capped chip whose step time drifts by
        # more than that between legs timed minutes apart (tools/exp_allreduce_overlap.py)
        t_no, t_gs = [], []
        for rnd in range(2):
            algo.grad_sync = None
            t_no.append(timed(step, steps, warm if rnd == 0 else 1, device, world))
            algo.grad_sync = gs
            if rnd == 0:
                gs.broadcast_parameters(model)
            t_gs.append(timed(step, steps, 1, device, world))
        ms_no_ar, ms_local_bn = sum(t_no) / len(t_no), sum(t_gs) / len(t_gs)
        gs.broadcast_parameters(model)
        flat = torch.zeros(sum(p.numel() for p in model.parameters()), device=device)
        allreduce_ms = timed(lambda: dist.all_reduce(flat), 10, 3, device, world)   # the 213.7 MB flat gradient, alone
        del flat
        model.convert_sync_batchnorm(BatchNormSync())
        ms_nccl_bn = timed(step, steps, warm, device, world)
        bn_kind = "syncbn-nccl"
        try:
            peer = PeerBatchNormSync()
            ok = torch.ones(1, device=device)
        except Exception as e:  # noqa: BLE001  (no IPC peer access on this box: keep NCCL)
            peer, ok = None, torch.zeros(1, device=device)
            if rank == 0:
                print(f"PeerBatchNormSync unavailable ({e}); SyncBatchNorm stays on NCCL", file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        peer_everywhere = float(ok.item()) > 0
        if not peer_everywhere and peer is not None:
            peer.close(collective=False)   # some other rank could not map the buffers: nobody will use them
            peer = None
        if peer_everywhere:
            model.convert_sync_batchnorm(peer)
            bn_kind = "syncbn-peer"
    ms = timed(step, steps, warm, device, world)
    model.convert_sync_batchnorm(None)
    if world > 1 and peer is not None:
        torch.cuda.synchronize(device)
        peer.close()   # collective: unmap everywhere, barrier, then free
    # batch-independent cost of a step (launch latency, small GEMMs, optimiser, re-pack): the same step on 24 boards
    obs_s, mb_s = _update_batch(24, 99 + rank, device)
    gs, algo.grad_sync = algo.grad_sync, None

    def small_step():
        algo._step_fused(km, obs_s, mb_s, None)
        algo._optimizer_tail()
    fixed_ms = timed(small_step, 5, 3, device, world)
    algo.grad_sync = gs
    # GAE over the reference-shaped buffer T=128 x N=64 (+ normalisation)
    from keisei_b200 import gae as G
    T, N = 128, 64
    r, v = torch.randn(T, N, device=device), 0.3 * torch.randn(T, N, device=device)
    term = torch.rand(T, N, device=device) < 0.02
    nv = torch.randn(N, device=device)

    def gae_step():
        adv = G.compute_gae_gpu(r, v, term, nv, 0.99, 0.95).reshape(-1)
        G.normalize_advantages_(adv)
    gae_ms = timed(gae_step, 20, 3, device, world)
    sps = UPDATE_GLOBAL_B / (ms * 1e-3)
    frac = (22.97e9 * UPDATE_GLOBAL_B / world / (ms * 1e-3) / 1e12) / peaks()["bf16_tflops_sustained"]
    compact = {"value": round(sps, 1), "unit": "samples/s", "ms_per_step": round(ms, 3), "batch": UPDATE_GLOBAL_B,
               "frac": round(frac, 4), "bn": bn_kind,
               "fixed_ms": round(fixed_ms, 3), "gae_ms": round(gae_ms, 4)}
    if world > 1:
        compact.update({"allreduce_ms": round(allreduce_ms, 3), "allreduce_exposed_ms": round(ms_local_bn - ms_no_ar, 3),
                        "syncbn_ms": round(ms - ms_local_bn, 3), "ms_local_bn": round(ms_local_bn, 3), "ms_nccl_bn": None if ms_nccl_bn is None else round(ms_nccl_bn, 3)})
    if world == 1:
        e2e = update_e2e(algo, device, rank, world)
        e2e_dev = update_e2e(algo, device, rank, world, device_buffer=True)
        ref_sched = update_e2e(algo, device, rank, world, device_buffer=True, batch_size=256, epochs=4)
        detail.update({"e2e": e2e, "e2e_device_buffer": e2e_dev, "reference_schedule_256x32x4": ref_sched})
        compact.update({"e2e": round(e2e["value"], 1), "e2e_devbuf": round(e2e_dev["value"], 1),
                        "ref_sched_256x32x4": round(ref_sched["value"], 1)})
    detail.update({"includes": "fwd+losses+bwd+allreduce+unscale+clip+Adam", "batchnorm": bn_kind, "ms_no_allreduce": ms_no_ar,
                   "ms_local_bn": ms_local_bn, "ms_nccl_syncbn": ms_nccl_bn, "allreduce_alone_ms": allreduce_ms, "fixed_ms_b24": fixed_ms})
    return compact, detail


def bench_resnet_update(args, device, rank, world) -> dict:
    """BASELINE.json configs[3]: ResNet baseline (hidden 256 x 40 layers, 50-ch obs, scalar value) standard-PPO
    update, global batch 4096 split over the ranks, bf16: forward (batch-stat BN) + fused policy loss + MSE value
    loss + backward + gradient all-reduce + unscale/clip/Adam. Loss composition unpinned (see keisei_b200/ppo.py)."""
    from keisei_b200.algorithm_registry import PPOParams
    from keisei_b200.distributed import GradSync
    from keisei_b200.models import ResNetModel, ResNetParams
    from keisei_b200.ppo import PPOAlgorithm
    GB = 4096
    Bu = GB // world
    torch.manual_seed(1)
    model = ResNetModel(ResNetParams(hidden_size=256, num_layers=40)).to(device)
    algo = PPOAlgorithm(PPOParams(batch_size=Bu), model, use_amp=True)
    g = torch.Generator().manual_seed(17 + rank)
    obs = torch.randn(Bu, 50, 9, 9, generator=g).to(device)
    mask = torch.zeros(Bu, A, dtype=torch.bool)
    actions = torch.randint(0, A, (Bu,), generator=g)
    mask.scatter_(1, torch.randint(0, A, (Bu, 80), generator=g), True)
    mask[torch.arange(Bu), actions] = True
    mb = (mask.to(device), actions.to(device), (-3 * torch.rand(Bu, generator=g)).to(device), torch.randn(Bu, generator=g).to(device),
          None, None, torch.randn(Bu, generator=g).clamp(-1, 1).to(device))
    if world > 1:
        algo.grad_sync = GradSync()
        algo.grad_sync.broadcast_parameters(model)
    model.train()
    km = algo._kernel_model(device)

    def step():
        algo._step_fused(km, obs, mb, None)
        algo._optimizer_tail()

    steps, warm = max(2, min(args.steps, 5)), 3
    ms = timed(step, steps, warm, device, world)
    # rollout on the same model: select_actions at 4096 boards per GPU
    r_obs, r_mask = synth_boards(GB, 300 + rank, device)
    ms_roll = timed(lambda: algo.select_actions(r_obs, r_mask), steps, warm, device, world)
    flop_fwd = 18.66e6 + 80 * 95.55e6
    out = {"metric": "standard-PPO update samples/s (ResNet 40x256, batch 4096)", "value": GB / (ms * 1e-3), "unit": "samples/s",
           "ms_per_step": ms, "steps": steps, "warmup": warm, "global_batch": GB, "per_gpu_batch": Bu, "scaling": "strong",
           "frac_of_tensor_roofline": ((3 * flop_fwd - 18.66e6) * Bu / (ms * 1e-3) / 1e12) / peaks()["bf16_tflops_sustained"],
           "rollout_positions_per_s": world * GB / (ms_roll * 1e-3), "rollout_ms_per_step": ms_roll,
           "includes": "fwd+losses+bwd+allreduce+unscale+clip+Adam", "parity": "model pinned (tests/golden/resnet_tiny.npz); loss composition unpinned"}
    del model, algo
    torch.cuda.empty_cache()
    return out


def bench_league(algo, device, rank, world) -> dict:
    """BASELINE.json configs[4]: league-scale rollout, 512 envs per GPU (the `num_games` cap, config.py:574), 128
    consecutive select_actions steps timed end to end incl. launch overhead; and the split variant 256 learner +
    4 x 64 opponent sub-batches per step (katago_loop.py:284-431). No communication between ranks."""
    Bl = 512
    obs, mask = synth_boards(Bl, 500 + rank, device)
    def run128():
        for _ in range(128):
            algo.select_actions(obs, mask)
    ms = timed(run128, 2, 3, device, world)
    parts = [(0, 256)] + [(256 + 64 * k, 320 + 64 * k) for k in range(4)]
    def split128():
        for _ in range(128):
            for lo, hi in parts:
                algo.select_actions(obs[lo:hi], mask[lo:hi])
    ms_split = timed(split128, 2, 3, device, world)
    subs = [(obs[lo:hi], mask[lo:hi]) for lo, hi in parts]
    def grouped128():   # the same five sub-batches as parallel branches of one replayed graph
        for _ in range(128):
            algo.select_actions_many(subs)
    ms_grouped = timed(grouped128, 2, 3, device, world)
    # the whole split-merge environment step through the reference's own entry point (katago_loop.py:284-431, swapped in
    # by dropin.install_into_reference): host-side partition, gather, grouped forwards, sampling, merged scatter, and the
    # actions read back on the host as the loop does for vecenv.step(); fixed 256 + 4 x 64 partition, then a fresh random
    # partition every step (sub-batch sizes vary -> bucketed graphs)
    import numpy as np
    from keisei_b200.split_merge import split_merge_step
    base = algo.forward_model
    opponents = {k: base for k in range(4)}
    players_fixed = np.concatenate([np.zeros(256, np.uint8), np.ones(256, np.uint8)])
    ids_fixed = np.concatenate([np.zeros(256, np.int64), np.repeat(np.arange(4), 64)])
    rng = np.random.default_rng(5 + rank)
    randoms = [(rng.integers(0, 2, Bl).astype(np.uint8), rng.integers(0, 4, Bl).astype(np.int64)) for _ in range(16)]
    def sm128(parts_of):
        def run():
            for i in range(128):
                players, ids = parts_of(i)
                r = split_merge_step(obs, mask, players, base, opponent_models=opponents, env_opponent_ids=ids, learner_side=0)
                r.actions.cpu()
        return run
    ms_sm = timed(sm128(lambda i: (players_fixed, ids_fixed)), 2, 3, device, world)
    ms_sm_rand = timed(sm128(lambda i: randoms[i % 16]), 2, 3, device, world)
    sm = {"what": "split_merge_step (partition + gather + grouped forwards + sampling + scatter + actions D2H) per 512-env step",
          "fixed_256_4x64": world * Bl * 128 / (ms_sm * 1e-3), "random_partition_each_step": world * Bl * 128 / (ms_sm_rand * 1e-3),
          "unit": "positions/s", "ms_per_128_steps": [ms_sm, ms_sm_rand]}
    return {"split_merge_step": sm, "metric": "league rollout positions/s (512 envs per GPU, 128 consecutive steps)", "value": world * Bl * 128 / (ms * 1e-3),
            "unit": "positions/s", "ms_per_128_steps": ms, "envs_per_gpu": Bl, "scaling": "weak",
            "split_variant": {"value": world * Bl * 128 / (ms_split * 1e-3), "unit": "positions/s", "ms_per_128_steps": ms_split,
                              "sub_batches": "256 learner + 4 x 64 opponents (same weights: synthetic)",
                              "grouped_value": world * Bl * 128 / (ms_grouped * 1e-3), "grouped_ms_per_128_steps": ms_grouped,
                              "grouped": "select_actions_many: the five sub-batches as parallel branches of one CUDA graph"}}


def update_e2e(algo, device, rank, world, device_buffer: bool = False, batch_size: int | None = None, epochs: int = 1) -> dict:
    """The public call a user makes: KataGoPPOAlgorithm.update(buffer, next_values) on a host-resident
    KataGoRolloutBuffer of T=128 x N=64 = 8192 samples (the reference's profiled update shape,
    scripts/profile_hotpath.py:411-455), one epoch, one minibatch of 8192: buffer flatten, H2D of
    observations + masks (225 MB), GAE + normalisation, shuffle/gather, fwd+bwd+clip+Adam, metrics D2H."""
    import dataclasses
    from keisei_b200.katago_ppo import KataGoRolloutBuffer
    T, N = 128, 64
    g = torch.Generator().manual_seed(11 + rank)
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A, device=device if device_buffer else None)
    steps = []
    for t in range(T):
        obs = torch.randn(N, 50, 9, 9, generator=g)
        mask = torch.zeros(N, A, dtype=torch.bool)
        actions = torch.randint(0, A, (N,), generator=g)
        mask.scatter_(1, torch.randint(0, A, (N, 80), generator=g), True)
        mask[torch.arange(N), actions] = True
        term = torch.rand(N, generator=g) < 0.02
        steps.append((obs, actions, -3 * torch.rand(N, generator=g), 0.3 * torch.randn(N, generator=g), term.float(), term, term, mask,
                      torch.where(term, torch.randint(0, 3, (N,), generator=g), torch.full((N,), -1)), torch.randn(N, generator=g).clamp(-1.5, 1.5)))
    old_params = algo.params
    algo.params = dataclasses.replace(old_params, batch_size=batch_size or T * N, epochs_per_batch=epochs)
    nv = torch.randn(N, generator=g).to(device)
    times = []
    for rep in range(3 if epochs == 1 else 2):
        for st in steps:
            buf.add(*st)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        algo.update(buf, nv)
        torch.cuda.synchronize(device)
        times.append(time.perf_counter() - t0)
    algo.params = old_params
    best = min(times[1:])
    if epochs != 1 or batch_size is not None:
        n_steps = epochs * ((T * N + (batch_size or T * N) - 1) // (batch_size or T * N))
        return {"value": epochs * T * N / best, "unit": "samples/s", "ms_per_update": best * 1e3, "optimizer_steps": n_steps,
                "ms_per_minibatch": best * 1e3 / n_steps, "note": f"update() with batch_size={batch_size}, {epochs} epochs over 8192 samples "
                "(the reference's profiled schedule, profile_hotpath.py:411-455), device buffer, wall clock"}
    if device_buffer:
        return {"value": T * N / best, "unit": "samples/s", "ms_per_update": best * 1e3, "samples": T * N,
                "h2d_bytes_per_update": 0, "d2h_bytes_per_update": 9 * 8,
                "note": "update(buffer, next_values) with KataGoRolloutBuffer(device=cuda): the rollout steps were stored in "
                        "HBM as they were produced, so the update ships nothing; wall clock, best of 2 after 1 warm-up"}
    return {"value": T * N / best, "unit": "samples/s", "ms_per_update": best * 1e3, "samples": T * N,
            "h2d_bytes_per_update": T * N * (50 * 81 * 4 + A + 8 + 4 * 4 + 8), "d2h_bytes_per_update": 9 * 8,
            "note": "update(buffer, next_values): host buffer -> metrics dict, wall clock, best of 2 after 1 warm-up"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-update", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also run configs[3] (ResNet PPO), configs[4] (league rollout) and the "
                    "informational torch-eager port on the GPU; results go to the detail record")
    ap.add_argument("--no-extra", action="store_true", help="(default; kept for compatibility)")
    ap.add_argument("--detail-file", default=None, help="also write the full detail record (JSON) here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback on the product path)")
        run_ours(args)


if __name__ == "__main__":
    main()
