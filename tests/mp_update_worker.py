"""Worker of tests/test_multi_gpu.py: one rank of a real multi-process, multi-GPU KataGo-PPO step (torchrun env).

Each rank: same seed-0 model (ranks != 0 perturb their copy first so the parameter broadcast is exercised), NCCL
`GradSync`, SyncBatchNorm through `PeerBatchNormSync` (CUDA IPC over NVLink) or `BatchNormSync` (NCCL), its own
contiguous shard of a fixed global batch; `_step_fused` + `_optimizer_tail`. Asserts inside the run that every rank
ends with bit-identical parameters and BatchNorm running statistics; rank 0 stores the averaged gradient and the
updated parameters for the launcher to compare with a single-process full-batch step.

usage: torchrun ... mp_update_worker.py <fp32|bf16> <peer|nccl> <out.pt> <local_batch>
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))

CFG = dict(num_blocks=3, channels=128, se_reduction=8, global_pool_channels=32, policy_channels=16, value_fc_size=32,
           score_fc_size=32)
A = 11259


def global_batch(n, seed=31):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(n, 50, 9, 9, generator=g)
    mask = torch.rand(n, A, generator=g) < 0.01
    acts = torch.randint(0, A, (n,), generator=g)
    mask[torch.arange(n), acts] = True
    # every W/D/L target valid: the cross-entropy is a mean over valid rows, so equal valid counts per rank make the
    # average of the per-rank losses equal the full-batch loss (the reference's DDP has the same property)
    return (obs, mask, acts, -3 * torch.rand(n, generator=g), torch.randn(n, generator=g), torch.randint(0, 3, (n,), generator=g),
            torch.randn(n, generator=g).clamp(-1.5, 1.5))


def main() -> None:
    mode, kind, out_path, bl = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    from keisei_b200.distributed import BatchNormSync, GradSync, PeerBatchNormSync
    from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
    from keisei_b200.models import SEResNetModel, SEResNetParams

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = SEResNetModel(SEResNetParams(**CFG)).to(dev)
    if rank != 0:
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.05 * torch.randn_like(p))
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=(mode == "bf16"), batch_size=bl), model)
    algo.grad_sync = GradSync(bucket_bytes=1 << 20)     # small buckets: several overlapped all-reduces even on this small model
    algo.grad_sync.broadcast_parameters(model)
    sync = PeerBatchNormSync() if kind == "peer" else BatchNormSync()
    model.convert_sync_batchnorm(sync)
    batch = [t[rank * bl:(rank + 1) * bl].to(dev) for t in global_batch(world * bl)]
    obs, mask, acts, old, adv, cats, score_t = batch
    model.train()
    pl, vl, sl, ent, _ = algo._step_fused(model, obs, (mask, acts, old, adv, cats, score_t, adv), None)
    assert algo.grad_sync.last_overlap_buckets >= 3, algo.grad_sync.last_overlap_buckets   # bucketed, overlapped exchange ran
    scale = float(algo.scaler.get_scale()) if algo.scaler.is_enabled() else 1.0
    flat = (algo._flat_grad / scale / getattr(algo, "_grad_div", 1.0)).clone()   # overlapped buckets come back summed over the ranks
    algo._optimizer_tail()
    torch.cuda.synchronize(dev)
    # identical state on every rank, bit for bit
    for name, t in list(model.named_parameters()) + [(n, b) for n, b in model.named_buffers() if b.is_floating_point()]:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.detach().contiguous())
        for r in range(1, world):
            assert torch.equal(parts[0], parts[r]), f"{name}: rank {r} differs from rank 0 after update"
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    for r in range(1, world):
        assert torch.equal(parts[0], parts[r]), f"averaged gradient differs on rank {r}"
    losses = torch.stack([pl, vl, sl, ent]).float()
    dist.all_reduce(losses)
    if rank == 0:
        torch.save({"flat": flat.cpu(), "params": {n: p.detach().cpu() for n, p in model.named_parameters()},
                    "buffers": {n: b.detach().cpu() for n, b in model.named_buffers()}, "losses": (losses / world).cpu(),
                    "world": world}, out_path)
    model.convert_sync_batchnorm(None)
    if kind == "peer":
        sync.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
