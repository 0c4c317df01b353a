"""Gradient exchange variants of the data-parallel update step, interleaved in one process so that clock drift cancels:
none / overlapped buckets (default) / one all-reduce after the backward / other bucket sizes.
usage: torchrun --nproc-per-node N tools/exp_allreduce_overlap.py"""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
import bench
from keisei_b200.distributed import GradSync
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
Bu = bench.UPDATE_GLOBAL_B // world
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True, batch_size=Bu), model)
obs, mb = bench._update_batch(Bu, 7 + rank, dev)
model.train()
km = algo._kernel_model(dev)
variants = {"none": None, "overlap_48MB": GradSync(), "overlap_16MB": GradSync(bucket_bytes=16 << 20), "overlap_110MB": GradSync(bucket_bytes=110 << 20),
            "after_backward": GradSync(overlap=False)}
variants["overlap_48MB"].broadcast_parameters(model)
def step():
    algo._step_fused(km, obs, mb, None)
    algo._optimizer_tail()
acc = {k: [] for k in variants}
for rnd in range(4):
    for name, gs in variants.items():
        algo.grad_sync = gs
        acc[name].append(bench.timed(step, 4, 2 if rnd else 3, dev, world))
if rank == 0:
    print(json.dumps({"world": world, "per_gpu_batch": Bu, "ms": {k: [round(x, 2) for x in v] for k, v in acc.items()},
                      "mean_minus_none": {k: round(sum(v) / len(v) - sum(acc["none"]) / len(acc["none"]), 2) for k, v in acc.items()}}))
dist.barrier()
dist.destroy_process_group()
