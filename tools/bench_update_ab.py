"""A/B of the update step (40x256, bf16): env switches are read once per process -> one process per variant.
usage: python tools/bench_update_ab.py [label] [batch]"""
import os, sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench

dev = torch.device("cuda:0")
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams
torch.manual_seed(0)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True, batch_size=B), model)
obs, mb = bench._update_batch(B, 7, dev)
model.train()
km = algo._kernel_model(dev)
def step():
    algo._step_fused(km, obs, mb, None)
    algo._optimizer_tail()
ms = bench.timed(step, 6, 3, dev, 1)
print(json.dumps({"label": sys.argv[1] if len(sys.argv) > 1 else "", "batch": B, "ms": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1)}), flush=True)
