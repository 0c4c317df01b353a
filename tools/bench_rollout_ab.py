"""A/B of the rollout step (40x256, bf16, 4096 boards): env switches are read once per process, so each variant runs in
its own process: python tools/bench_rollout_ab.py [label]"""
import os, sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench

dev = torch.device("cuda:0")
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams
torch.manual_seed(0)
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), model)
out = {"label": sys.argv[1] if len(sys.argv) > 1 else "", "env": {k: v for k, v in os.environ.items() if k.startswith("KB_")}}
for B in (4096, 512):
    obs, mask = bench.synth_boards(B, 100, dev)
    ms = bench.timed(lambda: algo.select_actions(obs, mask), 10, 4, dev, 1)
    out[f"B{B}"] = {"ms": round(ms, 3), "pos_per_s": round(B / ms * 1e3, 1)}
print(json.dumps(out), flush=True)
