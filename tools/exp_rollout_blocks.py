"""Rollout step time against the number of blocks (4096 boards, bf16): the slope is the live cost of one block, to compare
with the sum of its kernels timed alone (conv1 0.264 + fused conv2 0.307 + global_fc 0.013 ms)."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
obs, mask = bench.synth_boards(B, 100, dev)
res = {}
for nb in (40, 20, 10, 0 + 1, 40):
    torch.manual_seed(0)
    cfg = dict(bench.MODEL_CFG); cfg["num_blocks"] = nb
    model = SEResNetModel(SEResNetParams(**cfg)).to(dev)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), model)
    ms = bench.timed(lambda: algo.select_actions(obs, mask), 10, 4, dev, 1)
    res.setdefault(nb, []).append(round(ms, 3))
    del model, algo
print(json.dumps({"B": B, "ms_by_blocks": res}))
