"""bench.py's league leg alone (512 envs per GPU; sequential sub-batches vs grouped graph branches)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams  # noqa: E402
from keisei_b200.models import SEResNetModel, SEResNetParams  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), model)
print(json.dumps(bench.bench_league(algo, dev, 0, 1)))
