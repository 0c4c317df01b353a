"""One rollout step or one update step bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off` launch lists / full captures (profiles/README.md)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams  # noqa: E402
from keisei_b200.models import SEResNetModel, SEResNetParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="rollout", choices=["rollout", "update"])
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--blocks", type=int, default=40)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = dict(bench.MODEL_CFG); cfg["num_blocks"] = args.blocks
model = SEResNetModel(SEResNetParams(**cfg)).to(dev)
B = args.batch or (4096 if args.mode == "rollout" else 8192)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True, batch_size=B), model)
obs, mask = bench.synth_boards(B, 1, dev)
if args.mode == "rollout":
    step = lambda: algo.select_actions(obs, mask)
else:
    g = torch.Generator().manual_seed(2)
    actions = torch.randint(0, bench.A, (B,), generator=g).to(dev)
    mask[torch.arange(B, device=dev), actions] = True
    mb = (mask, actions, (-3 * torch.rand(B, generator=g)).to(dev), torch.randn(B, generator=g).to(dev),
          torch.randint(-1, 3, (B,), generator=g).to(dev), torch.randn(B, generator=g).clamp(-1.5, 1.5).to(dev))
    km = algo._kernel_model(dev)
    def step():
        algo._step_fused(km, obs, mb, None)
        algo._optimizer_tail()
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", args.mode, B)
