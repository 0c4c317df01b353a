"""Time the fused block-tail kernel (se_apply.cu) alone with CUDA events: B boards of 256 channels, inputs
alternated between two buffer sets (each set > L2). Prints GB/s of algorithmic traffic (3 tensors)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from keisei_b200 import model_ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
C, S = 256, 16
TRAIN = len(sys.argv) > 3 and sys.argv[3] == 'train'
g = torch.Generator().manual_seed(0)
sets = []
for _ in range(2):
    z = torch.randn(B, 81, C, generator=g).bfloat16().to(dev)
    res = torch.randn(B, 81, C, generator=g).clamp_min(0).bfloat16().to(dev)
    sets.append((z, res, z.float().mean(dim=1)))
w1 = (torch.randn(S, C, generator=g) / 16).to(dev); b1 = torch.zeros(S, device=dev)
w2 = (torch.randn(2 * C, S, generator=g) / 4).to(dev); b2 = torch.zeros(2 * C, device=dev)
for i in range(3):
    model_ops.se_block_tail(*sets[i & 1], w1, b1, w2, b2, want_ties=TRAIN, se_raw=TRAIN)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    model_ops.se_block_tail(*sets[i & 1], w1, b1, w2, b2, want_ties=TRAIN, se_raw=TRAIN)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
gb = 3 * B * 81 * C * 2 / 1e9
print(f"se_block_tail B={B}: {ms * 1e3:.1f} us/launch, {gb / (ms * 1e-3):.0f} GB/s algorithmic (3 x {gb / 3 * 1e3:.0f} MB)")
