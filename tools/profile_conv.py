"""A few launches of the CTA-pair conv (eval epilogue: folded BN + ReLU + gpool bias) at B=4096 for
`ncu --set full -k regex:conv3x3_tc2 -s 3 -c 1` (the roofline kernel of bench.py)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from keisei_b200 import model_ops

dev = torch.device("cuda:0")
C, B = 256, 4096
torch.manual_seed(0)
wf = model_ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / 48, torch.bfloat16)
sc, sh = torch.ones(C, device=dev), torch.zeros(C, device=dev)
gb = torch.zeros(B, C, device=dev)
xs = [torch.randn(B, 81, C, device=dev).bfloat16() for _ in range(2)]
for i in range(5):
    model_ops.conv3x3(xs[i & 1], wf, backend=3, scale=sc, shift=sh, relu=True, gbias=gb)
torch.cuda.synchronize()
print("done")
