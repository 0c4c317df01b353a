"""A few launches of the fused-tail conv2 (CTA-pair kernel) for `ncu --set full -k regex:conv3x3_tc2 -s 3 -c 1`."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from keisei_b200 import model_ops

dev = torch.device("cuda:0")
C, S, B = 256, 16, 4096
torch.manual_seed(0)
wf = model_ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / 48, torch.bfloat16)
sc, sh = torch.ones(C, device=dev), torch.zeros(C, device=dev)
w1, b1 = torch.randn(S, C, device=dev) / 16, torch.zeros(S, device=dev)
w2, b2 = torch.randn(2 * C, S, device=dev) / 4, torch.zeros(2 * C, device=dev)
xs = [torch.randn(B, 81, C, device=dev).bfloat16() for _ in range(2)]
rs = [torch.randn(B, 81, C, device=dev).bfloat16() for _ in range(2)]
for i in range(5):
    model_ops.conv3x3_se_tail(xs[i & 1], wf, sc, sh, rs[i & 1], w1, b1, w2, b2)
torch.cuda.synchronize()
print("done")
