"""Isolated 256->256 convolution: single-CTA vs CTA-pair (cta_group::2) kernel, CUDA events, alternating inputs > L2."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from keisei_b200 import model_ops

dev = torch.device("cuda:0")
C = 256
torch.manual_seed(0)
wf = model_ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / 48, torch.bfloat16)
sc, sh = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for B in (512, 1024, 2048, 4096, 8192):
    xs = [torch.randn(B, 81, C, device=dev).bfloat16() for _ in range(2 if B >= 2048 else 8)]
    gb = torch.zeros(B, C, device=dev)
    for backend, name in ((2, "single"), (3, "pair")):
        for kw, tag in ((dict(scale=sc, shift=sh, relu=True, gbias=gb), "eval1"), (dict(want_sums=True), "train")):
            for i in range(3):
                model_ops.conv3x3(xs[i % len(xs)], wf, backend=backend, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for i in range(reps):
                model_ops.conv3x3(xs[i % len(xs)], wf, backend=backend, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"B={B:5d} {name:6s} {tag:5s} {ms*1000:8.1f} us  {2*81*256*2304*B/ms/1e9:8.1f} TFLOP/s", flush=True)
