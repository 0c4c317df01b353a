"""Weight-gradient convolution alone (tcgen05 kernel + the slice reduction), CUDA events over alternating inputs > L2.
usage: python tools/bench_wgrad.py [label] [B]"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from keisei_b200 import model_ops

dev = torch.device("cuda:0")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
g = torch.Generator(device=dev).manual_seed(0)
sets = [(torch.randn(B, 81, 256, device=dev, generator=g).bfloat16(), torch.randn(B, 81, 256, device=dev, generator=g).bfloat16())
        for _ in range(2)]
def run(i):
    return model_ops.conv3x3_wgrad(sets[i & 1][0], sets[i & 1][1], backend=1)
for i in range(4):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 40
e0.record()
for i in range(reps):
    run(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 2.0 * B * 81 * 256 * 256 * 9
print(json.dumps({"label": sys.argv[1] if len(sys.argv) > 1 else "", "B": B, "us": round(ms * 1e3, 1), "tflops": round(flops / ms / 1e9, 1)}))
