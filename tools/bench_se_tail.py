"""Isolated fused-tail conv2 (CTA-pair kernel) vs the plain eval conv2 epilogue, CUDA events."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from keisei_b200 import model_ops

dev = torch.device("cuda:0")
C, S = 256, 16
torch.manual_seed(0)
wf = model_ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / 48, torch.bfloat16)
sc, sh = torch.ones(C, device=dev), torch.zeros(C, device=dev)
w1, b1 = torch.randn(S, C, device=dev) / 16, torch.zeros(S, device=dev)
w2, b2 = torch.randn(2 * C, S, device=dev) / 4, torch.zeros(2 * C, device=dev)


def timeit(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000


for B in (2048, 4096):
    xs = [torch.randn(B, 81, C, device=dev).bfloat16() for _ in range(2)]
    rs = [torch.randn(B, 81, C, device=dev).bfloat16() for _ in range(2)]
    t_plain = timeit(lambda i: model_ops.conv3x3(xs[i & 1], wf, backend=3, scale=sc, shift=sh, want_board_mean=True))
    t_fused = timeit(lambda i: model_ops.conv3x3_se_tail(xs[i & 1], wf, sc, sh, rs[i & 1], w1, b1, w2, b2))
    print(f"B={B}: plain conv2 (pair, affine+board mean) {t_plain:7.1f} us   fused tail {t_fused:7.1f} us", flush=True)
