"""Isolated timings of the rollout sampling pieces (CUDA events)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from keisei_b200 import policy_ops
dev = torch.device("cuda:0")
A = 11259
def timeit(fn, reps=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
for B in (512, 4096):
    g = torch.Generator(device=dev).manual_seed(1)
    sets = []
    for _ in range(2):
        lg = torch.randn(B, 11264, device=dev, generator=g).bfloat16()[:, :A]
        mk = torch.zeros(B, A, dtype=torch.bool, device=dev)
        mk.scatter_(1, torch.randint(0, A, (B, 80), device=dev, generator=g), True)
        sets.append((lg, mk, policy_ops.pack_mask_bits(mk)))
    vl = torch.randn(B, 3, device=dev)
    print(B, "pack        %.1f us" % timeit(lambda i: policy_ops.pack_mask_bits(sets[i & 1][1])))
    print(B, "sparse      %.1f us" % timeit(lambda i: policy_ops.policy_sample(sets[i & 1][0], sets[i & 1][2], vl)))
    print(B, "dense       %.1f us" % timeit(lambda i: policy_ops.policy_sample(sets[i & 1][0], sets[i & 1][1], vl, dense=True)))
    print(B, "pack+sparse %.1f us" % timeit(lambda i: policy_ops.policy_sample(sets[i & 1][0], sets[i & 1][1], vl)), flush=True)
