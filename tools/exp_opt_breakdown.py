"""Wall / GPU time of the optimiser tail of an update step (unscale_, clip_grad_norm_, Adam step, scaler.update)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams  # noqa: E402
from keisei_b200.models import SEResNetModel, SEResNetParams  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True, batch_size=64), model)
flat = torch.randn(sum(p.numel() for p in model.parameters()), device=dev) * 1e-3
off = 0
for p in model.parameters():
    p.grad = flat[off:off + p.numel()].view(p.shape)
    off += p.numel()
opt, sc = algo.optimizer, algo.scaler


def timeit(name, fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:34s} cpu {1e3 * (t1 - t0) / n:7.3f} ms   cpu+gpu {1e3 * (t2 - t0) / n:7.3f} ms")


def full():
    loss = torch.zeros((), device=dev, requires_grad=True)
    sc.scale(loss)           # marks the scaler as used for this step
    sc.unscale_(opt)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    sc.step(opt)
    sc.update()


timeit("unscale+clip+step+update", full)
timeit("clip_grad_norm_(params)", lambda: torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0))
timeit("flat norm + clamp + mul", lambda: flat.mul_(torch.clamp(1.0 / (torch.linalg.vector_norm(flat) + 1e-6), max=1.0)))
timeit("optimizer.step (fused Adam)", lambda: opt.step())
inv, found = torch.ones((), device=dev), torch.zeros((), device=dev)
timeit("foreach unscale (576 tensors)", lambda: torch._amp_foreach_non_finite_check_and_unscale_([p.grad for p in model.parameters()], found, inv))
timeit("foreach unscale (flat)", lambda: torch._amp_foreach_non_finite_check_and_unscale_([flat], found, inv))
