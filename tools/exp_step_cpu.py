"""How much of an update step is host enqueue time? (per batch size: enqueue-only wall time vs synchronised wall time)"""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
for B in (256, 1024, 4096):
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True, batch_size=B), model)
    obs, mb = bench._update_batch(B, 7, dev)
    model.train()
    km = algo._kernel_model(dev)
    def step():
        algo._step_fused(km, obs, mb, None)
        algo._optimizer_tail()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    enq, tot = [], []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
    print(json.dumps({"B": B, "enqueue_ms": round(min(enq), 2), "total_ms": round(min(tot), 2)}), flush=True)
    algo.strict_guards = False
    for _ in range(2):
        step()
    enq, tot = [], []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
    print(json.dumps({"B": B, "strict_guards": False, "enqueue_ms": round(min(enq), 2), "total_ms": round(min(tot), 2)}), flush=True)
