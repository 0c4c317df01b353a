"""Host-side phases of the end-to-end rollout step (bench.RolloutIngest.step) at 4096 boards."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), model)
obs, mask = bench.synth_boards(B, 1, dev)
ri = bench.RolloutIngest(algo, obs.cpu(), mask.cpu(), dev)
for _ in range(4): ri.step()
acc = {}
def mark(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
reps = 10
torch.cuda.synchronize(); w0 = time.perf_counter()
for _ in range(reps):
    t = time.perf_counter()
    cur = ri.slot; ri.k += 1
    ri.slot = ri.ingest.submit(*ri.src[ri.k & 1]); t = mark("submit", t)
    d_obs, d_mask = ri.ingest.get(cur); t = mark("get", t)
    a, lp, v = algo.select_actions(d_obs, d_mask); t = mark("select_actions", t)
    ri.ingest.release(cur)
    ri.h_out[0].copy_(a, non_blocking=True); ri.h_out[1].copy_(lp, non_blocking=True); ri.h_out[2].copy_(v, non_blocking=True); t = mark("d2h_enqueue", t)
    torch.cuda.current_stream(dev).synchronize(); t = mark("sync", t)
wall = (time.perf_counter() - w0) / reps * 1e3
ms = bench.timed(lambda: algo.select_actions(obs, mask), 10, 3, dev, 1)
print(json.dumps({"B": B, "e2e_ms": round(wall, 3), "device_resident_ms": round(ms, 3), "phases_ms": {k: round(v / reps * 1e3, 3) for k, v in acc.items()}}))
