"""Where a small-batch select_actions step spends its time: host time of each phase (no syncs in between), then the same
phases with a device sync after each (host + device)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from keisei_b200 import policy_ops
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), model)
obs, mask = bench.synth_boards(B, 1, dev)
for _ in range(4):
    algo.select_actions(obs, mask)
torch.cuda.synchronize()
def run(sync):
    acc = {}
    def mark(name, t0):
        if sync: torch.cuda.synchronize()
        t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return time.perf_counter()
    reps = 20
    for _ in range(reps):
        t = time.perf_counter()
        model.eval(); t = mark("eval()", t)
        with torch.no_grad():
            out = model.rollout_forward(obs); t = mark("rollout_forward", t)
            flat = out.policy_logits.reshape(B, -1)
            a, lp, v, legal, flags = policy_ops.policy_sample(flat, mask, out.value_logits, out.score_lead, 0.0, seed=1); t = mark("policy_sample", t)
            bad = int(flags[0].item()); t = mark("flags.item()", t)
        model.train(); t = mark("train()", t)
    return {k: round(v / reps * 1e3, 3) for k, v in acc.items()}
print(json.dumps({"B": B, "host_only_ms": run(False), "with_sync_ms": run(True)}))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): algo.select_actions(obs, mask)
torch.cuda.synchronize(); print("select_actions ms/step", (time.perf_counter() - t0) / 20 * 1e3)
