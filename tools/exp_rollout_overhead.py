import sys, time, torch
sys.path.insert(0, '.')
import bench
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams
dev = torch.device('cuda:0')
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), model)
obs, mask = bench.synth_boards(4096, 1, dev)
for strict in (True, False):
    algo.strict_guards = strict
    for _ in range(3): algo.select_actions(obs, mask)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cpu = 0.0
    for _ in range(10):
        c0 = time.perf_counter(); algo.select_actions(obs, mask); cpu += time.perf_counter() - c0
    e1.record(); torch.cuda.synchronize()
    print('strict', strict, 'wall ms/step', (time.perf_counter() - t0) * 100, 'event ms/step', e0.elapsed_time(e1) / 10, 'cpu call ms', cpu * 100)
# pure forward only
model.eval()
with torch.no_grad():
    for _ in range(3): model(obs)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): model(obs)
    e1.record(); torch.cuda.synchronize()
print('forward only ms', e0.elapsed_time(e1) / 10)
