import sys, time, torch
sys.path.insert(0, '.')
import bench
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams
dev = torch.device('cuda:0')
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev)
algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), model)
model.eval()
for B in (4096, 512):
    obs, mask = bench.synth_boards(B, 1, dev)
    with torch.no_grad():
        for _ in range(3): model(obs)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter(); model(obs); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    print('B', B, 'cpu issue ms / total ms per forward:', [(round(a, 2), round(b, 2)) for a, b in ts])
# raw launch cost of a trivial torch op and of an empty-ish library call
x = torch.zeros(8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): x.add_(1)
t1 = time.perf_counter(); torch.cuda.synchronize()
print('torch add_ issue us/launch', (t1 - t0) / 200 * 1e6)
import os
print('cpu count', os.cpu_count(), open('/proc/cpuinfo').read().split('model name')[1].split('\n')[0])
