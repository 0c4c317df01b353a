"""Grouped multi-model rollout: one sub-batch alone on its SM share vs all branches in one graph (40x256, bf16)."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from keisei_b200 import model_ops
from keisei_b200.models import SEResNetModel, SEResNetParams
from keisei_b200.models.se_resnet import rollout_forward_many

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SEResNetModel(SEResNetParams(**bench.MODEL_CFG)).to(dev).eval()
tables = model._ptr_tables(); dtype = torch.bfloat16
wpack = model._packed(tables.params, tables.buffers, dtype)
res = {}
def graph_time(fn, reps=20):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(dev); s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            keep = fn()
    torch.cuda.current_stream(dev).wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 3)
for B, sms in ((64, 18), (64, 148), (256, 74), (256, 148), (512, 148)):
    obs = torch.randn(B, 50, 9, 9, device=dev)
    res[f"B{B}_sms{sms}"] = graph_time(lambda: model_ops.seresnet_forward_raw(obs, tables, wpack, False, 1, True, num_sms=sms))
def branches(specs):
    """specs: list of (B, num_sms); one graph, one branch per spec."""
    obs = [torch.randn(b, 50, 9, 9, device=dev) for b, _ in specs]
    def fn():
        cur = torch.cuda.current_stream(dev)
        side = [torch.cuda.Stream(dev) for _ in specs[1:]]
        for st in side:
            st.wait_stream(cur)
        keep = [model_ops.seresnet_forward_raw(obs[0], tables, wpack, False, 1, True, num_sms=specs[0][1])]
        for st, o, (_, sms) in zip(side, obs[1:], specs[1:]):
            with torch.cuda.stream(st):
                keep.append(model_ops.seresnet_forward_raw(o, tables, wpack, False, 1, True, num_sms=sms))
        for st in side:
            cur.wait_stream(st)
        return keep
    return graph_time(fn)
res["2x256_sms74"] = branches([(256, 74), (256, 74)])
res["2x256_sms148"] = branches([(256, 148), (256, 148)])
res["256+4x64_shares"] = branches([(256, 74), (64, 20), (64, 18), (64, 18), (64, 18)])
res["256+4x64_sms148"] = branches([(256, 148)] + [(64, 148)] * 4)
res["8x64_sms18"] = branches([(64, 18)] * 8)
print(json.dumps(res))
