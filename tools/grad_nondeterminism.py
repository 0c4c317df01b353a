"""Run-to-run nondeterminism of the flat gradient of one fp32 step (same model, same batch, six runs): last bits only
(3.5e-8 relative L2 — float atomics in the SIMT weight gradient and the BatchNorm sums), with packed and byte masks."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.model_registry import build_model
from keisei_b200 import policy_ops
DEV = "cuda:0"
torch.manual_seed(0)
cfg = dict(num_blocks=2, channels=64, se_reduction=8, global_pool_channels=16, policy_channels=8, value_fc_size=16, score_fc_size=16)
a = build_model("se_resnet", dict(cfg)).to(DEV)
ta = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=12, epochs_per_batch=1, use_amp=False), a)
A = 11259
g = torch.Generator().manual_seed(3)
B = 12
obs = torch.randn(B, 50, 9, 9, generator=g).to(DEV)
mask = (torch.rand(B, A, generator=g) < 0.01)
acts = torch.randint(0, A, (B,), generator=g); mask[torch.arange(B), acts] = True
mask, acts = mask.to(DEV), acts.to(DEV)
old = (-2 * torch.rand(B, generator=g)).to(DEV); adv = torch.randn(B, generator=g).to(DEV)
cats = torch.randint(0, 3, (B,), generator=g).to(DEV); sc = torch.randn(B, generator=g).clamp(-1, 1).to(DEV)
km = ta._kernel_model(torch.device(DEV))
a.train()
for variant in ("bits", "bytes"):
    grads = []
    for rep in range(6):
        mk = policy_ops.pack_mask_bits(mask) if variant == "bits" else mask
        if variant == "bytes":
            # bypass the trainer's packing: call the loss op with byte masks directly
            orig = policy_ops.pack_mask_bits
            policy_ops.pack_mask_bits = lambda m: m
        ta._step_fused(km, obs, (mk, acts, old, adv, cats, sc, adv), None)
        if variant == "bytes":
            policy_ops.pack_mask_bits = orig
        torch.cuda.synchronize()
        grads.append(ta._flat_grad.clone())
    ref = grads[0]
    for r in grads[1:]:
        d = (r - ref).abs()
        print(variant, "max abs diff %.3e  max|g| %.3e  rel L2 %.3e  n_diff %d" % (float(d.max()), float(ref.abs().max()), float(d.norm() / ref.norm()), int((d > 0).sum())))
