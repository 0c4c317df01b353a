"""The path bench.py times — bf16 activations, tcgen05 forward / data-gradient / weight-gradient convolutions inside
the C schedule, the bf16 column kernels, `se_mlp_bwd`, the TF32 head / MLP backward GEMMs and the side-stream weight
gradients — checked against the CPU ORACLE (fp32 restatement of the reference, oracle/keisei_oracle.py), not against
this library's own fp32 twin.

Bars (north_star): bf16 logits / values / losses within 2e-2 relative of the reference on identical inputs and
weights. Parameter gradients after a bf16 backward are compared by direction (cosine) and norm ratio, as the reference's
own bf16 autocast differs from its fp32 gradients by rounding noise that grows towards the stem
(reference se_resnet.py:68-90; tests/test_resnet_gpu.py uses the same bars for the plain ResNet).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import keisei_oracle as O
from keisei_b200 import _lib, model_ops, policy_ops
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
DEV = "cuda:0"
A = 11259

CFGS = {
    "3x128": dict(num_blocks=3, channels=128, se_reduction=8, global_pool_channels=32, policy_channels=16,
                  value_fc_size=32, score_fc_size=32),
    "2x256": dict(num_blocks=2, channels=256),  # reference default heads: SE hidden 16, gpool 128, policy 32, value 256, score 128
}


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


def cos_ratio(got, want):
    g = np.asarray(got, np.float64).ravel(); w = np.asarray(want, np.float64).ravel()
    ng, nw = np.linalg.norm(g), np.linalg.norm(w)
    return float(g @ w / max(ng * nw, 1e-300)), float(ng / max(nw, 1e-300))


def nhwc(x):
    return x.permute(0, 2, 3, 1).reshape(x.shape[0], 81, x.shape[1]).contiguous()


def nchw(x):
    return x.reshape(x.shape[0], 9, 9, x.shape[2]).permute(0, 3, 1, 2).contiguous()


def make_batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(B, 50, 9, 9, generator=g)
    mask = torch.rand(B, A, generator=g) < 0.01
    acts = torch.randint(0, A, (B,), generator=g)
    mask[torch.arange(B), acts] = True
    old = -3 * torch.rand(B, generator=g)
    adv = torch.randn(B, generator=g)
    cats = torch.randint(-1, 3, (B,), generator=g)
    score_t = torch.randn(B, generator=g).clamp(-1.5, 1.5)
    return obs, mask, acts, old, adv, cats, score_t


def oracle_step(model, batch, num_blocks, dtype=torch.float32):
    """Training-mode forward + KataGo-PPO loss + backward of the oracle on the model's weights.
    Returns (policy, value, score, losses dict, {name: grad}). dtype=float64 gives the rounding-free yardstick."""
    obs, mask, acts, old, adv, cats, score_t = batch
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    if dtype != torch.float32:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
        obs, old, adv, score_t = obs.to(dtype), old.to(dtype), adv.to(dtype), score_t.to(dtype)
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    wp, wv, ws = O.seresnet_forward(sd, obs, num_blocks, training=True)
    want = O.ppo_losses(wp, wv, ws, mask, acts, old, adv, cats, score_t)
    want["loss"].backward()
    grads = {k: v.grad for k, v in sd.items() if v.requires_grad}
    return wp.detach(), wv.detach(), ws.detach(), {k: float(v) for k, v in want.items() if v.ndim == 0}, grads


def check_grads(named_grads, want_grads, cos_min, lo=0.8, hi=1.25, skip_small=1e-7):
    """Every parameter gradient: direction and norm vs the oracle. Gradients whose oracle norm is below `skip_small`
    relative to the largest gradient norm carry no signal (rounding noise on both sides) and are only required to be
    small as well."""
    top = max(float(np.linalg.norm(g.numpy())) for g in want_grads.values())
    bad = {}
    for name, got in named_grads:
        want = want_grads[name].numpy()
        nw = float(np.linalg.norm(want))
        if nw < skip_small * top:
            if float(np.linalg.norm(got)) > 10 * skip_small * top:
                bad[name] = ("noise-level oracle gradient but large kernel gradient", float(np.linalg.norm(got)), nw)
            continue
        c, r = cos_ratio(got, want)
        if not (c > cos_min and lo < r < hi):
            bad[name] = (round(c, 4), round(r, 4))
    return bad


# ---------------------------------------------------------------------------------------------------------------------
# (a) whole model, bf16, tensor-core schedule: forward, losses and every parameter gradient vs the oracle
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg_name,B", [("3x128", 26), ("2x256", 24), ("2x256", 49)])
def test_bf16_autograd_op_forward_loss_and_all_gradients_vs_oracle(cfg_name, B):
    cfg = CFGS[cfg_name]
    torch.manual_seed(11)
    m = SEResNetModel(SEResNetParams(**cfg))
    batch = make_batch(B, 12 + B)
    wp, wv, ws, want, want_g = oracle_step(m, batch, cfg["num_blocks"])
    obs, mask, acts, old, adv, cats, score_t = [t.to(DEV) for t in batch]
    m = m.to(DEV).train()
    m.configure_amp(True, torch.bfloat16, "cuda")
    n0 = _lib.launch_count()
    o = m(obs)
    assert o.policy_logits.dtype == torch.bfloat16 and _lib.launch_count() > n0
    assert rel(o.policy_logits.detach().float().cpu().numpy(), wp.numpy()) < 2e-2
    assert rel(o.value_logits.detach().cpu().numpy(), wv.numpy()) < 2e-2
    assert rel(o.score_lead.detach().cpu().numpy(), ws.numpy()) < 2e-2
    flat = m.last_policy_buffer[:, :A]     # the padded logits buffer the trainer consumes in place
    assert flat.requires_grad
    out2, *_ = policy_ops.ppo_policy_loss(flat, mask, acts, old, adv, 0.2)
    out3 = policy_ops.value_losses(o.value_logits, cats, o.score_lead, score_t)
    loss = out2[0] + 1.5 * out3[0] + 0.02 * out3[1] - 0.01 * out2[1]
    for got, key in ((out2[0], "policy_loss"), (out2[1], "entropy"), (out3[0], "value_loss"), (out3[1], "score_loss"), (loss, "loss")):
        assert abs(got.item() - want[key]) <= 2e-2 * max(abs(want[key]), 1e-3), (key, got.item(), want[key])
    loss.backward()
    named = [(n, p.grad.float().cpu().numpy()) for n, p in m.named_parameters()]
    assert all(np.isfinite(g).all() for _, g in named)
    bad = check_grads(named, want_g, cos_min=0.97)
    assert not bad, bad


@pytest.mark.parametrize("cfg_name,B", [("3x128", 25), ("2x256", 30)])
def test_bf16_step_fused_flat_gradient_vs_oracle(cfg_name, B):
    """The trainer's own step (`_step_fused`: raw C calls, flat gradient, GradScaler-scaled loss) — what bench.py times."""
    cfg = CFGS[cfg_name]
    torch.manual_seed(21)
    m = SEResNetModel(SEResNetParams(**cfg))
    batch = make_batch(B, 22 + B)
    _, _, _, want, want_g = oracle_step(m, batch, cfg["num_blocks"])
    obs, mask, acts, old, adv, cats, score_t = [t.to(DEV) for t in batch]
    m = m.to(DEV)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True, batch_size=B), m)
    m.train()
    km = algo._kernel_model(torch.device(DEV))
    assert km is m
    pl, vl, sl, ent, _ = algo._step_fused(km, obs, (mask, acts, old, adv, cats, score_t, adv), None)
    for got, key in ((pl, "policy_loss"), (vl, "value_loss"), (sl, "score_loss"), (ent, "entropy")):
        assert abs(got.item() - want[key]) <= 2e-2 * max(abs(want[key]), 1e-3), (key, got.item(), want[key])
    scale = float(algo.scaler.get_scale())
    assert scale > 1.0   # AMP on CUDA: the gradients below are loss-scaled
    named = [(n, (p.grad.float() / scale).cpu().numpy()) for n, p in m.named_parameters()]
    bad = check_grads(named, want_g, cos_min=0.97)
    assert not bad, bad
    # the optimiser tail unscales, clips to grad_clip and steps: the norm it reports is the oracle's gradient norm
    before = [p.detach().clone() for p in m.parameters()]
    gn = float(algo._optimizer_tail())
    want_norm = float(np.sqrt(sum(float((g.double() ** 2).sum()) for g in want_g.values())))
    assert abs(gn - want_norm) <= 5e-2 * want_norm, (gn, want_norm)
    assert any(not torch.equal(a, b) for a, b in zip(before, m.parameters()))


# ---------------------------------------------------------------------------------------------------------------------
# (b) tcgen05 data gradient, directly: conv3x3(dy, flipped pack, backend=1) == conv_transpose2d in float64
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C", [(3, 128), (7, 256), (50, 256), (301, 256)])
def test_tcgen05_dgrad_vs_float64_conv_transpose(B, C):
    g = torch.Generator().manual_seed(B + C)
    dy = torch.randn(B, C, 9, 9, generator=g).bfloat16()
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).bfloat16()
    want = F.conv_transpose2d(dy.double(), w.double(), padding=1)     # dL/dx of y = conv2d(x, w, padding=1)
    _, wd = model_ops.pack_conv_weight(w.float().to(DEV), torch.bfloat16, with_dgrad=True)
    n0 = _lib.launch_count()
    dx, *_ = model_ops.conv3x3(nhwc(dy).to(DEV), wd, backend=1)
    torch.cuda.synchronize()
    assert _lib.launch_count() > n0
    got = nchw(dx.float().cpu())
    assert rel(got.numpy(), want.numpy()) < 1e-2
    assert rel(got[:, :, 0, :].numpy(), want[:, :, 0, :].numpy()) < 1e-2      # board edges: padding taps
    assert rel(got[-1].numpy(), want[-1].numpy()) < 1e-2                      # last (partial) 3-board tile
    # and it is the adjoint of the forward kernel on the same weights: <conv(x), dy> == <x, dgrad(dy)>
    x = torch.randn(B, C, 9, 9, generator=g).bfloat16()
    wf = model_ops.pack_conv_weight(w.float().to(DEV), torch.bfloat16)
    y, *_ = model_ops.conv3x3(nhwc(x).to(DEV), wf, backend=1)
    lhs = float((nchw(y.float().cpu()).double() * dy.double()).sum())
    rhs = float((x.double() * got.double()).sum())
    assert abs(lhs - rhs) <= 1e-2 * max(abs(lhs), abs(rhs), 1.0)


# ---------------------------------------------------------------------------------------------------------------------
# (c) the real 40 x 256 network in TRAINING mode
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big_model_cpu():
    torch.manual_seed(0)
    return SEResNetModel(SEResNetParams())   # 40 x 256, reference defaults


BIG_TENSORS = ("input_conv.weight", "blocks.0.conv1.weight", "blocks.19.conv2.weight", "blocks.39.conv2.weight",
               "blocks.20.global_fc.0.weight", "policy_conv2.weight", "value_fc1.weight")


def test_fp32_40x256_training_forward_and_gradients_vs_cpu_oracle(big_model_cpu):
    """All 81 convolutions, 82 batch-statistics BatchNorms and their backward in fp32 against the CPU oracle at B = 4."""
    import copy
    m = copy.deepcopy(big_model_cpu)
    B = 4
    batch = make_batch(B, 5)
    wp, wv, ws, want, want_g = oracle_step(m, batch, 40)
    obs, mask, acts, old, adv, cats, score_t = [t.to(DEV) for t in batch]
    m = m.to(DEV).train()
    m.configure_amp(False)
    o = m(obs)
    assert rel(o.policy_logits.detach().cpu().numpy(), wp.numpy()) < 1e-4
    assert rel(o.value_logits.detach().cpu().numpy(), wv.numpy()) < 1e-4
    assert rel(o.score_lead.detach().cpu().numpy(), ws.numpy()) < 1e-4
    out2, *_ = policy_ops.ppo_policy_loss(m.last_policy_buffer[:, :A], mask, acts, old, adv, 0.2)
    out3 = policy_ops.value_losses(o.value_logits, cats, o.score_lead, score_t)
    loss = out2[0] + 1.5 * out3[0] + 0.02 * out3[1] - 0.01 * out2[1]
    assert abs(loss.item() - want["loss"]) <= 1e-4 * max(abs(want["loss"]), 1e-3)
    loss.backward()
    grads = dict((n, p.grad.cpu().numpy()) for n, p in m.named_parameters())
    # Gradients: fp32 on both sides, so what differs is summation order (81 convolutions, 82 batch-statistics layers over
    # only B*81 = 324 elements per channel). The yardstick is the SAME oracle in float64: the kernels must be as close to
    # it as the reference's own fp32 arithmetic is (within a factor 4, floor 1e-4) — measured 1e-4 .. 5e-4 for both.
    _, _, _, _, exact_g = oracle_step(big_model_cpu, batch, 40, dtype=torch.float64)
    report = {}
    for name in BIG_TENSORS:
        exact = exact_g[name].numpy()
        report[name] = (rel_l2(grads[name], exact), rel_l2(want_g[name].numpy(), exact))
    bad = {k: v for k, v in report.items() if not v[0] < max(4 * v[1], 1e-4)}
    assert not bad, report
    bad = check_grads(list(grads.items()), want_g, cos_min=0.9999, lo=0.999, hi=1.001)
    assert not bad, bad


def test_bf16_40x256_step_fused_8192_samples_vs_64_sample_oracle(big_model_cpu):
    """BASELINE configs[2] at full size through the trainer's own step. Size-independent property: a batch made of 128
    copies of a 64-sample batch has the SAME BatchNorm batch statistics, mean losses and mean-loss gradient as the
    64-sample batch — so the 8192-sample bf16 step is compared with the fp32 CPU oracle evaluated on the 64 samples."""
    import copy
    m = copy.deepcopy(big_model_cpu)
    base, reps = 64, 128
    batch = make_batch(base, 9)
    _, _, _, want, want_g = oracle_step(m, batch, 40)
    big = [t.repeat(reps, *([1] * (t.ndim - 1))).to(DEV) for t in batch]
    obs, mask, acts, old, adv, cats, score_t = big
    assert obs.shape[0] == 8192
    m = m.to(DEV)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True, batch_size=8192), m)
    m.train()
    n0 = _lib.launch_count()
    pl, vl, sl, ent, v_logits = algo._step_fused(algo._kernel_model(torch.device(DEV)), obs, (mask, acts, old, adv, cats, score_t, adv), None)
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 > 500
    for got, key in ((pl, "policy_loss"), (vl, "value_loss"), (sl, "score_loss"), (ent, "entropy")):
        assert np.isfinite(got.item())
        assert abs(got.item() - want[key]) <= 2e-2 * max(abs(want[key]), 1e-3), (key, got.item(), want[key])
    # every copy of a sample gets the same value logits (batch-statistics BatchNorm is permutation invariant)
    v = v_logits.float().view(reps, base, 3)
    assert float((v - v[0:1]).abs().max()) <= 2e-2 * float(v.abs().max())
    scale = float(algo.scaler.get_scale())
    flat = algo._flat_grad
    assert bool(torch.isfinite(flat).all())
    grads = dict((n, (p.grad.float() / scale).cpu().numpy()) for n, p in m.named_parameters())
    # Calibration: what the REFERENCE ITSELF produces on this GPU under AMP — the oracle's PyTorch ops under bf16
    # autocast (cuDNN / cuBLAS kernels) on the same 64 samples. 40 bf16 residual blocks turn rounding noise into
    # gradient-direction noise that grows towards the stem; the kernels must stay as close to the fp32 gradient as the
    # reference's own bf16 path does (minus a small margin), and never below 0.85.
    sd_amp = {k: v.detach().clone().to(DEV) for k, v in big_model_cpu.state_dict().items()}
    for t in sd_amp.values():
        if t.is_floating_point() and t.ndim > 0:
            t.requires_grad_(True)
    b64 = [t.to(DEV) for t in batch]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ap, av, asc = O.seresnet_forward(sd_amp, b64[0], 40, training=True)
        amp_losses = O.ppo_losses(ap.float(), av.float(), asc.float(), *b64[1:])
    amp_losses["loss"].backward()
    report = {}
    for name in BIG_TENSORS:
        ours = cos_ratio(grads[name], want_g[name].numpy())
        ref_amp = cos_ratio(sd_amp[name].grad.float().cpu().numpy(), want_g[name].numpy())
        report[name] = {"ours": ours, "reference_bf16_autocast": ref_amp}
    print("gradient cosine / norm ratio vs fp32 oracle:", report)
    low = {k: v for k, v in report.items()
           if not (v["ours"][0] > max(0.85, v["reference_bf16_autocast"][0] - 0.05) and 0.8 < v["ours"][1] < 1.25)}
    assert not low, report
    gn = float(torch.linalg.vector_norm(flat)) / scale
    want_norm = float(np.sqrt(sum(float((g.double() ** 2).sum()) for g in want_g.values())))
    assert abs(gn - want_norm) <= 0.1 * want_norm, (gn, want_norm)


def test_policy_loss_8192_rows_illegal_logits_get_exactly_zero_gradient():
    B = 8192
    g = torch.Generator(device=DEV).manual_seed(3)
    logits = torch.randn(B, 11264, device=DEV, generator=g).bfloat16()[:, :A].requires_grad_(True)
    mask = torch.rand(B, A, device=DEV, generator=g) < 0.007
    acts = torch.randint(0, A, (B,), device=DEV, generator=g)
    mask[torch.arange(B, device=DEV), acts] = True
    out2, *_ = policy_ops.ppo_policy_loss(logits, mask, acts, -3 * torch.rand(B, device=DEV, generator=g),
                                          torch.randn(B, device=DEV, generator=g), 0.2)
    (out2[0] - 0.01 * out2[1]).backward()
    gl = logits.grad
    assert bool(torch.isfinite(gl).all())
    assert bool((gl[~mask] == 0).all())                 # bit-exact bar: illegal actions never receive gradient
    assert bool((gl[mask] != 0).any())
