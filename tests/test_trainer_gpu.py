"""The drop-in trainer on CUDA: the fused update() (C-ABI forward/backward, flat gradients)
against the reference's own update() golden run; rollout select_actions semantics."""
import numpy as np
import pytest
import torch

from keisei_b200 import _lib
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams, KataGoRolloutBuffer
from keisei_b200.model_registry import build_model
from keisei_b200.value_adapter import get_value_adapter
from test_trainer_host import TINY, check_update, run_update_against_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_update_matches_reference_golden_cuda_fp32():
    n0 = _lib.launch_count()
    g, model, metrics = run_update_against_golden(DEV)
    assert _lib.launch_count() - n0 > 50  # the kernels ran
    check_update(g, model, metrics, 2e-4)


def test_update_env_ids_layout_and_adapter_cuda():
    torch.manual_seed(0)
    model = build_model("se_resnet", dict(TINY)).to(DEV)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=5, epochs_per_batch=2), model)
    N, A = 3, 11259
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
    for t in range(4):
        n = 2 if t % 2 else 3  # ragged: split-merge style env subsets
        ids = torch.arange(n)
        obs = torch.randn(n, 50, 9, 9, device=DEV)
        mask = torch.rand(n, A, device=DEV) < 0.01; mask[:, 3] = True
        a, lp, v = algo.select_actions(obs, mask, get_value_adapter("multi_head", score_blend_alpha=0.2))
        assert mask[torch.arange(n), a].all()
        term = torch.zeros(n, dtype=torch.bool)
        buf.add(obs, a, lp, v, torch.zeros(n), term, term, mask, torch.full((n,), -1), torch.zeros(n), env_ids=ids,
                next_value_override=torch.full((n,), float("nan")))
    before = [p.detach().clone() for p in model.parameters()]
    m = algo.update(buf, torch.randn(N, device=DEV), value_adapter=get_value_adapter("multi_head"))
    assert all(np.isfinite(v) for v in m.values()) and m["score_loss"] == 0.0
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    algo.flush_timings()
    assert len(algo.timings["update_forward_backward_ms"]) == 2 * 2 and len(algo.timings["select_actions_forward_ms"]) == 4


def test_select_actions_zero_legal_raises_cuda():
    model = build_model("se_resnet", dict(TINY)).to(DEV)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(), model)
    obs = torch.randn(4, 50, 9, 9, device=DEV)
    mask = torch.ones(4, 11259, dtype=torch.bool, device=DEV); mask[2] = False
    with pytest.raises(RuntimeError, match=r"Environments \[2\] have zero legal actions"):
        algo.select_actions(obs, mask)
    assert model.training


def test_update_bf16_amp_runs_tcgen05_and_tracks_fp32():
    torch.manual_seed(1)
    cfg = dict(num_blocks=2, channels=128, se_reduction=8, global_pool_channels=32, policy_channels=16,
               value_fc_size=32, score_fc_size=32)
    m32 = build_model("se_resnet", dict(cfg)).to(DEV)
    m16 = build_model("se_resnet", dict(cfg)).to(DEV)
    m16.load_state_dict(m32.state_dict())
    N, T, A = 6, 4, 11259
    g = torch.Generator().manual_seed(2)
    def fill(buf):
        gg = torch.Generator().manual_seed(3)
        for t in range(T):
            obs = torch.randn(N, 50, 9, 9, generator=gg)
            mask = torch.rand(N, A, generator=gg) < 0.01
            a = torch.randint(0, A, (N,), generator=gg); mask[torch.arange(N), a] = True
            term = torch.rand(N, generator=gg) < 0.2
            buf.add(obs, a, -2 * torch.rand(N, generator=gg), 0.3 * torch.randn(N, generator=gg), term.float(), term, term, mask,
                    torch.where(term, torch.randint(0, 3, (N,), generator=gg), torch.full((N,), -1)), torch.randn(N, generator=gg).clamp(-1, 1))
    out = {}
    for name, model, amp in (("fp32", m32, False), ("bf16", m16, True)):
        buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
        fill(buf)
        algo = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=N * T, epochs_per_batch=1, use_amp=amp), model)
        out[name] = algo.update(buf, torch.zeros(N, device=DEV))
    for k in ("policy_loss", "value_loss", "score_loss", "entropy"):
        a, b = out["bf16"][k], out["fp32"][k]
        assert abs(a - b) <= 2e-2 * max(abs(b), 1e-3), (k, a, b)


def _fill(buf, N, T, A, seed, dev=None, with_override=False):
    gg = torch.Generator().manual_seed(seed)
    for t in range(T):
        obs = torch.randn(N, 50, 9, 9, generator=gg)
        mask = torch.rand(N, A, generator=gg) < 0.01
        a = torch.randint(0, A, (N,), generator=gg); mask[torch.arange(N), a] = True
        term = torch.rand(N, generator=gg) < 0.2
        fields = [obs, a, -2 * torch.rand(N, generator=gg), 0.3 * torch.randn(N, generator=gg), term.float(), term, term, mask,
                  torch.where(term, torch.randint(0, 3, (N,), generator=gg), torch.full((N,), -1)),
                  torch.randn(N, generator=gg).clamp(-1, 1)]
        ov = torch.where(torch.rand(N, generator=gg) < 0.3, torch.randn(N, generator=gg), torch.full((N,), float("nan")))
        if dev is not None:
            fields = [f.to(dev) for f in fields]
            ov = ov.to(dev)
        buf.add(*fields, next_value_override=ov if with_override else None)


def test_device_resident_buffer_matches_host_buffer():
    """KataGoRolloutBuffer(device=cuda): same flatten() contents, same perspective overrides, and a
    bit-identical update() (same seeds -> same shuffles) with no host<->device copy of observations or masks."""
    N, T, A = 5, 6, 11259
    host, devb = KataGoRolloutBuffer(N, (50, 9, 9), A), KataGoRolloutBuffer(N, (50, 9, 9), A, device=DEV)
    _fill(host, N, T, A, 7)
    _fill(devb, N, T, A, 7, dev=DEV)
    host.fill_alternating_perspective_overrides(); devb.fill_alternating_perspective_overrides()
    fh, fd = host.flatten(), devb.flatten()
    assert set(fh.keys()) | {"legal_masks_packed"} == set(fd.keys())   # the device buffer keeps its masks bit-packed
    for k in fh:
        assert fd[k].device.type == "cuda" and fd[k].dtype == fh[k].dtype and fd[k].shape == fh[k].shape, k
        assert torch.equal(fd[k].cpu().nan_to_num(7.0), fh[k].nan_to_num(7.0)), k
    results = []
    for buf in (host, devb):
        torch.manual_seed(5)
        model = build_model("se_resnet", dict(TINY)).to(DEV)
        algo = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=10, epochs_per_batch=2), model)
        torch.manual_seed(6)
        m = algo.update(buf, torch.linspace(-1, 1, N, device=DEV))
        assert buf.size == 0
        results.append((m, [p.detach().clone() for p in model.parameters()]))
    assert results[0][0].keys() == results[1][0].keys()
    for k, v in results[0][0].items():
        assert abs(v - results[1][0][k]) <= 1e-5 * max(1.0, abs(v)), k
    for a, b in zip(results[0][1], results[1][1]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)


def test_device_resident_buffer_guards():
    N, A = 3, 11259
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A, device=DEV)
    z = lambda *s, **k: torch.zeros(*s, device=DEV, **k)  # noqa: E731
    ok = dict(obs=z(N, 50, 9, 9), actions=z(N, dtype=torch.long), log_probs=z(N), values=z(N), rewards=z(N),
              dones=z(N, dtype=torch.bool), terminated=z(N, dtype=torch.bool), legal_masks=z(N, A, dtype=torch.bool),
              value_categories=z(N, dtype=torch.long), score_targets=z(N))
    with pytest.raises(ValueError, match="Cannot flatten an empty buffer"):
        buf.flatten()
    with pytest.raises(AssertionError, match="terminated must be a subset of dones"):
        buf.add(**{**ok, "terminated": torch.ones(N, dtype=torch.bool, device=DEV)})
    with pytest.raises(ValueError, match=r"invalid values \{5\}"):
        buf.add(**{**ok, "value_categories": torch.tensor([0, 5, -1], device=DEV)})
    with pytest.raises(ValueError, match="contains NaN"):
        buf.add(**{**ok, "score_targets": torch.tensor([0.0, float("nan"), 0.0], device=DEV)})
    with pytest.raises(ValueError, match="unnormalized"):
        buf.add(**{**ok, "score_targets": torch.tensor([0.0, 40.0, 0.0], device=DEV)})
    assert buf.size == 0
    buf.add(**ok)
    assert buf.size == 1 and buf.flatten()["observations"].is_cuda


def _two_trainers(use_amp):
    torch.manual_seed(0)
    cfg = dict(num_blocks=2, channels=64, se_reduction=8, global_pool_channels=16, policy_channels=8, value_fc_size=16, score_fc_size=16)
    a = build_model("se_resnet", dict(cfg)).to(DEV)
    b = build_model("se_resnet", dict(cfg)).to(DEV)
    b.load_state_dict(a.state_dict())
    ta = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=12, epochs_per_batch=1, use_amp=use_amp), a)
    tb = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=12, epochs_per_batch=1, use_amp=use_amp), b)
    tb.fused_optimizer_tail = False          # stock path: foreach unscale + norm + PyTorch's fused Adam
    return a, b, ta, tb


@pytest.mark.parametrize("use_amp", [False, True])
def test_fused_optimizer_tail_matches_torch_adam_and_keeps_state_dict(use_amp):
    """global-norm + clip + Adam in two launches (csrc/optim.cu) against the stock sequence
    (reference katago_ppo.py:926-933) over several steps, and the optimizer state stays a plain torch.optim.Adam state."""
    a, b, ta, tb = _two_trainers(use_amp)
    N, T, A = 6, 4, 11259
    for rep in range(3):
        for algo in (ta, tb):
            buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
            _fill(buf, N, T, A, 40 + rep)
            torch.manual_seed(100 + rep)
            algo.update(buf, torch.zeros(N, device=DEV))
    # Two independent trajectories: the gradients of two runs differ in their last bits (3.5e-8 relative: atomics in the
    # weight-gradient / BatchNorm sums, tools/grad_nondeterminism.py), and Adam divides by sqrt(v) + eps — an element whose gradient is
    # below eps = 1e-8 moves by lr * g / eps, so last-bit noise there becomes a step of order lr, after which every
    # gradient of the two runs differs at the 1e-3 level. The trajectories are therefore compared loosely (an algorithmic
    # error in the tail — bias correction, eps placement, clip coefficient — is an O(1) relative error); the tight check
    # of the tail's arithmetic is test_fused_optimizer_tail_equals_stock_tail_on_identical_gradients below.
    tol = dict(rtol=2e-2, atol=2e-3) if use_amp else dict(rtol=5e-3, atol=2e-5)
    lr, steps = ta.params.learning_rate, 6
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        d = (p - q).abs()
        bad = d > tol["atol"] + tol["rtol"] * q.abs()
        assert float(bad.float().mean()) <= 0.02 and float(d.max()) <= 2.2 * steps * lr, (n, float(bad.float().mean()), float(d.max()))
    sa, sb = ta.optimizer.state_dict(), tb.optimizer.state_dict()
    assert sa["param_groups"] == sb["param_groups"] and sa["state"].keys() == sb["state"].keys()
    for k in sa["state"]:
        assert set(sa["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 6.0   # 2 minibatches x 3 updates
    ea = torch.cat([sa["state"][k]["exp_avg"].flatten() for k in sa["state"]])
    eb = torch.cat([sb["state"][k]["exp_avg"].flatten() for k in sb["state"]])
    assert float((ea - eb).norm()) <= (0.25 if use_amp else 5e-3) * float(eb.norm()), float((ea - eb).norm() / eb.norm())
    # the state is interchangeable: a fresh torch Adam loads it (checkpoint.py:123 restores positionally) ...
    fresh = torch.optim.Adam(a.parameters(), lr=ta.params.learning_rate, fused=True)
    fresh.load_state_dict(sa)
    # ... and the trainer keeps going on a replaced optimizer (katago_loop.py:1859 swaps `.optimizer` at seat rotation)
    ta.optimizer = fresh
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
    _fill(buf, N, T, A, 77)
    before = [p.detach().clone() for p in a.parameters()]
    m = ta.update(buf, torch.zeros(N, device=DEV))
    assert all(np.isfinite(v) for v in m.values())
    assert any(not torch.equal(x, y) for x, y in zip(before, a.parameters()))
    assert float(fresh.state_dict()["state"][0]["step"]) == 8.0


@pytest.mark.parametrize("use_amp", [False, True])
def test_fused_optimizer_tail_equals_stock_tail_on_identical_gradients(use_amp):
    """The tail's arithmetic in isolation: both trainers are handed the SAME flat gradient every step (the second trainer's
    buffer is overwritten with the first one's), so unscale + global-norm clip + Adam of csrc/optim.cu must reproduce
    GradScaler.unscale_ + clip_grad_norm_ + torch.optim.Adam(fused) to fp32 rounding, step after step."""
    a, b, ta, tb = _two_trainers(use_amp)
    dev = torch.device(DEV)
    N, A = 12, 11259
    g = torch.Generator().manual_seed(9)
    for step in range(5):
        obs = torch.randn(N, 50, 9, 9, generator=g).to(DEV)
        mask = torch.rand(N, A, generator=g) < 0.01
        acts = torch.randint(0, A, (N,), generator=g); mask[torch.arange(N), acts] = True
        mb = (mask.to(DEV), acts.to(DEV), (-2 * torch.rand(N, generator=g)).to(DEV), torch.randn(N, generator=g).to(DEV),
              torch.randint(0, 3, (N,), generator=g).to(DEV), torch.randn(N, generator=g).clamp(-1, 1).to(DEV),
              torch.randn(N, generator=g).to(DEV))
        for model, algo in ((a, ta), (b, tb)):
            model.train()
            algo._step_fused(algo._kernel_model(dev), obs, mb, None)
        tb._flat_grad.copy_(ta._flat_grad)
        ta._optimizer_tail(); tb._optimizer_tail()
        for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-8), (step, n, float((p - q).abs().max()))
    sa, sb = ta.optimizer.state_dict(), tb.optimizer.state_dict()
    for k in sa["state"]:
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 5.0
        for key in ("exp_avg", "exp_avg_sq"):
            x, y = sa["state"][k][key], sb["state"][k][key]   # torch: lerp_ / addcmul_; here: fused multiply-adds
            assert float((x - y).abs().max()) <= 1e-5 * float(y.abs().max()) + 1e-30, (k, key, float((x - y).abs().max()))


def test_fused_optimizer_tail_skips_non_finite_gradients():
    """GradScaler semantics in the fused tail: a non-finite gradient leaves parameters, moments and step untouched and
    halves the loss scale."""
    a, _, ta, _ = _two_trainers(True)
    N, T, A = 6, 2, 11259
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
    _fill(buf, N, T, A, 5)
    ta.update(buf, torch.zeros(N, device=DEV))              # one good step: state exists
    scale0 = ta.scaler.get_scale()
    before = [p.detach().clone() for p in a.parameters()]
    step0 = float(ta.optimizer.state_dict()["state"][0]["step"])
    km = ta._kernel_model(torch.device(DEV))
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(12, 50, 9, 9, generator=g).to(DEV)
    mask = torch.ones(12, A, dtype=torch.bool, device=DEV)
    z = torch.zeros(12, device=DEV)
    mb = (mask, torch.zeros(12, dtype=torch.long, device=DEV), z, z + 1, torch.zeros(12, dtype=torch.long, device=DEV), z, z)
    a.train()
    ta._step_fused(km, obs, mb, None)
    ta._flat_grad[123] = float("inf")
    ta._optimizer_tail()
    torch.cuda.synchronize()
    for x, y in zip(before, a.parameters()):
        assert torch.equal(x, y)
    assert float(ta.optimizer.state_dict()["state"][0]["step"]) == step0
    assert ta.scaler.get_scale() == scale0 * 0.5
