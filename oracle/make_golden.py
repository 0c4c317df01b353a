"""Generate tests/golden/*.npz by running the REAL reference (imported read-only from
/root/reference) on seeded inputs. Run in the build container only:

    python oracle/make_golden.py

The GPU box has no /root/reference; tests read the committed .npz files. TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

TINY = dict(num_blocks=2, channels=32, se_reduction=4, global_pool_channels=16, policy_channels=8,
            value_fc_size=16, score_fc_size=16, obs_channels=50)


def _np(t):
    return t.detach().cpu().numpy()


def make_inputs(B: int, seed: int, A: int = 11259):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(B, 50, 9, 9, generator=g)
    mask = torch.rand(B, A, generator=g) < 0.01
    actions = torch.randint(0, A, (B,), generator=g)
    mask[torch.arange(B), actions] = True
    old_logp = -3.0 * torch.rand(B, generator=g)
    adv = torch.randn(B, generator=g)
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    cats = torch.randint(-1, 3, (B,), generator=g)
    score_t = torch.randn(B, generator=g).clamp(-1.5, 1.5)
    return obs, mask, actions, old_logp, adv, cats, score_t


def golden_model():
    from keisei.training.model_registry import build_model
    from keisei.training.katago_ppo import ppo_clip_loss, wdl_cross_entropy_loss
    import torch.nn.functional as F

    torch.manual_seed(0)
    model = build_model("se_resnet", dict(TINY))
    # make BN affine/running stats non-trivial
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
        for name, p in model.named_parameters():
            if ".bn" in name or "_bn" in name:
                if name.endswith("weight"):
                    p.copy_(0.5 + torch.rand(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    B = 6
    obs, mask, actions, old_logp, adv, cats, score_t = make_inputs(B, seed=2)

    out = {f"sd/{k}": _np(v) for k, v in sd0.items()}
    out.update(obs=_np(obs), mask=_np(mask), actions=_np(actions), old_logp=_np(old_logp), adv=_np(adv),
               cats=_np(cats), score_t=_np(score_t))

    model.eval()
    with torch.no_grad():
        o = model(obs)
    out.update(eval_policy=_np(o.policy_logits.contiguous()), eval_value=_np(o.value_logits), eval_score=_np(o.score_lead))

    model.train()
    o = model(obs)
    flat = o.policy_logits.reshape(B, -1)
    masked = flat.masked_fill(~mask, float("-inf"))
    logp_all = F.log_softmax(masked, dim=-1)
    new_logp = logp_all.gather(1, actions.unsqueeze(1)).squeeze(1)
    pl = ppo_clip_loss(new_logp, old_logp, adv, 0.2)
    probs = logp_all.exp()
    ent = -(probs * logp_all.masked_fill(~mask, 0.0)).sum(-1).mean()
    vl = wdl_cross_entropy_loss(o.value_logits, cats)
    sl = F.mse_loss(o.score_lead.squeeze(-1), score_t)
    loss = 1.0 * pl + 1.5 * vl + 0.02 * sl - 0.01 * ent
    loss.backward()
    out.update(train_policy=_np(o.policy_logits.contiguous()), train_value=_np(o.value_logits), train_score=_np(o.score_lead),
               loss=_np(loss), policy_loss=_np(pl), value_loss=_np(vl), score_loss=_np(sl), entropy=_np(ent),
               new_logp=_np(new_logp))
    for k, p in model.named_parameters():
        out[f"grad/{k}"] = _np(p.grad)
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out[f"sd_after/{k}"] = _np(v)
    np.savez_compressed(OUT / "seresnet_tiny.npz", **out)
    print("seresnet_tiny.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


def golden_gae():
    from keisei.training.gae import compute_gae, compute_gae_gpu, compute_gae_padded, compute_gae_padded_gpu
    g = torch.Generator().manual_seed(3)
    T, N = 37, 7
    r = torch.randn(T, N, generator=g)
    v = 0.3 * torch.randn(T, N, generator=g)
    term = torch.rand(T, N, generator=g) < 0.1
    nv = torch.randn(N, generator=g)
    ov = torch.full((T, N), float("nan"))
    sel = torch.rand(T, N, generator=g) < 0.15
    ov[sel] = torch.randn(int(sel.sum()), generator=g)
    ov[-1, 0] = 0.77  # override at the last step too
    out = dict(r=_np(r), v=_np(v), term=_np(term), nv=_np(nv), ov=_np(ov))
    out["adv_plain"] = _np(compute_gae(r, v, term, nv, 0.99, 0.95))
    out["adv_plain_gpufn"] = _np(compute_gae_gpu(r, v, term, nv, 0.99, 0.95))
    out["adv_override"] = _np(compute_gae(r, v, term, nv, 0.99, 0.95, next_value_override=ov))
    out["adv_override_gpufn"] = _np(compute_gae_gpu(r, v, term, nv, 0.99, 0.95, next_value_override=ov))
    out["adv_termfloat"] = _np(compute_gae(r, v, term.float(), nv, 0.97, 0.9))
    out["adv_1d"] = _np(compute_gae(r[:, 0], v[:, 0], term[:, 0], nv[0], 0.99, 0.95))
    lengths = torch.tensor([37, 5, 1, 20, 36, 12, 37])
    termp = term.float().clone()
    for i, L in enumerate(lengths.tolist()):
        termp[L:, i] = 1.0
    out["lengths"] = _np(lengths)
    out["termp"] = _np(termp)
    out["adv_padded"] = _np(compute_gae_padded(r, v, termp, nv, lengths, 0.99, 0.95))
    out["adv_padded_gpufn"] = _np(compute_gae_padded_gpu(r, v, termp, nv, lengths, 0.99, 0.95))
    out["adv_padded_override"] = _np(compute_gae_padded(r, v, termp, nv, lengths, 0.99, 0.95, next_value_override=ov))
    a = out["adv_override"].reshape(-1)
    at = torch.from_numpy(a.copy())
    out["adv_override_normalized"] = _np((at - at.mean()) / (at.std() + 1e-8))
    # the reference's hand-computed known answers (tests/test_gae.py:10-40, test_katago_ppo.py:449-489)
    out["known_adv_3step"] = _np(compute_gae(torch.tensor([1.0, 1.0, 1.0]), torch.tensor([0.5, 0.5, 0.5]),
                                             torch.tensor([False, False, False]), torch.tensor(0.5), 0.99, 0.95))
    np.savez_compressed(OUT / "gae.npz", **out)
    print("gae.npz ok")


def golden_rollout():
    """select_actions log-prob semantics for given actions, fp32 and bf16 (katago_ppo.py:599-613)."""
    from keisei.training.katago_ppo import KataGoPPOAlgorithm
    from keisei.training.value_adapter import get_value_adapter
    g = torch.Generator().manual_seed(4)
    B, A = 8, 11259
    logits = 2.0 * torch.randn(B, A, generator=g)
    mask = torch.rand(B, A, generator=g) < 0.01
    mask[0] = False
    mask[0, 1234] = True  # single legal action -> log_prob ~ 0
    logits[1, :] = -5.0
    mask[1, 100] = True
    logits[1, 100] = 12.0  # near-certain action (exercises the 1-eps clamp)
    actions = torch.stack([torch.nonzero(mask[i])[torch.randint(0, int(mask[i].sum()), (1,), generator=g)][0, 0] for i in range(B)])
    actions[1] = 100
    out = dict(logits=_np(logits), mask=_np(mask), actions=_np(actions))
    for name, lg in (("f32", logits), ("bf16", logits.to(torch.bfloat16))):
        masked = lg.masked_fill(~mask, float("-inf"))
        probs = torch.softmax(masked, dim=-1)
        dist = torch.distributions.Categorical(probs, validate_args=False)
        out[f"logp_{name}"] = _np(dist.log_prob(actions).float())
        # low-probability action as well (exercises the eps clamp from below)
        worst = torch.where(mask, lg.float(), torch.full_like(lg.float(), float("inf"))).argmin(dim=-1)
        out[f"logp_worst_{name}"] = _np(dist.log_prob(worst).float())
        out["worst"] = _np(worst)
    vl = torch.randn(B, 3, generator=g)
    sc = 2.0 * torch.randn(B, 1, generator=g)
    out.update(value_logits=_np(vl), score_lead=_np(sc))
    out["scalar_value"] = _np(KataGoPPOAlgorithm.scalar_value(vl))
    ad = get_value_adapter("multi_head", 1.5, 0.02, 0.3)
    out["scalar_value_blend03"] = _np(ad.scalar_value_blended(vl, sc))
    np.savez_compressed(OUT / "rollout.npz", **out)
    print("rollout.npz ok")


def golden_update():
    """One reference KataGoPPOAlgorithm.update() over a tiny (T,N) buffer, single minibatch."""
    from keisei.training.model_registry import build_model
    from keisei.training.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams, KataGoRolloutBuffer
    torch.manual_seed(5)
    model = build_model("se_resnet", dict(TINY))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    T, N, A = 6, 4, 11259
    g = torch.Generator().manual_seed(6)
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
    steps = []
    for t in range(T):
        obs = torch.randn(N, 50, 9, 9, generator=g)
        mask = torch.rand(N, A, generator=g) < 0.01
        actions = torch.randint(0, A, (N,), generator=g)
        mask[torch.arange(N), actions] = True
        logp = -3.0 * torch.rand(N, generator=g)
        values = 0.3 * torch.randn(N, generator=g)
        term = torch.rand(N, generator=g) < 0.2
        rewards = torch.where(term, torch.sign(torch.randn(N, generator=g)), torch.zeros(N))
        cats = torch.where(term, torch.randint(0, 3, (N,), generator=g), torch.full((N,), -1))
        score_t = torch.randn(N, generator=g).clamp(-1.5, 1.5)
        ov = torch.full((N,), float("nan"))
        if t == 2:
            ov[1] = 0.25
        steps.append(dict(obs=obs, actions=actions, logp=logp, values=values, rewards=rewards, term=term,
                          mask=mask, cats=cats, score_t=score_t, ov=ov))
        buf.add(obs, actions, logp, values, rewards, term, term, mask, cats, score_t, next_value_override=ov)
    next_values = torch.randn(N, generator=g)
    params = KataGoPPOParams(batch_size=T * N, epochs_per_batch=1, learning_rate=1e-3)
    algo = KataGoPPOAlgorithm(params, model)
    metrics = algo.update(buf, next_values)
    out = {f"sd/{k}": _np(v) for k, v in sd0.items()}
    for k in steps[0]:
        out[f"steps/{k}"] = np.stack([_np(s[k]) for s in steps])
    out["next_values"] = _np(next_values)
    for k, v in metrics.items():
        out[f"metrics/{k}"] = np.float64(v)
    for k, v in model.state_dict().items():
        out[f"sd_after/{k}"] = _np(v)
    np.savez_compressed(OUT / "update_tiny.npz", **out)
    print("update_tiny.npz ok", metrics)


def golden_resnet():
    """ResNetModel (the `resnet` registry entry) forward train/eval, BN running stats, and the gradients of the
    restated scalar-PPO loss — model + autograd from the REAL reference classes, the loss from the surviving
    reference functions (ppo_clip_loss, ScalarValueAdapter)."""
    from keisei.training.model_registry import build_model
    from keisei.training.katago_ppo import ppo_clip_loss
    from keisei.training.value_adapter import ScalarValueAdapter
    import torch.nn.functional as F

    torch.manual_seed(3)
    model = build_model("resnet", dict(hidden_size=32, num_layers=2))
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
        for name, p in model.named_parameters():
            if ".bn" in name or "_bn" in name:
                if name.endswith("weight"):
                    p.copy_(0.5 + torch.rand(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
        # keep the fixture small: the 11259 x 162 policy_fc matrix snapped to a 1/64 grid compresses ~8x
        model.policy_fc.weight.copy_(torch.round(model.policy_fc.weight * 64) / 64)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    B = 6
    obs, mask, actions, old_logp, adv, _cats, _score = make_inputs(B, seed=5)
    returns = torch.randn(B, generator=g).clamp(-1, 1)
    out = {f"sd/{k}": _np(v) for k, v in sd0.items()}
    out.update(obs=_np(obs), mask=_np(mask), actions=_np(actions), old_logp=_np(old_logp), adv=_np(adv), returns=_np(returns))
    model.eval()
    with torch.no_grad():
        pl_e, v_e = model(obs)
    out.update(eval_policy=_np(pl_e), eval_value=_np(v_e))
    model.train()
    logits, value = model(obs)
    logp_all = F.log_softmax(logits.masked_fill(~mask, float("-inf")), dim=-1)
    new_logp = logp_all.gather(1, actions.unsqueeze(1)).squeeze(1)
    pl = ppo_clip_loss(new_logp, old_logp, adv, 0.2)
    ent = -(logp_all.exp() * logp_all.masked_fill(~mask, 0.0)).sum(-1).mean()
    vl = ScalarValueAdapter().compute_value_loss(value, returns, None, None)
    loss = pl + 0.5 * vl - 0.01 * ent
    loss.backward()
    out.update(train_policy=_np(logits), train_value=_np(value), loss=_np(loss), policy_loss=_np(pl), value_loss=_np(vl),
               entropy=_np(ent), new_logp=_np(new_logp))
    for k, p in model.named_parameters():
        out[f"grad/{k}"] = _np(p.grad)
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out[f"sd_after/{k}"] = _np(v)
    np.savez_compressed(OUT / "resnet_tiny.npz", **out)
    print("resnet_tiny.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


def _buffer_steps(seed: int, T: int, N: int, ragged: bool, A: int = 11259):
    """Seeded rollout steps; `ragged`: split-merge style env subsets with env_ids (katago_loop.py:284-431)."""
    g = torch.Generator().manual_seed(seed)
    steps = []
    for t in range(T):
        ids = torch.arange(N) if not ragged else torch.sort(torch.randperm(N, generator=g)[: max(1, N - (t % 3))]).values
        n = ids.numel()
        obs = torch.randn(n, 50, 9, 9, generator=g)
        mask = torch.rand(n, A, generator=g) < 0.01
        actions = torch.randint(0, A, (n,), generator=g)
        mask[torch.arange(n), actions] = True
        term = torch.rand(n, generator=g) < 0.15
        trunc = (torch.rand(n, generator=g) < 0.1) & ~term
        dones = term | trunc
        rewards = torch.where(term, torch.randint(-1, 2, (n,), generator=g).float(), torch.zeros(n))
        cats = torch.where(term, torch.randint(0, 3, (n,), generator=g), torch.full((n,), -1))
        ov = torch.where(trunc, torch.randn(n, generator=g), torch.full((n,), float("nan")))
        steps.append(dict(obs=obs, actions=actions, logp=-3 * torch.rand(n, generator=g), values=0.5 * torch.randn(n, generator=g),
                          rewards=rewards, dones=dones, term=term, mask=mask, cats=cats,
                          score_t=torch.randn(n, generator=g).clamp(-1.5, 1.5), ids=ids, ov=ov))
    return steps


def golden_buffer():
    """KataGoRolloutBuffer (katago_ppo.py:128-388): flatten() of the (T, N) grid layout after
    fill_alternating_perspective_overrides(), flatten() of the ragged env_ids layout, and the reference update() on the
    ragged buffer (per-env padded GAE, katago_ppo.py:649-773) — metrics and a parameter checksum."""
    from keisei.training.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams, KataGoRolloutBuffer
    from keisei.training.model_registry import build_model

    out = {}
    T, N, A = 5, 4, 11259
    for name, ragged in (("grid", False), ("ragged", True)):
        steps = _buffer_steps(21 if not ragged else 22, T, N, ragged)
        buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
        for st in steps:
            buf.add(st["obs"], st["actions"], st["logp"], st["values"], st["rewards"], st["dones"], st["term"], st["mask"],
                    st["cats"], st["score_t"], env_ids=st["ids"] if ragged else None, next_value_override=st["ov"])
        buf.fill_alternating_perspective_overrides()
        flat = buf.flatten()
        for k, v in flat.items():
            if k in ("observations", "legal_masks"):
                # content is checked through a checksum (the raw tensors are regenerated from the seed by the test)
                out[f"{name}/{k}_sum"] = np.array(v.double().sum().item() if v.dtype != torch.bool else int(v.sum().item()))
            else:
                out[f"{name}/{k}"] = _np(v)
        out[f"{name}/size"] = np.array(buf.size)
    # the reference update() on the ragged buffer
    steps = _buffer_steps(22, T, N, True)
    torch.manual_seed(0)
    model = build_model("se_resnet", dict(TINY))
    for k, v in model.state_dict().items():
        out["sd/" + k] = _np(v).copy()   # a copy: update() below changes the parameters in place
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
    for st in steps:
        buf.add(st["obs"], st["actions"], st["logp"], st["values"], st["rewards"], st["dones"], st["term"], st["mask"],
                st["cats"], st["score_t"], env_ids=st["ids"], next_value_override=st["ov"])
    total = sum(st["ids"].numel() for st in steps)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=total, epochs_per_batch=1, learning_rate=1e-3), model)
    nv = torch.linspace(-0.5, 0.5, N)
    torch.manual_seed(5)
    metrics = algo.update(buf, nv)
    out["update/next_values"] = _np(nv)
    for k, v in metrics.items():
        out["update/metric/" + k] = np.array(v)
    for k, v in model.state_dict().items():
        if v.is_floating_point():
            out["update/after/" + k] = _np(v)
    np.savez_compressed(OUT / "buffer.npz", **out)
    print("buffer.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


def golden_split_merge():
    """The reference's `split_merge_step` (katago_loop.py:284-431) on CPU with three real tiny models: cohort mode with a
    per-env learner side and a blending value adapter, and legacy single-opponent mode. The sampled actions depend on
    torch's CPU generator (seeded right before each call)."""
    from keisei.training.katago_loop import split_merge_step
    from keisei.training.model_registry import build_model
    from keisei.training.value_adapter import get_value_adapter
    N, A = 12, 11259
    out = {}
    models = []
    for k in range(3):
        torch.manual_seed(10 + k)
        m = build_model("se_resnet", dict(TINY))
        m.train()
        with torch.no_grad():
            m(torch.randn(16, 50, 9, 9))      # non-trivial BatchNorm running statistics
        m.eval()
        models.append(m)
        for name, v in m.state_dict().items():
            out[f"sd{k}/{name}"] = _np(v)
    g = torch.Generator().manual_seed(21)
    obs = torch.randn(N, 50, 9, 9, generator=g)
    mask = torch.rand(N, A, generator=g) < 0.01
    mask[torch.arange(N), torch.randint(0, A, (N,), generator=g)] = True
    players = torch.randint(0, 2, (N,), generator=g).numpy().astype(np.uint8)
    side = torch.randint(0, 2, (N,), generator=g).numpy().astype(np.uint8)
    opp_ids = torch.randint(0, 2, (N,), generator=g).numpy().astype(np.int64)
    out.update(obs=_np(obs), mask=_np(mask), players=players, side=side, opp_ids=opp_ids)
    ad = get_value_adapter("multi_head", 1.5, 0.02, 0.3)
    torch.manual_seed(123)
    r = split_merge_step(obs=obs, legal_masks=mask, current_players=players, learner_model=models[0],
                         opponent_models={0: models[1], 1: models[2], 7: models[1]}, env_opponent_ids=opp_ids, learner_side=side,
                         value_adapter=ad)
    for f in ("actions", "learner_mask", "opponent_mask", "learner_log_probs", "learner_values", "learner_indices"):
        out[f"cohort/{f}"] = _np(getattr(r, f))
    torch.manual_seed(321)
    r = split_merge_step(obs=obs, legal_masks=mask, current_players=players, learner_model=models[0], opponent_model=models[2],
                         learner_side=1)
    for f in ("actions", "learner_mask", "opponent_mask", "learner_log_probs", "learner_values", "learner_indices"):
        out[f"legacy/{f}"] = _np(getattr(r, f))
    np.savez_compressed(OUT / "split_merge.npz", **out)
    print("split_merge.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    if not REF.exists():
        sys.exit("needs /root/reference (build container only)")
    sys.path.insert(0, str(REF))
    torch.set_num_threads(4)
    OUT.mkdir(parents=True, exist_ok=True)
    golden_gae()
    golden_rollout()
    golden_model()
    golden_update()
    golden_resnet()
    golden_buffer()
    golden_split_merge()
