"""Pin the CPU oracle (oracle/keisei_oracle.py) against vectors produced by the REAL reference
(oracle/make_golden.py) and against the reference's own hand-computed known answers."""
import numpy as np
import pytest
import torch

from oracle import keisei_oracle as O
from conftest import load_golden, state_dict_from


def test_gae_known_answers_from_reference_tests():
    # reference tests/test_gae.py:10-40: single step, gamma=.99 lam=.95 -> 1 + .99*.5 - .5 ... chain 2.5 / 4.34625 / 5.081
    g = load_golden("gae.npz")
    adv = O.gae_numpy(np.ones((3, 1)), np.full((3, 1), 0.5), np.zeros((3, 1)), np.array([0.5]), 0.99, 0.95)
    np.testing.assert_array_equal(adv[:, 0], g["known_adv_3step"])
    # terminal step: advantage = r - v (reference tests/test_gae.py:42-60 semantics)
    adv = O.gae_numpy(np.array([[1.0]]), np.array([[0.25]]), np.array([[1.0]]), np.array([9.0]), 0.99, 0.95)
    assert adv[0, 0] == np.float32(0.75)


@pytest.mark.parametrize("key,kw", [
    ("adv_plain", {}), ("adv_plain_gpufn", {}),
    ("adv_override", {"ov": True}), ("adv_override_gpufn", {"ov": True}),
])
def test_gae_oracle_bit_exact_vs_reference(key, kw):
    g = load_golden("gae.npz")
    adv = O.gae_numpy(g["r"], g["v"], g["term"], g["nv"], 0.99, 0.95, override=g["ov"] if kw.get("ov") else None)
    np.testing.assert_array_equal(adv, g[key])


def test_gae_oracle_variants():
    g = load_golden("gae.npz")
    np.testing.assert_array_equal(O.gae_numpy(g["r"], g["v"], g["term"].astype(np.float32), g["nv"], 0.97, 0.9), g["adv_termfloat"])
    np.testing.assert_array_equal(O.gae_numpy(g["r"][:, :1], g["v"][:, :1], g["term"][:, :1], g["nv"][:1], 0.99, 0.95)[:, 0], g["adv_1d"])
    for key, ov in (("adv_padded", None), ("adv_padded_gpufn", None), ("adv_padded_override", g["ov"])):
        adv = O.gae_numpy(g["r"], g["v"], g["termp"], g["nv"], 0.99, 0.95, override=ov, lengths=g["lengths"])
        np.testing.assert_array_equal(adv, g[key])
    np.testing.assert_allclose(O.normalize_advantages(g["adv_override"].reshape(-1)), g["adv_override_normalized"], atol=1e-6)


def test_model_oracle_vs_reference_eval_and_train():
    g = load_golden("seresnet_tiny.npz")
    sd = state_dict_from(g)
    obs = torch.from_numpy(g["obs"])
    with torch.no_grad():
        p, v, s = O.seresnet_forward(sd, obs, 2, training=False)
    np.testing.assert_allclose(p.numpy(), g["eval_policy"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(v.numpy(), g["eval_value"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(s.numpy(), g["eval_score"], rtol=1e-5, atol=1e-6)
    new_stats = {}
    with torch.no_grad():
        p, v, s = O.seresnet_forward(sd, obs, 2, training=True, new_stats=new_stats)
    np.testing.assert_allclose(p.numpy(), g["train_policy"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(v.numpy(), g["train_value"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(s.numpy(), g["train_score"], rtol=1e-4, atol=1e-5)
    for k, val in new_stats.items():
        np.testing.assert_allclose(val.numpy(), g["sd_after/" + k], rtol=1e-5, atol=1e-6)


def test_loss_and_gradient_oracle_vs_reference():
    g = load_golden("seresnet_tiny.npz")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in state_dict_from(g).items()}
    obs = torch.from_numpy(g["obs"])
    p, v, s = O.seresnet_forward(sd, obs, 2, training=True)
    out = O.ppo_losses(p, v, s, torch.from_numpy(g["mask"]), torch.from_numpy(g["actions"]),
                       torch.from_numpy(g["old_logp"]), torch.from_numpy(g["adv"]),
                       torch.from_numpy(g["cats"]), torch.from_numpy(g["score_t"]))
    for k in ("loss", "policy_loss", "value_loss", "score_loss", "entropy"):
        np.testing.assert_allclose(out[k].item(), g[k], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["new_log_probs"].detach().numpy(), g["new_logp"], rtol=1e-5, atol=1e-5)
    out["loss"].backward()
    for k, t in sd.items():
        if t.requires_grad:
            ref = g["grad/" + k]
            scale = max(np.abs(ref).max(), 1e-6)
            assert np.abs(t.grad.numpy() - ref).max() <= 1e-4 * scale + 1e-7, k


def test_rollout_logprob_and_value_oracle_vs_reference():
    g = load_golden("rollout.npz")
    logits, mask = torch.from_numpy(g["logits"]), torch.from_numpy(g["mask"])
    a = torch.from_numpy(g["actions"])
    np.testing.assert_allclose(O.rollout_log_prob(logits, mask, a).numpy(), g["logp_f32"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(O.rollout_log_prob(logits.bfloat16(), mask, a).float().numpy(), g["logp_bf16"])
    # finding 5: bf16 clamp range [-4.852, -0.00784]
    assert g["logp_bf16"].max() <= -0.0078 and g["logp_worst_bf16"].min() >= -4.86
    assert abs(g["logp_f32"][0]) < 1e-5  # single legal action -> log-prob ~ 0 (tests/test_katago_ppo.py:297-306)
    vl, sc = torch.from_numpy(g["value_logits"]), torch.from_numpy(g["score_lead"])
    np.testing.assert_allclose(O.scalar_value(vl).numpy(), g["scalar_value"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(O.scalar_value(vl, sc, 0.3).numpy(), g["scalar_value_blend03"], rtol=1e-6, atol=1e-7)


def test_action_indexing_matches_nhwc_reshape():
    # flat = (row*9+col)*139 + move_type is exactly reshape(B,-1) of a (B,9,9,139) tensor
    t = torch.arange(9 * 9 * 139).reshape(1, 9, 9, 139)
    for (r, c, m) in [(0, 0, 0), (8, 8, 138), (3, 5, 77)]:
        assert t[0, r, c, m].item() == O.action_index(r, c, m)
    assert O.ACTION_SPACE == 11259


def test_amax_tie_and_std_zero_semantics():
    x = torch.zeros(1, 1, 3, 3, requires_grad=True)
    x.amax(dim=(-2, -1)).sum().backward()
    np.testing.assert_allclose(x.grad.numpy().reshape(-1), O.amax_tie_grad(np.zeros(9)), atol=1e-7)
    y = torch.full((1, 1, 3, 3), 2.0, requires_grad=True)
    y.std(dim=(-2, -1), correction=0).sum().backward()
    assert torch.all(y.grad == 0)
    # reference tests/test_se_resnet.py:173-219 known values: mean 2.5, max 4.0, std sqrt(1.25)
    z = torch.tensor([1.0, 2.0, 3.0, 4.0]).reshape(1, 1, 2, 2)
    gp = O.global_pool(z)
    np.testing.assert_allclose(gp.numpy().reshape(-1), [2.5, 4.0, np.sqrt(1.25)], rtol=1e-6)
