"""Plain ResNet baseline (BASELINE.json configs[3]) on the CPU: oracle pinned against the reference's golden
vectors, the host mirror of ResNetModel / registry, and the scalar-PPO trainer's host logic."""
import numpy as np
import pytest
import torch

from conftest import load_golden, state_dict_from
from oracle import keisei_oracle as O
from keisei_b200.algorithm_registry import PPOParams
from keisei_b200.katago_ppo import KataGoRolloutBuffer
from keisei_b200.model_registry import build_model, get_model_contract, get_obs_channels, validate_model_params
from keisei_b200.models import ResNetModel, ResNetParams
from keisei_b200.ppo import PPOAlgorithm
from keisei_b200.value_adapter import ScalarValueAdapter, get_value_adapter


def test_resnet_oracle_vs_reference_golden_forward_and_stats():
    g = load_golden("resnet_tiny.npz")
    sd = state_dict_from(g)
    obs = torch.from_numpy(g["obs"])
    with torch.no_grad():
        p, v = O.resnet_forward(sd, obs, 2, training=False)
    np.testing.assert_allclose(p.numpy(), g["eval_policy"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(v.numpy(), g["eval_value"], rtol=1e-5, atol=1e-6)
    new_stats = {}
    with torch.no_grad():
        p, v = O.resnet_forward(sd, obs, 2, training=True, new_stats=new_stats)
    np.testing.assert_allclose(p.numpy(), g["train_policy"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(v.numpy(), g["train_value"], rtol=1e-4, atol=1e-5)
    for k, val in new_stats.items():
        np.testing.assert_allclose(val.numpy(), g["sd_after/" + k], rtol=1e-5, atol=1e-6)


def test_resnet_oracle_loss_and_gradients_vs_reference_golden():
    g = load_golden("resnet_tiny.npz")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in state_dict_from(g).items()}
    p, v = O.resnet_forward(sd, torch.from_numpy(g["obs"]), 2, training=True)
    out = O.scalar_ppo_losses(p, v, torch.from_numpy(g["mask"]), torch.from_numpy(g["actions"]), torch.from_numpy(g["old_logp"]),
                              torch.from_numpy(g["adv"]), torch.from_numpy(g["returns"]))
    for k in ("loss", "policy_loss", "value_loss", "entropy"):
        np.testing.assert_allclose(out[k].item(), g[k], rtol=1e-5, atol=1e-6)
    out["loss"].backward()
    for k, t in sd.items():
        if t.requires_grad:
            ref = g["grad/" + k]
            assert np.abs(t.grad.numpy() - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-6) + 1e-7, k


def test_resnet_host_model_matches_golden_and_state_dict_layout():
    g = load_golden("resnet_tiny.npz")
    m = ResNetModel(ResNetParams(hidden_size=32, num_layers=2))
    assert list(m.state_dict().keys()) == [k[3:] for k in g if k.startswith("sd/")]
    m.load_state_dict(state_dict_from(g), strict=True)
    m.eval()
    with torch.no_grad():
        p, v = m(torch.from_numpy(g["obs"]))
    np.testing.assert_allclose(p.numpy(), g["eval_policy"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(v.numpy(), g["eval_value"], rtol=1e-5, atol=1e-6)
    assert p.shape == (6, 11259) and v.shape == (6, 1)


def test_resnet_registry_and_validation_errors():
    # reference model_registry.py:25,72-76; tests/test_models.py:28-80
    assert get_model_contract("resnet") == "scalar" and get_obs_channels("resnet") == 50
    assert isinstance(build_model("resnet", {"hidden_size": 16, "num_layers": 1}), ResNetModel)
    with pytest.raises(ValueError):
        validate_model_params("resnet", {"hidden_size": 0, "num_layers": 1})
    with pytest.raises(ValueError):
        validate_model_params("resnet", {"hidden_size": 8, "num_layers": -1})
    with pytest.raises(TypeError, match="Invalid params"):
        validate_model_params("resnet", {"hidden": 8})
    m = build_model("resnet", {"hidden_size": 8, "num_layers": 0})
    with pytest.raises(ValueError, match="Expected obs shape"):
        m(torch.zeros(2, 46, 9, 9))
    with pytest.raises(ValueError, match="NHWC"):
        m(torch.zeros(2, 9, 9, 50))
    assert isinstance(get_value_adapter("scalar"), ScalarValueAdapter)


def test_scalar_ppo_update_on_cpu_matches_manual_step():
    """One minibatch covering the whole buffer: the trainer's step must equal a hand-rolled step with the oracle
    loss (GAE -> returns, normalised advantages, clip 1.0, Adam)."""
    torch.manual_seed(0)
    N, T, A = 4, 3, 11259
    model = build_model("resnet", dict(hidden_size=16, num_layers=1))
    twin = build_model("resnet", dict(hidden_size=16, num_layers=1))
    twin.load_state_dict(model.state_dict())
    algo = PPOAlgorithm(PPOParams(batch_size=N * T, epochs_per_batch=1), model)
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
    g = torch.Generator().manual_seed(1)
    rows = []
    for t in range(T):
        obs = torch.randn(N, 50, 9, 9, generator=g)
        mask = torch.rand(N, A, generator=g) < 0.01
        mask[:, 7] = True
        a, lp, v = algo.select_actions(obs, mask)
        assert mask[torch.arange(N), a].all() and v.shape == (N,)
        term = torch.tensor([t == T - 1] * N)
        rew = torch.randn(N, generator=g)
        buf.add(obs, a, lp, v, rew, term, term, mask, torch.full((N,), -1), torch.zeros(N))
        rows.append((obs, mask, a, lp, v, rew, term))
    nv = torch.zeros(N)
    # manual reference step on the twin
    obs = torch.cat([r[0] for r in rows]); mask = torch.cat([r[1] for r in rows]); acts = torch.cat([r[2] for r in rows])
    old = torch.cat([r[3] for r in rows]); vals = torch.stack([r[4] for r in rows]); rew = torch.stack([r[5] for r in rows])
    term = torch.stack([r[6] for r in rows])
    adv = torch.from_numpy(O.gae_numpy(rew.numpy(), vals.numpy(), term.numpy(), nv.numpy(), 0.99, 0.95))
    returns = (adv + vals).reshape(-1)
    advn = torch.from_numpy(O.normalize_advantages(adv.reshape(-1).numpy()))
    twin.train()
    opt = torch.optim.Adam(twin.parameters(), lr=3e-4)
    logits, value = twin(obs)
    out = O.scalar_ppo_losses(logits, value, mask, acts, old, advn, returns)
    out["loss"].backward()
    torch.nn.utils.clip_grad_norm_(twin.parameters(), 1.0)
    opt.step()
    # trainer step (single minibatch: the permutation does not change a full-batch mean; BN statistics neither)
    metrics = algo.update(buf, nv)
    np.testing.assert_allclose(metrics["policy_loss"], out["policy_loss"].item(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(metrics["value_loss"], out["value_loss"].item(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(metrics["entropy"], out["entropy"].item(), rtol=1e-5)
    assert metrics["score_loss"] == 0.0 and buf.size == 0
    for (n, p), (_, q) in zip(model.named_parameters(), twin.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-3, atol=2e-5), n
