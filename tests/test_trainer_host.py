"""Host logic of the drop-in trainer on CPU tensors: the update() composition (GAE -> advantage
normalisation -> losses -> clip -> Adam) against the reference's own update() golden run, the
buffer guards, registries and adapters. No CUDA needed."""
import numpy as np
import pytest
import torch

from conftest import load_golden, state_dict_from
from keisei_b200 import gae as G
from keisei_b200.algorithm_registry import validate_algorithm_params
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams, KataGoRolloutBuffer
from keisei_b200.model_registry import build_model, get_model_contract, get_obs_channels, validate_model_params
from keisei_b200.value_adapter import MultiHeadValueAdapter, ScalarValueAdapter, get_value_adapter

TINY = dict(num_blocks=2, channels=32, se_reduction=4, global_pool_channels=16, policy_channels=8,
            value_fc_size=16, score_fc_size=16, obs_channels=50)


def fill_buffer_from_golden(g, buf):
    T = g["steps/obs"].shape[0]
    for t in range(T):
        s = {k[len("steps/"):]: torch.from_numpy(np.array(v[t])) for k, v in g.items() if k.startswith("steps/")}
        buf.add(s["obs"], s["actions"], s["logp"], s["values"], s["rewards"], s["term"], s["term"], s["mask"], s["cats"],
                s["score_t"], next_value_override=s["ov"])


def run_update_against_golden(device):
    g = load_golden("update_tiny.npz")
    model = build_model("se_resnet", dict(TINY))
    model.load_state_dict(state_dict_from(g), strict=True)
    model = model.to(device)
    T, N = g["steps/obs"].shape[:2]
    buf = KataGoRolloutBuffer(N, (50, 9, 9), 11259)
    fill_buffer_from_golden(g, buf)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=T * N, epochs_per_batch=1, learning_rate=1e-3), model)
    metrics = algo.update(buf, torch.from_numpy(g["next_values"]).to(device))
    assert buf.size == 0 and model.training
    return g, model, metrics


def check_update(g, model, metrics, tol):
    for k in ("policy_loss", "value_loss", "score_loss", "entropy", "gradient_norm"):
        ref = float(g["metrics/" + k])
        assert abs(metrics[k] - ref) <= tol * max(abs(ref), 1e-6) + 1e-7, (k, metrics[k], ref)
    sd = model.state_dict()
    for k, v in sd.items():
        ref = g["sd_after/" + k]
        if "num_batches_tracked" in k:
            assert int(v) == int(ref)
            continue
        # Adam's first step moves every weight by ~lr * sign(grad): compare the *update*, not the weight
        before = g["sd/" + k]
        d_ref, d_got = ref - before, v.detach().cpu().numpy() - before
        scale = max(np.abs(d_ref).max(), 1e-12)
        assert np.abs(d_got - d_ref).max() <= 0.02 * scale + 1e-7, k


def test_update_matches_reference_golden_cpu():
    g, model, metrics = run_update_against_golden("cpu")
    check_update(g, model, metrics, 1e-4)


def test_buffer_guards_and_roundtrip():
    buf = KataGoRolloutBuffer(2, (50, 9, 9), 11259)
    with pytest.raises(ValueError, match="Cannot flatten an empty buffer"):
        buf.flatten()
    obs = torch.zeros(2, 50, 9, 9); a = torch.zeros(2, dtype=torch.long); z = torch.zeros(2)
    mask = torch.ones(2, 11259, dtype=torch.bool)
    with pytest.raises(AssertionError, match="terminated must be a subset of dones"):
        buf.add(obs, a, z, z, z, torch.tensor([False, False]), torch.tensor([True, False]), mask, torch.tensor([-1, -1]), z)
    with pytest.raises(ValueError, match="invalid values"):
        buf.add(obs, a, z, z, z, z.bool(), z.bool(), mask, torch.tensor([3, 0]), z)
    with pytest.raises(ValueError, match="NaN"):
        buf.add(obs, a, z, z, z, z.bool(), z.bool(), mask, torch.tensor([0, 0]), torch.tensor([float("nan"), 0.0]))
    with pytest.raises(ValueError, match="unnormalized"):
        buf.add(obs, a, z, z, z, z.bool(), z.bool(), mask, torch.tensor([0, 0]), torch.tensor([10.0, 0.0]))
    for t in range(3):
        buf.add(obs + t, a + t, z, z + t, z, z.bool(), z.bool(), mask, torch.tensor([-1, 1]), z)
    assert buf.size == 3
    d = buf.flatten()
    assert d["observations"].shape == (6, 50, 9, 9) and d["legal_masks"].shape == (6, 11259)
    assert d["values"].tolist() == [0, 0, 1, 1, 2, 2]
    buf.fill_alternating_perspective_overrides()
    ov = buf.flatten()["next_value_override"].reshape(3, 2)
    assert ov[0].tolist() == [-1.0, -1.0] and ov[1].tolist() == [-2.0, -2.0] and torch.isnan(ov[2]).all()
    buf.clear()
    assert buf.size == 0


def test_registries_and_adapters():
    assert get_model_contract("se_resnet") == "multi_head" and get_obs_channels("se_resnet") == 50
    with pytest.raises(ValueError, match="Unknown architecture"):
        build_model("nope", {})
    with pytest.raises(TypeError, match="Invalid params"):
        validate_model_params("se_resnet", {"bogus": 1})
    with pytest.raises(ValueError):
        validate_model_params("se_resnet", {"channels": 8, "se_reduction": 16})
    with pytest.raises(ValueError, match="Unknown algorithm"):
        validate_algorithm_params("ppo", {})
    assert isinstance(validate_algorithm_params("katago_ppo", {"batch_size": 64}), KataGoPPOParams)
    with pytest.raises(ValueError):
        KataGoPPOParams(batch_size=0)
    assert isinstance(get_value_adapter("scalar"), ScalarValueAdapter)
    ad = get_value_adapter("multi_head", 1.5, 0.02, 0.25)
    assert isinstance(ad, MultiHeadValueAdapter)
    with pytest.raises(ValueError, match="Unknown model contract"):
        get_value_adapter("x")
    with pytest.raises(ValueError):
        MultiHeadValueAdapter(score_blend_alpha=1.5)
    vl = torch.tensor([[2.0, 0.0, -1.0]]); sc = torch.tensor([[3.0]])
    p = torch.softmax(vl, -1)
    assert torch.allclose(ad.scalar_value_blended(vl, sc), 0.75 * (p[:, 0] - p[:, 2]) + 0.25 * 1.0)
    with pytest.raises(ValueError, match="requires value_cats"):
        ad.compute_value_loss(vl)
    loss = ad.compute_value_loss(vl.requires_grad_(), value_cats=torch.tensor([-1]), score_targets=torch.tensor([0.0]), score_pred=sc)
    loss.backward()
    assert torch.all(vl.grad == 0)  # all-ignored -> graph-connected zero


def test_entropy_schedule_and_select_actions_guards_cpu():
    model = build_model("se_resnet", dict(TINY))
    algo = KataGoPPOAlgorithm(KataGoPPOParams(entropy_decay_epochs=10, lambda_entropy=0.01), model, warmup_epochs=5, warmup_entropy=0.05)
    assert algo.get_entropy_coeff(0) == 0.05 and algo.get_entropy_coeff(15) == 0.01
    assert abs(algo.get_entropy_coeff(10) - 0.03) < 1e-12
    obs = torch.randn(3, 50, 9, 9)
    mask = torch.zeros(3, 11259, dtype=torch.bool); mask[:, 7] = True
    a, lp, v = algo.select_actions(obs, mask)
    assert a.tolist() == [7, 7, 7] and lp.abs().max() < 1e-5 and v.abs().max() <= 1.0 and model.training
    mask[1] = False
    with pytest.raises(RuntimeError, match=r"Environments \[1\] have zero legal actions"):
        algo.select_actions(obs, mask)


def test_gae_host_path_matches_golden():
    g = load_golden("gae.npz")
    t = lambda k: torch.from_numpy(g[k])
    np.testing.assert_array_equal(G.compute_gae(t("r"), t("v"), t("term"), t("nv"), 0.99, 0.95).numpy(), g["adv_plain"])
    np.testing.assert_array_equal(G.compute_gae_gpu(t("r"), t("v"), t("term"), t("nv"), 0.99, 0.95, next_value_override=t("ov")).numpy(), g["adv_override"])
    np.testing.assert_array_equal(G.compute_gae_padded(t("r"), t("v"), t("termp"), t("nv"), t("lengths"), 0.99, 0.95, next_value_override=t("ov")).numpy(), g["adv_padded_override"])
    np.testing.assert_array_equal(G.compute_gae(t("r")[:, 0], t("v")[:, 0], t("term")[:, 0], t("nv")[0], 0.99, 0.95).numpy(), g["adv_1d"])
    with pytest.raises(ValueError, match="only supports 2D"):
        G.compute_gae_gpu(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(()), 0.99, 0.95)


def test_select_actions_many_and_device_kwarg_on_cpu():
    """The CUDA-only fast paths degrade to the reference behaviour on CPU tensors: select_actions_many = one
    select_actions per sub-batch; KataGoRolloutBuffer(device="cpu") is the reference buffer; SyncBatchNorm conversion is
    accepted and ignored by the CPU path (torch's own SyncBatchNorm is GPU-only)."""
    torch.manual_seed(0)
    model = build_model("se_resnet", dict(TINY))
    algo = KataGoPPOAlgorithm(KataGoPPOParams(), model)
    batches = []
    for n in (3, 5):
        mask = torch.rand(n, 11259) < 0.01
        mask[:, 7] = True
        batches.append((torch.randn(n, 50, 9, 9), mask))
    res = algo.select_actions_many(batches)
    assert len(res) == 2
    for (a, lp, v), (o, k) in zip(res, batches):
        assert a.shape == (o.shape[0],) and k[torch.arange(o.shape[0]), a].all() and (lp <= 0).all() and v.abs().max() <= 1
    assert model.training
    with pytest.raises(ValueError, match="one model per batch"):
        algo.select_actions_many(batches, models=[model])
    buf = KataGoRolloutBuffer(2, (50, 9, 9), 11259, device="cpu")
    z = torch.zeros(2)
    buf.add(torch.zeros(2, 50, 9, 9), z.long(), z, z, z, z.bool(), z.bool(), torch.ones(2, 11259, dtype=torch.bool), z.long(), z)
    assert buf.size == 1 and buf.flatten()["observations"].device.type == "cpu"

    class FakeSync:
        world_size = 2

        def all_reduce_(self, t):  # pragma: no cover - never reached on CPU
            raise AssertionError("the CPU path must not exchange BatchNorm statistics")

    model.convert_sync_batchnorm(FakeSync())
    out = model(torch.randn(2, 50, 9, 9))
    assert out.policy_logits.shape == (2, 9, 9, 139)


def _fill_from_steps(buf, steps, ragged):
    for st in steps:
        buf.add(st["obs"], st["actions"], st["logp"], st["values"], st["rewards"], st["dones"], st["term"], st["mask"],
                st["cats"], st["score_t"], env_ids=st["ids"] if ragged else None, next_value_override=st["ov"])


@pytest.mark.parametrize("layout", ["grid", "ragged"])
def test_buffer_flatten_matches_reference_buffer_golden(layout):
    """KataGoRolloutBuffer against the REAL reference buffer (tests/golden/buffer.npz, oracle/make_golden.py:golden_buffer):
    the (T, N) grid layout after fill_alternating_perspective_overrides() and the ragged split-merge env_ids layout."""
    from oracle.make_golden import _buffer_steps
    g = load_golden("buffer.npz")
    ragged = layout == "ragged"
    steps = _buffer_steps(22 if ragged else 21, 5, 4, ragged)
    buf = KataGoRolloutBuffer(4, (50, 9, 9), 11259)
    _fill_from_steps(buf, steps, ragged)
    buf.fill_alternating_perspective_overrides()
    flat = buf.flatten()
    assert buf.size == int(g[f"{layout}/size"])
    want_keys = {k.split("/", 1)[1] for k in g if k.startswith(layout + "/")} - {"size"}
    assert {k if k not in ("observations", "legal_masks") else k + "_sum" for k in flat} == want_keys
    for k, v in flat.items():
        if k == "observations":
            assert abs(v.double().sum().item() - float(g[f"{layout}/observations_sum"])) < 1e-6
        elif k == "legal_masks":
            assert int(v.sum().item()) == int(g[f"{layout}/legal_masks_sum"])
        else:
            want = g[f"{layout}/{k}"]
            assert v.numpy().dtype == want.dtype and v.shape == want.shape, k
            np.testing.assert_array_equal(v.numpy(), want, err_msg=k)   # NaN == NaN position-wise


def test_update_on_ragged_env_ids_buffer_matches_reference_golden():
    """update() through the per-env padded GAE path (reference katago_ppo.py:649-773) on the ragged buffer: metrics and
    every parameter after the Adam step against the reference's own run."""
    from oracle.make_golden import _buffer_steps
    g = load_golden("buffer.npz")
    steps = _buffer_steps(22, 5, 4, True)
    model = build_model("se_resnet", dict(TINY))
    model.load_state_dict({k[3:]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith("sd/")}, strict=True)
    buf = KataGoRolloutBuffer(4, (50, 9, 9), 11259)
    _fill_from_steps(buf, steps, True)
    total = sum(st["ids"].numel() for st in steps)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=total, epochs_per_batch=1, learning_rate=1e-3), model)
    torch.manual_seed(5)
    metrics = algo.update(buf, torch.from_numpy(g["update/next_values"]))
    for k in ("policy_loss", "value_loss", "score_loss", "entropy", "gradient_norm"):
        want = float(g["update/metric/" + k])
        assert abs(metrics[k] - want) <= 2e-4 * max(1.0, abs(want)), (k, metrics[k], want)
    for name, v in model.state_dict().items():
        if v.is_floating_point():
            want = g["update/after/" + name]
            assert np.abs(v.numpy() - want).max() <= 2e-4 * max(1e-2, np.abs(want).max()), name


def test_model_deepcopy_pickle_and_weight_cache_invalidation():
    """Opponent snapshots deep-copy the learner (league / tournament code); captured rollout graphs and their locks must
    not travel. The packed-weight cache is invalidated explicitly by load_state_dict and by the trainer's optimiser tail
    (Adam(fused=True) does not bump Tensor._version)."""
    import copy
    import pickle
    m = build_model("se_resnet", dict(TINY))
    m2 = copy.deepcopy(m)
    assert m2._graphs is not m._graphs and len(m2._graphs) == 0
    m3 = pickle.loads(pickle.dumps(m))
    assert list(m3.state_dict()) == list(m.state_dict())
    m._wpack_key = ("stale",)
    m.load_state_dict(m2.state_dict())
    assert m._wpack_key is None
    m._wpack_key = ("stale",)
    m.invalidate_packed_weights()
    assert m._wpack_key is None


def test_sm_shares_of_a_grouped_rollout_graph():
    """Every branch of the grouped league graph gets whole CTA pairs in proportion to its boards, at least one pair, and
    together they use the device once (keisei_b200/models/se_resnet.py: _sm_shares)."""
    from keisei_b200.models.se_resnet import _sm_shares
    assert _sm_shares([512], 148) == [148]
    sh = _sm_shares([256, 64, 64, 64, 64], 148)
    assert sum(sh) == 148 and all(x % 2 == 0 and x >= 2 for x in sh) and sh[0] == 74 and max(sh[1:]) - min(sh[1:]) <= 2
    assert _sm_shares([64, 64], 148) == [74, 74]
    sh = _sm_shares([3000, 8, 8], 148)
    assert sum(sh) <= 150 and sh[1] == sh[2] == 2
    sh = _sm_shares([8] * 100, 148)           # more branches than CTA pairs: one pair each, the graph serialises the rest
    assert sh == [2] * 100
