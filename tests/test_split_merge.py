"""`keisei_b200.split_merge.split_merge_step` (reference katago_loop.py:284-431).

CPU: bit-for-bit against a golden run of the REFERENCE function with three real tiny SE-ResNets (cohort mode with a
per-env learner side + blending value adapter, legacy single-opponent mode; `oracle/make_golden.py: golden_split_merge`),
plus the guards the reference's own tests pin (tests/test_split_merge.py). GPU: the grouped CUDA-graph path on the same
golden inputs, every learner log-prob / value checked against the CPU oracle for the actions the kernel drew."""
import numpy as np
import pytest
import torch

from conftest import load_golden, state_dict_from
from keisei_b200.model_registry import build_model
from keisei_b200.split_merge import SplitMergeResult, split_merge_step
from keisei_b200.value_adapter import get_value_adapter

TINY = dict(num_blocks=2, channels=32, se_reduction=4, global_pool_channels=16, policy_channels=8,
            value_fc_size=16, score_fc_size=16, obs_channels=50)
FIELDS = ("actions", "learner_mask", "opponent_mask", "learner_log_probs", "learner_values", "learner_indices")


def golden_setup(device="cpu"):
    g = load_golden("split_merge.npz")
    models = []
    for k in range(3):
        m = build_model("se_resnet", dict(TINY))
        m.load_state_dict(state_dict_from(g, f"sd{k}/"), strict=True)
        models.append(m.to(device).eval())
    obs, mask = torch.from_numpy(g["obs"]).to(device), torch.from_numpy(g["mask"]).to(device)
    return g, models, obs, mask


def run_cohort(g, models, obs, mask, **kw):
    return split_merge_step(obs=obs, legal_masks=mask, current_players=g["players"], learner_model=models[0],
                            opponent_models={0: models[1], 1: models[2], 7: models[1]}, env_opponent_ids=g["opp_ids"],
                            learner_side=g["side"], value_adapter=get_value_adapter("multi_head", 1.5, 0.02, 0.3), **kw)


def run_legacy(g, models, obs, mask, **kw):
    return split_merge_step(obs=obs, legal_masks=mask, current_players=g["players"], learner_model=models[0],
                            opponent_model=models[2], learner_side=1, **kw)


def check_against_golden(r, g, prefix):
    assert isinstance(r, SplitMergeResult) or type(r).__name__ == "SplitMergeResult"
    for f in ("actions", "learner_mask", "opponent_mask", "learner_indices"):
        assert np.array_equal(getattr(r, f).cpu().numpy(), g[f"{prefix}/{f}"]), f
    assert r.actions.dtype == torch.int64 and r.learner_mask.dtype == torch.bool and r.learner_indices.dtype == torch.int64
    for f in ("learner_log_probs", "learner_values"):
        np.testing.assert_allclose(getattr(r, f).cpu().numpy(), g[f"{prefix}/{f}"], rtol=1e-4, atol=1e-5, err_msg=f)


def test_cohort_mode_matches_reference_golden_cpu():
    g, models, obs, mask = golden_setup()
    torch.manual_seed(123)
    check_against_golden(run_cohort(g, models, obs, mask), g, "cohort")
    assert not models[0].training        # the learner is left in eval mode (katago_loop.py:330-333)


def test_legacy_mode_matches_reference_golden_cpu():
    g, models, obs, mask = golden_setup()
    torch.manual_seed(321)
    check_against_golden(run_legacy(g, models, obs, mask), g, "legacy")


def test_guards_and_degenerate_partitions_cpu():
    g, models, obs, mask = golden_setup()
    N = obs.shape[0]
    with pytest.raises(ValueError, match="Must provide either opponent_model or opponent_models"):
        split_merge_step(obs=obs, legal_masks=mask, current_players=g["players"], learner_model=models[0])
    # all learner / all opponent
    r = split_merge_step(obs=obs, legal_masks=mask, current_players=np.zeros(N, np.uint8), learner_model=models[0], opponent_model=models[1])
    assert bool(r.learner_mask.all()) and r.learner_log_probs.shape == (N,) and np.array_equal(r.learner_indices.numpy(), np.arange(N))
    r = split_merge_step(obs=obs, legal_masks=mask, current_players=np.ones(N, np.uint8), learner_model=models[0], opponent_model=models[1])
    assert bool(r.opponent_mask.all()) and r.learner_log_probs.shape == (0,) and r.learner_values.shape == (0,)
    assert bool(mask[torch.arange(N), r.actions].all())
    # an opponent id nobody plays against is skipped; envs of an id without a model keep action 0
    players = np.array([0, 1] * (N // 2), np.uint8)
    ids = np.where(np.arange(N) % 4 == 1, 5, 0).astype(np.int64)
    calls = []

    class Spy(torch.nn.Module):
        def forward(self, x):
            calls.append(x.shape[0])
            return models[1](x)

    r = split_merge_step(obs=obs, legal_masks=mask, current_players=players, learner_model=models[0],
                         opponent_models={0: models[1], 9: Spy()}, env_opponent_ids=ids)
    assert calls == [] and bool((r.actions[torch.from_numpy((ids == 5) & (players == 1))] == 0).all())
    # zero legal actions: the error names the ENVIRONMENT ids of the offending rows, learner and opponent alike
    bad = mask.clone()
    bad[4] = False
    with pytest.raises(RuntimeError, match=r"Learner envs \[4\] have zero legal actions"):
        split_merge_step(obs=obs, legal_masks=bad, current_players=players, learner_model=models[0], opponent_model=models[1])
    bad = mask.clone()
    bad[7] = False
    with pytest.raises(RuntimeError, match=r"Opponent envs \[7\] have zero legal actions"):
        split_merge_step(obs=obs, legal_masks=bad, current_players=players, learner_model=models[0], opponent_model=models[1])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["cohort", "legacy"])
def test_grouped_graph_path_against_oracle_gpu(mode):
    from oracle import keisei_oracle as O
    dev = torch.device("cuda:0")
    g, models, obs, mask = golden_setup(dev)
    sds = [state_dict_from(g, f"sd{k}/") for k in range(3)]
    run = run_cohort if mode == "cohort" else run_legacy
    ad = get_value_adapter("multi_head", 1.5, 0.02, 0.3) if mode == "cohort" else None
    replayed0 = models[0].graph_replayed_kernels
    seen = set()
    for it in range(4):          # second sighting of the bucket signature captures the grouped graph, later calls replay it
        r = run(g, models, obs, mask, seed=100 + it)
        assert np.array_equal(r.learner_mask.cpu().numpy(), g[f"{mode}/learner_mask"])
        assert np.array_equal(r.learner_indices.cpu().numpy(), g[f"{mode}/learner_indices"])
        a = r.actions.cpu()
        covered = torch.from_numpy(g[f"{mode}/learner_mask"]) | torch.from_numpy(np.isin(g["opp_ids"], [0, 1]) if mode == "cohort"
                                                                                  else np.ones(len(a), bool))
        assert bool(mask.cpu()[torch.arange(len(a)), a][covered].all()), "illegal action"
        assert bool((a[~covered] == 0).all())
        li = r.learner_indices.cpu()
        with torch.no_grad():
            p_ref, v_ref, s_ref = O.seresnet_forward(sds[0], obs.cpu()[li], TINY["num_blocks"], training=False)
        lp_ref = O.rollout_log_prob(p_ref.reshape(len(li), -1), mask.cpu()[li], a[li])
        assert torch.allclose(r.learner_log_probs.cpu(), lp_ref, rtol=1e-3, atol=1e-4)
        val_ref = ad.scalar_value_blended(v_ref, s_ref) if ad is not None else O.scalar_value(v_ref)
        assert torch.allclose(r.learner_values.cpu(), val_ref, rtol=1e-3, atol=1e-4)
        seen.add(tuple(a.tolist()))
    assert models[0].graph_replayed_kernels > replayed0, "the grouped CUDA graph never replayed"
    assert len(seen) > 1, "different seeds drew identical actions for every env"


@pytest.mark.gpu
def test_zero_legal_guard_gpu():
    dev = torch.device("cuda:0")
    g, models, obs, mask = golden_setup(dev)
    players = np.array([0, 1] * (obs.shape[0] // 2), np.uint8)
    bad = mask.clone()
    bad[7] = False
    with pytest.raises(RuntimeError, match=r"Opponent envs \[7\] have zero legal actions"):
        split_merge_step(obs=obs, legal_masks=bad, current_players=players, learner_model=models[0], opponent_model=models[1])
    bad = mask.clone()
    bad[4] = False
    with pytest.raises(RuntimeError, match=r"Learner envs \[4\] have zero legal actions"):
        split_merge_step(obs=obs, legal_masks=bad, current_players=players, learner_model=models[0], opponent_model=models[1])
