"""CUDA-graph replay of the rollout forward (SEResNetModel.rollout_forward): bit-identical to the ordinary
launch sequence, follows parameter updates, and is what select_actions uses for small batches."""
import pytest
import torch

from keisei_b200 import _lib
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams
from keisei_b200.models import SEResNetModel, SEResNetParams

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(num_blocks=2, channels=128, se_reduction=8, global_pool_channels=32, policy_channels=16, value_fc_size=32,
           score_fc_size=32)


@pytest.mark.parametrize("amp", [False, True])
def test_graph_replay_matches_plain_forward_and_tracks_weight_updates(amp):
    torch.manual_seed(0)
    m = SEResNetModel(SEResNetParams(**CFG)).to(DEV).eval()
    if amp:
        m.configure_amp(True, torch.bfloat16, "cuda")
    obs = [torch.randn(24, 50, 9, 9, device=DEV) for _ in range(3)]
    with torch.no_grad():
        want = [m(o) for o in obs]
        want = [(w.policy_logits.clone(), w.value_logits.clone(), w.score_lead.clone()) for w in want]
    n_graphs = 0
    for rep in range(2):
        for o, w in zip(obs, want):
            got = m.rollout_forward(o)
            assert torch.equal(got.policy_logits, w[0]) and torch.equal(got.value_logits, w[1]) and torch.equal(got.score_lead, w[2])
        n_graphs = len(m._graphs)
    assert n_graphs == 1
    # an optimiser step changes the weights: the replay must see the re-packed weights
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.01 * torch.randn_like(p))
        w2 = m(obs[0])
        w2 = (w2.policy_logits.clone(), w2.value_logits.clone())
    got = m.rollout_forward(obs[0])
    assert torch.equal(got.policy_logits, w2[0]) and torch.equal(got.value_logits, w2[1])
    assert not torch.equal(got.policy_logits, want[0][0])
    # training mode and oversize batches fall through to the ordinary path
    m.graph_max_batch = 8
    n_before = len(m._graphs)
    with torch.no_grad():
        w3 = m(obs[1]).value_logits.clone()
    assert torch.equal(m.rollout_forward(obs[1]).value_logits, w3)
    assert len(m._graphs) == n_before   # 24 boards > graph_max_batch: no capture, the plain launch sequence ran


def test_select_actions_uses_graph_and_stays_correct():
    torch.manual_seed(1)
    m = SEResNetModel(SEResNetParams(**CFG)).to(DEV)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), m)
    algo._sample_seed = 123
    obs = torch.randn(16, 50, 9, 9, device=DEV)
    mask = torch.rand(16, 11259, device=DEV) < 0.01
    mask[:, 5] = True
    a1, lp1, v1 = algo.select_actions(obs, mask)
    assert len(m._graphs) == 0          # a bucket is captured on its second sighting
    a1, lp1, v1 = algo.select_actions(obs, mask)
    assert len(m._graphs) == 1
    n0 = _lib.launch_count()
    a2, lp2, v2 = algo.select_actions(obs, mask)
    # the replay launches the network from the graph: only the sampling kernel goes through the library counter
    assert _lib.launch_count() - n0 <= 2
    assert mask[torch.arange(16, device=DEV), a2].all()
    assert torch.equal(v1, v2)
    m.graph_max_batch = 0
    a3, lp3, v3 = algo.select_actions(obs, mask)
    assert torch.equal(v3, v1)


@pytest.mark.parametrize("B", [7, 300, 601])
def test_two_branch_split_rollout_is_bit_identical(B):
    """Graph-replayed batches >= rollout_split_min run as two half batches on two captured branches
    (SEResNetModel._captured_forward): boards are independent in eval mode, so the result is bit-identical."""
    torch.manual_seed(2)
    m = SEResNetModel(SEResNetParams(**CFG)).to(DEV).eval()
    m.configure_amp(True, torch.bfloat16, "cuda")
    obs = torch.randn(B, 50, 9, 9, device=DEV)
    m.rollout_split_min = 0
    with torch.no_grad():
        w = m.rollout_forward(obs)
        want = (w.policy_logits.clone(), w.value_logits.clone(), w.score_lead.clone())
    m._graphs.clear()
    m.rollout_split_min = 2
    for _ in range(3):
        got = m.rollout_forward(obs)
        assert torch.equal(got.policy_logits, want[0]) and torch.equal(got.value_logits, want[1]) and torch.equal(got.score_lead, want[2])
    assert m.last_policy_buffer[:, 11259:].abs().sum().item() == 0


def test_grouped_rollout_many_models_is_bit_identical_and_select_actions_many():
    """rollout_forward_many / select_actions_many: learner + opponents' sub-batches as parallel branches of one graph."""
    from keisei_b200.models import rollout_forward_many
    torch.manual_seed(3)
    ms = [SEResNetModel(SEResNetParams(**CFG)).to(DEV).eval() for _ in range(3)]
    for m in ms:
        m.configure_amp(True, torch.bfloat16, "cuda")
    sizes = [40, 9, 17, 9]
    models = [ms[0], ms[1], ms[2], ms[1]]                      # one model appears twice
    obs = [torch.randn(b, 50, 9, 9, device=DEV) for b in sizes]
    with torch.no_grad():
        want = []
        for m, o in zip(models, obs):
            w = m(o)
            want.append((w.policy_logits.clone(), w.value_logits.clone(), w.score_lead.clone()))
    for _ in range(3):
        got = rollout_forward_many(list(zip(models, obs)))
        for g, w in zip(got, want):
            assert torch.equal(g.policy_logits, w[0]) and torch.equal(g.value_logits, w[1]) and torch.equal(g.score_lead, w[2])
    # weights change -> the replay follows (in-place re-pack)
    with torch.no_grad():
        for p in ms[1].parameters():
            p.add_(0.01 * torch.randn_like(p))
        w1 = ms[1](obs[1]).policy_logits.clone()
    got = rollout_forward_many(list(zip(models, obs)))
    assert torch.equal(got[1].policy_logits, w1) and torch.equal(got[0].policy_logits, want[0][0])
    # trainer-level call
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), ms[0])
    algo._sample_seed = 7
    masks = [torch.rand(b, 11259, device=DEV) < 0.01 for b in sizes]
    for k in masks:
        k[:, 5] = True
    res = algo.select_actions_many(list(zip(obs, masks)), models=models)
    assert len(res) == 4
    for (a, lp, v), k, o in zip(res, masks, obs):
        assert a.shape == (o.shape[0],) and k[torch.arange(o.shape[0], device=DEV), a].all() and (lp <= 0).all() and v.abs().max() <= 1
    single = algo.select_actions(obs[0], masks[0])
    assert torch.equal(single[2], res[0][2])                  # same network output -> same scalar values
    assert ms[0].training                                      # forward model left in train mode, as select_actions does
    bad = masks[2].clone(); bad[3] = False
    with pytest.raises(RuntimeError, match=r"Environments \[3\] have zero legal actions"):
        algo.select_actions_many([(obs[0], masks[0]), (obs[2], bad)], models=[ms[0], ms[2]])


def test_bucketed_batches_share_one_graph_and_stay_bit_identical():
    """Sub-batch sizes vary from step to step in a split-merge rollout (katago_loop.py:337-344): sizes of one bucket replay
    ONE captured graph (padded static input, outputs sliced back), with results bit-identical to the plain forward."""
    from keisei_b200.models.se_resnet import rollout_bucket
    assert [rollout_bucket(b) for b in (1, 8, 9, 64, 65, 96, 512, 513, 2048, 2049, 4096)] == [8, 8, 16, 64, 96, 96, 512, 640, 2048, 2304, 4096]
    torch.manual_seed(4)
    m = SEResNetModel(SEResNetParams(**CFG)).to(DEV).eval()
    m.configure_amp(True, torch.bfloat16, "cuda")
    sizes = [70, 65, 96, 81, 70, 90]                      # all in bucket 96
    for i, b in enumerate(sizes):
        obs = torch.randn(b, 50, 9, 9, device=DEV)
        with torch.no_grad():
            w = m(obs)
            want = (w.policy_logits.clone(), w.value_logits.clone(), w.score_lead.clone())
        n0 = _lib.launch_count()
        got = m.rollout_forward(obs)
        assert got.policy_logits.shape == (b, 9, 9, 139) and got.value_logits.shape == (b, 3)
        assert torch.equal(got.policy_logits, want[0]) and torch.equal(got.value_logits, want[1]) and torch.equal(got.score_lead, want[2])
        if i >= 2:
            assert _lib.launch_count() == n0            # replayed: no direct launch through the library
    assert len(m._graphs) == 1 and m._graphs.captures == 1
    # LRU bounded by bytes: a tiny budget keeps only the most recent graph
    m._graphs.max_bytes = 1
    for b in (8, 8, 16, 16, 24, 24):
        m.rollout_forward(torch.randn(b, 50, 9, 9, device=DEV))
    assert len(m._graphs) == 1


def test_graph_buffers_are_per_thread():
    """A tournament thread / DynamicTrainer may run the same model concurrently (dynamic_trainer.py:44-50): its replay
    must not overwrite the static output buffers the training thread is still reading."""
    import threading
    torch.manual_seed(5)
    m = SEResNetModel(SEResNetParams(**CFG)).to(DEV).eval()
    a, b = torch.randn(16, 50, 9, 9, device=DEV), torch.randn(16, 50, 9, 9, device=DEV)
    with torch.no_grad():
        want_a, want_b = m(a).value_logits.clone(), m(b).value_logits.clone()
    m.rollout_forward(a)
    mine = m.rollout_forward(a)                           # replayed: views of this thread's static buffers
    assert torch.equal(mine.value_logits, want_a)
    other = {}

    def worker():
        with torch.cuda.stream(torch.cuda.Stream(DEV)):
            m.rollout_forward(b)
            other["out"] = m.rollout_forward(b).value_logits.clone()
            torch.cuda.current_stream().synchronize()

    t = threading.Thread(target=worker)
    t.start(); t.join(120)
    assert torch.equal(other["out"], want_b)
    assert torch.equal(mine.value_logits, want_a)        # untouched by the other thread's replay
    assert len(m._graphs) == 2


def test_select_actions_never_flips_the_training_flags_on_the_kernel_path(monkeypatch):
    """`select_actions` evaluates in inference mode by passing the mode to the kernels (`rollout_forward(eval_mode=True)`),
    not by walking the module tree twice per step; the model is in train mode afterwards like in the reference
    (katago_ppo.py:575-617), and the result is what an eval-mode forward gives — also for the grouped league step."""
    torch.manual_seed(3)
    m = SEResNetModel(SEResNetParams(**CFG)).to(DEV)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(use_amp=True), m)
    A = 11259
    obs = torch.randn(12, 50, 9, 9, device=DEV)
    mask = torch.rand(12, A, device=DEV) < 0.01
    mask[:, 5] = True
    m.eval()
    with torch.no_grad():
        want = m.rollout_forward(obs).policy_logits.clone()
    m.train()
    calls = {"eval": 0, "train": 0}
    real_eval, real_train = m.eval, m.train
    monkeypatch.setattr(m, "eval", lambda: (calls.__setitem__("eval", calls["eval"] + 1), real_eval())[1])
    monkeypatch.setattr(m, "train", lambda mode=True: (calls.__setitem__("train", calls["train"] + 1), real_train(mode))[1])
    for _ in range(3):   # plain launch, capture, replay
        a, lp, v = algo.select_actions(obs, mask)
        assert m.training and mask[torch.arange(12, device=DEV), a].all()
        assert torch.equal(m.last_policy_buffer[:, :A].reshape(want.shape), want)
    res = algo.select_actions_many([(obs[:8], mask[:8]), (obs[8:], mask[8:])])
    assert m.training and len(res) == 2
    assert calls == {"eval": 0, "train": 0}
    # a model that was left in eval mode comes back in train mode (the reference's `finally: train()`)
    real_eval()                                   # (Module.eval() is train(False): counted by the patched train)
    calls["train"] = 0
    algo.select_actions(obs, mask)
    assert m.training and calls["eval"] == 0 and calls["train"] == 1
    # training-mode `rollout_forward` without eval_mode is still the ordinary training forward (batch statistics)
    monkeypatch.undo()
    m.train()
    with torch.no_grad():
        tr = m.rollout_forward(obs).policy_logits
    assert not torch.equal(tr, want)
