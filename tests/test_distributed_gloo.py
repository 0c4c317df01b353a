"""N>1 path on CPU: world_size 2 over gloo (reference strategy: tests/integration/test_ddp_training.py).
Each rank runs the drop-in trainer's update() on its own rollout; gradients are averaged by
GradSync; non-BatchNorm-buffer weights must end identical across ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

TINY = dict(num_blocks=1, channels=16, se_reduction=4, global_pool_channels=8, policy_channels=8,
            value_fc_size=8, score_fc_size=8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    from keisei_b200.distributed import BatchNormSync, GradSync, init_from_env, shutdown
    from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams, KataGoRolloutBuffer
    from keisei_b200.model_registry import build_model
    torch.set_num_threads(1)
    env = init_from_env(backend="gloo")
    assert env.launched and env.world_size == world and env.rank == rank and env.local_rank == rank
    torch.manual_seed(100 + rank)            # different init + different rollouts per rank
    model = build_model("se_resnet", dict(TINY))
    algo = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=8, epochs_per_batch=1), model)
    algo.grad_sync = GradSync()
    algo.grad_sync.broadcast_parameters(model)   # every rank starts from rank 0's weights
    flat = torch.full((5,), float(rank + 1))
    algo.grad_sync.all_reduce_flat(flat)
    assert torch.allclose(flat, torch.full((5,), 1.5))
    # the SyncBatchNorm exchange object (NCCL variant; gloo here): a plain SUM of the (2*C,) float64 statistics
    bn = BatchNormSync()
    sums = torch.arange(6, dtype=torch.float64) * (rank + 1)
    assert bn.world_size == world and torch.equal(bn.all_reduce_(sums), torch.arange(6, dtype=torch.float64) * 3)
    N, A = 4, 11259
    buf = KataGoRolloutBuffer(N, (50, 9, 9), A)
    for t in range(2):
        obs = torch.randn(N, 50, 9, 9)
        mask = torch.rand(N, A) < 0.01; mask[:, 5] = True
        a, lp, v = algo.select_actions(obs, mask)
        term = torch.zeros(N, dtype=torch.bool)
        buf.add(obs, a, lp, v, torch.randn(N), term, term, mask, torch.full((N,), -1), torch.zeros(N))
    metrics = algo.update(buf, torch.zeros(N))
    assert all(v == v for v in metrics.values())
    sd = {k: v for k, v in model.state_dict().items() if "running_" not in k and "num_batches" not in k}
    torch.save(sd, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    shutdown()


@pytest.mark.timeout(180)
def test_two_rank_gloo_update_keeps_weights_in_sync(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_rank_env_without_and_with_partial_launcher_env(monkeypatch):
    from keisei_b200.distributed import init_from_env, rank_env
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    env = rank_env()
    assert not env.launched and env.world_size == 1 and env.rank == 0
    assert init_from_env() == env  # no launcher: nothing to join
    monkeypatch.setenv("RANK", "0")
    with pytest.raises(RuntimeError, match="LOCAL_RANK"):
        rank_env()
