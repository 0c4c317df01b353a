"""SyncBatchNorm on the CUDA path (SEResNetModel.convert_sync_batchnorm; reference katago_loop.py:494-497 wraps the
model in torch.nn.SyncBatchNorm under DDP by default).

Two "ranks" are emulated on ONE GPU by two threads, each with its own model replica and CUDA stream; the BatchNorm
exchange object sums the per-layer statistics across the threads (the NCCL version is `distributed.BatchNormSync`,
exercised by bench.py at N > 1). Property: with synchronised statistics the two half-batches reproduce the single
process full-batch run — outputs, running statistics, and (summed over ranks) every parameter gradient — and the
full-batch forward is itself checked against the oracle."""
import threading

import numpy as np
import pytest
import torch

from keisei_b200 import model_ops
from keisei_b200.models import SEResNetModel, SEResNetParams
from oracle import keisei_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class ThreadSum:
    """all_reduce_(sum) across `n` in-process ranks; `handle(r)` is rank r's exchange object (the backward runs on an
    autograd worker thread, so the rank is bound to the object, not to the calling thread)."""

    def __init__(self, n: int) -> None:
        self.n = n
        self.barrier = threading.Barrier(n, timeout=60)
        self.slots: list = [None] * n
        self.calls = 0

    def handle(self, rank: int) -> "RankHandle":
        return RankHandle(self, rank)


class RankHandle:
    def __init__(self, shared: ThreadSum, rank: int) -> None:
        self.shared, self.rank, self.world_size = shared, rank, shared.n

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        sh = self.shared
        torch.cuda.current_stream().synchronize()
        sh.slots[self.rank] = t
        sh.barrier.wait()
        total = torch.stack(sh.slots).sum(0)
        torch.cuda.current_stream().synchronize()
        sh.barrier.wait()
        t.copy_(total)
        if self.rank == 0:
            sh.calls += 1
        return t


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def _peer_handles(world: int, slot_doubles: int):
    """`world` in-process ranks of the NVLink peer-memory exchange (csrc/peer_sync.cu): the 'peer' buffers are plain
    device buffers of this process, one PeerBatchNormSync per rank."""
    from keisei_b200 import _lib
    from keisei_b200.distributed import PeerBatchNormSync
    nbytes = _lib.load().kb_peer_buffer_bytes(world, 4, slot_doubles)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=DEV) for _ in range(world)]
    torch.cuda.synchronize()
    return bufs, [PeerBatchNormSync.from_local_buffers([b.data_ptr() for b in bufs], r, slot_doubles, timeout_s=30) for r in range(world)]


def test_peer_memory_allreduce_kernel_three_ranks_many_rounds():
    """The NVLink peer-memory exchange kernel (csrc/peer_sync.cu): sums in rank order, bit-identical on every rank, slots
    reused over many exchanges. The three ranks run as ONE cooperative launch (block r = rank r on its own data and
    buffers): separate launches that spin on one another are not guaranteed to be co-resident on a single GPU."""
    import ctypes
    from keisei_b200 import _lib
    W, n = 3, 512
    bufs, handles = _peer_handles(W, n)
    g = torch.Generator().manual_seed(0)
    rounds = [[torch.randn(n - 7 * k, generator=g, dtype=torch.float64) for _ in range(W)] for k in range(11)]
    ctx_ptrs = (ctypes.c_void_p * W)(*[ctypes.addressof(h.ctx) for h in handles])
    lib = _lib.load()
    for k, parts in enumerate(rounds):
        dev_parts = [q.to(DEV) for q in parts]
        ptrs = (ctypes.c_void_p * W)(*[t.data_ptr() for t in dev_parts])
        _lib.check(lib.kb_peer_allreduce_emulate(ptrs, dev_parts[0].numel(), ctx_ptrs, W, torch.cuda.current_stream().cuda_stream),
                   "kb_peer_allreduce_emulate")
        torch.cuda.synchronize()
        want = parts[0].clone()
        for q in parts[1:]:
            want += q                      # rank order
        for r in range(W):
            assert torch.equal(dev_parts[r].cpu(), want), (k, r)
            handles[r].check()             # no exchange timed out
    assert all(h.ctx.seq == len(rounds) for h in handles)
    del bufs


# The exchange here is the host-synchronised Python hook (each "rank" thread drains its stream before the statistics are
# summed), so no kernel ever waits for another launch. The peer-memory exchange inside the full schedule is asserted on
# REAL multi-process runs in tests/test_multi_gpu.py (2 GPUs) and its kernel in the cooperative-launch test above.
@pytest.mark.parametrize("kind", ["python_hook"])
@pytest.mark.parametrize("amp", [False, True])
def test_two_ranks_with_sync_bn_match_full_batch(amp, kind):
    torch.manual_seed(3)
    p = SEResNetParams(num_blocks=2, channels=64, se_reduction=8, global_pool_channels=16, policy_channels=8,
                       value_fc_size=16, score_fc_size=16)
    ref = SEResNetModel(p)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    W, Bh = 2, 5
    g = torch.Generator().manual_seed(4)
    obs = torch.randn(W * Bh, 50, 9, 9, generator=g)
    wp = torch.randn(W * Bh, 9, 9, 139, generator=g); wv = torch.randn(W * Bh, 3, generator=g); wsc = torch.randn(W * Bh, 1, generator=g)

    def loss_of(out, lo, hi):
        return ((out.policy_logits.float() * wp[lo:hi].to(DEV)).sum() + (out.value_logits * wv[lo:hi].to(DEV)).sum()
                + (out.score_lead * wsc[lo:hi].to(DEV)).sum()) / (W * Bh)

    # single process, whole batch, ordinary BatchNorm
    ref = ref.to(DEV).train()
    if amp:
        ref.configure_amp(True, torch.bfloat16, "cuda")
    full = ref(obs.to(DEV))
    loss_of(full, 0, W * Bh).backward()
    if not amp:  # the full-batch run itself against the oracle (fp32 bar 1e-4)
        with torch.no_grad():
            op, ov, osc = O.seresnet_forward({k: v.clone() for k, v in sd.items()}, obs, p.num_blocks, training=True)
        assert rel(full.policy_logits.detach().cpu().numpy(), op.numpy()) < 1e-4
        assert rel(full.value_logits.detach().cpu().numpy(), ov.numpy()) < 1e-4

    sync = ThreadSum(W)
    peer_bufs, peer_handles = _peer_handles(W, 2 * 64) if kind == "peer_memory" else (None, None)
    results: list = [None] * W
    errors: list = []

    def rank_main(r: int) -> None:
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream(DEV)):
                m = SEResNetModel(p)
                m.load_state_dict(sd)
                m = m.to(DEV).train().convert_sync_batchnorm(peer_handles[r] if kind == "peer_memory" else sync.handle(r))
                if amp:
                    m.configure_amp(True, torch.bfloat16, "cuda")
                lo, hi = r * Bh, (r + 1) * Bh
                # the trainer's path (KataGoPPOAlgorithm._step_fused): raw C-ABI forward / backward on THIS thread.
                # (autograd would run both ranks' backward on the one per-device engine thread: the first rank to
                # reach an exchange would wait forever for the second — an artefact of emulating ranks with threads)
                tables = m._ptr_tables()
                dtype = m._act_dtype(torch.device(DEV))
                code = 0 if dtype == torch.float32 else 1
                wpack = m._packed(tables.params, tables.buffers, dtype)
                pol, val, sco, ws, new_stats = model_ops.seresnet_forward_raw(
                    obs[lo:hi].to(DEV), tables, wpack, True, code, bool(m.use_tensor_cores), m.bn_sync)
                m._store_running_stats(tables.buffers, new_stats)
                dpol = torch.zeros_like(pol)
                dpol[:, :model_ops.POLICY_A] = (wp[lo:hi].reshape(Bh, -1) / (W * Bh)).to(DEV)
                flat = model_ops.seresnet_backward_raw(tables, wpack, ws, dpol, (wv[lo:hi] / (W * Bh)).to(DEV),
                                                       (wsc[lo:hi] / (W * Bh)).to(DEV), code, bool(m.use_tensor_cores),
                                                       m._grad_sizes, m.bn_sync)
                torch.cuda.current_stream().synchronize()
                grads, off = {}, 0
                for n, q in m.named_parameters():
                    grads[n] = flat[off:off + q.numel()].view(q.shape).clone()
                    off += q.numel()
                out = type(full)(policy_logits=pol[:, :model_ops.POLICY_A].view(Bh, 9, 9, 139), value_logits=val, score_lead=sco)
                results[r] = (out, grads, {n: b.clone() for n, b in m.named_buffers()})
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            sync.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(W)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
    n_exchanges = 2 * (2 * p.num_blocks + 2)   # one exchange per BatchNorm layer, forward and backward
    assert (peer_handles[0].ctx.seq if kind == "peer_memory" else sync.calls) == n_exchanges

    tol_out, tol_grad = (3e-2, 6e-2) if amp else (1e-4, 1e-3)
    for r in range(W):
        out = results[r][0]
        lo, hi = r * Bh, (r + 1) * Bh
        assert rel(out.policy_logits.detach().float().cpu(), full.policy_logits[lo:hi].detach().float().cpu()) < tol_out
        assert rel(out.value_logits.detach().cpu(), full.value_logits[lo:hi].detach().cpu()) < tol_out
        assert rel(out.score_lead.detach().cpu(), full.score_lead[lo:hi].detach().cpu()) < tol_out
        # running statistics: every rank holds the GLOBAL batch statistics (unbiased variance over W*Bh*81)
        for n, b in ref.named_buffers():
            if b.is_floating_point():
                assert rel(results[r][2][n].cpu(), b.cpu()) < (1e-2 if amp else 1e-5), n
    bad = {}
    for n, q in ref.named_parameters():
        tot = sum(results[r][1][n] for r in range(W))
        e = rel(tot.cpu(), q.grad.cpu())
        if e > tol_grad:
            bad[n] = e
    assert not bad, bad


def test_sync_hook_error_surfaces_as_exception():
    class Broken:
        world_size = 2

        def all_reduce_(self, t):
            raise RuntimeError("link down")

    m = SEResNetModel(SEResNetParams(num_blocks=1, channels=32, se_reduction=4, global_pool_channels=8, policy_channels=8,
                                     value_fc_size=8, score_fc_size=8)).to(DEV).train().convert_sync_batchnorm(Broken())
    with pytest.raises(RuntimeError, match="link down"):
        m(torch.randn(2, 50, 9, 9, device=DEV))
    m.convert_sync_batchnorm(None)
    m(torch.randn(2, 50, 9, 9, device=DEV))


def test_lost_peer_is_reported_to_the_host_and_not_committed():
    """A peer that never arrives: the exchange gives up after its (configurable) timeout, the sums come back NaN, the
    host-visible status word makes `check()` raise PeerLostError, and BatchNorm finalize leaves the running statistics
    untouched (no silent NaN poisoning)."""
    from keisei_b200 import _lib
    from keisei_b200.distributed import PeerBatchNormSync, PeerLostError
    nbytes = _lib.load().kb_peer_buffer_bytes(2, 4, 64)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=DEV) for _ in range(2)]
    h = PeerBatchNormSync.from_local_buffers([b.data_ptr() for b in bufs], 0, 64, timeout_s=0.3)
    h.check()                                   # nothing pending
    t = torch.ones(8, dtype=torch.float64, device=DEV)
    h.all_reduce_(t)                            # rank 1 never shows up
    torch.cuda.synchronize()
    assert bool(t.isnan().all())
    with pytest.raises(PeerLostError, match="waiting for rank 1"):
        h.check()
    h.check()                                   # reported once, then cleared
    # NaN statistics are not committed: a training forward on such sums keeps the running buffers
    m = SEResNetModel(SEResNetParams(num_blocks=1, channels=32, se_reduction=4, global_pool_channels=8, policy_channels=8,
                                     value_fc_size=8, score_fc_size=8)).to(DEV).train()
    before = {n: b.clone() for n, b in m.named_buffers() if b.is_floating_point()}

    class Poison:
        world_size = 2

        def all_reduce_(self, sums):
            return sums.fill_(float("nan"))

    m.convert_sync_batchnorm(Poison())
    with torch.no_grad():
        m(torch.randn(4, 50, 9, 9, device=DEV))
    for n, b in m.named_buffers():
        if b.is_floating_point():
            assert torch.equal(b, before[n]), n
    h.close(collective=False)
