"""Bit-packed legal masks (1,408 B instead of 11,259 B per sample) through the rollout buffer, the
sampler, the PPO loss kernels and the fused minibatch gather; the no-mask mode used by the supervised loss
(reference sl/trainer.py:147-158). Integer / index work: bit-exact against the byte-mask kernels and torch indexing."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from keisei_b200 import _lib, policy_ops, sl
from keisei_b200.katago_ppo import KataGoPPOAlgorithm, KataGoPPOParams, KataGoRolloutBuffer
from keisei_b200.model_registry import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
A = 11259
TINY = dict(num_blocks=1, channels=32, se_reduction=4, global_pool_channels=8, policy_channels=8, value_fc_size=8, score_fc_size=8)


def _rows(B, seed, dtype=torch.bfloat16, density=0.01):
    g = torch.Generator(device=DEV).manual_seed(seed)
    logits = torch.randn(B, 11264, device=DEV, generator=g).to(dtype)[:, :A]
    mask = torch.rand(B, A, device=DEV, generator=g) < density
    acts = torch.randint(0, A, (B,), device=DEV, generator=g)
    mask[torch.arange(B, device=DEV), acts] = True
    return logits, mask, acts


@pytest.mark.parametrize("A_", [11259, 64, 33, 1])
def test_pack_unpack_roundtrip_and_bit_order(A_):
    g = torch.Generator(device=DEV).manual_seed(A_)
    mask = torch.rand(37, A_, device=DEV, generator=g) < 0.3
    bits = policy_ops.pack_mask_bits(mask)
    assert bits.dtype == torch.int32 and bits.shape == (37, (A_ + 31) // 32)
    assert torch.equal(policy_ops.unpack_mask_bits(bits, A_), mask)
    # action i = bit (i & 31) of word (i >> 5); padding bits are zero
    m = mask.cpu().numpy()
    want = np.zeros((37, (A_ + 31) // 32), dtype=np.uint32)
    for i in range(A_):
        want[:, i >> 5] |= (m[:, i].astype(np.uint32) << np.uint32(i & 31))
    np.testing.assert_array_equal(bits.cpu().numpy().view(np.uint32), want)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("padded", [True, False])
def test_loss_and_sampler_kernels_packed_equals_bytes(dtype, padded):
    """Bit-packed masks run their own kernels (streaming NaN scan + walk over the legal entries; write-only backward):
    same integers (flags, zero gradient on illegal entries), floats equal to the byte-mask kernels up to summation order.
    `padded=False` gives rows that are not 16-byte aligned: the scalar tails and the generic backward."""
    B = 70
    logits, mask, acts = _rows(B, 3, dtype)
    if not padded:
        logits = logits.contiguous()
    mask[5] = False; mask[5, 77] = True                      # single legal action
    mask[6] = False; mask[6, 11258] = True; mask[6, 0] = True   # first and last action of the row
    acts[6] = 11258
    acts[7] = int((~mask[7]).nonzero()[0])                   # an ILLEGAL action: log-prob -inf in both
    bits = policy_ops.pack_mask_bits(mask)
    g = torch.Generator(device=DEV).manual_seed(4)
    old, adv = -3 * torch.rand(B, device=DEV, generator=g), torch.randn(B, device=DEV, generator=g)
    outs = {}
    for name, mk in (("bytes", mask), ("bits", bits)):
        lg = logits.detach().clone().requires_grad_(True)
        out2, new_lp, row_ent, row_lse, dlogp, flags = policy_ops.ppo_policy_loss(lg, mk, acts, old, adv, 0.2)
        ok = torch.ones(B, dtype=torch.bool, device=DEV); ok[7] = False
        (new_lp[ok].sum() * 0.01 + row_ent.sum() * 0.002).backward()
        outs[name] = (new_lp.detach(), row_ent.detach(), row_lse.detach(), lg.grad.clone(), flags)
    by, bi = outs["bytes"], outs["bits"]
    assert torch.equal(by[4], bi[4]) and int(bi[4][0]) == 0 and int(bi[4][1]) == 0
    assert bool(torch.isneginf(bi[0][7])) and bool(torch.isneginf(by[0][7]))
    ok = torch.ones(B, dtype=torch.bool, device=DEV); ok[7] = False
    for k in (0, 1, 2):
        assert torch.allclose(by[k][ok], bi[k][ok], rtol=2e-6, atol=2e-6), k
    assert bool((bi[3][~mask] == 0).all())
    gd = (by[3].float() - bi[3].float()).abs()
    tol = 1e-7 if dtype == torch.float32 else 4e-3 * by[3].float().abs().max()      # one bf16 ulp of the largest entry
    assert float(gd.max()) <= float(tol), float(gd.max())
    vl = torch.randn(B, 3, device=DEV)
    for forced in (None, acts):
        r0 = policy_ops.policy_sample(logits, mask, vl, seed=11, offset=5, forced_actions=forced, dense=True)
        r1 = policy_ops.policy_sample(logits, bits, vl, seed=11, offset=5, forced_actions=forced)
        r2 = policy_ops.policy_sample(logits, mask, vl, seed=11, offset=5, forced_actions=forced, dense=False)
        for k in (0, 2, 3, 4):
            assert torch.equal(r0[k], r1[k]), k
        for a, b in zip(r1, r2):                                   # packing on the fly == pre-packed
            assert torch.equal(a, b) or bool((a.isnan() == b.isnan()).all() and torch.equal(a.nan_to_num(), b.nan_to_num()))
        fin = ~torch.isinf(r0[1]) & ~torch.isnan(r0[1])
        assert float((r0[1][fin] - r1[1][fin]).abs().max()) <= (2e-5 if dtype == torch.float32 else 0.0625)
        if forced is None:
            assert mask[torch.arange(B, device=DEV), r1[0]].all()
            assert int(r1[0][5]) == 77                             # the row with a single legal action
    # zero-legal row is flagged identically
    bad = mask.clone(); bad[9] = False
    f0 = policy_ops.policy_sample(logits, bad, vl, seed=1, offset=1, dense=True)[4]
    f1 = policy_ops.policy_sample(logits, policy_ops.pack_mask_bits(bad), vl, seed=1, offset=1)[4]
    assert int(f0[0]) == 1 and torch.equal(f0, f1)
    outb = policy_ops.ppo_policy_loss(logits, policy_ops.pack_mask_bits(bad), acts, old, adv, 0.2)
    assert int(outb[5][0]) == 1 and float(outb[1][9]) == 0.0 and float(outb[2][9]) == 0.0
    # NaN in an ILLEGAL raw logit is still reported (the reference checks the raw model output, katago_ppo.py:860-862)
    poisoned = logits.clone()
    poisoned[11, int((~mask[11]).nonzero()[3])] = float("nan")
    poisoned[12, A - 1] = float("nan")                             # in the scalar tail of the row
    for mk in (mask, bits):
        assert int(policy_ops.ppo_policy_loss(poisoned, mk, acts, old, adv, 0.2)[5][1]) == 2


@pytest.mark.parametrize("rows", [1, 3, 64])
def test_pack_rows_at_every_alignment_and_tensor_tail(rows):
    """Rows of 11,259 bytes start at every byte alignment; the last row ends at the tensor's last byte (no over-read that
    changes the result): compare with a host-side numpy packbits."""
    g = torch.Generator(device=DEV).manual_seed(rows)
    mask = (torch.rand(rows, A, device=DEV, generator=g) < 0.5)
    mask[-1, -5:] = True
    as_u8 = (mask.to(torch.uint8) * 255)                           # any non-zero byte is a set bit
    want = np.packbits(np.pad(mask.cpu().numpy(), ((0, 0), (0, 352 * 32 - A))), axis=1, bitorder="little").view(np.uint32)
    for m in (mask, as_u8):
        got = policy_ops.pack_mask_bits(m).cpu().numpy().view(np.uint32)
        np.testing.assert_array_equal(got, want)
    if rows > 1:                                                   # a row slice starts at an arbitrary byte address
        got = policy_ops.pack_mask_bits(mask[1:]).cpu().numpy().view(np.uint32)
        np.testing.assert_array_equal(got, want[1:])


def test_no_mask_mode_is_plain_log_softmax_and_sl_losses_match_torch():
    B = 33
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(B, A, generator=g)
    value = torch.randn(B, 3, generator=g); score = torch.randn(B, 1, generator=g)
    pt, vt, st = torch.randint(0, A, (B,), generator=g), torch.randint(0, 3, (B,), generator=g), torch.randn(B, generator=g)
    ref_in = [t.clone().requires_grad_(True) for t in (logits, value, score)]
    want = 1.0 * F.cross_entropy(ref_in[0], pt) + 1.5 * F.cross_entropy(ref_in[1], vt) + 0.02 * F.mse_loss(ref_in[2].squeeze(-1), st)
    want.backward()
    got_in = [t.to(DEV).requires_grad_(True) for t in (logits, value, score)]
    n0 = _lib.launch_count()
    loss, pl, vl, sc = sl.sl_losses(got_in[0], got_in[1], got_in[2], pt.to(DEV), vt.to(DEV), st.to(DEV))
    assert _lib.launch_count() > n0
    assert abs(loss.item() - want.item()) <= 1e-5 * abs(want.item())
    assert abs(pl.item() - F.cross_entropy(logits, pt).item()) <= 1e-5 * pl.item()
    loss.backward()
    for a, b in zip(got_in, ref_in):
        assert torch.allclose(a.grad.cpu(), b.grad, rtol=1e-4, atol=1e-7)
    # all-ones byte mask == no mask, bit for bit
    ones = torch.ones(B, A, dtype=torch.bool, device=DEV)
    z = torch.zeros(B, device=DEV)
    r0 = policy_ops.ppo_policy_loss(got_in[0].detach(), None, pt.to(DEV), z, z, 0.0)
    r1 = policy_ops.ppo_policy_loss(got_in[0].detach(), ones, pt.to(DEV), z, z, 0.0)
    assert torch.equal(r0[1], r1[1]) and torch.equal(r0[3], r1[3])


def test_sl_step_on_kernel_model_bf16_vs_oracle():
    from oracle import keisei_oracle as O
    torch.manual_seed(0)
    cfg = dict(num_blocks=2, channels=128, se_reduction=8, global_pool_channels=32, policy_channels=16, value_fc_size=32, score_fc_size=32)
    m = build_model("se_resnet", dict(cfg))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    B = 24
    g = torch.Generator().manual_seed(1)
    batch = {"observation": torch.randn(B, 50, 9, 9, generator=g), "policy_target": torch.randint(0, A, (B,), generator=g),
             "value_target": torch.randint(0, 3, (B,), generator=g), "score_target": torch.randn(B, generator=g).clamp(-1.5, 1.5)}
    with torch.no_grad():
        wp, wv, ws = O.seresnet_forward(sd, batch["observation"], 2, training=True)
    want = (F.cross_entropy(wp.reshape(B, -1), batch["policy_target"]).item(), F.cross_entropy(wv, batch["value_target"]).item(),
            F.mse_loss(ws.squeeze(-1), batch["score_target"]).item())
    m = m.to(DEV)
    step = sl.SLStep(m, torch.optim.Adam(m.parameters(), 1e-3), use_amp=True)
    before = [p.detach().clone() for p in m.parameters()]
    out = step(batch)
    for got, w, k in zip((out["policy_loss"], out["value_loss"], out["score_loss"]), want, ("policy", "value", "score")):
        assert abs(got.item() - w) <= 2e-2 * max(abs(w), 1e-3), (k, got.item(), w)
    assert any(not torch.equal(a, b) for a, b in zip(before, m.parameters()))


def test_gather_minibatch_equals_index_ops():
    N, M = 300, 128
    g = torch.Generator(device=DEV).manual_seed(9)
    obs = torch.randn(N, 50, 9, 9, device=DEV, generator=g)
    bits = policy_ops.pack_mask_bits(torch.rand(N, A, device=DEV, generator=g) < 0.02)
    acts, cats = torch.randint(0, A, (N,), device=DEV, generator=g), torch.randint(-1, 3, (N,), device=DEV, generator=g)
    f = [torch.randn(N, device=DEV, generator=g) for _ in range(4)]
    idx = torch.randperm(N, device=DEV, generator=g)[:M]
    n0 = _lib.launch_count()
    got = policy_ops.gather_minibatch(obs, bits, acts, f[0], f[1], cats, f[2], f[3], idx)
    assert _lib.launch_count() == n0 + 1
    want = (obs[idx], bits[idx], acts[idx], f[0][idx], f[1][idx], cats[idx], f[2][idx], f[3][idx])
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and torch.equal(a, b)


def test_device_buffer_stores_packed_masks_and_update_matches_host_buffer():
    N, T = 6, 5
    torch.manual_seed(0)
    host, devb = KataGoRolloutBuffer(N, (50, 9, 9), A), KataGoRolloutBuffer(N, (50, 9, 9), A, device=DEV)
    g = torch.Generator().manual_seed(2)
    masks = []
    for t in range(T):
        obs = torch.randn(N, 50, 9, 9, generator=g)
        mask = torch.rand(N, A, generator=g) < 0.01
        a = torch.randint(0, A, (N,), generator=g); mask[torch.arange(N), a] = True
        masks.append(mask)
        term = torch.rand(N, generator=g) < 0.2
        fields = [obs, a, -2 * torch.rand(N, generator=g), 0.3 * torch.randn(N, generator=g), term.float(), term, term, mask,
                  torch.where(term, torch.randint(0, 3, (N,), generator=g), torch.full((N,), -1)), torch.randn(N, generator=g).clamp(-1, 1)]
        host.add(*fields)
        devb.add(*[x.to(DEV) for x in fields])
    assert devb._storage["legal_masks"].dtype == torch.int32 and devb._storage["legal_masks"].shape[1] == 352
    flat = devb.flatten()
    assert dict.__contains__(flat, "legal_masks_packed") and not dict.__contains__(flat, "legal_masks")
    assert "legal_masks" in flat                                   # advertised, unpacked lazily
    assert torch.equal(flat["legal_masks"].cpu(), torch.cat(masks))
    assert set(host.flatten().keys()) | {"legal_masks_packed"} == set(flat.keys())
    # same seeds -> same shuffles -> identical update through the packed / gathered path
    res = []
    for buf in (host, devb):
        torch.manual_seed(5)
        m = build_model("se_resnet", dict(TINY)).to(DEV)
        algo = KataGoPPOAlgorithm(KataGoPPOParams(batch_size=16, epochs_per_batch=2), m)
        torch.manual_seed(6)
        metrics = algo.update(buf, torch.zeros(N, device=DEV))
        res.append((metrics, [p.detach().clone() for p in m.parameters()]))
    for k in ("policy_loss", "value_loss", "score_loss", "entropy"):
        assert res[0][0][k] == pytest.approx(res[1][0][k], rel=1e-6, abs=1e-8), k
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_pinned_ingest_double_buffer_delivers_each_step_intact():
    """keisei_b200.ingest.PinnedIngest (reference katago_loop.py:1529-1530): pinned, double-buffered
    host -> device transfer with optional bit-packing; slots are reused safely and `select_actions` consumes them."""
    from keisei_b200.ingest import PinnedIngest
    N = 40
    ing = PinnedIngest(DEV, N, (50, 9, 9), A, depth=2, pack_masks=True)
    g = torch.Generator().manual_seed(0)
    steps = []
    for k in range(5):
        n = N if k != 3 else 17                                    # a ragged step (fewer live envs)
        obs = torch.randn(n, 50, 9, 9, generator=g)
        mask = torch.rand(n, A, generator=g) < 0.01
        mask[:, k] = True
        steps.append((obs.numpy() if k % 2 else obs, mask.numpy() if k % 2 else mask, obs, mask))
    torch.manual_seed(0)
    m = build_model("se_resnet", dict(TINY)).to(DEV)
    algo = KataGoPPOAlgorithm(KataGoPPOParams(), m)
    slot = ing.submit(steps[0][0], steps[0][1])
    for k in range(5):
        cur = slot
        if k + 1 < 5:
            slot = ing.submit(steps[k + 1][0], steps[k + 1][1])   # next step in flight while this one is consumed
        d_obs, d_bits = ing.get(cur)
        assert d_bits.dtype == torch.int32 and d_obs.is_cuda
        assert torch.equal(d_obs.cpu(), steps[k][2])
        assert torch.equal(policy_ops.unpack_mask_bits(d_bits, A).cpu(), steps[k][3])
        a, lp, v = algo.select_actions(d_obs, d_bits)
        ing.release(cur)
        assert steps[k][3][torch.arange(d_obs.shape[0]), a.cpu()].all()
    with pytest.raises(ValueError, match="unexpected shapes"):
        ing.submit(torch.zeros(3, 46, 9, 9), torch.zeros(3, A, dtype=torch.bool))


@pytest.mark.parametrize("dtype,mode", [(torch.float32, 0), (torch.bfloat16, 1), (torch.bfloat16, 0)])
@pytest.mark.parametrize("density", [0.007, 0.4])
def test_sparse_warp_per_row_sampler_matches_the_dense_kernel(dtype, mode, density):
    """Bit-packed masks are sampled by a warp-per-row kernel that gathers only the legal logits (csrc/policy.cu:
    policy_sample_bits_kernel). Same Philox keys and tie-breaking as the dense one-CTA-per-row kernel: the drawn action,
    the legal counts and the flags are identical; log-probs agree up to summation order."""
    B = 333
    logits, mask, acts = _rows(B, 21, dtype, density)
    mask[7] = False                                               # zero legal actions
    mask[8] = False; mask[8, 11258] = True                        # single legal action in the last (partial) mask word
    logits = logits.clone(); logits[9][mask[9]] = float("-inf")   # every legal logit -inf: falls back to the first legal index
    vl = torch.randn(B, 3, device=DEV)
    sc = torch.randn(B, 1, device=DEV)
    for forced in (None, acts):
        d = policy_ops.policy_sample(logits, mask, vl, sc, 0.3, seed=5, offset=9, logprob_mode=mode, forced_actions=forced, dense=True)
        s_ = policy_ops.policy_sample(logits, policy_ops.pack_mask_bits(mask), vl, sc, 0.3, seed=5, offset=9, logprob_mode=mode,
                                      forced_actions=forced)
        assert torch.equal(d[0], s_[0])                           # actions
        assert torch.equal(d[3], s_[3]) and torch.equal(d[4], s_[4])   # legal counts, flags
        assert torch.equal(d[2], s_[2])                           # scalar values
        ok = torch.ones(B, dtype=torch.bool, device=DEV); ok[7] = False
        if forced is None:
            assert int(s_[0][8]) == 11258 and int(s_[0][9]) == int(mask[9].nonzero()[0])
        diff = (d[1][ok] - s_[1][ok]).abs()
        diff = diff[~diff.isnan()]
        if mode == 0:
            assert float(diff.max()) < 2e-5
        else:   # bf16-rounded log-probs: a last-bit difference of the normaliser may flip one rounding
            assert float((diff > 0).float().mean()) < 0.02 and float(diff.max()) <= 0.0625
    assert bool(s_[1][7].isnan())
