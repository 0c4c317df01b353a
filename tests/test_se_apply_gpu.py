"""The block tail — both implementations: 1 = TMA-bulk-staged kernel (csrc/se_apply.cu), 2 = SE-MLP kernel + column
streaming pass (csrc/se_apply_col.cu) — vs a float64 restatement of reference se_resnet.py:83-98 on the same
bf16-rounded inputs: SE MLP, scale/shift, residual, ReLU, pool statistics, tie counts."""
import numpy as np
import pytest
import torch

from keisei_b200 import model_ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("B,C,S,with_bn,train", [(1, 64, 4, True, True), (5, 128, 8, False, False), (300, 256, 16, True, True),
                                                 (149, 256, 16, False, False), (4096, 256, 16, False, False)])
def test_se_block_tail_vs_float64(B, C, S, with_bn, train, variant):
    g = torch.Generator().manual_seed(B * 7 + C)
    z = torch.randn(B, 81, C, generator=g).bfloat16()
    res = torch.randn(B, 81, C, generator=g).clamp_min(0).bfloat16()
    # a constant board/channel (std exactly 0, 81-way max tie) and a dead channel (all outputs 0)
    z[0, :, 0] = 0.5; res[0, :, 0] = 0.25
    z[0, :, 1] = -50.0; res[0, :, 1] = 0.0
    w1 = torch.randn(S, C, generator=g) / C ** 0.5; b1 = 0.1 * torch.randn(S, generator=g)
    w2 = torch.randn(2 * C, S, generator=g) / S ** 0.5; b2 = 0.1 * torch.randn(2 * C, generator=g)
    a = (0.5 + torch.rand(C, generator=g)) if with_bn else None
    b = (0.1 * torch.randn(C, generator=g)) if with_bn else None
    bmean = z.float().mean(dim=1)
    d = lambda t: None if t is None else t.to(DEV)  # noqa: E731
    out, pool, ties, se_in, seh, se = model_ops.se_block_tail(d(z), d(res), d(bmean), d(w1), d(b1), d(w2), d(b2), d(a), d(b),
                                                             want_ties=train, se_raw=train, variant=variant)
    # float64 restatement
    Z, R = z.double(), res.double()
    A = a.double() if with_bn else torch.ones(C, dtype=torch.float64)
    Bb = b.double() if with_bn else torch.zeros(C, dtype=torch.float64)
    sin = bmean.double() * A + Bb
    hid = torch.relu(sin @ w1.double().T + b1.double())
    sev = hid @ w2.double().T + b2.double()
    y = torch.relu((Z * A + Bb) * torch.sigmoid(sev[:, None, :C]) + sev[:, None, C:] + R)
    np.testing.assert_allclose(se_in.cpu().numpy(), sin.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(seh.cpu().numpy(), hid.numpy(), rtol=1e-4, atol=1e-5)
    sev_stored = sev.clone()
    if not train:  # eval hand-off: the scale half is stored with the sigmoid applied
        sev_stored[:, :C] = torch.sigmoid(sev[:, :C])
    np.testing.assert_allclose(se.cpu().numpy(), sev_stored.numpy(), rtol=1e-4, atol=1e-5)
    got = out.float().cpu().double()
    assert float((got - y).abs().max() / y.abs().max()) < 1e-2   # one bf16 rounding of the output
    # mean / std come from the fp32 outputs before the bf16 store; max (and ties) from the STORED values, exactly
    np.testing.assert_allclose(pool[:, :C].cpu().numpy(), y.mean(dim=1).numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(pool[:, C:2 * C].cpu().numpy(), got.amax(dim=1).float().numpy())
    np.testing.assert_allclose(pool[:, 2 * C:].cpu().numpy(), y.std(dim=1, correction=0).numpy(), rtol=1e-3, atol=1e-5)
    if train:
        np.testing.assert_array_equal(ties.cpu().numpy(), (got == got.amax(dim=1, keepdim=True)).sum(dim=1).float().numpy())
        assert ties[0, 0].item() == 81.0
    assert pool[0, 2 * C + 0].item() == 0.0                                   # constant board: std exactly 0
    assert pool[0, C + 1].item() == 0.0 and pool[0, 1].item() == 0.0           # dead channel
