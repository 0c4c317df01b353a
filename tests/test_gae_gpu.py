"""GAE kernel parity (K17/K18): bit-exact vs the oracle and the reference golden vectors,
then size-independent properties at BASELINE sizes. Bar: 1e-6 absolute (north_star)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import keisei_oracle as O
from keisei_b200 import gae as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a, dtype=None):
    t = torch.from_numpy(np.array(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def test_golden_all_variants_bit_exact():
    g = load_golden("gae.npz")
    r, v, term, nv, ov = _t(g["r"]), _t(g["v"]), _t(g["term"]), _t(g["nv"]), _t(g["ov"])
    np.testing.assert_array_equal(G.compute_gae(r, v, term, nv, 0.99, 0.95).cpu().numpy(), g["adv_plain"])
    np.testing.assert_array_equal(G.compute_gae_gpu(r, v, term, nv, 0.99, 0.95).cpu().numpy(), g["adv_plain_gpufn"])
    np.testing.assert_array_equal(G.compute_gae_gpu(r, v, term, nv, 0.99, 0.95, next_value_override=ov).cpu().numpy(), g["adv_override"])
    np.testing.assert_array_equal(G.compute_gae(r, v, term.float(), nv, 0.97, 0.9).cpu().numpy(), g["adv_termfloat"])
    np.testing.assert_array_equal(G.compute_gae(r[:, 0], v[:, 0], term[:, 0], nv[0], 0.99, 0.95).cpu().numpy(), g["adv_1d"])
    termp, lengths = _t(g["termp"]), torch.from_numpy(g["lengths"])
    np.testing.assert_array_equal(G.compute_gae_padded_gpu(r, v, termp, nv, lengths, 0.99, 0.95).cpu().numpy(), g["adv_padded"])
    np.testing.assert_array_equal(
        G.compute_gae_padded_gpu(r, v, termp, nv, lengths, 0.99, 0.95, next_value_override=ov).cpu().numpy(),
        g["adv_padded_override"])
    a = _t(g["adv_override"]).reshape(-1).clone()
    G.normalize_advantages_(a)
    np.testing.assert_allclose(a.cpu().numpy(), g["adv_override_normalized"], atol=1e-6)


@pytest.mark.parametrize("T,N,seed", [(128, 64, 42), (128, 64, 123), (1, 1, 7), (3, 33, 99), (257, 100, 55), (512, 512, 1)])
def test_vs_oracle_seeded(T, N, seed):
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((T, N)).astype(np.float32)
    v = (0.3 * rng.standard_normal((T, N))).astype(np.float32)
    term = rng.random((T, N)) < 0.02
    nv = rng.standard_normal(N).astype(np.float32)
    ov = np.where(rng.random((T, N)) < 0.05, rng.standard_normal((T, N)), np.nan).astype(np.float32)
    got = G.compute_gae_gpu(_t(r), _t(v), _t(term), _t(nv), 0.99, 0.95, next_value_override=_t(ov)).cpu().numpy()
    if T * N <= 128 * 64:
        want = O.gae_numpy(r, v, term, nv, 0.99, 0.95, override=ov)
        np.testing.assert_array_equal(got, want)
    # property (any size): the recurrence holds cell by cell, A[t] - decay*A[t+1] == delta
    nxt = np.concatenate([v[1:], nv[None]], 0)
    nxt = np.where(np.isnan(ov), nxt, ov)
    nd = 1.0 - term.astype(np.float32)
    delta = r + np.float32(0.99) * nxt * nd - v
    a_next = np.concatenate([got[1:], np.zeros((1, N), np.float32)], 0)
    np.testing.assert_allclose(got - np.float32(0.99 * 0.95) * nd * a_next, delta, atol=1e-5)


def test_edge_cases():
    z = torch.zeros(0, 4, device=DEV)
    assert G.compute_gae_gpu(z, z, z.bool(), torch.zeros(4, device=DEV), 0.99, 0.95).shape == (0, 4)
    with pytest.raises(ValueError):
        G.compute_gae_gpu(torch.zeros(5, device=DEV), torch.zeros(5, device=DEV), torch.zeros(5, device=DEV),
                          torch.zeros((), device=DEV), 0.99, 0.95)
    # int rewards, float64 values
    r = torch.ones(4, 2, dtype=torch.int64, device=DEV)
    v = torch.full((4, 2), 0.5, dtype=torch.float64, device=DEV)
    out = G.compute_gae(r, v, torch.zeros(4, 2, dtype=torch.bool, device=DEV), torch.full((2,), 0.5, dtype=torch.float64, device=DEV), 0.99, 0.95)
    assert out.dtype == torch.float64
    want = O.gae_numpy(np.ones((4, 2)), np.full((4, 2), 0.5), np.zeros((4, 2)), np.full(2, 0.5), 0.99, 0.95)
    np.testing.assert_allclose(out.cpu().numpy(), want, atol=1e-6)
    # no graph leaks out
    vv = torch.randn(4, 2, device=DEV, requires_grad=True)
    out = G.compute_gae(torch.randn(4, 2, device=DEV), vv, torch.zeros(4, 2, device=DEV), torch.zeros(2, device=DEV), 0.99, 0.95)
    assert not out.requires_grad
    # all terminated: A = r - v exactly
    r = torch.randn(9, 5, device=DEV); v = torch.randn(9, 5, device=DEV)
    out = G.compute_gae_gpu(r, v, torch.ones(9, 5, dtype=torch.bool, device=DEV), torch.randn(5, device=DEV), 0.99, 0.95)
    assert torch.equal(out, r - v)


def test_normalize_large_matches_torch():
    a = torch.randn(128 * 512, device=DEV) * 3 + 1
    want = (a - a.mean()) / (a.std() + 1e-8)
    G.normalize_advantages_(a)
    assert (a - want).abs().max().item() < 1e-5
    one = torch.tensor([2.5], device=DEV)
    assert G.normalize_advantages_(one).item() == 2.5
