"""Property tests (hypothesis) tying the two CPU restatements together on random shapes and configurations: the
closed-form numpy forward/backward (oracle/whitening_np.py) must agree with autograd through the operator-sequence
restatement (oracle/whitening_torch.py), including truncated / empty domain chunks, margins that clamp, and
upstream-gradient weights.  Both are separately pinned to the reference goldens (tests/test_oracle_golden.py)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import whitening_np as wnp, whitening_torch as wt


@settings(max_examples=40, deadline=None)
@given(B=st.integers(2, 9), H=st.integers(2, 7), W=st.integers(2, 7), n=st.integers(1, 4), K=st.integers(1, 4),
       margin=st.sampled_from([0.0, 0.01, 0.3, 5.0]), seed=st.integers(0, 10 ** 6),
       w=st.tuples(st.floats(0.1, 2.0), st.floats(0.1, 2.0), st.floats(0.1, 2.0)))
def test_closed_form_matches_autograd(B, H, W, n, K, margin, seed, w):
    g = torch.Generator().manual_seed(seed)
    z = (0.5 * torch.randn(B, 16, H, W, generator=g) + 0.3 * torch.randn(B, 16, 1, 1, generator=g)).double()
    zt = z.clone().requires_grad_(True)
    off, diag, dom, G = wt.whitening_terms(zt, n, K, margin)
    f = wnp.whitening_forward(z.numpy(), n, K, margin)
    assert np.allclose(G.detach().numpy(), f["gram"], rtol=1e-12, atol=1e-14)
    assert np.isclose(float(off), f["off"], rtol=1e-10, atol=1e-14) and np.isclose(float(diag), f["diag"], rtol=1e-10)
    dom_t = float(dom) if torch.is_tensor(dom) else float(dom)
    if np.isnan(f["dom"]):
        assert np.isnan(dom_t)
        return
    assert np.isclose(dom_t, f["dom"], rtol=1e-9, atol=1e-13)
    loss = w[0] * off + w[1] * diag + (w[2] * dom if torch.is_tensor(dom) else 0.0)
    loss.backward()
    dz, _ = wnp.whitening_backward(z.numpy(), f, n, K, w[0], w[1], w[2])
    scale = max(np.abs(dz).max(), 1e-30)
    assert np.abs(zt.grad.numpy() - dz).max() <= 1e-9 * scale


@settings(max_examples=25, deadline=None)
@given(B=st.integers(2, 12), n=st.integers(1, 5), K=st.integers(2, 4), seed=st.integers(0, 10 ** 6))
def test_mmd_is_invariant_to_samples_beyond_the_domain_chunks(B, n, K, seed):
    """compute_MMD ignores rows past K*n (algorithms.py:107): appending samples must not change it."""
    rng = np.random.RandomState(seed)
    v = rng.randn(B, 120) * 0.05
    base, _, _ = wnp.mmd_forward(v, n, K)
    if n * K <= B:
        more = np.concatenate([v, rng.randn(3, 120)], 0)
        again, _, _ = wnp.mmd_forward(more, n, K)
        assert (np.isnan(base) and np.isnan(again)) or np.isclose(base, again, rtol=1e-13)
