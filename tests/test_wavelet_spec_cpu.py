"""Track W: the self-authored float64 specification (oracle/wavelet_np.py) is internally consistent --
orthonormal (perfect reconstruction, Parseval), synthesis == adjoint, and its loss gradient matches finite
differences.  PARITY UNPINNED: nothing in the reference to compare with."""
import numpy as np
import pytest

from oracle import wavelet_np as wn


@pytest.mark.parametrize("wavelet", ["haar", "db2"])
@pytest.mark.parametrize("J", [1, 2, 3])
def test_spec_is_orthonormal(wavelet, J):
    rng = np.random.RandomState(J)
    x = rng.randn(2, 3, 16, 24)
    c = wn.dwt2d(x, wavelet, J)
    assert np.allclose(wn.idwt2d(c, wavelet, J), x, atol=1e-12)
    assert np.isclose((c ** 2).sum(), (x ** 2).sum(), rtol=1e-12)
    y = rng.randn(*x.shape)
    assert np.isclose((c * y).sum(), (x * wn.idwt2d(y, wavelet, J)).sum(), rtol=1e-10)
    M = wn.analysis_matrix(8, wn.FILTERS[wavelet])
    assert np.allclose(M @ M.T, np.eye(8), atol=1e-12)


def test_spec_known_answers():
    # constant map: all energy in the coarsest LL, every detail coefficient zero
    x = np.full((1, 8, 8), 2.0)
    for wv in ("haar", "db2"):
        c = wn.dwt2d(x, wv, 1)
        assert np.allclose(c[0, :4, :4], 4.0) and np.allclose(c[0][wn.detail_mask(8, 8, 1)], 0.0, atol=1e-12)
    # Haar on a 2x2 block: [[a,b],[c,d]] -> LL=(a+b+c+d)/2, LH=(a-b+c-d)/2, HL=(a+b-c-d)/2, HH=(a-b-c+d)/2
    c = wn.dwt2d(np.array([[1.0, 2.0], [3.0, 5.0]]), "haar", 1)
    assert np.allclose(c, [[5.5, -1.5], [-2.5, 0.5]])


def _analysis_1d(x, wavelet):
    """One level along the last axis through the specification's own 2-D entry point (a single row, H = 2 would mix
    rows, so the row is repeated: a constant column has a = sqrt2 * value, d = 0)."""
    x = np.asarray(x, dtype=np.float64)
    c = wn.dwt2d(np.tile(x, (2, 1)), wavelet, 1)          # 2 x N block: rows identical
    n = x.shape[0]
    return c[0, :n // 2] / np.sqrt(2.0), c[0, n // 2:] / np.sqrt(2.0)


def test_spec_hand_derived_known_answers():
    """The vectors derived by hand in oracle/wavelet_np.py's header (constant, unit impulse, ramp)."""
    s2, s3 = np.sqrt(2.0), np.sqrt(3.0)
    h = {"haar": np.array([1, 1]) / s2, "db2": np.array([1 + s3, 3 + s3, 3 - s3, 1 - s3]) / (4 * s2)}
    g = {"haar": np.array([1, -1]) / s2, "db2": np.array([1 - s3, -(3 - s3), 3 + s3, -(1 + s3)]) / (4 * s2)}
    N = 8
    for wv in ("haar", "db2"):
        F = len(h[wv])
        assert np.allclose(wn.FILTERS[wv], h[wv]) and np.allclose(wn.highpass(wn.FILTERS[wv]), g[wv])
        # constant
        a, d = _analysis_1d(np.full(N, 3.0), wv)
        assert np.allclose(a, 3.0 * s2) and np.allclose(d, 0.0, atol=1e-14)
        # unit impulse at every position m
        for m in range(N):
            a, d = _analysis_1d(np.eye(N)[m], wv)
            for n in range(N // 2):
                k = (m - 2 * n) % N
                assert np.isclose(a[n], h[wv][k] if k < F else 0.0) and np.isclose(d[n], g[wv][k] if k < F else 0.0)
    # ramp
    ramp = np.arange(N, dtype=np.float64)
    a, d = _analysis_1d(ramp, "haar")
    assert np.allclose(a, (4 * np.arange(N // 2) + 1) / s2) and np.allclose(d, -1 / s2)
    a, d = _analysis_1d(ramp, "db2")
    n = np.arange(N // 2 - 1)
    assert np.allclose(a[:-1], 2 * s2 * n + (3 - s3) / s2) and np.allclose(d[:-1], 0.0, atol=1e-13)
    assert np.isclose(a[-1], h["db2"][0] * (N - 2) + h["db2"][1] * (N - 1) + h["db2"][3])
    assert np.isclose(d[-1], g["db2"][0] * (N - 2) + g["db2"][1] * (N - 1) + g["db2"][3])
    # 2-D constant over J levels: LL_J = 2^J c
    for wv in ("haar", "db2"):
        c = wn.dwt2d(np.full((16, 16), 0.5), wv, 3)
        assert np.allclose(c[:2, :2], 8 * 0.5) and np.allclose(c[wn.detail_mask(16, 16, 1)], 0, atol=1e-13)


def test_spec_loss_gradient_matches_finite_differences():
    rng = np.random.RandomState(0)
    x = rng.rand(2, 2, 8, 8)
    loss, grad = wn.shape_loss(x, "db2", 2, (1.0, 0.5))
    eps = 1e-6
    for idx in [(0, 0, 1, 2), (1, 1, 7, 7), (0, 1, 4, 0)]:
        xp = x.copy(); xp[idx] += eps
        xm = x.copy(); xm[idx] -= eps
        fd = (wn.shape_loss(xp, "db2", 2, (1.0, 0.5))[0] - wn.shape_loss(xm, "db2", 2, (1.0, 0.5))[0]) / (2 * eps)
        assert abs(fd - grad[idx]) < 1e-7
