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
