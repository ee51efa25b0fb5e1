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


def test_factored_db2_equals_the_filter_bank():
    """The factored form the db2 level kernels compute (csrc/wavelet_db2.cu), restated in numpy, against the specification:
    p = b + a / sqrt3, q = b - sqrt3 a, lo = h1 (p + (h3 / h1) q_next), hi = -h2 (p + (h0 / h2) q_next) for the analysis and
    pbar = h1 lo - h2 hi, 3 qbar_next = 3 h3 lo - 3 h0 hi, a = (pbar - 3 qbar) / sqrt3, b = pbar + qbar for its adjoint; in 2-D the
    scales h1^2 (LL), h1 h2 (LH, HL; both with a sign flip) and h2^2 (HH) are folded into the level weight and the synthesis
    constants, so the kernels only ever hold the unscaled bands and their signs."""
    s3 = np.sqrt(3.0)
    h0, h1, h2, h3 = wn.FILTERS["db2"]
    i3, kl, kh = 1 / s3, h3 / h1, h0 / h2

    def row_stage(x):                                        # along W
        a, b = x[..., 0::2], x[..., 1::2]
        p, q = a * i3 + b, b - s3 * a
        qn = np.roll(q, -1, -1)
        return p + kl * qn, p + kh * qn

    def col_stage(u):                                        # along H
        a, b = u[..., 0::2, :], u[..., 1::2, :]
        p, q = a * i3 + b, b - s3 * a
        qn = np.roll(q, -1, -2)
        return p + kl * qn, p + kh * qn

    def analysis(x, sc):
        lt, ht = row_stage(x)
        LLt, HLt = col_stage(lt)
        LHt, HHt = col_stage(ht)
        loss = sc * (h1 * h2 * (np.abs(LHt) + np.abs(HLt)).sum() + h2 * h2 * np.abs(HHt).sum())
        return h1 * h1 * LLt, (np.sign(LHt), np.sign(HLt), np.sign(HHt)), loss

    def synthesis(g, sg, sc):
        sLH, sHL, sHH = sg
        cLL, c1, c2 = h1 * h1, h1 * h2 * sc, h2 * h2 * sc
        e = c1 * sHL
        pb, qn = cLL * g + e, 3 * kh * e + 3 * kl * cLL * g
        lam = c1 * sLH
        pbh, qnh = c2 * sHH + lam, 3 * kh * c2 * sHH + 3 * kl * lam
        q3lo, q3hi = np.roll(qn, 1, -2), np.roll(qnh, 1, -2)
        rows = ((i3 * (pb - q3lo), i3 * (pbh - q3hi)), (pb + q3lo / 3, pbh + q3hi / 3))
        out = np.empty(g.shape[:-2] + (2 * g.shape[-2], 2 * g.shape[-1]))
        for pr, (xlo, xhi) in enumerate(rows):
            p, ql = xlo + xhi, np.roll(3 * kh * xhi + 3 * kl * xlo, 1, -1)
            out[..., pr::2, 0::2] = i3 * (p - ql)
            out[..., pr::2, 1::2] = p + ql / 3
        return out

    rng = np.random.RandomState(3)
    N, H, W = 3, 16, 32
    x = rng.randn(N, H, W)
    for J in (1, 2, 3):
        wts = np.array([1.0, 0.7, 0.4])[:J]
        ref_loss, ref_grad = wn.shape_loss(x, "db2", J, wts)
        cur, signs, scales, loss = x, [], [], 0.0
        for j in range(1, J + 1):
            scales.append(wts[j - 1] / (3 * (H >> j) * (W >> j) * N))
            cur, sg, l = analysis(cur, scales[-1])
            signs.append(sg)
            loss += l
        assert np.allclose(cur, wn.dwt2d(x, "db2", J)[..., :H >> J, :W >> J], atol=1e-12)
        assert abs(loss - ref_loss) < 1e-12
        g = np.zeros_like(cur)
        for j in range(J, 0, -1):
            g = synthesis(g, signs[j - 1], scales[j - 1])
        assert np.abs(g - ref_grad).max() < 1e-13
