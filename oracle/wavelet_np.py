"""Self-authored float64 specification of the Track-W (wavelet) path.  PARITY UNPINNED.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference (tonyckc/WT-PSE-code) contains no wavelet
transform, no wavelet loss and no softmax OC/OD map; nothing here restates
reference code and nothing here can be checked against it.  The conventions below are this repository's own
choices, fixed so that the CUDA kernels have something exact to be compared with; the tests additionally
check the identities any correct orthonormal DWT must satisfy (perfect reconstruction, Parseval, adjoint).

Conventions
-----------
* Orthonormal filters, periodic extension, decimation phase 0:
      haar  h = [1, 1] / sqrt(2)
      db2   h = [1 + sqrt3, 3 + sqrt3, 3 - sqrt3, 1 - sqrt3] / (4 sqrt2)
      g[k] = (-1)^k h[L-1-k]
      a[n] = sum_k h[k] x[(2n + k) mod N],   d[n] = sum_k g[k] x[(2n + k) mod N]
* 2-D separable, one level on an h x w block:   LL | LH      (first letter: filter along rows/H,
                                                 ---+---       second: along columns/W; L = h, H = g)
                                                 HL | HH
  J levels recurse on the LL quadrant (Mallat packing, output has the input's shape).  H and W must be
  divisible by 2^J.
* Synthesis is the transpose (the transform is orthonormal), so inverse == adjoint.
* Shape-regularization loss on a batch of maps x[N][H][W] (N = batch * channels, e.g. the softmax OC/OD
  probabilities):   L = (1/N) sum_n sum_{j=1..J} w_j * mean(|detail coefficients of level j of map n|)
  i.e. an L1 sparsity penalty on the detail sub-bands (ragged boundaries cost more than smooth ones).

Relation to a published implementation (PyWavelets; NOT installed in this image, so stated, not executed)
---------------------------------------------------------------------------------------------------
``h`` above is PyWavelets' ``Wavelet(name).rec_lo`` and ``g`` its ``rec_hi`` (db2: rec_lo = [0.48296, 0.83652, 0.22414,
-0.12941], rec_hi = [-0.12941, -0.22414, 0.83652, -0.48296]); ``dec_lo`` / ``dec_hi`` are those reversed.
``pywt.dwt(x, name, mode='periodization')`` evaluates  cA[n] = sum_j dec_lo[j] x[(2n + F/2 - j) mod N]  (F = filter
length), i.e.  cA[n] = sum_k h[k] x[(2n + k - (F/2 - 1)) mod N]:  the SAME filters and decimation, input phase shifted
by F/2 - 1 samples.  Hence, per axis,

      haar:  dwt_here(x) == pywt.dwt(x, 'haar', mode='periodization')              (F/2 - 1 = 0)
      db2 :  dwt_here(x) == pywt.dwt(np.roll(x, -1), 'db2', mode='periodization')  (F/2 - 1 = 1)

The 2-D transform is the separable one: every row of the block is filtered along W, then every column along H (pywt.dwt
along axis -1, then along axis -2, with the roll above on each axis for db2).

Known-answer vectors that can be derived by hand from the formulas above (tests/test_wavelet_spec_cpu.py checks this
file against them, tests/test_gpu_wavelet.py the CUDA kernels), N = 8, 1-D unless noted:

  constant x = c          a[n] = c * sum(h) = c * sqrt2,  d[n] = c * sum(g) = 0;  2-D, J levels: LL_J = 2^J * c, details 0
  impulse  x = delta_m    a[n] = h[(m - 2n) mod N] if (m - 2n) mod N < F else 0,  d[n] likewise with g
  ramp     x[m] = m       haar:  a[n] = (4n + 1) / sqrt2,  d[n] = -1 / sqrt2
                          db2 :  two vanishing moments -> d[n] = 0 and a[n] = 2 sqrt2 n + (3 - sqrt3) / sqrt2 for every n
                                 whose taps do not wrap (n <= N/2 - 2); the last one (taps at N-2, N-1, 0, 1) is
                                 a = h0 (N-2) + h1 (N-1) + h3,   d = g0 (N-2) + g1 (N-1) + g3
"""
import numpy as np

SQ2, SQ3 = np.sqrt(2.0), np.sqrt(3.0)
FILTERS = {
    "haar": np.array([1.0, 1.0]) / SQ2,
    "db2": np.array([1 + SQ3, 3 + SQ3, 3 - SQ3, 1 - SQ3]) / (4 * SQ2),
}


def highpass(h):
    L = len(h)
    return np.array([(-1) ** k * h[L - 1 - k] for k in range(L)])


def analysis_matrix(n, h):
    """(n x n) orthonormal matrix [A; D]: first n/2 rows low-pass, last n/2 rows high-pass."""
    g = highpass(h)
    M = np.zeros((n, n))
    for i in range(n // 2):
        for k in range(len(h)):
            M[i, (2 * i + k) % n] += h[k]
            M[n // 2 + i, (2 * i + k) % n] += g[k]
    return M


def dwt2d(x, wavelet="haar", J=1):
    x = np.asarray(x, dtype=np.float64)
    H, W = x.shape[-2:]
    if H % (1 << J) or W % (1 << J):
        raise ValueError("H and W must be divisible by 2^J")
    h = FILTERS[wavelet]
    out = x.copy()
    hh, ww = H, W
    for _ in range(J):
        Mr, Mc = analysis_matrix(hh, h), analysis_matrix(ww, h)
        out[..., :hh, :ww] = Mr @ out[..., :hh, :ww] @ Mc.T
        hh, ww = hh // 2, ww // 2
    return out


def idwt2d(c, wavelet="haar", J=1):
    c = np.asarray(c, dtype=np.float64)
    H, W = c.shape[-2:]
    h = FILTERS[wavelet]
    out = c.copy()
    for j in reversed(range(J)):
        hh, ww = H >> j, W >> j
        Mr, Mc = analysis_matrix(hh, h), analysis_matrix(ww, h)
        out[..., :hh, :ww] = Mr.T @ out[..., :hh, :ww] @ Mc
    return out


def detail_mask(H, W, j):
    """Boolean mask of the three detail sub-bands of level j (1-based) in the Mallat layout."""
    m = np.zeros((H, W), dtype=bool)
    hh, ww = H >> (j - 1), W >> (j - 1)
    m[:hh, :ww] = True
    m[:hh // 2, :ww // 2] = False
    return m


def shape_loss(x, wavelet="haar", J=1, weights=None):
    """Returns (loss, dloss/dx)."""
    x = np.asarray(x, dtype=np.float64)
    N = int(np.prod(x.shape[:-2]))
    H, W = x.shape[-2:]
    w = np.ones(J) if weights is None else np.asarray(weights, dtype=np.float64)
    c = dwt2d(x, wavelet, J)
    gc = np.zeros_like(c)
    loss = 0.0
    for j in range(1, J + 1):
        m = detail_mask(H, W, j)
        cnt = m.sum()
        loss += w[j - 1] * np.abs(c[..., m]).sum() / cnt / N
        gc[..., m] = w[j - 1] * np.sign(c[..., m]) / cnt / N
    return loss, idwt2d(gc, wavelet, J)
