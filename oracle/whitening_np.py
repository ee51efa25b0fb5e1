"""Closed-form numpy restatement of the whitening / MMD / KD-MSE losses (forward AND backward).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Independent of autograd: the backward is the
analytic adjoint, so it cross-checks both the torch restatement and the CUDA kernels.

Follows, line by line:
  * Gram + instance/diagonal terms  -- algorithms.py:1277-1301, shape_networks.py:561-585
  * upper-triangle gather           -- algorithms.py:1305-1306, shape_networks.py:590-591
  * MMD (gaussian, gamma=[1])       -- algorithms.py:65-88,102-121, shape_networks.py:246-309
  * KD MSE                          -- shape_networks.py:596-597
Everything is evaluated in ``dtype`` (float64 by default = the "truth" value).
"""
import numpy as np

C = 16                 # self.dim, algorithms.py:1157
N_OFF = C * (C - 1) // 2   # 120 == torch.sum(reversal_i), algorithms.py:1168
N_DIAG = C             # 16  == torch.sum(diagonal),   algorithms.py:1166
TRIU_I, TRIU_J = np.triu_indices(C, 1)   # row-major (0,1),(0,2)...(14,15) == torch.triu_indices(16,16,1)


def chunk_bounds(B, n_per_domain, n_domains):
    """features[k] = inputs[n*k : n*(k+1)] with python slice truncation (algorithms.py:107)."""
    out = []
    for k in range(n_domains):
        lo = min(n_per_domain * k, B)
        hi = min(n_per_domain * (k + 1), B)
        out.append((lo, hi))
    return out


def gram(z, eps=1e-5, dtype=np.float64):
    """f_cor = bmm(f, f^T) / (HW - 1) + eps * I      (algorithms.py:1280-1283)."""
    B, c, H, W = z.shape
    assert c == C
    f = z.reshape(B, C, H * W).astype(dtype)
    g = np.einsum("bip,bjp->bij", f, f) / dtype(H * W - 1)
    return g + dtype(eps) * np.eye(C, dtype=dtype)


def mmd_forward(v, n_per_domain, n_domains):
    """compute_MMD.forward on v[B,120]; returns (loss, E) with E = exp(-D) over all sample pairs."""
    dtype = v.dtype.type
    sq = (v * v).sum(-1)
    D = np.maximum(sq[:, None] + sq[None, :] - 2.0 * (v @ v.T), dtype(1e-30))   # my_cdist, algorithms.py:65-71
    E = np.exp(-D)
    bounds = chunk_bounds(v.shape[0], n_per_domain, n_domains)
    penalty = dtype(0.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        for k in range(n_domains):
            for l in range(k + 1, n_domains):
                (a0, a1), (b0, b1) = bounds[k], bounds[l]
                kxx = E[a0:a1, a0:a1].mean() if a1 > a0 else dtype(np.nan)
                kyy = E[b0:b1, b0:b1].mean() if b1 > b0 else dtype(np.nan)
                kxy = E[a0:a1, b0:b1].mean() if (a1 > a0 and b1 > b0) else dtype(np.nan)
                penalty = penalty + (kxx + kyy - 2.0 * kxy)
    if n_domains > 1:
        penalty = penalty / dtype(n_domains * (n_domains - 1) / 2)
    return penalty, E, D


def mmd_backward(v, E, D, n_per_domain, n_domains):
    """d(loss)/dv for mmd_forward (gradient of exp(-max(D,1e-30)); zero where the clamp is active)."""
    dtype = v.dtype.type
    B = v.shape[0]
    bounds = chunk_bounds(B, n_per_domain, n_domains)
    # dL/dE as an explicit matrix over ordered pairs
    W = np.zeros((B, B), dtype=v.dtype)
    for k in range(n_domains):
        for l in range(k + 1, n_domains):
            (a0, a1), (b0, b1) = bounds[k], bounds[l]
            na, nb = a1 - a0, b1 - b0
            if na > 0:
                W[a0:a1, a0:a1] += dtype(1.0) / dtype(na * na)
            if nb > 0:
                W[b0:b1, b0:b1] += dtype(1.0) / dtype(nb * nb)
            if na > 0 and nb > 0:
                W[a0:a1, b0:b1] += dtype(-2.0) / dtype(na * nb)
    if n_domains > 1:
        W /= dtype(n_domains * (n_domains - 1) / 2)
    dD = -E * W
    dD[D <= dtype(1e-30)] = 0            # clamp_min_ passes gradient only where D > 1e-30... (strict: min is attained)
    # D_ac = |v_a|^2 + |v_c|^2 - 2 v_a.v_c  ->  dv_a += sum_c dD_ac (2 v_a - 2 v_c) ; dv_c += sum_a dD_ac (2 v_c - 2 v_a)
    S = dD + dD.T
    return 2.0 * (S.sum(1)[:, None] * v - S @ v)


def whitening_forward(z, n_per_domain, n_domains, margin=0.0, eps=1e-5, dtype=np.float64):
    """Returns dict(off, diag, dom, gram, off_b, diag_b, v, E, D).

    WT_PSE.compute_whitening_loss returns (off + diag, dom)          (algorithms.py:1301,1309)
    ShapeVariationalDist_x.compute_whitening_loss returns (off, diag, dom) (shape_networks.py:594)
    """
    B = z.shape[0]
    G = gram(z, eps, dtype)
    off_b = np.abs(G[:, TRIU_I, TRIU_J]).sum(-1) - dtype(margin)                  # :1289
    diag_b = np.abs(G[:, np.arange(C), np.arange(C)] - dtype(1.0)).sum(-1) - dtype(margin)   # :1297
    # torch.clamp(min=0) propagates NaN; np.maximum does too
    L_off = np.maximum(off_b / dtype(N_OFF), dtype(0)).sum() / dtype(B)            # :1290-1291
    L_diag = np.maximum(diag_b / dtype(N_DIAG), dtype(0)).sum() / dtype(B)         # :1298-1299
    v = np.ascontiguousarray(G[:, TRIU_I, TRIU_J])                                # :1305-1306
    dom, E, D = mmd_forward(v, n_per_domain, n_domains)
    return dict(off=L_off, diag=L_diag, dom=dom, gram=G, off_b=off_b, diag_b=diag_b, v=v, E=E, D=D)


def whitening_backward(z, fwd, n_per_domain, n_domains, g_off=1.0, g_diag=1.0, g_dom=1.0, dtype=np.float64):
    """dz for upstream gradients (g_off, g_diag, g_dom) of (L_off, L_diag, L_dom).  SURVEY Appendix A.2."""
    B, _, H, W = z.shape
    P = H * W
    G = fwd["gram"]
    S = np.zeros((B, C, C), dtype=dtype)
    # off-diagonal instance term: sign(G_ij) * [off_b/120 >= 0] / (B*120)
    act_off = (fwd["off_b"] / dtype(N_OFF) >= 0).astype(dtype)
    S[:, TRIU_I, TRIU_J] += dtype(g_off) / dtype(B * N_OFF) * np.sign(G[:, TRIU_I, TRIU_J]) * act_off[:, None]
    # domain term
    if n_domains > 1:
        dv = mmd_backward(fwd["v"], fwd["E"], fwd["D"], n_per_domain, n_domains)
        S[:, TRIU_I, TRIU_J] += dtype(g_dom) * dv
    # diagonal term
    act_diag = (fwd["diag_b"] / dtype(N_DIAG) >= 0).astype(dtype)
    ii = np.arange(C)
    S[:, ii, ii] += dtype(g_diag) / dtype(B * N_DIAG) * np.sign(G[:, ii, ii] - dtype(1.0)) * act_diag[:, None]
    M = (S + S.transpose(0, 2, 1)) / dtype(P - 1)
    f = z.reshape(B, C, P).astype(dtype)
    dz = np.einsum("bij,bjp->bip", M, f)
    return dz.reshape(z.shape), M


def mse_forward(a, b, dtype=np.float64):
    """nn.MSELoss(reduction='mean') -- shape_networks.py:434,596-597."""
    d = a.astype(dtype) - b.astype(dtype)
    return (d * d).mean()


def mse_backward(a, b, g=1.0, dtype=np.float64):
    d = a.astype(dtype) - b.astype(dtype)
    da = dtype(2.0 * g) * d / dtype(d.size)
    return da, -da


# ---- caller-level aggregation quirks -------------------------

def wt_pse_aggregate(per_embedding, num_embeddings=3):
    """WT_PSE.update, algorithms.py:1257-1267: sum over the first num_embeddings-1 embeddings of
    (off+diag, dom), each divided by num_embeddings (3, not 2)."""
    ins = sum(e["off"] + e["diag"] for e in per_embedding) / num_embeddings
    dom = sum(e["dom"] for e in per_embedding) / num_embeddings
    return ins, dom


def shape_aggregate(per_embedding, num_embeddings=3):
    """ShapeVariationalDist_x.update, shape_networks.py:539-554.  The tuple-unpack at :546 overwrites
    instance_wt_loss2 every iteration and :548 doubles it, so the diagonal term that survives is
    2 * L_diag(last embedding) / 3."""
    ins_off = sum(e["off"] for e in per_embedding) / num_embeddings
    ins_diag = 2.0 * per_embedding[-1]["diag"] / num_embeddings
    dom = sum(e["dom"] for e in per_embedding) / num_embeddings
    return ins_off + ins_diag, ins_off, ins_diag, dom
