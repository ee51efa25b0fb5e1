"""PyTorch-CPU restatement of the reference's shape-regularization losses (autograd backward).

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the "port" that ``bench.py`` times as the
CPU baseline when the real reference tree is not on the box: it issues the same ATen operator
sequence as the reference (one bmm over z, 16x16 masks, triu gather, addmm-form pairwise
distances), so its cost on host cores is the reference's cost.

Restates: algorithms.py:1277-1309 (WT_PSE.compute_whitening_loss),
          shape_networks.py:561-594 (ShapeVariationalDist_x.compute_whitening_loss),
          algorithms.py:59-121 (compute_MMD), shape_networks.py:596-597 (wasser_distance).
"""
import torch

DIM = 16


def _pairwise_sqdist(x, y):
    # algorithms.py:65-71 : |x|^2 + |y|^2 - 2 x.y via addmm, clamped at 1e-30
    xn = (x * x).sum(-1, keepdim=True)
    yn = (y * y).sum(-1, keepdim=True)
    d = torch.addmm(yn.t(), x, y.t(), alpha=-2) + xn
    return d.clamp_min(1e-30)


def _gauss_mean(x, y):
    # algorithms.py:73-80 with gamma=[1], then .mean() (:85-87)
    return torch.exp(-_pairwise_sqdist(x, y)).mean()


def mmd_penalty(v, n_per_domain, n_domains):
    """algorithms.py:102-121 (the .item() at :118-119 is a dead host sync and is not reproduced)."""
    groups = [v[n_per_domain * k:n_per_domain * (k + 1)] for k in range(n_domains)]
    total = 0
    for k in range(n_domains):
        for l in range(k + 1, n_domains):
            total = total + (_gauss_mean(groups[k], groups[k]) + _gauss_mean(groups[l], groups[l])
                             - 2 * _gauss_mean(groups[k], groups[l]))
    if n_domains > 1:
        total = total / (n_domains * (n_domains - 1) / 2)
    return total


def whitening_terms(z, n_per_domain, n_domains, margin=0.0, eps=1e-5):
    """Returns (L_off, L_diag, L_dom, G) as autograd-attached tensors."""
    B, C, H, W = z.shape
    if C != DIM:
        raise ValueError("whitening loss is defined for %d channels, got %d" % (DIM, C))
    f = z.contiguous().view(B, C, H * W)
    eye = torch.eye(C, dtype=z.dtype, device=z.device)
    upper = torch.ones(C, C, dtype=z.dtype, device=z.device).triu(1)
    G = torch.bmm(f, f.transpose(1, 2)) / (H * W - 1) + eps * eye
    G_off = G * upper
    G_diag = G * eye
    off_b = G_off.abs().sum(dim=(1, 2)) - margin
    L_off = (off_b / upper.sum()).clamp(min=0).sum() / B
    diag_b = (G_diag - eye).abs().sum(dim=(1, 2)) - margin
    L_diag = (diag_b / eye.sum()).clamp(min=0).sum() / B
    iu = torch.triu_indices(C, C, 1, device=z.device)
    v = G_off[:, iu[0], iu[1]]
    L_dom = mmd_penalty(v, n_per_domain, n_domains)
    return L_off, L_diag, L_dom, G


def wt_pse_whitening_loss(z, n_per_domain, n_domains, margin=0.0, eps=1e-5):
    """Two-value form of WT_PSE.compute_whitening_loss (algorithms.py:1301,1309)."""
    off, diag, dom, _ = whitening_terms(z, n_per_domain, n_domains, margin, eps)
    return off + diag, dom


def shape_whitening_loss(z, n_per_domain, margin=0.0, eps=1e-5):
    """Three-value form of ShapeVariationalDist_x.compute_whitening_loss; domain_num is the literal 3
    of shape_networks.py:448."""
    off, diag, dom, _ = whitening_terms(z, n_per_domain, 3, margin, eps)
    return off, diag, dom


def kd_mse(a, b):
    """shape_networks.py:596-597."""
    return torch.nn.functional.mse_loss(a, b, reduction="mean")


def fwd_bwd(z, n_per_domain, n_domains, margin=0.0, eps=1e-5):
    """One forward+backward of the two-value loss with unit upstream gradients; returns (ins, dom, dz)."""
    z = z.detach().requires_grad_(True)
    ins, dom = wt_pse_whitening_loss(z, n_per_domain, n_domains, margin, eps)
    (ins + dom).backward()
    return ins.detach(), dom.detach(), z.grad
