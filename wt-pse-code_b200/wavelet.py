"""Track W host side: multi-level 2-D DWT / IDWT and the L1 detail-coefficient shape loss.

PARITY UNPINNED -- the reference contains no wavelet code; the conventions are this
repository's own (Haar / db2, orthonormal, periodic, Mallat packing, see include/wtpse_b200.h).
"""
import ctypes

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _grad_ptr, _ptr, _require_cuda_f32, _stream_ptr

WAVELETS = {"haar": 0, "db2": 1}


def _prep(x, wavelet, J):
    _require_cuda_f32(x, "x")
    if x.dim() < 2:
        raise ValueError("x must be ... x H x W")
    if wavelet not in WAVELETS:
        raise ValueError("wavelet must be one of %s" % sorted(WAVELETS))
    H, W = x.shape[-2:]
    nmaps = x.numel() // (H * W) if H * W else 0
    return x.contiguous(), nmaps, H, W, WAVELETS[wavelet], int(J)


def _ws(lib, x, nmaps, H, W, J):
    nbytes = lib.wtpse_wavelet_workspace_bytes(nmaps, H, W, J)
    return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=x.device), nbytes


class _Dwt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wavelet, J, inverse):
        x, nmaps, H, W, wid, J = _prep(x, wavelet, J)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            out = torch.empty_like(x)
            ws, nbytes = _ws(lib, x, nmaps, H, W, J)
            if inverse:
                _lib.check(lib.wtpse_dwt2d_inverse(_ptr(x), nmaps, H, W, wid, J, _ptr(out), None, _ptr(ws), nbytes,
                                                   _stream_ptr(x.device)))
            else:
                _lib.check(lib.wtpse_dwt2d_forward(_ptr(x), nmaps, H, W, wid, J, _ptr(out), _ptr(ws), nbytes,
                                                   _stream_ptr(x.device)))
        ctx.cfg = (wavelet, J, inverse)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        wavelet, J, inverse = ctx.cfg
        # orthonormal transform: the adjoint of the analysis is the synthesis and vice versa
        return _Dwt.apply(g.contiguous(), wavelet, J, not inverse), None, None, None


def dwt2d(x, wavelet="haar", J=1):
    """Mallat-packed J-level coefficients of every H x W map of x (same shape as x)."""
    return _Dwt.apply(x, wavelet, J, False)


def idwt2d(coef, wavelet="haar", J=1):
    return _Dwt.apply(coef, wavelet, J, True)


class _WaveletLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wavelet, J, weights):
        x, nmaps, H, W, wid, J = _prep(x, wavelet, J)
        w = None
        if weights is not None:
            if len(weights) != J:
                raise ValueError("need one weight per level")
            w = (ctypes.c_float * J)(*[float(v) for v in weights])
        lib = _lib.load()
        with torch.cuda.device(x.device):
            gcoef = torch.empty_like(x)
            loss = torch.empty((), dtype=torch.float32, device=x.device)
            ws, nbytes = _ws(lib, x, nmaps, H, W, J)
            _lib.check(lib.wtpse_wavelet_loss_forward(_ptr(x), nmaps, H, W, wid, J, w, _ptr(loss), _ptr(gcoef), _ptr(ws),
                                                      nbytes, _stream_ptr(x.device)))
        ctx.save_for_backward(gcoef)
        ctx.cfg = (nmaps, H, W, wid, J)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        (gcoef,) = ctx.saved_tensors
        nmaps, H, W, wid, J = ctx.cfg
        if gout is None:
            return None, None, None, None
        lib = _lib.load()
        with torch.cuda.device(gcoef.device):
            dx = torch.empty_like(gcoef)
            ws, nbytes = _ws(lib, gcoef, nmaps, H, W, J)
            p_g, keep = _grad_ptr(gout, gcoef)
            _lib.check(lib.wtpse_dwt2d_inverse(_ptr(gcoef), nmaps, H, W, wid, J, _ptr(dx), p_g, _ptr(ws), nbytes,
                                               _stream_ptr(gcoef.device)))
            del keep
        return dx, None, None, None


def wavelet_shape_loss(maps, wavelet="haar", J=3, weights=None):
    """(1/N) sum_maps sum_j w_j mean|detail_j|: L1 sparsity of the detail sub-bands of every H x W map
    (e.g. softmax optic-cup / optic-disc probability maps, B x 2 x H x W)."""
    return _WaveletLoss.apply(maps, wavelet, J, None if weights is None else tuple(weights))
