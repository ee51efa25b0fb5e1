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


_WS_CACHE = {}


def _ws(lib, x, nmaps, H, W, J):
    """Scratch for one call.  One grow-only buffer per (device, stream): every launch that touches it is ordered on that
    stream, so consecutive calls can share it (a fresh 28 MB torch.empty per call is most of the host time otherwise)."""
    nbytes = lib.wtpse_wavelet_workspace_bytes(nmaps, H, W, J)
    key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=x.device)
        _WS_CACHE[key] = buf
    return buf, nbytes


def clear_workspace_cache():
    """Drops the cached scratch buffers (one per device and stream that has run a wavelet call)."""
    _WS_CACHE.clear()


class _on_device:
    """torch.cuda.device(dev) without the context-manager cost when dev is already current (the usual case)."""

    def __init__(self, device):
        self.ctx = None if torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


class _Dwt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wavelet, J, inverse):
        x, nmaps, H, W, wid, J = _prep(x, wavelet, J)
        lib = _lib.load()
        with _on_device(x.device):
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
    """Per-level path: forward writes dloss/dcoef, backward is one inverse transform scaled by the upstream gradient.
    Used when a map does not fit the cluster-resident kernel (wtpse_wavelet_resident_cluster == 0)."""

    @staticmethod
    def forward(ctx, x, wavelet, J, weights):
        x, nmaps, H, W, wid, J = _prep(x, wavelet, J)
        w = _weights(weights, J)
        lib = _lib.load()
        with _on_device(x.device):
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
        with _on_device(gcoef.device):
            dx = torch.empty_like(gcoef)
            ws, nbytes = _ws(lib, gcoef, nmaps, H, W, J)
            p_g, keep = _grad_ptr(gout, gcoef)
            _lib.check(lib.wtpse_dwt2d_inverse(_ptr(gcoef), nmaps, H, W, wid, J, _ptr(dx), p_g, _ptr(ws), nbytes,
                                               _stream_ptr(gcoef.device)))
            del keep
        return dx, None, None, None


def _weights(weights, J):
    if weights is None:
        return None
    if len(weights) != J:
        raise ValueError("need one weight per level")
    return (ctypes.c_float * J)(*[float(v) for v in weights])


def _resident_call(x, nmaps, H, W, wid, J, weights, upstream, want_grad):
    lib = _lib.load()
    with _on_device(x.device):
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if want_grad else None
        ws, nbytes = _ws(lib, x, nmaps, H, W, J)
        p_up, keep = (None, None) if upstream is None else _grad_ptr(upstream, x)
        _lib.check(lib.wtpse_wavelet_loss_resident(_ptr(x), nmaps, H, W, wid, J, _weights(weights, J), p_up, _ptr(loss),
                                                   None if grad is None else _ptr(grad), _ptr(ws), nbytes,
                                                   _stream_ptr(x.device)))
        del keep
    return loss, grad


class _WaveletLossResident(torch.autograd.Function):
    """Fused plans (wtpse_wavelet_loss_resident: cluster-resident maps, or streamed levels + a resident stage): the
    forward call already writes dloss/dx for upstream = 1 next to the loss, no coefficient buffer in HBM.  The backward
    then only rescales that buffer -- on the device, and only if the upstream gradient is not exactly 1."""

    @staticmethod
    def forward(ctx, x, wavelet, J, weights):
        x, nmaps, H, W, wid, J = _prep(x, wavelet, J)
        want_grad = ctx.needs_input_grad[0]
        loss, grad = _resident_call(x, nmaps, H, W, wid, J, weights, None, want_grad)
        ctx.cfg = (nmaps, H, W, wid, J, weights)
        ctx.unit_grad = grad                    # consumed (scaled in place and handed out) by the first backward
        ctx.save_for_backward(x)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        if gout is None:
            return None, None, None, None
        nmaps, H, W, wid, J, weights = ctx.cfg
        (x,) = ctx.saved_tensors
        grad, ctx.unit_grad = ctx.unit_grad, None
        if grad is None:
            # a second backward through a retained graph: recompute with the upstream gradient folded in
            _, grad = _resident_call(x, nmaps, H, W, wid, J, weights, gout, True)
            return grad, None, None, None
        lib = _lib.load()
        with _on_device(grad.device):
            p_g, keep = _grad_ptr(gout, grad)
            _lib.check(lib.wtpse_scale_unless_one(_ptr(grad), grad.numel(), p_g, _stream_ptr(grad.device)))
            del keep
        return grad, None, None, None


def resident_cluster_size(H, W, wavelet="haar", J=1):
    """Cluster size the resident kernel would use for H x W maps (0: not resident-capable, per-level path)."""
    return int(_lib.load().wtpse_wavelet_resident_cluster(int(H), int(W), WAVELETS[wavelet], int(J)))


def wavelet_shape_loss(maps, wavelet="haar", J=3, weights=None):
    """(1/N) sum_maps sum_j w_j mean|detail_j|: L1 sparsity of the detail sub-bands of every H x W map
    (e.g. softmax optic-cup / optic-disc probability maps, B x 2 x H x W)."""
    if wavelet not in WAVELETS:
        raise ValueError("wavelet must be one of %s" % sorted(WAVELETS))
    weights = None if weights is None else tuple(weights)
    if maps.is_cuda and maps.dim() >= 2 and resident_cluster_size(maps.shape[-2], maps.shape[-1], wavelet, J):
        return _WaveletLossResident.apply(maps, wavelet, J, weights)
    return _WaveletLoss.apply(maps, wavelet, J, weights)
