"""Standalone MMD domain loss on a B x 120 input: compute_MMD.forward
(algorithms.py:102-121, duplicate at shape_networks.py:283-309), forward + backward in one CTA each."""
import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _grad_ptr, _ptr, _require_cuda_f32, _stream_ptr


class _Mmd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, n_per_domain, n_domains):
        _require_cuda_f32(v, "inputs")
        if v.dim() != 2 or v.shape[1] != 120:
            raise ValueError("compute_MMD input must be B x 120, got %s" % (tuple(v.shape),))
        v = v.contiguous()
        B = v.shape[0]
        lib = _lib.load()
        with torch.cuda.device(v.device):
            ws_bytes = lib.wtpse_mmd_workspace_bytes(B)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=v.device)
            loss = torch.empty((), dtype=torch.float32, device=v.device)
            _lib.check(lib.wtpse_mmd_forward(_ptr(v), B, 120, int(n_per_domain), int(n_domains), _ptr(loss), _ptr(ws),
                                             ws_bytes, _stream_ptr(v.device)))
        ctx.save_for_backward(v)
        ctx.cfg = (int(n_per_domain), int(n_domains), ws_bytes)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        (v,) = ctx.saved_tensors
        n, K, ws_bytes = ctx.cfg
        if gout is None or not ctx.needs_input_grad[0]:
            return None, None, None
        lib = _lib.load()
        with torch.cuda.device(v.device):
            dv = torch.empty_like(v)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=v.device)
            p_g, keep = _grad_ptr(gout, v)
            _lib.check(lib.wtpse_mmd_backward(_ptr(v), p_g, v.shape[0], 120, n, K, _ptr(dv), _ptr(ws), ws_bytes,
                                              _stream_ptr(v.device)))
            del keep
        return dv, None, None


def mmd_penalty(v, n_per_domain, n_domains):
    return _Mmd.apply(v, n_per_domain, n_domains)
