"""Autograd-facing host side of the C ABI: the whitening (Gram) loss and the KD MSE.

PyTorch is plumbing here (device memory, current stream, autograd graph); all arithmetic runs in
``libwtpse_b200.so``.  Tensors must be CUDA float32; anything else raises -- there is no CPU path.

Reference statements replaced:
  whitening_terms        algorithms.py:1277-1309 / shape_networks.py:561-594 (+ compute_MMD.forward)
  kd_mse                 shape_networks.py:596-597
"""
import ctypes

import torch
from torch.autograd.function import once_differentiable

from . import _lib

CHANNELS = 16


def _require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %r" % (name, type(t)))
    if not t.is_cuda:
        raise RuntimeError("%s must live on a CUDA device: the shape-loss path has no CPU implementation" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _scalar_alias(buf, index):
    """0-dim tensor aliasing buf[index] without being an autograd view of it."""
    return torch.empty(0, dtype=buf.dtype, device=buf.device).set_(buf.untyped_storage(), buf.storage_offset() + index, ())


def _grad_ptr(g, like):
    if g is None:
        return None, None
    if g.dtype != torch.float32 or not g.is_cuda:
        g = g.to(device=like.device, dtype=torch.float32)
    g = g.contiguous()
    return ctypes.c_void_p(g.data_ptr()), g      # keep the tensor alive until the launch is enqueued


_WS_CACHE = {}


def _workspace(lib, B, P, device):
    """(buffer, bytes) of forward scratch for a [B][16][P] batch on the current stream of `device`.

    The C ABI's ticket contract (include/wtpse_b200.h, wtpse_whitening_ticket_bytes): the first bytes of the forward's
    workspace must be zero when a call is enqueued, and every call leaves them zero.  One grow-only buffer per (device,
    stream), zeroed when it is created, satisfies that for free: all launches that touch it are ordered on that stream.
    While a CUDA graph is being captured the buffer is a fresh `torch.zeros` instead (it then lives in the graph's own
    memory pool and its memset is part of the graph), so captured and eager work never share tickets."""
    nbytes = lib.wtpse_whitening_workspace_bytes(B, P)
    if torch.cuda.is_current_stream_capturing():
        return torch.zeros(nbytes, dtype=torch.uint8, device=device), nbytes
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _WS_CACHE[key] = buf
    return buf, nbytes


def clear_workspace_cache():
    """Drops the cached forward workspaces (one per device and stream that has run a whitening forward)."""
    _WS_CACHE.clear()


def _as_loss_input(z):
    """(tensor the kernels read, channels_last flag).  A dense channels-last z (what a channels-last backbone produces) is
    read in place by the *_cl entry points; anything else is made NCHW-contiguous, the reference's layout
    (`z.contiguous().view`, algorithms.py:1280)."""
    if z.dim() == 4 and not z.is_contiguous() and z.is_contiguous(memory_format=torch.channels_last) and z.data_ptr() % 16 == 0:
        return z, True
    return z.contiguous(), False


def _forward_outputs(B, device):
    """(losses [4], (gram [B][16][16], rowstat [B][2], domgrad [B][120])): the forward's outputs; the tuple is what it
    saves for the backward (include/wtpse_b200.h)."""
    losses = torch.empty(4, dtype=torch.float32, device=device)
    gram = torch.empty(B, CHANNELS, CHANNELS, dtype=torch.float32, device=device)
    rowstat = torch.empty(B, 2, dtype=torch.float32, device=device)
    domgrad = torch.empty(B, 120, dtype=torch.float32, device=device)
    return losses, (gram, rowstat, domgrad)


class _WhiteningLoss(torch.autograd.Function):
    """(L_off, L_diag, L_dom) or, with fold=True, (L_off + L_diag, L_dom)."""

    @staticmethod
    def forward(ctx, z, n_per_domain, n_domains, margin, eps, fold):
        _require_cuda_f32(z, "z")
        if z.dim() != 4:
            raise ValueError("z must be B x C x H x W, got shape %s" % (tuple(z.shape),))
        B, C, H, W = z.shape
        if C != CHANNELS:
            raise ValueError("whitening loss is defined for C == 16 feature maps (self.dim), got C == %d" % C)
        z, cl = _as_loss_input(z)
        P = H * W
        lib = _lib.load()
        with torch.cuda.device(z.device):
            ws, ws_bytes = _workspace(lib, B, P, z.device)
            losses, saved = _forward_outputs(B, z.device)
            gram, rowstat, domgrad = saved
            if cl:
                _lib.check(lib.wtpse_whitening_forward_cl(_ptr(z), None, B, C, P, int(n_per_domain), int(n_domains), float(margin),
                                                          float(eps), _ptr(losses), _ptr(gram), _ptr(rowstat), _ptr(domgrad),
                                                          _ptr(ws), ws_bytes, _stream_ptr(z.device)))
            else:
                _lib.check(lib.wtpse_whitening_forward(_ptr(z), B, C, P, int(n_per_domain), int(n_domains), float(margin),
                                                       float(eps), _ptr(losses), _ptr(gram), _ptr(rowstat), _ptr(domgrad),
                                                       _ptr(ws), ws_bytes, _stream_ptr(z.device)))
        ctx.save_for_backward(z, gram, rowstat, domgrad)
        ctx.cl = cl
        ctx.cfg = (int(n_per_domain), int(n_domains), bool(fold))
        if fold:
            return _scalar_alias(losses, 3), _scalar_alias(losses, 2)
        return _scalar_alias(losses, 0), _scalar_alias(losses, 1), _scalar_alias(losses, 2)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        z, gram, rowstat, domgrad = ctx.saved_tensors
        n, K, fold = ctx.cfg
        if fold:
            g_ins, g_dom = grads
            g_off, g_diag = g_ins, g_ins
        else:
            g_off, g_diag, g_dom = grads
        if not ctx.needs_input_grad[0] or (g_off is None and g_diag is None and g_dom is None):
            return None, None, None, None, None, None
        B, C, H, W = z.shape
        lib = _lib.load()
        with torch.cuda.device(z.device):
            dz = torch.empty_like(z)
            p_off, k0 = _grad_ptr(g_off, z)
            p_diag, k1 = _grad_ptr(g_diag, z)
            p_dom, k2 = _grad_ptr(g_dom, z)
            if ctx.cl:            # dz = empty_like(z) keeps the channels-last layout
                _lib.check(lib.wtpse_whitening_backward_cl(_ptr(z), None, _ptr(gram), _ptr(rowstat), _ptr(domgrad), p_off, p_diag,
                                                           p_dom, B, C, H * W, n, K, _ptr(dz), _stream_ptr(z.device)))
            else:
                _lib.check(lib.wtpse_whitening_backward(_ptr(z), _ptr(gram), _ptr(rowstat), _ptr(domgrad), p_off, p_diag, p_dom,
                                                        B, C, H * W, n, K, _ptr(dz), _stream_ptr(z.device)))
            del k0, k1, k2
        return dz, None, None, None, None, None


class _ReluWhiteningLoss(torch.autograd.Function):
    """relu(z) and the loss terms of z in one pass each way (include/wtpse_b200.h
    wtpse_whitening_relu_forward/backward).  Outputs: (relu(z), L_off, L_diag, L_dom) or, with fold=True,
    (relu(z), L_off + L_diag, L_dom)."""

    @staticmethod
    def forward(ctx, z, n_per_domain, n_domains, margin, eps, fold):
        _require_cuda_f32(z, "z")
        if z.dim() != 4:
            raise ValueError("z must be B x C x H x W, got shape %s" % (tuple(z.shape),))
        B, C, H, W = z.shape
        if C != CHANNELS:
            raise ValueError("whitening loss is defined for C == 16 feature maps (self.dim), got C == %d" % C)
        z, cl = _as_loss_input(z)
        P = H * W
        lib = _lib.load()
        with torch.cuda.device(z.device):
            ws, ws_bytes = _workspace(lib, B, P, z.device)
            losses, saved = _forward_outputs(B, z.device)
            gram, rowstat, domgrad = saved
            relu_out = torch.empty_like(z)                      # same layout as z
            fwd = lib.wtpse_whitening_forward_cl if cl else lib.wtpse_whitening_relu_forward
            _lib.check(fwd(_ptr(z), _ptr(relu_out), B, C, P, int(n_per_domain), int(n_domains), float(margin), float(eps),
                           _ptr(losses), _ptr(gram), _ptr(rowstat), _ptr(domgrad), _ptr(ws), ws_bytes, _stream_ptr(z.device)))
        ctx.save_for_backward(z, gram, rowstat, domgrad)
        ctx.cl = cl
        ctx.cfg = (int(n_per_domain), int(n_domains), bool(fold))
        if fold:
            return relu_out, _scalar_alias(losses, 3), _scalar_alias(losses, 2)
        return relu_out, _scalar_alias(losses, 0), _scalar_alias(losses, 1), _scalar_alias(losses, 2)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_relu, *grads):
        z, gram, rowstat, domgrad = ctx.saved_tensors
        n, K, fold = ctx.cfg
        if fold:
            g_ins, g_dom = grads
            g_off, g_diag = g_ins, g_ins
        else:
            g_off, g_diag, g_dom = grads
        none = (None,) * 6
        if not ctx.needs_input_grad[0]:
            return none
        no_loss_grad = g_off is None and g_diag is None and g_dom is None
        if g_relu is None and no_loss_grad:
            return none
        B, C, H, W = z.shape
        lib = _lib.load()
        with torch.cuda.device(z.device):
            dz = torch.empty_like(z)
            p_off, k0 = _grad_ptr(g_off, z)
            p_diag, k1 = _grad_ptr(g_diag, z)
            p_dom, k2 = _grad_ptr(g_dom, z)
            if ctx.cl:
                if g_relu is not None:
                    _require_cuda_f32(g_relu, "grad of relu(z)")
                    g_relu = g_relu.contiguous(memory_format=torch.channels_last)
                _lib.check(lib.wtpse_whitening_backward_cl(_ptr(z), _ptr(g_relu), _ptr(gram), _ptr(rowstat), _ptr(domgrad), p_off,
                                                           p_diag, p_dom, B, C, H * W, n, K, _ptr(dz), _stream_ptr(z.device)))
            elif g_relu is None:                     # only the loss was used downstream
                _lib.check(lib.wtpse_whitening_backward(_ptr(z), _ptr(gram), _ptr(rowstat), _ptr(domgrad), p_off, p_diag, p_dom,
                                                        B, C, H * W, n, K, _ptr(dz), _stream_ptr(z.device)))
            else:
                # loss unused (e.g. the teacher pass of the shape update): all-NULL upstream scalars make M_b = 0 and the
                # same kernel degenerates to the ReLU backward
                _require_cuda_f32(g_relu, "grad of relu(z)")
                g_relu = g_relu.contiguous()
                _lib.check(lib.wtpse_whitening_relu_backward(_ptr(z), _ptr(g_relu), _ptr(gram), _ptr(rowstat), _ptr(domgrad),
                                                             p_off, p_diag, p_dom, B, C, H * W, n, K, _ptr(dz),
                                                             _stream_ptr(z.device)))
            del k0, k1, k2
        return (dz,) + none[1:]


def relu_whitening_terms(z, n_per_domain, n_domains, margin=0.0, eps=1e-5):
    """(relu(z), L_off, L_diag, L_dom): the DeepWT tail `F.relu(z)` (algorithms.py:1105,1112) fused with
    ShapeVariationalDist_x.compute_whitening_loss(z)."""
    return _ReluWhiteningLoss.apply(z, n_per_domain, n_domains, margin, eps, False)


def relu_whitening_folded(z, n_per_domain, n_domains, margin=0.0, eps=1e-5):
    """(relu(z), L_off + L_diag, L_dom): the same for WT_PSE.compute_whitening_loss(z)."""
    return _ReluWhiteningLoss.apply(z, n_per_domain, n_domains, margin, eps, True)


def whitening_terms(z, n_per_domain, n_domains, margin=0.0, eps=1e-5):
    """(L_off, L_diag, L_dom): the three-value form ShapeVariationalDist_x.compute_whitening_loss returns."""
    return _WhiteningLoss.apply(z, n_per_domain, n_domains, margin, eps, False)


def whitening_folded(z, n_per_domain, n_domains, margin=0.0, eps=1e-5):
    """(L_off + L_diag, L_dom): the two-value form WT_PSE.compute_whitening_loss returns (algorithms.py:1301)."""
    return _WhiteningLoss.apply(z, n_per_domain, n_domains, margin, eps, True)


def gram_matrix(z, eps=1e-5):
    """f_cor of algorithms.py:1283 as a B x 16 x 16 tensor (no autograd); exposed for tests/diagnostics."""
    _require_cuda_f32(z, "z")
    B, C, H, W = z.shape
    z = z.contiguous()
    lib = _lib.load()
    with torch.cuda.device(z.device):
        ws, ws_bytes = _workspace(lib, B, H * W, z.device)
        losses, (gram, rowstat, domgrad) = _forward_outputs(B, z.device)
        _lib.check(lib.wtpse_whitening_forward(_ptr(z), B, C, H * W, 0, 0, 0.0, float(eps), _ptr(losses), _ptr(gram),
                                               _ptr(rowstat), _ptr(domgrad), _ptr(ws), ws_bytes, _stream_ptr(z.device)))
    return gram


class _KdMse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _require_cuda_f32(a, "a")
        _require_cuda_f32(b, "b")
        if a.shape != b.shape:
            raise ValueError("mse inputs must have the same shape, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
        a = a.contiguous()
        b = b.contiguous()
        N = a.numel()
        lib = _lib.load()
        with torch.cuda.device(a.device):
            ws_bytes = lib.wtpse_mse_workspace_bytes(N)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
            loss = torch.empty((), dtype=torch.float32, device=a.device)
            _lib.check(lib.wtpse_mse_forward(_ptr(a), _ptr(b), N, _ptr(loss), _ptr(ws), ws_bytes, _stream_ptr(a.device)))
        ctx.save_for_backward(a, b)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        need_a, need_b = ctx.needs_input_grad
        if gout is None or not (need_a or need_b):
            return None, None
        lib = _lib.load()
        with torch.cuda.device(a.device):
            da = torch.empty_like(a) if need_a else None
            db = torch.empty_like(b) if need_b else None
            p_g, keep = _grad_ptr(gout, a)
            _lib.check(lib.wtpse_mse_backward(_ptr(a), _ptr(b), p_g, a.numel(), _ptr(da), _ptr(db), _stream_ptr(a.device)))
            del keep
        return da, db


def kd_mse(a, b):
    """nn.MSELoss(reduction='mean')(a, b) -- ShapeVariationalDist_x.wasser_distance (shape_networks.py:596-597)."""
    return _KdMse.apply(a, b)


class HostPlan:
    """Plugin-facing host-buffer entry point (include/wtpse_b200.h: wtpse_host_plan_*).

    Takes HOST float32 buffers (numpy arrays or CPU tensors; pinned memory makes the copies
    asynchronous DMA), runs H2D -> forward -> backward -> D2H and returns when the host buffers are
    complete.  This is the call bench.py times as ``e2e``."""

    def __init__(self, B, H, W, device=None):
        self.B, self.H, self.W = int(B), int(H), int(W)
        self._lib = _lib.load()
        self._dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._h = ctypes.c_void_p()
        self._inflight = []           # (losses array, z_host, dz_host) of submitted steps: kept alive until wait()
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.wtpse_host_plan_create(self.B, self.H * self.W, ctypes.byref(self._h)))

    def run(self, z_host, n_per_domain, n_domains, margin=0.0, eps=1e-5, grad_w=(1.0, 1.0, 1.0), dz_host=None):
        if z_host.is_cuda or z_host.dtype != torch.float32 or not z_host.is_contiguous():
            raise ValueError("z_host must be a contiguous float32 CPU tensor")
        if tuple(z_host.shape) != (self.B, CHANNELS, self.H, self.W):
            raise ValueError("z_host shape %s does not match the plan" % (tuple(z_host.shape),))
        if dz_host is not None and (dz_host.is_cuda or dz_host.shape != z_host.shape or not dz_host.is_contiguous()):
            raise ValueError("dz_host must be a contiguous CPU tensor shaped like z_host")
        gw = (ctypes.c_float * 3)(*[float(x) for x in grad_w])
        out = (ctypes.c_float * 4)()
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.wtpse_host_plan_run(self._h, ctypes.c_void_p(z_host.data_ptr()), int(n_per_domain),
                                                     int(n_domains), float(margin), float(eps), gw, out,
                                                     ctypes.c_void_p(dz_host.data_ptr()) if dz_host is not None else None))
        return float(out[0]), float(out[1]), float(out[2])

    def submit(self, z_host, n_per_domain, n_domains, margin=0.0, eps=1e-5, grad_w=(1.0, 1.0, 1.0), dz_host=None):
        """Asynchronous run(): returns a ctypes float[4] that holds (L_off, L_diag, L_dom, L_off+L_diag) once wait()
        -- or the second-next submit() -- has returned.  Two device slots alternate, so consecutive steps overlap
        their PCIe transfers; every in-flight step needs its own host buffers."""
        if z_host.is_cuda or z_host.dtype != torch.float32 or not z_host.is_contiguous():
            raise ValueError("z_host must be a contiguous float32 CPU tensor")
        if tuple(z_host.shape) != (self.B, CHANNELS, self.H, self.W):
            raise ValueError("z_host shape %s does not match the plan" % (tuple(z_host.shape),))
        if dz_host is not None and (dz_host.is_cuda or dz_host.shape != z_host.shape or not dz_host.is_contiguous()):
            raise ValueError("dz_host must be a contiguous CPU tensor shaped like z_host")
        gw = (ctypes.c_float * 3)(*[float(x) for x in grad_w])
        out = (ctypes.c_float * 4)()
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.wtpse_host_plan_submit(self._h, ctypes.c_void_p(z_host.data_ptr()), int(n_per_domain),
                                                        int(n_domains), float(margin), float(eps), gw, out,
                                                        ctypes.c_void_p(dz_host.data_ptr()) if dz_host is not None else None))
        self._inflight.append((out, z_host, dz_host))
        del self._inflight[:-2]                       # the plan has two slots: older steps have left the device
        return out

    def wait(self):
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.wtpse_host_plan_wait(self._h))
        self._inflight.clear()

    def close(self):
        if self._h:
            self._lib.wtpse_host_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
