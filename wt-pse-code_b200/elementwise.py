"""Host side of the element-wise kernels around the loss (include/wtpse_b200.h):

  prepare_batch    custom_transforms.py:466-499 + :581-599  (uint8 image / raw mask -> fp32 image, OD/OC labels)
  od_roi           Trainer.py:842-853, 865-867              (threshold, in-place image += 1, ROI image, pos weight)
  attention_fuse   algorithms.py:1243-1249                  (sigmoid(conv1x1(z_post)) gate on the embedding), autograd
  upsample2x       algorithms.py:947                        (bilinear x2 of the decoder stages, channels-last), autograd
  conv_bias_act    algorithms.py:416-428, 1019-1030         (convolution bias + ReLU in one in-place pass, channels-last), autograd
  batch_norm_act   algorithms.py:877-962, 398-413           (training BatchNorm2d + ReLU, channels-last), autograd
  max_pool2        algorithms.py:897                        (F.max_pool2d(x, 2) of the encoder stages, channels-last), autograd
"""
import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _ptr, _require_cuda_f32, _stream_ptr


def _require_cuda_u8(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.uint8:
        raise TypeError("%s must be a CUDA uint8 tensor" % name)


def prepare_batch(raw_od, img_hwc=None, raw_oc=None):
    """raw_od: B x H x W uint8 ; img_hwc: B x H x W x 3 uint8 (optional).
    Returns (image B x 3 x H x W or None, label_od B x 1 x H x W, label_oc B x 1 x H x W), all float32."""
    _require_cuda_u8(raw_od, "raw_od")
    if raw_od.dim() != 3:
        raise ValueError("raw_od must be B x H x W")
    B, H, W = raw_od.shape
    raw_od = raw_od.contiguous()
    if img_hwc is not None:
        _require_cuda_u8(img_hwc, "img_hwc")
        if tuple(img_hwc.shape) != (B, H, W, 3):
            raise ValueError("img_hwc must be B x H x W x 3 matching raw_od")
        img_hwc = img_hwc.contiguous()
    if raw_oc is not None:
        _require_cuda_u8(raw_oc, "raw_oc")
        raw_oc = raw_oc.contiguous()
    lib = _lib.load()
    dev = raw_od.device
    with torch.cuda.device(dev):
        image = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev) if img_hwc is not None else None
        od = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
        oc = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
        _lib.check(lib.wtpse_prepare_batch(_ptr(img_hwc), _ptr(raw_od), _ptr(raw_oc), B, H, W, _ptr(image), _ptr(od), _ptr(oc),
                                           _stream_ptr(dev)))
    return image, od, oc


def od_roi(logits, image, target_oc=None, threshold=0.75):
    """Returns (od_pred, image_roi, sums) with sums = [sum(od_pred), sum(od_pred*target_oc), pos_weight] (device,
    no host sync).  `image` is incremented by 1 IN PLACE exactly as Trainer.py:850 does (its autograd version counter is
    bumped, so a graph that saved `image` earlier fails loudly instead of back-propagating through changed data).
    Precondition: `target_oc` holds {0, 1} labels (what custom_transforms.py:466-499 produces): sum(od_pred * target_oc) is
    counted as the number of non-zero products -- exact and order-independent for binary targets, wrong for soft ones."""
    _require_cuda_f32(logits, "logits")
    _require_cuda_f32(image, "image")
    if not image.is_contiguous():
        raise ValueError("image must be contiguous (it is updated in place)")
    B, C = image.shape[0], image.shape[1]
    HW = image.shape[2] * image.shape[3]
    if logits.numel() != B * HW:
        raise ValueError("logits must be B x 1 x H x W")
    logits = logits.detach().contiguous()
    if target_oc is not None:
        _require_cuda_f32(target_oc, "target_oc")
        target_oc = target_oc.contiguous()
    lib = _lib.load()
    dev = image.device
    with torch.cuda.device(dev):
        od_pred = torch.empty(B, 1, image.shape[2], image.shape[3], dtype=torch.float32, device=dev)
        roi = torch.empty_like(image)
        sums = torch.empty(3, dtype=torch.float32, device=dev)
        ws_bytes = lib.wtpse_od_roi_workspace_bytes()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.wtpse_od_roi(_ptr(logits), _ptr(target_oc), _ptr(image), _ptr(od_pred), _ptr(roi), B, C, HW,
                                    float(threshold), _ptr(sums), _ptr(ws), ws_bytes, _stream_ptr(dev)))
    torch.autograd.graph.increment_version(image)
    return od_pred, roi, sums


class _AttentionFuse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, z_post, weight, bias, coef, threshold):
        _require_cuda_f32(emb, "embedding")
        _require_cuda_f32(z_post, "z_posterior")
        if weight.numel() != 1 or bias.numel() != 1:
            raise ValueError("attention_layer is Conv2d(1, 1, kernel_size=1): one weight and one bias")
        B, Ce, H, W = emb.shape
        if z_post.shape[0] != B or z_post.numel() != B * H * W:
            raise ValueError("z_posterior must be B x 1 x H x W")
        emb = emb.contiguous()
        z_post = z_post.contiguous()
        wb = torch.cat([weight.detach().reshape(1), bias.detach().reshape(1)]).float()
        lib = _lib.load()
        dev = emb.device
        with torch.cuda.device(dev):
            fuse = torch.empty_like(emb)
            mask = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
            att = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
            _lib.check(lib.wtpse_attention_fuse_forward(_ptr(emb), _ptr(z_post), _ptr(wb), float(coef), B, Ce, H * W,
                                                        float(threshold), _ptr(fuse), _ptr(mask), _ptr(att), _stream_ptr(dev)))
        ctx.save_for_backward(emb, z_post, att, wb)
        ctx.coef = float(coef)
        ctx.wshape, ctx.bshape = weight.shape, bias.shape
        ctx.mark_non_differentiable(mask)
        return fuse, mask

    @staticmethod
    @once_differentiable
    def backward(ctx, g_fuse, _g_mask):
        emb, z_post, att, wb = ctx.saved_tensors
        if g_fuse is None:
            return None, None, None, None, None, None
        B, Ce, H, W = emb.shape
        need_emb, need_zp, need_w, need_b = ctx.needs_input_grad[:4]
        g_fuse = g_fuse.contiguous()
        lib = _lib.load()
        dev = emb.device
        with torch.cuda.device(dev):
            d_emb = torch.empty_like(emb) if need_emb else None
            d_zp = torch.empty_like(z_post) if need_zp else None
            d_wb = torch.empty(2, dtype=torch.float32, device=dev) if (need_w or need_b) else None
            ws_bytes = lib.wtpse_attention_fuse_workspace_bytes(B, H * W)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.wtpse_attention_fuse_backward(_ptr(g_fuse), _ptr(emb), _ptr(z_post), _ptr(att), _ptr(wb), ctx.coef,
                                                         B, Ce, H * W, _ptr(d_emb), _ptr(d_zp), _ptr(d_wb), _ptr(ws),
                                                         ws_bytes, _stream_ptr(dev)))
        d_w = d_wb[0].reshape(ctx.wshape) if need_w else None
        d_b = d_wb[1].reshape(ctx.bshape) if need_b else None
        return d_emb, d_zp, d_w, d_b, None, None


def attention_fuse(embedding, z_posterior, weight, bias, coef, threshold=0.75):
    """(fuse_embedding, z_posterior_attention_mask) of algorithms.py:1243-1249."""
    return _AttentionFuse.apply(embedding, z_posterior, weight, bias, coef, threshold)


def upsample2x_supported(x):
    """True if `x` can take the channels-last x2 kernel: CUDA float32, 4-D, C % 4 == 0, dense channels-last memory."""
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0
            and x.numel() > 0 and x.is_contiguous(memory_format=torch.channels_last))


def _upsample2x_call(t, N, H, W, C, adjoint):
    """H, W: low-resolution sizes.  Returns a channels-last N x C x (2H | H) x (2W | W) tensor."""
    lib = _lib.load()
    oh, ow = (H, W) if adjoint else (2 * H, 2 * W)
    with torch.cuda.device(t.device):
        out = torch.empty((N, C, oh, ow), dtype=torch.float32, device=t.device, memory_format=torch.channels_last)
        _lib.check(lib.wtpse_upsample2x_nhwc(_ptr(t), _ptr(out), N, H, W, C, 1 if adjoint else 0, _stream_ptr(t.device)))
    return out


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if not upsample2x_supported(x):
            raise ValueError("upsample2x needs a CUDA float32 channels-last N x C x H x W tensor with C % 4 == 0")
        N, C, H, W = x.shape
        ctx.dims = (N, C, H, W)
        return _upsample2x_call(x, N, H, W, C, False)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        N, C, H, W = ctx.dims
        g = g.contiguous(memory_format=torch.channels_last)
        return _upsample2x_call(g, N, H, W, C, True)


def upsample2x(x):
    """F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) for a channels-last tensor
    (ConvU.forward, algorithms.py:947), forward and backward as one-pass CUDA kernels."""
    return _Upsample2x.apply(x)


def channel_sum(g):
    """g.sum((0, 2, 3)) -- a convolution's bias gradient.  Dense channels-last CUDA float32 gradients with a power-of-two
    channel count take the two-stage CUDA reduction; anything else goes to ATen (this is the backbone, not the loss path)."""
    C = g.shape[1] if g.dim() == 4 else 0
    if not (g.is_cuda and g.dtype == torch.float32 and g.dim() == 4 and 4 <= C <= 1024 and (C & (C - 1)) == 0
            and g.is_contiguous(memory_format=torch.channels_last) and g.data_ptr() % 16 == 0):
        return g.sum((0, 2, 3))
    N, _, H, W = g.shape
    lib = _lib.load()
    with torch.cuda.device(g.device):
        ws_bytes = lib.wtpse_channel_sum_workspace_bytes(N * H * W, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
        out = torch.empty(C, dtype=torch.float32, device=g.device)
        _lib.check(lib.wtpse_channel_sum_nhwc(_ptr(g), N * H * W, C, _ptr(out), _ptr(ws), ws_bytes, _stream_ptr(g.device)))
    return out


class _BiasAct(torch.autograd.Function):
    """y (a bias-free convolution output, dense channels-last) -> act(y + bias), in place."""

    @staticmethod
    def forward(ctx, y, bias, relu):
        N, C, H, W = y.shape
        lib = _lib.load()
        with torch.cuda.device(y.device):
            _lib.check(lib.wtpse_bias_act_nhwc(_ptr(y), _ptr(bias), N * H * W, C, 1 if relu else 0, _stream_ptr(y.device)))
        ctx.mark_dirty(y)
        ctx.relu = bool(relu)
        if relu:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        if ctx.relu:
            (out,) = ctx.saved_tensors
            C = out.shape[1]
            if (ctx.needs_input_grad[1] and g.dtype == torch.float32 and (C & (C - 1)) == 0 and 4 <= C <= 1024
                    and g.is_contiguous(memory_format=torch.channels_last) and g.data_ptr() % 16 == 0):
                # ReLU backward and the bias gradient in one pass over g and the saved output
                N, _, H, W = out.shape
                lib = _lib.load()
                with torch.cuda.device(g.device):
                    gx = torch.empty_like(out)
                    gb = torch.empty(C, dtype=torch.float32, device=g.device)
                    ws_bytes = lib.wtpse_channel_sum_workspace_bytes(N * H * W, C)
                    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
                    _lib.check(lib.wtpse_relu_backward_channel_sum_nhwc(_ptr(g), _ptr(out), N * H * W, C, _ptr(gx), _ptr(gb),
                                                                        _ptr(ws), ws_bytes, _stream_ptr(g.device)))
                return gx, gb, None
            g = torch.ops.aten.threshold_backward(g, out, 0)          # what ReluBackward0 runs
        return g, (channel_sum(g) if ctx.needs_input_grad[1] else None), None


def conv_bias_act(conv, x, relu):
    """relu(conv(x)) or conv(x) for an nn.Conv2d with a bias: cuDNN convolution without the bias, then ONE in-place pass
    for bias (+ ReLU) instead of ATen's broadcast `add_` followed by a separate `relu_`.  Falls back to the plain
    operators when the output is not a dense channels-last CUDA float32 tensor with C % 4 == 0."""
    F = torch.nn.functional
    if (conv.bias is None or not x.is_cuda or x.dtype != torch.float32 or conv.out_channels % 4 != 0
            or conv.bias.data_ptr() % 16 != 0):
        y = conv(x)
        return F.relu(y, inplace=True) if relu else y
    y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
    if not (y.is_contiguous(memory_format=torch.channels_last) and y.data_ptr() % 16 == 0):
        y = y + conv.bias.view(1, -1, 1, 1)
        return F.relu(y, inplace=True) if relu else y
    return _BiasAct.apply(y, conv.bias, relu)


def batch_norm_act_supported(x, bn):
    """Can `bn` (training mode, affine, with running statistics) run on `x` through the channels-last CUDA kernels?"""
    C = x.shape[1] if x.dim() == 4 else 0
    return (bn.training and bn.affine and bn.track_running_stats and bn.momentum is not None and x.is_cuda
            and x.dtype == torch.float32 and x.dim() == 4 and 4 <= C <= 1024 and (C & (C - 1)) == 0 and x.numel() > 0
            and x.is_contiguous(memory_format=torch.channels_last) and x.data_ptr() % 16 == 0
            and bn.weight.dtype == torch.float32 and bn.weight.data_ptr() % 16 == 0 and bn.bias.data_ptr() % 16 == 0)


class _BatchNormAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, mean_shift, momentum, eps, relu):
        N, C, H, W = x.shape
        npix = N * H * W
        lib = _lib.load()
        dev = x.device
        with torch.cuda.device(dev):
            y = torch.empty_like(x)                                   # preserves the channels-last layout
            stats = torch.empty(3, C, dtype=torch.float32, device=dev)
            ws_bytes = lib.wtpse_batchnorm_workspace_bytes(npix, C)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.wtpse_batchnorm_relu_forward(_ptr(x), npix, C, _ptr(weight), _ptr(bias), _ptr(mean_shift), float(eps),
                                                        float(momentum), 1 if relu else 0, _ptr(running_mean), _ptr(running_var),
                                                        _ptr(y), _ptr(stats), _ptr(ws), ws_bytes, _stream_ptr(dev)))
        ctx.save_for_backward(x, weight, bias, stats)
        ctx.relu, ctx.ws_bytes = bool(relu), ws_bytes
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight, bias, stats = ctx.saved_tensors
        N, C, H, W = x.shape
        g = g.contiguous(memory_format=torch.channels_last)
        lib = _lib.load()
        dev = x.device
        with torch.cuda.device(dev):
            dx = torch.empty_like(x)
            dwb = torch.empty(2, C, dtype=torch.float32, device=dev)
            ws = torch.empty(ctx.ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.wtpse_batchnorm_relu_backward(_ptr(x), _ptr(g), N * H * W, C, _ptr(weight), _ptr(bias), _ptr(stats),
                                                         1 if ctx.relu else 0, _ptr(dx), _ptr(dwb[0]), _ptr(dwb[1]), _ptr(ws),
                                                         ctx.ws_bytes, _stream_ptr(dev)))
        return dx, dwb[0], dwb[1], None, None, None, None, None, None


def batch_norm_act(x, bn, relu, mean_shift=None):
    """relu?(bn(x)) for a training-mode nn.BatchNorm2d on a channels-last tensor (check batch_norm_act_supported first).
    Updates bn.running_mean / running_var / num_batches_tracked like the module; `mean_shift` (a [C] tensor) is added to the
    mean the running mean tracks -- the bias of the preceding convolution when its add was folded away."""
    if mean_shift is not None:
        mean_shift = mean_shift.detach()
    y = _BatchNormAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, mean_shift, bn.momentum, bn.eps, relu)
    bn.num_batches_tracked.add_(1)
    return y


def max_pool2_supported(x):
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0
            and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and x.numel() > 0
            and x.is_contiguous(memory_format=torch.channels_last) and x.data_ptr() % 16 == 0)


class _MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if not max_pool2_supported(x):
            raise ValueError("max_pool2 needs a CUDA float32 channels-last N x C x H x W tensor with C % 4 == 0 and even H, W")
        N, C, H, W = x.shape
        lib = _lib.load()
        with torch.cuda.device(x.device):
            y = torch.empty((N, C, H // 2, W // 2), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
            arg = torch.empty(y.numel(), dtype=torch.uint8, device=x.device)
            _lib.check(lib.wtpse_maxpool2_nhwc(_ptr(x), _ptr(y), _ptr(arg), N, H // 2, W // 2, C, 0, _stream_ptr(x.device)))
        ctx.save_for_backward(arg)
        ctx.dims = (N, C, H, W)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        N, C, H, W = ctx.dims
        g = g.contiguous(memory_format=torch.channels_last)
        lib = _lib.load()
        with torch.cuda.device(g.device):
            gx = torch.empty((N, C, H, W), dtype=torch.float32, device=g.device, memory_format=torch.channels_last)
            _lib.check(lib.wtpse_maxpool2_nhwc(_ptr(g), _ptr(gx), _ptr(arg), N, H // 2, W // 2, C, 1, _stream_ptr(g.device)))
        return gx


def max_pool2(x):
    """F.max_pool2d(x, 2) for a channels-last tensor with even H, W (ConvD.forward, algorithms.py:897)."""
    return _MaxPool2.apply(x)
