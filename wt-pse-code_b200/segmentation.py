"""The two entry-point classes of the path -- ``WT_PSE`` and ``ShapeVariationalDist_x`` -- with the
reference's constructor and ``update()`` / ``predict()`` signatures, tuple arities and tensor layouts
(SURVEY.md 8(a) rows R4/R6, 8(b)), so the reference's Trainer drives them unchanged.

Split of labour (north_star): the segmentation backbone -- U-Net stages, the DeepWT feature extractor,
the teacher/student shape networks -- stays ordinary PyTorch (cuDNN); the shape-regularization hot path
inside ``update()`` runs in the CUDA library:

    whitening + MMD loss   wtpse_whitening_forward/backward        (algorithms.py:1261-1264, shape_networks.py:545-549)
    KD MSE                 wtpse_mse_forward/backward              (shape_networks.py:529)
    attention fuse         wtpse_attention_fuse_forward/backward   (algorithms.py:1243-1249)

Module/parameter names follow the reference's state-dict keys (``inc.conv1.weight``,
``wt_model.DoubleConv.double_conv.0.weight``, ``prior_dist.mu_prior.4.bias`` ...), so checkpoints written by
either implementation load into the other with ``strict=True``.  No parameter or buffer is added for the loss.

Reference quirks that are reproduced on purpose: losses summed over 2
embeddings but divided by 3; the ``instance_wt_loss2`` overwrite-and-double in the shape update; the shape
network's MMD always using 3 domains; the teacher mean NOT detached in the KD loss; the
``normal(mu, std) * std + mu`` re-parameterisation.
"""
import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as wf
from .elementwise import (attention_fuse, batch_norm_act, batch_norm_act_supported, channel_sum, conv_bias_act, max_pool2,
                          max_pool2_supported, upsample2x, upsample2x_supported)

BASE_WIDTH = 16      # `n = 16` in algorithms.py:1159 / shape_networks.py:428; also the whitening loss' channel count


def _norm(planes):
    return nn.BatchNorm2d(planes)            # norm='bn' everywhere on the trained path (algorithms.py:1174)


class _PhantomBias(torch.autograd.Function):
    """y -> y, with the gradient `y + bias` would send to `bias` (see _conv_bn)."""

    @staticmethod
    def forward(ctx, y, bias):
        ctx.mark_dirty(y)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, (channel_sum(g) if ctx.needs_input_grad[1] else None)


def _conv_bn(conv, bn, x, fold, relu=False, cuda_bn=False):
    """relu?(bn(conv(x))) in training mode.

    fold=True: the convolution's bias add -- a separate pass over the activation in ATen (`add_` with a broadcast
    [1, C, 1, 1] operand: 25 ms of the 242 ms iteration at 15 x 512 x 512) -- is not executed: batch norm subtracts the
    batch mean, so bn(y + b) == bn(y) up to rounding.  What the bias does change is kept: the running mean tracks
    mean(y) + b, and the bias still gets the gradient sum autograd would give it (rounding noise, as in the reference).

    cuda_bn=True: batch norm and the ReLU run as the channels-last CUDA kernels (elementwise.batch_norm_act) when the
    convolution output allows it; otherwise cuDNN's batch norm and ATen's in-place ReLU."""
    foldable = fold and bn.training and conv.bias is not None and bn.running_mean is not None and bn.momentum is not None
    if foldable:
        y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        if conv.bias.requires_grad and torch.is_grad_enabled():
            y = _PhantomBias.apply(y, conv.bias)
    else:
        y = conv(x)
    if cuda_bn and batch_norm_act_supported(y, bn):
        return batch_norm_act(y, bn, relu, conv.bias if foldable else None)
    if foldable:
        b = conv.bias.detach()
        shifted_mean = bn.running_mean - b              # a temporary: batch_norm saves its running-mean argument for backward
        out = F.batch_norm(y, shifted_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
        bn.running_mean.copy_(shifted_mean + b)
        bn.num_batches_tracked.add_(1)
    else:
        out = bn(y)
    return F.relu(out, inplace=True) if relu else out


def set_cuda_pool(module, on=True):
    """Encoder stages below `module` pool with the channels-last 2x2 CUDA kernels when their input allows it (TrainStep)."""
    for m in module.modules():
        if hasattr(m, "cuda_pool"):
            m.cuda_pool = bool(on)
    return module


def set_cuda_batchnorm(module, on=True):
    """conv -> BatchNorm -> ReLU stages below `module` use the channels-last CUDA batch-norm kernels (TrainStep)."""
    for m in module.modules():
        if hasattr(m, "cuda_bn"):
            m.cuda_bn = bool(on)
    return module


class _ConvActSeq(nn.Sequential):
    """nn.Sequential of Conv2d / ReLU layers (same state-dict keys).  With ``fast_bias`` (TrainStep) every convolution
    runs without its bias and one in-place kernel applies bias (+ the ReLU that follows it): elementwise.conv_bias_act."""

    fast_bias = False

    def forward(self, x):
        if not self.fast_bias:
            return super().forward(x)
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                x = conv_bias_act(m, x, relu)
                i += 2 if relu else 1
            else:
                x = m(x)
                i += 1
        return x


def set_fast_bias(module, on=True):
    for m in module.modules():
        if isinstance(m, _ConvActSeq):
            m.fast_bias = bool(on)
    return module


def set_cuda_upsample(module, on=True):
    """Decoder stages below `module` up-sample with the channels-last CUDA kernel when their input allows it."""
    for m in module.modules():
        if hasattr(m, "cuda_upsample"):
            m.cuda_upsample = bool(on)
    return module


def set_conv_bias_folding(module, on=True):
    """Enable/disable _conv_bn's folding on every conv-BN stage below `module` (TrainStep turns it on)."""
    for m in module.modules():
        if hasattr(m, "fold_bias"):
            m.fold_bias = bool(on)
    return module


class ConvD(nn.Module):
    """Encoder stage: [maxpool] -> conv-bn -> conv-bn-relu -> conv-bn-relu (algorithms.py:877-917)."""

    def __init__(self, inplanes, planes, first=False):
        super().__init__()
        self.first = first
        self.conv1, self.bn1 = nn.Conv2d(inplanes, planes, 3, padding=1), _norm(planes)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes, 3, padding=1), _norm(planes)
        self.conv3, self.bn3 = nn.Conv2d(planes, planes, 3, padding=1), _norm(planes)
        self.fold_bias = False
        self.cuda_bn = False
        self.cuda_pool = False

    def forward(self, x):
        if not self.first:
            x = max_pool2(x) if (self.cuda_pool and max_pool2_supported(x)) else F.max_pool2d(x, 2)
        x = _conv_bn(self.conv1, self.bn1, x, self.fold_bias, False, self.cuda_bn)      # no activation after the first conv
        x = _conv_bn(self.conv2, self.bn2, x, self.fold_bias, True, self.cuda_bn)
        return _conv_bn(self.conv3, self.bn3, x, self.fold_bias, True, self.cuda_bn)


class ConvU(nn.Module):
    """Decoder stage: [conv-bn-relu] -> bilinear x2 -> 1x1 conv-bn-relu -> cat(skip) -> conv-bn-relu
    (algorithms.py:920-962)."""

    def __init__(self, planes, first=False):
        super().__init__()
        self.first = first
        if not first:
            self.conv1, self.bn1 = nn.Conv2d(2 * planes, planes, 3, padding=1), _norm(planes)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes // 2, 1), _norm(planes // 2)
        self.conv3, self.bn3 = nn.Conv2d(planes, planes, 3, padding=1), _norm(planes)
        self.fold_bias = False
        self.cuda_bn = False
        self.cuda_upsample = False        # TrainStep: channels-last x2 kernel instead of ATen's (elementwise.upsample2x)

    def forward(self, x, skip):
        if not self.first:
            x = _conv_bn(self.conv1, self.bn1, x, self.fold_bias, True, self.cuda_bn)
        if self.cuda_upsample and upsample2x_supported(x):
            x = upsample2x(x)
        else:
            x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x = _conv_bn(self.conv2, self.bn2, x, self.fold_bias, True, self.cuda_bn)
        x = torch.cat([skip, x], 1)
        return _conv_bn(self.conv3, self.bn3, x, self.fold_bias, True, self.cuda_bn)


class _UNetTrunk(nn.Module):
    """down1..down4 / up1..up4 shared by all three U-Nets (widths 16-32-64-128-256)."""

    def _build_trunk(self, n=BASE_WIDTH):
        self.down1, self.down2 = ConvD(n, 2 * n), ConvD(2 * n, 4 * n)
        self.down3, self.down4 = ConvD(4 * n, 8 * n), ConvD(8 * n, 16 * n)
        self.up1 = ConvU(16 * n, first=True)
        self.up2, self.up3, self.up4 = ConvU(8 * n), ConvU(4 * n), ConvU(2 * n)

    def _trunk(self, x1):
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        x4 = self.down3(x3)
        x5 = self.down4(x4)
        x = self.up1(x5, x4)
        x = self.up2(x, x3)
        x = self.up3(x, x2)
        return self.up4(x, x1)


class _DoubleConv(nn.Module):
    """conv-bn-relu x2; state-dict keys double_conv.{0,1,3,4} (algorithms.py:398-413)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                                         nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))
        self.fold_bias = False
        self.cuda_bn = False

    def forward(self, x):
        dc = self.double_conv
        x = _conv_bn(dc[0], dc[1], x, self.fold_bias, True, self.cuda_bn)
        return _conv_bn(dc[3], dc[4], x, self.fold_bias, True, self.cuda_bn)


class _DoubleConvWT(nn.Module):
    """conv-relu-conv, no norm; keys double_conv.{0,2} (algorithms.py:416-428)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = _ConvActSeq(nn.Conv2d(cin, cout, 3, padding=1), nn.ReLU(inplace=True),
                                       nn.Conv2d(cout, cout, 3, padding=1))

    def forward(self, x):
        return self.double_conv(x)


class DeepWT(nn.Module):
    """Whitening feature extractor (algorithms.py:1080-1117): returns [z0, z1, relu(z1)]; z0 and z1 -- both
    PRE-activation, the InstanceNorm of the reference is constructed but commented out -- feed the
    whitening loss, the last one feeds the shape networks."""

    def __init__(self, input_channel, out_channel, whitening=True):
        super().__init__()
        self.whitening = whitening
        self.fused_loss = None            # see enable_relu_fusion / deepwt_forward
        self._pending_terms = _PendingTerms()
        if whitening:
            self.DoubleConv = _DoubleConvWT(input_channel, out_channel)
            self.DoubleConv2 = _DoubleConvWT(out_channel, out_channel)

    def forward(self, x):
        return deepwt_forward(self, x)


class _PendingTerms(dict):
    """id(embedding) -> (weakref to it, its loss terms).  Never copied or pickled with the module: the terms are
    non-leaf autograd tensors of ONE forward pass."""

    def __deepcopy__(self, memo):
        return _PendingTerms()

    def __reduce__(self):
        return (_PendingTerms, ())


def fused_config(wt_model):
    """The fused tail's loss configuration, read from the owning network AT CALL TIME (so later changes to
    ``owner.margin`` / ``owner.mmd_operator`` are honoured), or None when the fusion is off."""
    cfg = getattr(wt_model, "fused_loss", None)
    if cfg is None:
        return None
    owner = cfg["owner"]()
    if owner is None:
        return None
    op = owner.mmd_operator
    return {"fold": cfg["fold"], "n_per_domain": int(op.batch_size), "n_domains": int(op.domain_num),
            "margin": float(owner.margin), "eps": float(owner.eps)}


def deepwt_forward(self, x):
    """DeepWT.forward (algorithms.py:1091-1117, shape_networks.py:215-239), also bindable on the reference's own DeepWT
    instances (dropin.bind(..., fuse_relu=True)).

    With ``self.fused_loss`` set by the owning network each embedding z_k and the ``F.relu(z_k)`` that
    follows it take ONE pass over z_k forward (Gram + ReLU write, wtpse_whitening_relu_forward) and ONE backward
    (M_b z + ReLU backward + the sum of both gradients, wtpse_whitening_relu_backward).  The loss terms are parked in
    ``self._pending_terms`` (keyed by the embedding's identity, overwritten by the next forward); the owner's
    compute_whitening_loss(z_k) takes them from there instead of reading z_k again.  Nothing is attached to the tensors
    themselves: an attribute on z_k holding scalars whose grad_fn saves z_k would be a reference cycle through the
    autograd graph that Python's collector cannot see.  Values and gradients are those of the unfused sequence
    (tests/test_gpu_fusion.py)."""
    if not self.whitening:
        return [x]
    cfg = fused_config(self)
    z0 = self.DoubleConv(x)
    if cfg is None or not (torch.is_grad_enabled() and z0.requires_grad and z0.is_cuda):
        self._pending_terms = _PendingTerms()
        z1 = self.DoubleConv2(F.relu(z0))
        return [z0, z1, F.relu(z1)]
    fn = wf.relu_whitening_folded if cfg["fold"] else wf.relu_whitening_terms
    args = (cfg["n_per_domain"], cfg["n_domains"], cfg["margin"], cfg["eps"])
    r0, *terms0 = fn(z0, *args)
    z1 = self.DoubleConv2(r0)
    r1, *terms1 = fn(z1, *args)
    self._pending_terms = _PendingTerms({id(z0): (weakref.ref(z0), tuple(terms0)), id(z1): (weakref.ref(z1), tuple(terms1))})
    return [z0, z1, r1]


def enable_relu_fusion(owner, on=True):
    """Turn the DeepWT-tail fusion on/off for a WT_PSE or ShapeVariationalDist_x (ours or the reference's)."""
    fold = type(owner).__name__ == "WT_PSE"           # two-value form (algorithms.py:1301) vs three values
    owner.wt_model.fused_loss = None if not on else {"fold": fold, "owner": weakref.ref(owner)}
    owner.wt_model._pending_terms = _PendingTerms()
    return owner


def fused_terms(owner, z, arity):
    """Loss terms ``owner.wt_model``'s fused forward computed for the embedding z, or None.  Taken once: the entry is
    removed, so nothing keeps the autograd graph alive after the caller drops its own references."""
    table = getattr(getattr(owner, "wt_model", None), "_pending_terms", None)
    if not table:
        return None
    hit = table.get(id(z))
    if hit is None or hit[0]() is not z or len(hit[1]) != arity:
        return None
    del table[id(z)]
    return hit[1]


def _head(cin, mid, cout):
    return _ConvActSeq(nn.Conv2d(cin, cin, 1), nn.ReLU(), nn.Conv2d(cin, mid, 1), nn.ReLU(), nn.Conv2d(mid, cout, 1))


class ShapeVariationalDist_y_x(_UNetTrunk):
    """Teacher shape network p(shape | mask, features) (algorithms.py:979-1075)."""

    def __init__(self, hparams, device, n_channels, bilinear, n_classes, wt=True, prior=True, number_source_domain=3):
        super().__init__()
        self.device, self.prior, self.wt = device, prior, hparams["whitening"]
        self.number_source_domain = number_source_domain
        n = BASE_WIDTH
        if self.wt:
            self.inc = _DoubleConv(n_channels, n)
            self.fusion = _ConvActSeq(nn.Conv2d(2 * n, n, 1), nn.ReLU())
        else:
            self.inc = _DoubleConv(n_channels + 1, n)
        self._build_trunk(n)
        self.mu_prior = _head(2 * n, 8, n_classes)
        self.logvar_prior = _head(2 * n, 8, n_classes)

    def unet_extractor(self, inputs, mask):
        if self.wt:
            x1 = self.fusion(torch.cat([self.inc(mask), inputs], 1))
        else:
            x1 = self.inc(torch.cat([mask, inputs], 1))
        return self._trunk(x1)

    def sample_forward(self, inputs, mask=None, training=True):
        fm = self.unet_extractor(inputs, mask)
        mu = self.mu_prior(fm)
        if not training:
            return mu
        logvar = self.logvar_prior(fm)
        return self.reparameterization(mu, logvar), mu

    def reparameterization(self, mu, logvar):
        std = torch.exp(logvar / 2)
        return mu + std * torch.randn_like(std)          # algorithms.py:1068-1075


class attention_layer(nn.Module):  # noqa: N801  (reference name; key attention_layer.layer1.*)
    def __init__(self, channel1, channel2):
        super().__init__()
        self.layer1 = nn.Conv2d(channel1, channel2, kernel_size=1)

    def forward(self, x):
        x1 = self.layer1(x)
        return torch.sigmoid(x1), x1


class _MmdConfig:
    """Stands in for the reference's compute_MMD instance: only (domain_num, batch_size) are state."""

    def __init__(self, domain_num, batch_size):
        self.domain_num, self.batch_size = domain_num, batch_size

    def forward(self, inputs, **kwargs):
        from .mmd import mmd_penalty
        return mmd_penalty(inputs, self.batch_size, self.domain_num)


class WT_PSE(_UNetTrunk):
    """Segmentation network + shape regularisation (algorithms.py:1134-1353)."""

    def __init__(self, n_channels, n_classes, hparams, device, two_step, per_domain_batch=8, source_domain_num=3,
                 feature_dim=8, bilinear=True):
        super().__init__()
        self.n_channels, self.n_classes, self.device, self.hparams = n_channels, n_classes, device, hparams
        self.eps = 1e-5
        self.two_step, self.per_domain_batch, self.number_source_domain = two_step, per_domain_batch, source_domain_num
        self.whitening, self.cat_shape, self.margin = hparams["whitening"], hparams["cat_shape"], hparams["margin"]
        self.dim = BASE_WIDTH
        self.mmd_operator = _MmdConfig(domain_num=source_domain_num, batch_size=per_domain_batch)
        n = BASE_WIDTH
        self.wt_model = DeepWT(3, n, whitening=self.whitening)
        self.inc = ConvD(n_channels, n, first=True)
        self._build_trunk(n)
        fuse_dim = feature_dim
        if hparams["shape_prior"]:
            self.prior_dist = ShapeVariationalDist_y_x(hparams, device, 1, bilinear, n_classes=1, wt=self.whitening,
                                                       prior=True, number_source_domain=source_domain_num)
            if self.cat_shape:
                fuse_dim = feature_dim + 1
        self.mu = _ConvActSeq(nn.Conv2d(2 * n, 2 * n, 1), nn.ReLU(), nn.Conv2d(2 * n, feature_dim, 1))
        self.outc = nn.Sequential(nn.Conv2d(fuse_dim, n_classes, 1))
        self.attention_layer = attention_layer(1, 1)

    # -- backbone (PyTorch) -----------------------------------------------------------------------
    def embed(self, inputs):
        return self.mu(self._trunk(self.inc(inputs)))

    # -- hot path ---------------------------------------------------------------------------------
    def compute_whitening_loss(self, z):
        """(instance_loss, domain_loss) -- algorithms.py:1277-1309, one CUDA forward + one fused backward."""
        pre = fused_terms(self, z, 2)
        if pre is not None:
            return pre
        return wf.whitening_folded(z, self.mmd_operator.batch_size, self.mmd_operator.domain_num, float(self.margin),
                                   float(self.eps))

    def update(self, inputs, mask, step=0, plot_show=0, two_stage_inputs=None, sp_mask=None, two_step=False):
        embedding = self.embed(inputs)
        attention_mask = 0
        feats = None
        if self.hparams["shape_prior"]:
            feats = self.wt_model(two_stage_inputs if two_step else inputs)
            z_post, _z_post_mu = self.prior_dist.sample_forward(feats[-1], mask, training=True)
            if self.hparams["shape_attention"]:
                lay = self.attention_layer.layer1
                embedding_f, attention_mask = attention_fuse(embedding, z_post, lay.weight, lay.bias,
                                                             self.hparams["shape_attention_coeffient"])
            else:
                embedding_f = embedding
            embedding = torch.cat([embedding_f, z_post], 1) if self.cat_shape else embedding_f
        instance_wt_loss, domain_wt_loss = 0, 0
        if self.hparams["whitening"]:
            num_embeddings = len(feats)
            for k in range(num_embeddings - 1):                       # two embeddings ...
                ins_k, dom_k = self.compute_whitening_loss(feats[k])
                instance_wt_loss = instance_wt_loss + ins_k
                domain_wt_loss = domain_wt_loss + dom_k
            instance_wt_loss = instance_wt_loss / num_embeddings      # ... divided by three (algorithms.py:1266-1267)
            domain_wt_loss = domain_wt_loss / num_embeddings
        output = self.outc(embedding)
        if self.hparams["shape_prior"]:
            return output, attention_mask, attention_mask, instance_wt_loss, domain_wt_loss
        return output, 0, 0, 0, 0

    def predict(self, learn_x_network, inputs_all):
        if self.two_step:
            inputs, two_stage_inputs = inputs_all[0], inputs_all[1]
        else:
            inputs = two_stage_inputs = inputs_all
        embedding = self.embed(inputs)
        pre_sigmoid = None
        if self.hparams["shape_prior"]:
            feats = learn_x_network.wt_model(two_stage_inputs)
            z_post = learn_x_network.sample_forward(feats[-1], training=False)
            if self.hparams["shape_attention"]:
                att, pre_sigmoid = self.attention_layer(z_post)
                fuse = self.hparams["shape_attention_coeffient"] * embedding + att * embedding
            else:
                fuse = embedding
            embedding = torch.cat([fuse, z_post], 1) if self.cat_shape else fuse
        return self.outc(embedding), pre_sigmoid


class ShapeVariationalDist_x(_UNetTrunk):
    """Student shape network p(shape | features) with its distillation update (shape_networks.py:415-597)."""

    def __init__(self, hparams, device, n_classes, number_source_domain=3, batch_size=3):
        super().__init__()
        self.device, self.batch_size, self.hparams = device, batch_size, hparams
        self.wt = self.whitening = hparams["whitening"]
        self.number_source_domain = number_source_domain
        self.eps, self.margin, self.dim = 1e-5, hparams["margin"], BASE_WIDTH
        n = BASE_WIDTH
        self.wt_model = DeepWT(3, n, whitening=self.whitening)
        if not self.wt:
            self.inc = _DoubleConv(3, n)
        self._build_trunk(n)
        self.mmd_operator = _MmdConfig(domain_num=3, batch_size=batch_size)      # literal 3, shape_networks.py:448
        self.mu_prior = _head(2 * n, 8, n_classes)
        self.logvar_prior = _head(2 * n, 8, n_classes)
        # True = the reference's behaviour: the KD loss back-propagates into the teacher (`main_network`), whose
        # gradients the trainer throws away at its next zero_grad (shape_networks.py:524, Trainer.py:768).  False runs
        # the teacher forward (BatchNorm statistics still update) without recording it: every student gradient and
        # every weight after the optimizer steps is unchanged, one U-Net backward per update is saved.
        self.teacher_grad = True

    def unet_extractor(self, inputs):
        return self._trunk(inputs if self.wt else self.inc(inputs))

    @staticmethod
    def _scrub_nan(t):
        # shape_networks.py:490-492 / :504-506: `if isnan(t).any(): t = nan_to_num(t); t[t == inf] = 0`.
        # Same result without the host sync of the python `if`: the whole tensor is replaced only when it
        # contains a NaN (the follow-up `t[t == inf] = 0` can never fire after nan_to_num).
        return torch.where(torch.isnan(t).any(), torch.nan_to_num(t), t)

    def sample_forward(self, inputs, training):
        fm = self.unet_extractor(inputs)
        mu = self._scrub_nan(self.mu_prior(fm))
        if not training:
            return mu
        logvar = self.logvar_prior(fm)
        return self.reparameterization(mu, logvar), mu

    def reparameterization(self, mu, logvar):
        std = self._scrub_nan(torch.exp(logvar / 2))
        # shape_networks.py:507-509: `torch.normal(mu, std) * std + mu`.  torch.normal(mean, std) draws randn * std + mean
        # from the same generator stream and passes no gradient to its arguments; written out
        # that way it has no host-side `std >= 0` check, so the iteration stays CUDA-graph capturable.
        sampled = (torch.randn_like(std) * std + mu).detach()
        return sampled * std + mu

    def compute_whitening_loss(self, z):
        """(off_diagonal_loss, diagonal_loss, domain_loss) -- shape_networks.py:561-594."""
        pre = fused_terms(self, z, 3)
        if pre is not None:
            return pre
        return wf.whitening_terms(z, self.mmd_operator.batch_size, self.mmd_operator.domain_num, float(self.margin),
                                  float(self.eps))

    def wasser_distance(self, prior_space_mu, posterior_space_mu):
        return wf.kd_mse(prior_space_mu, posterior_space_mu)            # nn.MSELoss(reduction='mean'), :596-597

    def update(self, main_network, inputs, mask, step=0, plot_show=0, two_stage_inputs=None, two_step=False):
        if not self.hparams["whitening"]:
            return 0, 0, 0, 0, 0
        x = two_stage_inputs if two_step else inputs
        # teacher: features + mask (its mean is NOT detached: gradients reach main_network and are discarded
        # by the caller's zero_grad, shape_networks.py:524); student: features only
        with torch.set_grad_enabled(self.teacher_grad and torch.is_grad_enabled()):
            teacher_feats = main_network.wt_model(x)
            _z_post, z_post_mu = main_network.prior_dist.sample_forward(teacher_feats[-1], mask, training=True)
        student_feats = self.wt_model(x)
        _z_pre, z_pre_mu = self.sample_forward(student_feats[-1], training=True)
        kd_loss = self.wasser_distance(z_post_mu, z_pre_mu)
        # (the two attention_layer forwards at shape_networks.py:531-535 have no effect on any output)
        off_sum, diag_last, dom_sum = 0, 0, 0
        num_embeddings = len(student_feats)
        for k in range(num_embeddings - 1):
            off_k, diag_k, dom_k = self.compute_whitening_loss(student_feats[k])
            off_sum = off_sum + off_k
            diag_last = diag_k + diag_k           # `a, instance_wt_loss2, b = f(); instance_wt_loss2 += instance_wt_loss2`
            dom_sum = dom_sum + dom_k
        instance_ij = off_sum / num_embeddings
        instance_ii = diag_last / num_embeddings
        domain = dom_sum / num_embeddings
        return kd_loss, instance_ij + instance_ii, instance_ij, instance_ii, domain
