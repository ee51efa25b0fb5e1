"""Drop-in layer: the reference's own method surface for the shape-regularization path.

The reference has no plugin/FFI interface; its boundary is the set of Python methods that
``Trainer.train_epoch`` reaches.  The functions below have the reference's names,
argument meaning, return arity and return conventions, and ``install()`` rebinds them on the
reference's classes so ``Trainer.py`` / ``train.py`` run unmodified:

    compute_whitening_loss(self, z)            WT_PSE                   algorithms.py:1277   -> (ins, dom)
    compute_whitening_loss(self, z)            ShapeVariationalDist_x   shape_networks.py:561 -> (off, diag, dom)
    wasser_distance(self, a, b)                ShapeVariationalDist_x   shape_networks.py:596 -> mse
    compute_MMD.forward(self, inputs)          both modules             algorithms.py:102 / shape_networks.py:283

Outputs are 0-dim CUDA tensors attached to autograd (Trainer calls ``.item()`` on them and adds
them into ``loss_main``).  No parameters or buffers are introduced, so checkpoints keep their keys.
The dead ``penalty.item()`` host sync of algorithms.py:118-119 is not reproduced, nor are the
per-call H2D copies of ``eye``/``triu_indices`` (algorithms.py:1296,1305).
"""
import types

from . import functional as F


def _domain_cfg(obj):
    op = obj.mmd_operator
    return int(op.batch_size), int(op.domain_num)


def _fused(obj, z, arity):
    from .segmentation import fused_terms
    return fused_terms(obj, z, arity)


def wt_pse_compute_whitening_loss(self, z):
    """Replacement for WT_PSE.compute_whitening_loss (algorithms.py:1277-1309)."""
    pre = _fused(self, z, 2)           # parked by the fused DeepWT tail (bind(..., fuse_relu=True)), if that is on
    if pre is not None:
        return pre
    n, K = _domain_cfg(self)
    return F.whitening_folded(z, n, K, float(self.margin), float(self.eps))


def shape_compute_whitening_loss(self, z):
    """Replacement for ShapeVariationalDist_x.compute_whitening_loss (shape_networks.py:561-594)."""
    pre = _fused(self, z, 3)
    if pre is not None:
        return pre
    n, K = _domain_cfg(self)          # K is the literal 3 of shape_networks.py:448
    return F.whitening_terms(z, n, K, float(self.margin), float(self.eps))


def shape_wasser_distance(self, prior_space_mu, posterior_space_mu):
    """Replacement for ShapeVariationalDist_x.wasser_distance (shape_networks.py:596-597)."""
    return F.kd_mse(prior_space_mu, posterior_space_mu)


def mmd_forward(self, inputs, **kwargs):
    """Replacement for compute_MMD.forward (algorithms.py:102-121): inputs is B x 120."""
    from . import mmd
    return mmd.mmd_penalty(inputs, int(self.batch_size), int(self.domain_num))


def install(algorithms=None, shape_networks=None):
    """Rebind the hot-path methods on the reference's classes (pass the imported modules).

    Returns a dict of the original attributes so ``uninstall`` can restore them."""
    saved = {}
    if algorithms is not None:
        saved[(algorithms.WT_PSE, "compute_whitening_loss")] = algorithms.WT_PSE.compute_whitening_loss
        algorithms.WT_PSE.compute_whitening_loss = wt_pse_compute_whitening_loss
        saved[(algorithms.compute_MMD, "forward")] = algorithms.compute_MMD.forward
        algorithms.compute_MMD.forward = mmd_forward
    if shape_networks is not None:
        cls = shape_networks.ShapeVariationalDist_x
        saved[(cls, "compute_whitening_loss")] = cls.compute_whitening_loss
        cls.compute_whitening_loss = shape_compute_whitening_loss
        saved[(cls, "wasser_distance")] = cls.wasser_distance
        cls.wasser_distance = shape_wasser_distance
        saved[(shape_networks.compute_MMD, "forward")] = shape_networks.compute_MMD.forward
        shape_networks.compute_MMD.forward = mmd_forward
    return saved


def uninstall(saved):
    for (cls, name), fn in saved.items():
        setattr(cls, name, fn)


def bind(obj, fuse_relu=False):
    """Rebind on ONE instance (e.g. only the OD model) instead of the class.

    fuse_relu=True additionally rebinds ``obj.wt_model.forward`` (DeepWT.forward, algorithms.py:1091-1117 /
    shape_networks.py:215-239) to the fused tail: Gram + ReLU in one pass over each embedding, and one backward pass
    (SURVEY.md 8(f).1).  Same outputs, same gradients; no parameters or buffers change."""
    name = type(obj).__name__
    if fuse_relu and name in ("WT_PSE", "ShapeVariationalDist_x"):
        from . import segmentation as seg
        seg.enable_relu_fusion(obj, True)
        obj.wt_model.forward = types.MethodType(seg.deepwt_forward, obj.wt_model)
    if name == "WT_PSE":
        obj.compute_whitening_loss = types.MethodType(wt_pse_compute_whitening_loss, obj)
    elif name == "ShapeVariationalDist_x":
        obj.compute_whitening_loss = types.MethodType(shape_compute_whitening_loss, obj)
        obj.wasser_distance = types.MethodType(shape_wasser_distance, obj)
    else:
        raise TypeError("bind() expects a WT_PSE or ShapeVariationalDist_x instance, got %s" % name)
    return obj
