"""Import shim for the *unmodified* reference at /root/reference (build container only).

The reference is a flat script collection whose hot-path modules import packages that
are not installed here and are unused on this path (``matplotlib`` at
algorithms.py:16 / shape_networks.py:16 / custom_transforms.py:9,13 and ``torchfile`` at
algorithms.py:11).  They get empty stand-ins.  The reference also hard-codes ``.cuda()``
(algorithms.py:1162-1164,1296,1305; shape_networks.py:449-455,581,590); on a CPU-only
box that call is made the identity for the duration of the import/use.

Nothing here travels to the GPU box: /root/reference does not exist there.  Only
``oracle/make_golden.py`` and the (skipped-when-absent) CPU cross-checks use it.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "algorithms.py"))


def _stub(name, **attrs):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
    return sys.modules[name]


def load():
    """Return (algorithms, shape_networks, custom_transforms) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch

    noop = lambda *a, **k: None
    pyplot = _stub("matplotlib.pyplot", imshow=noop, imsave=noop)
    _stub("matplotlib", pyplot=pyplot)
    _stub("torchfile")
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mods = [importlib.import_module(m) for m in ("algorithms", "shape_networks", "custom_transforms")]
    return tuple(mods)


DEFAULT_HPARAMS = {
    # hparams_registry.py:75-93 (WT_PSE defaults)
    "whitening": True, "margin": 0, "shape_prior": True, "shape_attention": True,
    "cat_shape": False, "shape_attention_coeffient": 0.3, "shape_start": 0.5,
    "instance_wt_gm": 1, "domain_wt_gm": 1, "multi-turn": 1,
}
