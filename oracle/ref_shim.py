"""Import shim for the *unmodified* reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference modules are imported from ``/root/reference`` when that
tree exists (build container) and otherwise from ``oracle/_ref`` -- the byte-for-byte copy ``oracle/build_ref.py``
makes at build time, which is git-ignored but travels to the GPU box with the snapshot; its SHA-256s are checked
against the committed ``oracle/ref_manifest.json`` before the import, so whatever runs is the reference as published.

The reference is a flat script collection whose hot-path modules import packages that are not installed in this
image and are unused on this path (``matplotlib`` at algorithms.py:16 / shape_networks.py:16 /
custom_transforms.py:9,13 and ``torchfile`` at algorithms.py:11).  They get empty stand-ins.  The reference also
hard-codes ``.cuda()`` (algorithms.py:1162-1164,1296,1305; shape_networks.py:449-455,581,590): on a box with a GPU
those calls run as written; on a CPU-only box ``Tensor.cuda`` is made the identity.
"""
import contextlib
import importlib
import os
import sys
import types

from . import build_ref

REFERENCE_ROOT = build_ref.REFERENCE_ROOT
_LOADED = None


def root():
    """Directory the reference modules are imported from, or None."""
    if build_ref.reference_present():
        return REFERENCE_ROOT
    if build_ref.vendored_present():
        return build_ref.VENDOR_DIR
    return None


def available():
    return root() is not None


def _stub(name, **attrs):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
    return sys.modules[name]


def load():
    """Return (algorithms, shape_networks, custom_transforms) of the reference."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    src = root()
    if src is None:
        raise RuntimeError("reference modules not found: neither %s nor %s (run `python -m oracle.build_ref` in the "
                           "build container)" % (REFERENCE_ROOT, build_ref.VENDOR_DIR))
    build_ref.verify(src)                         # unmodified, or refuse
    import torch

    try:
        importlib.import_module("matplotlib.pyplot")
    except Exception:
        noop = lambda *a, **k: None
        pyplot = _stub("matplotlib.pyplot", imshow=noop, imsave=noop)
        _stub("matplotlib", pyplot=pyplot)
    try:
        importlib.import_module("torchfile")
    except Exception:
        _stub("torchfile")
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if src not in sys.path:
        sys.path.insert(0, src)
    _LOADED = tuple(importlib.import_module(m) for m in ("algorithms", "shape_networks", "custom_transforms"))
    return _LOADED


@contextlib.contextmanager
def cpu_only():
    """Run the reference on HOST tensors on a box that has a GPU (bench.py's cpu_baseline / --impl reference legs).

    The reference hard-codes ``.cuda()`` for its loss constants (algorithms.py:1162-1164, 1296, 1305); with CPU inputs
    those calls would move one operand to the device and the next operator would fail.  Inside this context
    ``Tensor.cuda`` is the identity -- exactly what load() installs on a CPU-only box, and what the golden vectors
    were generated under -- and it is restored on exit.  Construct the model AND call it inside the context."""
    import torch

    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = saved


DEFAULT_HPARAMS = {
    # hparams_registry.py:75-93 (WT_PSE defaults)
    "whitening": True, "margin": 0, "shape_prior": True, "shape_attention": True,
    "cat_shape": False, "shape_attention_coeffient": 0.3, "shape_start": 0.5,
    "instance_wt_gm": 1, "domain_wt_gm": 1, "multi-turn": 1,
}
