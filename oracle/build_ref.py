"""Vendor the UNMODIFIED reference modules of the hot path into ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.build_ref

The reference is pure Python: there is nothing to compile, so "building" ``oracle/_ref`` means a byte-for-byte
copy of the three modules the path lives in, made in the build container where ``/root/reference`` exists, so
that they travel to the GPU box with the snapshot the same way the built ``.so`` does.  Nothing under
``oracle/_ref`` enters the git history (``.gitignore``), no reference source is committed.  What IS committed is
``oracle/ref_manifest.json``: the SHA-256 of each file as it lies under ``/root/reference``.  ``verify()`` checks a
vendored tree against it, so a GPU test that imports ``oracle/_ref`` proves it ran the unmodified reference.

Called by ``__graft_entry__.build()`` whenever ``/root/reference`` is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
VENDOR_DIR = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "ref_manifest.json")
# the modules the hot path lives in; Trainer.py / train.py need sconf, tensorboardX, medpy, pytz and
# skimage (all absent from the image) and are represented by oracle/ref_iteration.py instead
FILES = ("algorithms.py", "shape_networks.py", "custom_transforms.py")


def _sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def reference_present():
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, f)) for f in FILES)


def vendored_present():
    return all(os.path.isfile(os.path.join(VENDOR_DIR, f)) for f in FILES)


def manifest():
    with open(MANIFEST) as f:
        return json.load(f)


def verify(root=VENDOR_DIR):
    """Raise unless every module under ``root`` is byte-identical to the reference the manifest was taken from."""
    want = manifest()["sha256"]
    for name in FILES:
        got = _sha256(os.path.join(root, name))
        if got != want[name]:
            raise RuntimeError("%s differs from the reference (sha256 %s, manifest %s)" % (os.path.join(root, name), got, want[name]))
    return True


def build(write_manifest=False):
    """Copy the modules; returns the vendored directory.  ``write_manifest=True`` (maintainer action, build container
    only) refreshes oracle/ref_manifest.json from /root/reference."""
    if not reference_present():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if write_manifest or not os.path.exists(MANIFEST):
        with open(MANIFEST, "w") as f:
            json.dump({"source": REFERENCE_ROOT, "note": "sha256 of the reference modules oracle/_ref must be a copy of",
                       "sha256": {n: _sha256(os.path.join(REFERENCE_ROOT, n)) for n in FILES}}, f, indent=1, sort_keys=True)
            f.write("\n")
    os.makedirs(VENDOR_DIR, exist_ok=True)
    for name in FILES:
        dst = os.path.join(VENDOR_DIR, name)
        src = os.path.join(REFERENCE_ROOT, name)
        if not os.path.exists(dst) or _sha256(dst) != _sha256(src):
            shutil.copyfile(src, dst)
    verify(VENDOR_DIR)
    return VENDOR_DIR


if __name__ == "__main__":
    print(build(write_manifest="--write-manifest" in sys.argv))
