"""Build ``libwtpse_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every ``csrc/*.cu`` is compiled to an object under ``build/`` (git-ignored) only when it or a header changed, several at a
time, then linked; an edit to one kernel file costs one compile.
"""
import concurrent.futures
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libwtpse_b200.so")
SOURCES = ["api.cu", "whitening_gram.cu", "whitening_epilogue.cu", "whitening_apply.cu", "whitening_apply_relu.cu", "whitening_apply_cl.cu",
           "whitening_cl_tma.cu", "mse.cu", "profile.cu", "elementwise.cu", "backbone_elementwise.cu", "batchnorm.cu",
           "wavelet.cu", "wavelet_resident.cu", "wavelet_stream.cu", "wavelet_tiles.cu", "wavelet_db2.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libwtpse_b200.so")
    return exe


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(os.path.dirname(PKG_DIR), "include", h) for h in ("wtpse_b200.h", "wtpse_b200_debug.h")]
    return [h for h in hs if os.path.exists(h)]


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB_PATH, sources() + _headers())


def build(force=False, verbose=False, extra_flags=()):
    if not force and not extra_flags and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc, headers = _nvcc(), _headers()
    flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
    todo = [s for s in sources() if force or extra_flags or _stale(_obj(s), [s] + headers)]

    def compile_one(src):
        return src, subprocess.run([nvcc] + flags + ["-c", src, "-o", _obj(src)], capture_output=True, text=True)

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        for src, res in pool.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (os.path.basename(src), res.stdout, res.stderr))
            if verbose:
                print(os.path.basename(src), res.stderr)
    res = subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + [_obj(s) for s in sources()], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
