"""ctypes binding of the C ABI declared in include/wtpse_b200.h.

There is NO fallback: if ``libwtpse_b200.so`` is missing or a call fails, an exception is raised.
"""
import ctypes
import os

from . import _build

_c = ctypes
_LIB = None

WTPSE_OK = 0
ABI_VERSION = 2

EXPORTS = {
    # name: (restype, argtypes)
    "wtpse_abi_version": (_c.c_int, []),
    "wtpse_last_error": (_c.c_char_p, []),
    "wtpse_sm_count": (_c.c_int, []),
    "wtpse_whitening_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int64]),
    "wtpse_whitening_ticket_bytes": (_c.c_size_t, [_c.c_int]),
    # z, B, C, P, n, K, margin, eps, losses, gram, rowstat, domgrad, workspace, workspace_bytes, stream
    "wtpse_whitening_forward": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_float,
                                           _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                           _c.c_size_t, _c.c_void_p]),
    # z, gram, rowstat, domgrad, g_off, g_diag, g_dom, B, C, P, n, K, dz, stream
    "wtpse_whitening_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                            _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_void_p,
                                            _c.c_void_p]),
    # z, relu_out, B, C, P, n, K, margin, eps, losses, gram, rowstat, domgrad, workspace, workspace_bytes, stream
    "wtpse_whitening_relu_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_int,
                                                _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    # z, grad_relu, gram, rowstat, domgrad, g_off, g_diag, g_dom, B, C, P, n, K, dz, stream
    "wtpse_whitening_relu_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                 _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64,
                                                 _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "wtpse_upsample2x_nhwc": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "wtpse_bias_act_nhwc": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_void_p]),
    "wtpse_channel_sum_workspace_bytes": (_c.c_size_t, [_c.c_int64, _c.c_int]),
    "wtpse_channel_sum_nhwc": (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_batchnorm_workspace_bytes": (_c.c_size_t, [_c.c_int64, _c.c_int]),
    "wtpse_batchnorm_relu_forward": (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_float,
                                                _c.c_float, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                _c.c_size_t, _c.c_void_p]),
    "wtpse_batchnorm_relu_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                 _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_maxpool2_nhwc": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "wtpse_whitening_forward_cl": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_int,
                                              _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                              _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_whitening_backward_cl": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                               _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64,
                                               _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "wtpse_relu_backward_channel_sum_nhwc": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_void_p,
                                                        _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_mmd_workspace_bytes": (_c.c_size_t, [_c.c_int]),
    "wtpse_mmd_forward": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                     _c.c_size_t, _c.c_void_p]),
    "wtpse_mmd_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                      _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_mse_workspace_bytes": (_c.c_size_t, [_c.c_int64]),
    "wtpse_mse_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_size_t,
                                     _c.c_void_p]),
    "wtpse_mse_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p,
                                      _c.c_void_p]),
    "wtpse_prepare_batch": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                       _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "wtpse_od_roi_workspace_bytes": (_c.c_size_t, []),
    "wtpse_od_roi": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                _c.c_int64, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_attention_fuse_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int64]),
    "wtpse_attention_fuse_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_float, _c.c_int, _c.c_int,
                                                _c.c_int64, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                _c.c_void_p]),
    "wtpse_attention_fuse_backward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                 _c.c_float, _c.c_int, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_void_p,
                                                 _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_wavelet_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    "wtpse_dwt2d_forward": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                       _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_dwt2d_inverse": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                       _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_wavelet_loss_forward": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                              _c.POINTER(_c.c_float), _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                              _c.c_size_t, _c.c_void_p]),
    "wtpse_profile_enable": (None, [_c.c_int]),
    "wtpse_profile_reset": (None, []),
    "wtpse_profile_kernel_count": (_c.c_int, []),
    "wtpse_profile_kernel_name": (_c.c_char_p, [_c.c_int]),
    "wtpse_profile_launches": (_c.c_longlong, [_c.c_int]),
    "wtpse_profile_read": (_c.c_int, [_c.c_int, _c.POINTER(_c.c_longlong), _c.POINTER(_c.c_double)]),
    "wtpse_wavelet_resident_cluster": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    "wtpse_wavelet_loss_resident": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                               _c.POINTER(_c.c_float), _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                               _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "wtpse_scale_unless_one": (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p]),
    # include/wtpse_b200_debug.h (diagnostics: not part of the product ABI)
    "wtpse_debug_set": (_c.c_int, [_c.c_char_p, _c.c_int]),
    "wtpse_debug_get": (_c.c_int, [_c.c_char_p, _c.POINTER(_c.c_int)]),
    "wtpse_debug_pdl_slow_copy": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p]),
    "wtpse_host_plan_create": (_c.c_int, [_c.c_int, _c.c_int64, _c.POINTER(_c.c_void_p)]),
    "wtpse_host_plan_destroy": (None, [_c.c_void_p]),
    "wtpse_host_plan_run": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_float, _c.c_float,
                                       _c.POINTER(_c.c_float), _c.POINTER(_c.c_float), _c.c_void_p]),
    "wtpse_host_plan_submit": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_float, _c.c_float,
                                          _c.POINTER(_c.c_float), _c.POINTER(_c.c_float), _c.c_void_p]),
    "wtpse_host_plan_wait": (_c.c_int, [_c.c_void_p]),
}


class WtpseError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (once) and return the ctypes handle; raises if the library is absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise WtpseError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.wtpse_abi_version() != ABI_VERSION:
        raise WtpseError("ABI mismatch: library %d, binding %d" % (lib.wtpse_abi_version(), ABI_VERSION))
    _LIB = lib
    return lib


def debug_set(name, value):
    """Diagnostic switch of include/wtpse_b200_debug.h (tests, bench.py, tools -- never the product path)."""
    check(load().wtpse_debug_set(name.encode(), int(value)))


def debug_get(name):
    out = _c.c_int(0)
    check(load().wtpse_debug_get(name.encode(), _c.byref(out)))
    return out.value


def check(rc):
    if rc != WTPSE_OK:
        msg = load().wtpse_last_error()
        raise WtpseError("wtpse call failed (code %d): %s" % (rc, msg.decode() if msg else "?"))
