"""wt-pse-code_b200: B200-native (sm_100a CUDA) shape-regularization hot path of WT-PSE.

Import as ``wtpse_b200`` (alias package at the repo root).  The arithmetic lives in
``libwtpse_b200.so`` (C ABI: include/wtpse_b200.h); this package is the thin PyTorch-facing host
side.  There is no CPU or eager-PyTorch fallback: calls raise if the library is missing.
"""
from . import _build, _lib  # noqa: F401
from .functional import (HostPlan, gram_matrix, kd_mse, relu_whitening_folded, relu_whitening_terms,  # noqa: F401
                         whitening_folded, whitening_terms)
from .mmd import mmd_penalty  # noqa: F401
from .elementwise import attention_fuse, od_roi, prepare_batch, upsample2x  # noqa: F401
from .wavelet import dwt2d, idwt2d, wavelet_shape_loss  # noqa: F401
from . import dropin  # noqa: F401
from . import dp, segmentation, synthetic, train_step  # noqa: F401
from .segmentation import ShapeVariationalDist_x, WT_PSE  # noqa: F401
from .train_step import TrainStep  # noqa: F401

__all__ = ["whitening_terms", "whitening_folded", "relu_whitening_terms", "relu_whitening_folded", "gram_matrix", "kd_mse", "mmd_penalty", "HostPlan", "dropin",
           "prepare_batch", "od_roi", "attention_fuse", "WT_PSE", "ShapeVariationalDist_x", "TrainStep"]
