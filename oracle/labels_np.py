"""numpy restatement of the integer OC/OD label path and the coarse-to-fine ROI step.

TEST INFRASTRUCTURE (see oracle/__init__.py).

  * raw uint8 mask -> {0,1} OD / OC labels : custom_transforms.py:466-499 (Normalize_tf.__call__)
    and the test-time twin fundus_dataloader.py:112-134.
  * od_pred = (sigmoid(logits) > 0.75) and image_roi = (image + 1) * od_pred - 1 : Trainer.py:842-853.
  * oc_pos_weight = sum(od_pred) / sum(od_pred * target_oc), 1.0 when inf/nan : Trainer.py:865-867.
"""
import numpy as np


def trilevel(raw):
    """custom_transforms.py:473-477: 255 above 200, 128 in (50, 201), 0 otherwise (float64 image)."""
    raw = np.asarray(raw, dtype=np.uint8)
    out = np.zeros(raw.shape, dtype=np.float64)
    out[raw > 200] = 255
    out[(raw > 50) & (raw < 201)] = 128
    return out


def labels_from_raw(raw_od, raw_oc=None):
    """Returns (label_od, label_oc) as uint8 HxWx1 arrays, exactly as Normalize_tf leaves them.

    Quirk kept (custom_transforms.py:493-494): label_oc is derived from the tri-level image of the
    *OD* raw mask, not of raw_oc -- raw_oc only supplies the output buffer."""
    raw_od = np.array(raw_od, dtype=np.uint8)
    raw_oc = raw_od.copy() if raw_oc is None else np.array(raw_oc, dtype=np.uint8)
    tri = trilevel(raw_od)
    od = raw_od.copy()
    od[tri < 255] = 1
    od[tri == 255] = 0
    oc = raw_oc.copy()
    oc[tri > 0] = 0
    oc[tri == 0] = 1
    return od[..., None], oc[..., None]


def normalize_image(img_u8):
    """custom_transforms.py:468-472: float32(img) / 127.5 - 1.0 (two separate fp32 roundings)."""
    x = np.asarray(img_u8).astype(np.float32)
    x /= np.float32(127.5)
    x -= np.float32(1.0)
    return x


def sigmoid_f32(x):
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x, dtype=np.float32))).astype(np.float32)


def od_threshold(logits, thr=0.75):
    """Trainer.py:842 -- (sigmoid(output) > 0.75).float()"""
    return (sigmoid_f32(logits) > np.float32(thr)).astype(np.float32)


def od_threshold_ambiguous(logits, thr=0.75):
    """True where the exact sigmoid lies within 2 fp32 ulps of the threshold.  There the comparison
    depends on the expf implementation (torch-CPU/sleef, numpy and CUDA libdevice differ in the last
    bit), so no CPU restatement can be bit-authoritative; on the GPU box the parity test compares
    against the reference statement itself, ``torch.sigmoid(x) > 0.75``, executed on the same device."""
    s = 1.0 / (1.0 + np.exp(-np.asarray(logits, dtype=np.float64)))
    return np.abs(s - thr) <= 2.0 * 2.0 ** -24


def roi_image(image, od_pred):
    """Trainer.py:850-852 -- image += 1 ; image_roi = image * od_pred ; image_roi -= 1 (fp32, three roundings)."""
    t = (np.asarray(image, dtype=np.float32) + np.float32(1.0)).astype(np.float32)
    t = (t * np.asarray(od_pred, dtype=np.float32)).astype(np.float32)
    return (t - np.float32(1.0)).astype(np.float32)


def oc_pos_weight(od_pred, target_oc):
    """Trainer.py:865-867."""
    num = np.float32(np.sum(od_pred, dtype=np.float64))
    den = np.float32(np.sum(od_pred * target_oc, dtype=np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.float32(num) / np.float32(den)
    if np.isinf(w) or np.isnan(w):
        w = np.float32(1.0)
    return w
