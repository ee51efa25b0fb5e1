"""Pin the oracle (oracle/*.py) to the golden vectors produced by the unmodified reference.

CPU-only.  Tolerances: 1e-5 relative for Gram entries, instance terms, KD and
gradients (max-abs normalised); the domain (MMD) scalar is compared with
|delta| <= 1e-5 * max(|ref|, 1) because Kxx + Kyy - 2Kxy cancels (each K is O(1)).
"""
import numpy as np
import pytest
import torch

from conftest import golden, golden_files, rel_err
from oracle import labels_np, whitening_np as wnp, whitening_torch as wt

TOL = 1e-5


def _close(a, b, tol=TOL, scale=None):
    if np.isnan(b):
        return np.isnan(a)
    s = max(abs(b), scale or 0.0, 1e-300)
    return abs(a - b) <= tol * s


@pytest.mark.parametrize("fname", golden_files("whitening_"))
def test_numpy_closed_form_matches_reference(fname):
    g = golden(fname)
    z, n, K, margin = g["z"], int(g["n"]), int(g["K"]), float(g["margin"])
    w = g["weights"]
    f = wnp.whitening_forward(z, n, K, margin, 1e-5, np.float64)
    assert rel_err(f["gram"], g["f64_gram"]) < 1e-12
    assert rel_err(f["gram"], g["f32_gram"]) < TOL
    for ref in ("f64_", "f32_"):
        assert _close(f["off"] + f["diag"], float(g[ref + "wt_ins"]))
        assert _close(f["dom"], float(g[ref + "wt_dom"]), scale=1.0)
    dz, _ = wnp.whitening_backward(z, f, n, K, w[0], w[0], w[2])
    assert rel_err(dz, g["f64_wt_dz"]) < 1e-6          # stored rounded to fp32
    assert rel_err(dz, g["f32_wt_dz"]) < TOL
    # three-value form: the shape network hard-codes 3 domains (shape_networks.py:448)
    f3 = wnp.whitening_forward(z, n, 3, margin, 1e-5, np.float64)
    assert _close(f3["off"], float(g["f64_sh_off"]))
    assert _close(f3["diag"], float(g["f64_sh_diag"]))
    assert _close(f3["dom"], float(g["f64_sh_dom"]), scale=1.0)
    if not np.isnan(float(g["f64_sh_dom"])):
        dz3, _ = wnp.whitening_backward(z, f3, n, 3, w[0], w[1], w[2])
        assert rel_err(dz3, g["f32_sh_dz"]) < TOL


@pytest.mark.parametrize("fname", golden_files("whitening_"))
def test_torch_restatement_matches_reference(fname):
    g = golden(fname)
    z = torch.from_numpy(g["z"]).requires_grad_(True)
    n, K, margin = int(g["n"]), int(g["K"]), float(g["margin"])
    w = g["weights"]
    ins, dom = wt.wt_pse_whitening_loss(z, n, K, margin)
    (float(w[0]) * ins + float(w[2]) * dom).backward()
    assert _close(float(ins), float(g["f32_wt_ins"]))
    assert _close(float(dom), float(g["f32_wt_dom"]), scale=1.0)
    assert rel_err(z.grad.numpy(), g["f32_wt_dz"]) < TOL
    off, diag, dom3 = wt.shape_whitening_loss(torch.from_numpy(g["z"]), n, margin)
    assert _close(float(off), float(g["f32_sh_off"]))
    assert _close(float(diag), float(g["f32_sh_diag"]))
    assert _close(float(dom3), float(g["f32_sh_dom"]), scale=1.0)


@pytest.mark.parametrize("fname", golden_files("mmd_"))
def test_mmd_matches_reference(fname):
    g = golden(fname)
    v, n, K = g["v"].astype(np.float64), int(g["n"]), int(g["K"])
    loss, E, D = wnp.mmd_forward(v, n, K)
    dv = wnp.mmd_backward(v, E, D, n, K)
    for tag in ("alg", "sn"):
        assert _close(loss, float(g[tag + "_f64_loss"]), tol=1e-12, scale=1.0)
        assert _close(loss, float(g[tag + "_f32_loss"]), scale=1.0)
        assert rel_err(dv, g[tag + "_f64_dv"]) < 1e-10
        assert rel_err(dv, g[tag + "_f32_dv"]) < 1e-4     # fp32 addmm-form cancellation in the reference itself
    x = torch.from_numpy(g["v"]).requires_grad_(True)
    out = wt.mmd_penalty(x, n, K)
    out.backward()
    assert _close(float(out), float(g["alg_f32_loss"]), scale=1.0)


@pytest.mark.parametrize("fname", golden_files("mse_"))
def test_mse_matches_reference(fname):
    g = golden(fname)
    a, b, gout = g["a"], g["b"], float(g["gout"])
    assert _close(wnp.mse_forward(a, b), float(g["f64_loss"]), tol=1e-12)
    da, db = wnp.mse_backward(a, b, gout)
    assert rel_err(da, g["f64_da"]) < 1e-12 and rel_err(db, g["f64_db"]) < 1e-12
    assert _close(float(wt.kd_mse(torch.from_numpy(a), torch.from_numpy(b))), float(g["f32_loss"]))


def test_update_aggregation_quirks():
    """/3-not-/2 and the `instance_wt_loss2 += instance_wt_loss2` overwrite."""
    g = golden("update_b6_16x16.npz")
    n, K = int(g["n"]), int(g["K"])
    per = [wnp.whitening_forward(g[k], n, K, 0.0, 1e-5) for k in ("main_z0", "main_z1")]
    ins, dom = wnp.wt_pse_aggregate(per)
    assert _close(ins, float(g["wt_ins"])) and _close(dom, float(g["wt_dom"]), scale=1.0)
    per = [wnp.whitening_forward(g[k], n, 3, 0.0, 1e-5) for k in ("shape_z0", "shape_z1")]
    tot, ij, ii, dom = wnp.shape_aggregate(per)
    assert _close(tot, float(g["sh_total"])) and _close(ij, float(g["sh_ij"])) and _close(ii, float(g["sh_ii"]))
    assert _close(dom, float(g["sh_dom"]), scale=1.0)
    assert _close(wnp.mse_forward(g["mu_teacher"], g["mu_student"]), float(g["sh_kd"]))
    assert bool(g["att_equal"])


def test_label_path_bit_exact():
    g = golden("labels_24x32.npz")
    od, oc = labels_np.labels_from_raw(g["raw_od"])
    assert od.dtype == np.uint8 and np.array_equal(od, g["same_od"]) and np.array_equal(oc, g["same_oc"])
    od2, oc2 = labels_np.labels_from_raw(g["raw_od"], g["raw_oc"])
    assert np.array_equal(od2, g["diff_od"]) and np.array_equal(oc2, g["diff_oc"])
    # closed form: OD = [raw <= 200], OC = [raw_od <= 50]
    assert np.array_equal(od[..., 0], (g["raw_od"] <= 200).astype(np.uint8))
    assert np.array_equal(oc[..., 0], (g["raw_od"] <= 50).astype(np.uint8))
    assert np.array_equal(labels_np.normalize_image(g["img"]), g["same_image"])
    assert np.array_equal(labels_np.roi_image(g["image"], g["od_pred"]), g["image_roi"])
    pred = labels_np.od_threshold(g["logits"])
    amb = labels_np.od_threshold_ambiguous(g["logits"])
    assert 0 < amb.sum() <= 4                       # the ln(3) probes planted by make_golden
    assert ((pred != g["od_pred"]) & ~amb).sum() == 0
    assert _close(float(labels_np.oc_pos_weight(g["od_pred"], g["target_oc"])), float(g["oc_pos_weight"]))
