"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.make_golden

The reference has no tests, fixtures or golden vectors of its own, so the only
thing that pins results is executing its code.  Every array stored here is produced by reference
code objects imported from /root/reference through oracle/ref_shim.py:
  whitening_*.npz  WT_PSE.compute_whitening_loss (algorithms.py:1277-1309) and
                   ShapeVariationalDist_x.compute_whitening_loss (shape_networks.py:561-594),
                   forward values, input gradient, and the same code re-run in float64.
  mmd_*.npz        compute_MMD.forward (algorithms.py:102-121 / shape_networks.py:283-309).
  mse_*.npz        ShapeVariationalDist_x.wasser_distance (shape_networks.py:596-597).
  update_*.npz     WT_PSE.update (algorithms.py:1216-1275) and ShapeVariationalDist_x.update
                   (shape_networks.py:512-558) on seeded weights: the whitening embeddings they
                   consumed and the loss tuples they returned (pins the /3 and += quirks).
  labels_*.npz     custom_transforms.Normalize_tf.__call__ (custom_transforms.py:466-499) and the
                   threshold/ROI statements of Trainer.py:842-853,865-867 executed verbatim.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def synth_z(B, H, W, seed=1234, offset=True, scale=0.3, dtype=torch.float32):
    """z = 0.3*randn(B,16,H,W) + 0.2*randn(B,16,1,1); iid variant drops the offset."""
    g = torch.Generator().manual_seed(seed)
    z = scale * torch.randn(B, 16, H, W, generator=g)
    if offset:
        z = z + 0.2 * torch.randn(B, 16, 1, 1, generator=g)
    return z.to(dtype)


def build_models(alg, sn, hp, n, K, dtype):
    torch.set_default_dtype(dtype)
    try:
        torch.manual_seed(0)
        main = alg.WT_PSE(3, 1, hp, "cpu", False, per_domain_batch=n, source_domain_num=K)
        shape = sn.ShapeVariationalDist_x(hp, "cpu", 1, number_source_domain=K, batch_size=n)
    finally:
        torch.set_default_dtype(torch.float32)
    return main, shape


def run_whitening(alg, sn, z32, n, K, margin, weights, dtype):
    hp = dict(ref_shim.DEFAULT_HPARAMS)
    hp["margin"] = margin
    main, shape = build_models(alg, sn, hp, n, K, dtype)
    out = {}
    torch.set_default_dtype(dtype)
    try:
        z = z32.to(dtype).clone().requires_grad_(True)
        ins, dom = main.compute_whitening_loss(z)
        dom_t = dom if torch.is_tensor(dom) else torch.zeros((), dtype=dtype)
        (weights[0] * ins + weights[2] * dom_t).backward()      # WT_PSE form: off and diag share one weight
        out["wt_ins"], out["wt_dom"] = float(ins.detach()), float(dom_t.detach())
        out["wt_dz"] = z.grad.detach().numpy().copy()
        # 3-value form (domain_num literal 3 inside ShapeVariationalDist_x)
        z2 = z32.to(dtype).clone().requires_grad_(True)
        off, diag, dom3 = shape.compute_whitening_loss(z2)
        (weights[0] * off + weights[1] * diag + weights[2] * dom3).backward()
        out["sh_off"], out["sh_diag"], out["sh_dom"] = float(off.detach()), float(diag.detach()), float(dom3.detach())
        out["sh_dz"] = z2.grad.detach().numpy().copy()
        # the Gram the reference forms (same statement as algorithms.py:1283)
        f = z32.to(dtype).view(z32.shape[0], 16, -1)
        out["gram"] = (torch.bmm(f, f.transpose(1, 2)).div(f.shape[-1] - 1) + main.eps * main.i.to(dtype)).numpy()
    finally:
        torch.set_default_dtype(torch.float32)
    return out


WHITENING_CASES = [
    # name,            B, H,  W,  n, K, margin, offset, seed, (w_off, w_diag, w_dom), scale
    ("b6_16x16",       6, 16, 16, 2, 3, 0.0,  True,  1234, (1.0, 1.0, 1.0), 0.3),
    ("b8_trailing",    8, 16, 16, 2, 3, 0.0,  True,  1235, (1.0, 1.0, 1.0), 0.3),
    ("iid_12x20",      6, 12, 20, 2, 3, 0.0,  False, 1236, (1.0, 1.0, 1.0), 0.3),
    ("margin_small",   6, 16, 16, 2, 3, 0.01, True,  1237, (1.0, 1.0, 1.0), 0.3),
    ("margin_clamps",  6, 16, 16, 2, 3, 5.0,  True,  1238, (1.0, 1.0, 1.0), 0.3),
    ("k2_n3",          6, 16, 16, 3, 2, 0.0,  True,  1239, (0.7, 1.3, 2.5), 0.3),
    ("ragged_5x7",     6, 5,  7,  2, 3, 0.0,  True,  1240, (1.0, 1.0, 1.0), 0.3),
    ("big_scale",      6, 16, 16, 2, 3, 0.0,  True,  1241, (1.0, 0.5, 0.25), 2.0),
    ("b9_n3",          9, 16, 16, 3, 3, 0.0,  True,  1242, (1.0, 1.0, 1.0), 0.3),
    ("b15_n5_12x12",  15, 12, 12, 5, 3, 0.0,  True,  1243, (1.0, 1.0, 1.0), 1.0),
]


def gen_whitening(alg, sn):
    for name, B, H, W, n, K, margin, offset, seed, w, scale in WHITENING_CASES:
        z = synth_z(B, H, W, seed, offset, scale)
        r32 = run_whitening(alg, sn, z, n, K, margin, w, torch.float32)
        r64 = run_whitening(alg, sn, z, n, K, margin, w, torch.float64)
        arrays = dict(z=z.numpy(), n=n, K=K, margin=margin, eps=1e-5, weights=np.array(w))
        for k, v in r32.items():
            arrays["f32_" + k] = v
        for k, v in r64.items():
            # float64 truth; the big gradient arrays are stored rounded to float32 (6e-8 relative,
            # far inside the 1e-5 tolerance) to keep the fixtures small
            arrays["f64_" + k] = v.astype(np.float32) if k.endswith("_dz") else v
        np.savez_compressed(os.path.join(OUT, "whitening_%s.npz" % name), **arrays)
        print("whitening_%-14s ins=%.6g dom=%.6g | f64 ins=%.6g dom=%.6g" %
              (name, r32["wt_ins"], r32["wt_dom"], r64["wt_ins"], r64["wt_dom"]))


def gen_mmd(alg, sn):
    g = torch.Generator().manual_seed(77)
    for name, B, n, K, spread in [("sep", 9, 3, 3, 0.5), ("iid", 12, 4, 3, 0.0), ("k2", 8, 4, 2, 0.3), ("trail", 8, 2, 3, 0.4)]:
        v = 0.05 * torch.randn(B, 120, generator=g)
        v = v + spread * torch.randn(B, 1, generator=g) * 0.2
        res = {}
        for tag, cls in (("alg", alg.compute_MMD), ("sn", sn.compute_MMD)):
            for dt, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
                x = v.to(dt).clone().requires_grad_(True)
                out = cls(domain_num=K, batch_size=n).forward(x)
                out.backward()
                res["%s_%s_loss" % (tag, dn)] = float(out)
                res["%s_%s_dv" % (tag, dn)] = x.grad.numpy().copy()
        np.savez_compressed(os.path.join(OUT, "mmd_%s.npz" % name), v=v.numpy(), n=n, K=K, **res)
        print("mmd_%-6s loss32=%.6g loss64=%.6g" % (name, res["alg_f32_loss"], res["alg_f64_loss"]))


def gen_mse(alg, sn):
    hp = dict(ref_shim.DEFAULT_HPARAMS)
    _, shape = build_models(alg, sn, hp, 2, 3, torch.float32)
    g = torch.Generator().manual_seed(5)
    for name, shp in [("b6_16x16", (6, 1, 16, 16)), ("odd", (5, 1, 7, 9))]:
        a = torch.randn(*shp, generator=g)
        b = torch.randn(*shp, generator=g)
        res = {}
        for dt, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
            aa = a.to(dt).clone().requires_grad_(True)
            bb = b.to(dt).clone().requires_grad_(True)
            out = shape.wasser_distance(aa, bb)
            (1.7 * out).backward()
            res[dn + "_loss"] = float(out)
            res[dn + "_da"] = aa.grad.numpy().copy()
            res[dn + "_db"] = bb.grad.numpy().copy()
        np.savez_compressed(os.path.join(OUT, "mse_%s.npz" % name), a=a.numpy(), b=b.numpy(), gout=1.7, **res)
        print("mse_%-10s %.6g" % (name, res["f32_loss"]))


def gen_update(alg, sn):
    """Run both update() entry points; record what the whitening loss saw and what came back."""
    hp = dict(ref_shim.DEFAULT_HPARAMS)
    n, K, H = 2, 3, 16
    B = n * K
    main, shape = build_models(alg, sn, hp, n, K, torch.float32)
    main.train(); shape.train()
    g = torch.Generator().manual_seed(99)
    image = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    image = image + 0.3 * torch.arange(B).view(B, 1, 1, 1).div(n, rounding_mode="floor")   # per-domain shift
    mask = (torch.rand(B, 1, H, H, generator=g) > 0.5).float()
    # predict() (algorithms.py:1311-1353) in eval mode, BEFORE any train-mode forward touches the BatchNorm
    # running statistics: pure backbone, pins the PyTorch side of the port from the seed alone
    main.eval(); shape.eval()
    with torch.no_grad():
        pred_logits, pred_pre = main.predict(shape, image)
    main.train(); shape.train()
    with torch.no_grad():
        emb_main = [t.clone() for t in main.wt_model.forward(image)]
        emb_shape = [t.clone() for t in shape.wt_model.forward(image)]
    torch.manual_seed(7)
    logits, att1, att2, ins, dom = main.update(image, mask, step=0, plot_show=0, two_stage_inputs=image,
                                               sp_mask=mask, two_step=True)
    torch.manual_seed(7)
    kd, ins_total, ins_ij, ins_ii, dom_s = shape.update(main, image, mask, step=0, plot_show=0,
                                                        two_stage_inputs=image, two_step=True)
    # teacher / student mu that the KD loss compared (RNG-free: mu precedes sampling)
    with torch.no_grad():
        fm_t = main.prior_dist.unet_extractor(emb_main[-1], mask)
        mu_t = main.prior_dist.mu_prior(fm_t)
        mu_s = shape.mu_prior(shape.unet_extractor(emb_shape[-1]))
    np.savez_compressed(
        os.path.join(OUT, "update_b6_16x16.npz"),
        n=n, K=K, margin=0.0, eps=1e-5, image=image.numpy(), mask=mask.numpy(),
        predict_logits=pred_logits.numpy(), predict_pre_sigmoid=pred_pre.numpy(), logits=logits.detach().numpy(),
        main_z0=emb_main[0].numpy(), main_z1=emb_main[1].numpy(),
        shape_z0=emb_shape[0].numpy(), shape_z1=emb_shape[1].numpy(),
        mu_teacher=mu_t.numpy(), mu_student=mu_s.numpy(),
        wt_ins=float(ins), wt_dom=float(dom),
        sh_kd=float(kd), sh_total=float(ins_total), sh_ij=float(ins_ij), sh_ii=float(ins_ii), sh_dom=float(dom_s),
        att_equal=bool(torch.equal(att1, att2)), logits_shape=np.array(logits.shape),
    )
    print("update: wt(ins=%.6g dom=%.6g) shape(kd=%.6g tot=%.6g ij=%.6g ii=%.6g dom=%.6g)" %
          (float(ins), float(dom), float(kd), float(ins_total), float(ins_ij), float(ins_ii), float(dom_s)))


def gen_labels(ct):
    rng = np.random.RandomState(3)
    H, W = 24, 32
    raw = rng.randint(0, 256, size=(H, W)).astype(np.uint8)
    raw.flat[:256] = np.arange(256, dtype=np.uint8)          # every byte value appears
    raw_oc = rng.randint(0, 256, size=(H, W)).astype(np.uint8)
    img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    res = {}
    for tag, oc_src in (("same", raw), ("diff", raw_oc)):
        sample = {"image": img.copy(), "label_od": raw.copy(), "label_oc": oc_src.copy()}
        out = ct.Normalize_tf()(sample)
        res[tag + "_image"] = out["image"]
        res[tag + "_od"] = out["label_od"]
        res[tag + "_oc"] = out["label_oc"]
    # Trainer.py:842-853 and :865-867, statements executed verbatim on seeded tensors
    g = torch.Generator().manual_seed(11)
    B = 6
    output = 3.0 * torch.randn(B, 1, H, W, generator=g)
    output.view(-1)[:8] = torch.tensor([1.0986, 1.0987, 1.09861, 1.098612, 1.0986123, 1.0986124, -1.0986, 40.0])
    image = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    target_oc = (torch.rand(B, 1, H, W, generator=g) > 0.7).float()
    image_in = image.clone()
    od_pred = (torch.sigmoid(output) > 0.75).float().detach().float()
    image += 1
    image_roi = image * od_pred
    image_roi -= 1
    oc_pos_weight = torch.sum(od_pred) / torch.sum(od_pred * target_oc)
    if torch.isinf(oc_pos_weight) or torch.isnan(oc_pos_weight):
        oc_pos_weight = torch.tensor(1.)
    np.savez_compressed(os.path.join(OUT, "labels_24x32.npz"), raw_od=raw, raw_oc=raw_oc, img=img,
                        logits=output.numpy(), image=image_in.numpy(), target_oc=target_oc.numpy(),
                        od_pred=od_pred.numpy(), image_roi=image_roi.numpy(),
                        oc_pos_weight=float(oc_pos_weight), **res)
    print("labels: od ones=%d oc ones=%d pos_weight=%.6g" %
          (int(res["same_od"].sum()), int(res["same_oc"].sum()), float(oc_pos_weight)))


def main():
    os.makedirs(OUT, exist_ok=True)
    alg, sn, ct = ref_shim.load()
    gen_whitening(alg, sn)
    gen_mmd(alg, sn)
    gen_mse(alg, sn)
    gen_update(alg, sn)
    gen_labels(ct)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("wrote %s (%.1f KiB total)" % (OUT, tot / 1024))


if __name__ == "__main__":
    main()
