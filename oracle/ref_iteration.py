"""One ``Trainer.train_epoch`` iteration (Trainer.py:762-925) driven on REAL model objects.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``Trainer.py`` itself cannot be imported in this image (it needs
sconf, tensorboardX, medpy, pytz, skimage -- none installed, no network), so the body of its hot loop is restated
here statement by statement, each block citing the lines it follows.  The four networks passed in are the
reference's own ``algorithms.WT_PSE`` / ``shape_networks.ShapeVariationalDist_x`` instances (stock, or after
``wtpse_b200.dropin.install`` / ``bind``): this file adds no arithmetic of its own beyond the trainer's loss sums.

``host_syncs=True`` keeps the trainer's per-loss ``.item()`` reads (Trainer.py:788-800, 828-832, 874-885, 917-919) --
that is what the reference's iteration costs; bench.py times it that way as ``train_step.reference_eager``.
"""
import math

import torch
import torch.nn.functional as F


def build_reference_models(alg, sn, hparams, n_per_domain, n_domains, device, seed=0):
    """train.py:91-138: two WT_PSE (OD: two_step=False, OC: two_step=True), two shape networks, four Adam(5e-4, (0.9, 0.99))."""
    torch.manual_seed(seed)
    model = alg.WT_PSE(3, 1, hparams, device, False, per_domain_batch=n_per_domain, source_domain_num=n_domains)
    model_shape = sn.ShapeVariationalDist_x(hparams, device, 1, number_source_domain=n_domains, batch_size=n_per_domain)
    model_oc = alg.WT_PSE(3, 1, hparams, device, True, per_domain_batch=n_per_domain, source_domain_num=n_domains)
    model_shape_oc = sn.ShapeVariationalDist_x(hparams, device, 1, number_source_domain=n_domains, batch_size=n_per_domain)
    nets = [m.to(device).train() for m in (model, model_shape, model_oc, model_shape_oc)]
    optims = [torch.optim.Adam(m.parameters(), lr=5e-4, betas=(0.9, 0.99)) for m in nets]
    return nets, optims


def _item(t, host_syncs):
    return t.item() if (host_syncs and torch.is_tensor(t)) else t


def trainer_iteration(nets, optims, image, target_od, target_oc, hparams, epoch=0, host_syncs=True, step_optim=True):
    """Returns a dict of the iteration's loss tensors (detached).  ``image`` is modified in place (``image += 1``,
    Trainer.py:850), as in the reference.  ``step_optim=False`` leaves the weights alone (gradient comparisons)."""
    model, model_shape, model_oc, model_shape_oc = nets
    optim, optim_shape, optim_oc, optim_shape_oc = optims
    wi, wd = hparams["instance_wt_gm"], hparams["domain_wt_gm"]
    out = {}

    # ---- step 1, OD segmentation network: Trainer.py:766-805 ------------------------------------------------------
    optim.zero_grad()
    model.zero_grad()
    output, _sp, _sp_mask, loss_ins_wt, loss_dom_wt = model.update(image, target_od, step=epoch, plot_show=0,
                                                                 two_stage_inputs=image, sp_mask=target_od, two_step=True)
    loss_seg = F.binary_cross_entropy(torch.sigmoid(output), target_od)          # bceloss = nn.BCELoss(), Trainer.py:28
    _item(loss_seg, host_syncs)
    if hparams["whitening"]:
        _item(loss_ins_wt, host_syncs)
        _item(loss_dom_wt, host_syncs)
        loss_data = _item((loss_seg + loss_ins_wt + loss_dom_wt).data, host_syncs)
    else:
        loss_data = _item(loss_seg.data, host_syncs)
    if host_syncs and math.isnan(loss_data):
        raise ValueError("loss is nan while training")
    loss_main = loss_seg + wi * loss_ins_wt + wd * loss_dom_wt
    loss_main.backward()
    if step_optim:
        optim.step()
    out.update(loss_seg=loss_seg, ins_wt=loss_ins_wt, dom_wt=loss_dom_wt)

    # ---- step 2, OD shape network: Trainer.py:810-841 -------------------------------------------------------------
    if hparams["whitening"]:
        for _ in range(hparams["multi-turn"]):
            optim_shape.zero_grad()
            model_shape.zero_grad()
            loss_kd, loss_ins_s, loss_ij, loss_ii, loss_dom_s = model_shape.update(model, image, target_od, step=epoch, plot_show=0,
                                                                                   two_stage_inputs=image, two_step=True)
            loss_shape = loss_kd + wi * loss_ins_s + wd * loss_dom_s
            loss_shape.backward()
            if step_optim:
                optim_shape.step()
        for t in (loss_kd, loss_ins_s, loss_dom_s, loss_ii, loss_ij):
            _item(t, host_syncs)
        out.update(kd=loss_kd, ins_wt_shape=loss_ins_s, ins_ij=loss_ij, ins_ii=loss_ii, dom_wt_shape=loss_dom_s)

    # ---- coarse-to-fine ROI: Trainer.py:842-853 -------------------------------------------------------------------
    od_pred = (torch.sigmoid(output) > 0.75).float().detach().float()
    optim_oc.zero_grad()
    model_oc.zero_grad()
    image += 1
    image_roi = image * od_pred
    image_roi -= 1

    # ---- step 3, OC segmentation network: Trainer.py:856-892 ------------------------------------------------------
    output_oc, _sp, _sp_mask, loss_ins_oc, loss_dom_oc = model_oc.update(image_roi, target_oc, step=epoch, plot_show=0,
                                                                         two_stage_inputs=image_roi, two_step=True)
    oc_pos_weight = torch.sum(od_pred) / torch.sum(od_pred * target_oc)
    if host_syncs:
        if torch.isinf(oc_pos_weight) or torch.isnan(oc_pos_weight):               # Trainer.py:866-867 (a host sync)
            oc_pos_weight = torch.tensor(1.).cuda()
    else:
        oc_pos_weight = torch.where(torch.isfinite(oc_pos_weight), oc_pos_weight, torch.ones_like(oc_pos_weight))
    loss_seg_oc = F.binary_cross_entropy_with_logits(output_oc * od_pred, target_oc, pos_weight=oc_pos_weight.to(output_oc.dtype))   # .to: a no-op in float32 (float64 runs of the tests)
    _item(loss_seg_oc, host_syncs)
    if hparams["whitening"]:
        _item(loss_ins_oc, host_syncs)
        _item(loss_dom_oc, host_syncs)
        loss_data_oc = _item((loss_seg_oc + loss_ins_oc + loss_dom_oc).data, host_syncs)
    else:
        loss_data_oc = _item(loss_seg_oc.data, host_syncs)
    if host_syncs and math.isnan(loss_data_oc):
        raise ValueError("loss is nan while training")
    loss_main_oc = loss_seg_oc + wi * loss_ins_oc + wd * loss_dom_oc
    loss_main_oc.backward()
    if step_optim:
        optim_oc.step()
    out.update(loss_seg_oc=loss_seg_oc, ins_wt_oc=loss_ins_oc, dom_wt_oc=loss_dom_oc, od_pred=od_pred, pos_weight=oc_pos_weight)

    # ---- step 4, OC shape network: Trainer.py:894-925 -------------------------------------------------------------
    if hparams["whitening"]:
        for _ in range(hparams["multi-turn"]):
            optim_shape_oc.zero_grad()
            model_shape_oc.zero_grad()
            loss_kd_oc, loss_ins_s_oc, _ij, _ii, loss_dom_s_oc = model_shape_oc.update(model_oc, image_roi, target_oc, step=epoch,
                                                                                       plot_show=0, two_stage_inputs=image_roi,
                                                                                       two_step=True)
            loss_shape_oc = loss_kd_oc + wi * loss_ins_s_oc + wd * loss_dom_s_oc
            loss_shape_oc.backward()
            if step_optim:
                optim_shape_oc.step()
        for t in (loss_kd_oc, loss_ins_s_oc, loss_dom_s_oc):
            _item(t, host_syncs)
        out.update(kd_oc=loss_kd_oc, ins_wt_shape_oc=loss_ins_s_oc, dom_wt_shape_oc=loss_dom_s_oc)
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
