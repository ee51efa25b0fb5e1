"""CPU checks of the PyTorch side of the entry-point classes (wt-pse-code_b200/segmentation.py): same
state-dict keys and initialisation stream as the reference, same backbone outputs.  The loss kernels
themselves need a GPU and are covered by tests/test_gpu_*.py."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err

HP = {"whitening": True, "margin": 0, "shape_prior": True, "shape_attention": True, "cat_shape": False,
      "shape_attention_coeffient": 0.3, "shape_start": 0.5, "instance_wt_gm": 1, "domain_wt_gm": 1, "multi-turn": 1}


def _ours(n=2, K=3):
    from wtpse_b200 import segmentation as seg

    torch.manual_seed(0)
    main = seg.WT_PSE(3, 1, dict(HP), "cpu", False, per_domain_batch=n, source_domain_num=K)
    shape = seg.ShapeVariationalDist_x(dict(HP), "cpu", 1, number_source_domain=K, batch_size=n)
    return main, shape


def test_predict_matches_reference_golden():
    """Seed-0 construction consumes the RNG in the reference's order, so the weights -- and therefore
    predict() (algorithms.py:1311-1353, pure backbone) -- reproduce what the reference produced."""
    g = golden("update_b6_16x16.npz")
    main, shape = _ours(int(g["n"]), int(g["K"]))
    main.eval(); shape.eval()
    with torch.no_grad():
        logits, pre = main.predict(shape, torch.from_numpy(g["image"]))
    assert rel_err(logits.numpy(), g["predict_logits"]) < 1e-5
    assert rel_err(pre.numpy(), g["predict_pre_sigmoid"]) < 1e-5


def test_whitening_features_match_reference_golden():
    g = golden("update_b6_16x16.npz")
    main, shape = _ours(int(g["n"]), int(g["K"]))
    with torch.no_grad():
        fm = main.wt_model(torch.from_numpy(g["image"]))
        fs = shape.wt_model(torch.from_numpy(g["image"]))
    assert len(fm) == 3 and torch.equal(fm[2], torch.relu(fm[1]))
    assert rel_err(fm[0].numpy(), g["main_z0"]) < 1e-6 and rel_err(fm[1].numpy(), g["main_z1"]) < 1e-6
    assert rel_err(fs[0].numpy(), g["shape_z0"]) < 1e-6 and rel_err(fs[1].numpy(), g["shape_z1"]) < 1e-6
    main.train(); shape.train()
    with torch.no_grad():
        mu_t = main.prior_dist.mu_prior(main.prior_dist.unet_extractor(fm[-1], torch.from_numpy(g["mask"])))
        mu_s = shape.mu_prior(shape.unet_extractor(fs[-1]))
    assert rel_err(mu_t.numpy(), g["mu_teacher"]) < 1e-5 and rel_err(mu_s.numpy(), g["mu_student"]) < 1e-5


def test_parameter_counts_match_survey():
    main, shape = _ours()
    assert sum(p.numel() for p in main.parameters()) == 6378661
    assert sum(p.numel() for p in shape.parameters()) == 3189570


def test_state_dict_is_interchangeable_with_the_reference():
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not on this box")
    alg, sn, _ = ref_shim.load()
    torch.manual_seed(0)
    ref_main = alg.WT_PSE(3, 1, dict(HP), "cpu", False, per_domain_batch=2, source_domain_num=3)
    ref_shape = sn.ShapeVariationalDist_x(dict(HP), "cpu", 1, number_source_domain=3, batch_size=2)
    main, shape = _ours()
    for ours, ref in ((main, ref_main), (shape, ref_shape)):
        a, b = ours.state_dict(), ref.state_dict()
        assert list(sorted(a)) == list(sorted(b))
        for k in a:
            assert torch.equal(a[k], b[k]), k                 # same init stream under the same seed
        ours.load_state_dict(b, strict=True)
        ref.load_state_dict(a, strict=True)


def test_no_extra_state_for_the_loss():
    main, shape = _ours()
    keys = list(main.state_dict()) + list(shape.state_dict())
    assert not any(k.split(".")[0] in ("i", "reversal_i", "diagonal", "mmd_operator") for k in keys)


def test_update_refuses_cpu_tensors():
    main, _ = _ours()
    x = torch.randn(6, 3, 16, 16)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        main.update(x, (x[:, :1] > 0).float(), two_stage_inputs=x, two_step=True)


def test_dropin_install_rebinds_the_reference_classes():
    """dropin.install() swaps the hot-path methods on the reference's own classes (and uninstall restores them)."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not on this box")
    import wtpse_b200 as wb

    alg, sn, _ = ref_shim.load()
    torch.manual_seed(0)
    ref_main = alg.WT_PSE(3, 1, dict(HP), "cpu", False, per_domain_batch=2, source_domain_num=3)
    ref_shape = sn.ShapeVariationalDist_x(dict(HP), "cpu", 1, number_source_domain=3, batch_size=2)
    z = torch.randn(6, 16, 8, 8)
    before = ref_main.compute_whitening_loss(z)
    saved = wb.dropin.install(alg, sn)
    try:
        assert alg.WT_PSE.compute_whitening_loss is wb.dropin.wt_pse_compute_whitening_loss
        assert sn.ShapeVariationalDist_x.wasser_distance is wb.dropin.shape_wasser_distance
        # the rebound methods read the reference's own attributes and route into the CUDA path (which refuses CPU input)
        for call in (lambda: ref_main.compute_whitening_loss(z), lambda: ref_shape.compute_whitening_loss(z),
                     lambda: ref_shape.wasser_distance(z[:, :1], z[:, 1:2]),
                     lambda: ref_main.mmd_operator.forward(torch.randn(6, 120))):
            with pytest.raises(RuntimeError, match="no CPU implementation"):
                call()
    finally:
        wb.dropin.uninstall(saved)
    after = ref_main.compute_whitening_loss(z)
    assert float(before[0]) == float(after[0]) and float(before[1]) == float(after[1])


def test_conv_bias_folding_keeps_outputs_gradients_and_running_statistics():
    """segmentation._conv_bn(fold=True) -- what TrainStep runs -- against the plain conv(+bias) -> BatchNorm sequence of the
    reference (algorithms.py:877-962): same activations, input/weight/BN gradients and BatchNorm buffers up to fp32 rounding;
    the conv biases in front of a BatchNorm only ever see rounding noise as their gradient (in both modes)."""
    import copy

    from wtpse_b200 import segmentation as seg

    torch.manual_seed(0)
    cases = [(seg.ConvD(8, 16), False), (seg.ConvD(8, 16, first=True), False), (seg.ConvU(16), True), (seg._DoubleConv(8, 16), False)]
    for m0, is_up in cases:
        m0.train()
        for p in m0.parameters():
            p.data.normal_(0, 0.5)
        x = torch.randn(4, 32 if is_up else 8, 16, 16)
        skip = torch.randn(4, 8, 32, 32)
        res = []
        for fold in (False, True):
            m = seg.set_conv_bias_folding(copy.deepcopy(m0), fold)
            xx = x.clone().requires_grad_()
            for _ in range(2):                                     # two updates of the running statistics
                y = m(xx, skip) if is_up else m(xx)
            m.zero_grad()
            xx.grad = None
            (y * torch.linspace(0, 1, y.numel()).view_as(y)).sum().backward()
            res.append((y.detach(), xx.grad, dict((n, p.grad.clone()) for n, p in m.named_parameters()),
                        dict((n, b.clone()) for n, b in m.named_buffers())))
        a, b = res
        conv_biases = {name + ".bias" for name, mod in m0.named_modules() if isinstance(mod, torch.nn.Conv2d)}
        assert float((a[0] - b[0]).abs().max()) <= 1e-5 * float(a[0].abs().max())
        assert float((a[1] - b[1]).abs().max()) <= 1e-5 * float(a[1].abs().max())
        for n in a[2]:
            if n in conv_biases:
                scale = max(float(g.abs().max()) for k, g in a[2].items() if k.endswith("weight"))
                assert float(a[2][n].abs().max()) <= 1e-4 * scale and float(b[2][n].abs().max()) <= 1e-4 * scale   # noise only
            else:
                assert float((a[2][n] - b[2][n]).abs().max()) <= 1e-5 * max(float(a[2][n].abs().max()), 1.0), n
        for n in a[3]:
            assert torch.allclose(a[3][n].float(), b[3][n].float(), rtol=1e-5, atol=1e-6), n
            if n.endswith("num_batches_tracked"):
                assert int(b[3][n]) == 2


def test_host_side_switches_of_the_train_step_kernels():
    """Host logic around the backbone / fusion kernels that needs no GPU: layout dispatch of the loss input, the
    per-network fusion config, the flag setters, state-dict keys of the Sequential subclass, ATen fall-backs."""
    from wtpse_b200 import elementwise as ew
    from wtpse_b200 import functional as wf
    from wtpse_b200 import segmentation as seg

    # loss input: dense channels-last -> read in place by the *_cl entry points; everything else -> NCHW contiguous
    z = torch.randn(2, 16, 6, 5)
    t, cl = wf._as_loss_input(z)
    assert not cl and t.is_contiguous()
    t, cl = wf._as_loss_input(z.contiguous(memory_format=torch.channels_last))
    assert cl and t.is_contiguous(memory_format=torch.channels_last)
    t, cl = wf._as_loss_input(z.contiguous(memory_format=torch.channels_last)[:, :, ::2])       # strided view: copy
    assert not cl and t.is_contiguous()
    t, cl = wf._as_loss_input(torch.randn(2, 16, 1, 1))                                          # ambiguous strides: NCHW
    assert not cl

    hp = dict(HP)
    main = seg.WT_PSE(3, 1, hp, "cpu", True, per_domain_batch=4, source_domain_num=2)
    shape = seg.ShapeVariationalDist_x(hp, "cpu", 1, number_source_domain=2, batch_size=4)
    assert main.wt_model.fused_loss is None and shape.teacher_grad is True
    seg.enable_relu_fusion(main)
    seg.enable_relu_fusion(shape)
    assert seg.fused_config(main.wt_model) == {"fold": True, "n_per_domain": 4, "n_domains": 2, "margin": 0.0, "eps": 1e-5}
    assert shape.wt_model.fused_loss["fold"] is False and seg.fused_config(shape.wt_model)["n_domains"] == 3     # literal 3
    main.margin = 0.25                                          # read at call time, not snapshotted
    assert seg.fused_config(main.wt_model)["margin"] == 0.25
    main.margin = hp["margin"]
    import copy
    assert copy.deepcopy(main).wt_model._pending_terms == {}
    seg.enable_relu_fusion(main, False)
    assert main.wt_model.fused_loss is None
    assert seg.fused_terms(main, z, 2) is None                        # no tag -> compute_whitening_loss runs the kernels

    for setter, attr in ((seg.set_conv_bias_folding, "fold_bias"), (seg.set_cuda_batchnorm, "cuda_bn"),
                         (seg.set_cuda_pool, "cuda_pool"), (seg.set_cuda_upsample, "cuda_upsample")):
        setter(main, True)
        mods = [m for m in main.modules() if hasattr(m, attr)]
        assert mods and all(getattr(m, attr) for m in mods)
        setter(main, False)
        assert not any(getattr(m, attr) for m in mods)
    seg.set_fast_bias(main, True)
    assert all(m.fast_bias for m in main.modules() if isinstance(m, seg._ConvActSeq))
    seg.set_fast_bias(main, False)

    head = seg._head(32, 8, 1)
    assert isinstance(head, torch.nn.Sequential) and list(head.state_dict()) == ["0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias"]
    x = torch.randn(2, 32, 5, 5)
    want = head(x)
    head.fast_bias = True                                       # CPU input: conv_bias_act takes the plain operators
    assert torch.allclose(head(x), want)

    g = torch.randn(3, 8, 4, 4)
    assert torch.allclose(ew.channel_sum(g), g.sum((0, 2, 3)))  # ATen route off the GPU
    bn = torch.nn.BatchNorm2d(8).train()
    assert not ew.batch_norm_act_supported(g, bn) and not ew.upsample2x_supported(g) and not ew.max_pool2_supported(g)
    with pytest.raises(ValueError):
        ew.upsample2x(g)
    with pytest.raises(ValueError):
        ew.max_pool2(g)
