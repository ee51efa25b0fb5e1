"""SURVEY.md 8(f).1 -- producer fusion with the DeepWT tail: wtpse_whitening_relu_forward/backward against the
unfused sequence (ATen relu + the plain loss kernels) and, end to end, against the oracle's operator-sequence restatement
of the reference (oracle/whitening_torch.py, pinned to the reference goldens in tests/test_oracle_golden.py)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

HP = {"whitening": True, "margin": 0, "shape_prior": True, "shape_attention": True, "cat_shape": False,
      "shape_attention_coeffient": 0.3, "shape_start": 0.5, "instance_wt_gm": 1, "domain_wt_gm": 1, "multi-turn": 1}


def _z(B, H, W, seed, dev):
    g = torch.Generator(device="cpu").manual_seed(seed)
    z = 0.3 * torch.randn(B, 16, H, W, generator=g) + 0.2 * torch.randn(B, 16, 1, 1, generator=g)
    return z.to(dev)


@pytest.mark.parametrize("B,H,W,n,K,fold", [
    (6, 64, 64, 2, 3, True),          # TMA path, several tiles per sample
    (6, 64, 64, 2, 3, False),
    (8, 256, 256, 2, 3, True),        # BASELINE configs[0] size (samples 6, 7 outside the MMD)
    (5, 37, 29, 2, 2, False),         # P % 4 != 0: generic kernels
    (3, 30, 30, 1, 3, True),          # P = 900: one partial tile
    (30, 48, 40, 10, 3, True),
])
def test_fused_relu_whitening_equals_the_unfused_sequence(B, H, W, n, K, fold):
    import wtpse_b200 as wb

    dev = torch.device("cuda:0")
    z_a = _z(B, H, W, 7, dev).requires_grad_()
    z_b = z_a.detach().clone().requires_grad_()
    z_b.data[0, 3, 0, :4] = 0.0                                         # exact zeros: relu'(0) = 0
    z_a.data.copy_(z_b.data)
    g = torch.Generator(device="cpu").manual_seed(11)
    g_relu = torch.randn(B, 16, H, W, generator=g).to(dev)
    weights = [1.0, 0.5, 2.0][: 2 if fold else 3]

    # unfused: ATen relu + the plain loss kernels, gradients summed by autograd
    plain = (wb.whitening_folded if fold else wb.whitening_terms)(z_a, n, K)
    r_a = torch.relu(z_a)
    (sum(w * t for w, t in zip(weights, plain)) + (r_a * g_relu).sum()).backward()

    fused = (wb.relu_whitening_folded if fold else wb.relu_whitening_terms)(z_b, n, K)
    r_b, terms = fused[0], fused[1:]
    (sum(w * t for w, t in zip(weights, terms)) + (r_b * g_relu).sum()).backward()

    assert torch.equal(r_b, r_a)                                        # bit-exact activation
    for t_f, t_p in zip(terms, plain):
        assert torch.equal(t_f, t_p), (float(t_f), float(t_p))          # same kernels, same summation order
    # dz = (M z) + [z > 0] g: the same two addends autograd sums, one rounding either way
    assert torch.equal(z_b.grad, z_a.grad), float((z_b.grad - z_a.grad).abs().max())


def test_fused_backward_when_only_one_branch_is_used():
    import wtpse_b200 as wb

    dev = torch.device("cuda:0")
    z0 = _z(6, 32, 32, 3, dev)
    # only the loss is used downstream
    z = z0.clone().requires_grad_()
    out = wb.relu_whitening_folded(z, 2, 3)
    (out[1] + out[2]).backward()
    zz = z0.clone().requires_grad_()
    a, b = wb.whitening_folded(zz, 2, 3)
    (a + b).backward()
    assert torch.equal(z.grad, zz.grad)
    # only the activation is used downstream (the teacher pass of the shape update): plain ReLU backward
    z = z0.clone().requires_grad_()
    out = wb.relu_whitening_terms(z, 2, 3)
    (out[0] * 3.0).sum().backward()
    assert torch.equal(z.grad, 3.0 * (z0 > 0).float())


def test_fused_entry_points_reject_bad_arguments():
    import ctypes

    import wtpse_b200 as wb

    lib = wb._lib.load()
    dev = torch.device("cuda:0")
    z = _z(2, 8, 8, 1, dev)
    ws = torch.empty(lib.wtpse_whitening_workspace_bytes(2, 64), dtype=torch.uint8, device=dev)
    out = torch.empty(4 + 2 * 256 + 4 + 2 * 120, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.wtpse_whitening_relu_forward(p(z), p(z), 2, 16, 64, 1, 2, 0.0, 1e-5, p(out), p(out[4:]), p(out[516:]), p(out[520:]),
                                          p(ws), ws.numel(), None)
    assert rc != 0 and b"alias" in lib.wtpse_last_error()
    rc = lib.wtpse_whitening_relu_forward(p(z), None, 2, 16, 64, 1, 2, 0.0, 1e-5, p(out), p(out[4:]), p(out[516:]), p(out[520:]),
                                          p(ws), ws.numel(), None)
    assert rc != 0
    with pytest.raises(RuntimeError):
        wb.relu_whitening_folded(z.cpu(), 1, 2)                          # no CPU path


def test_deepwt_tail_fusion_matches_the_reference_operator_sequence():
    """update() of both entry points with the fused DeepWT tail: losses and every parameter gradient equal the unfused
    path, which tests/test_gpu_update.py pins to the reference's goldens."""
    from wtpse_b200 import segmentation as seg

    dev = torch.device("cuda:0")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        main0 = seg.WT_PSE(3, 1, HP, dev, True, per_domain_batch=2, source_domain_num=3).to(dev).train()
        shape0 = seg.ShapeVariationalDist_x(HP, dev, 1, number_source_domain=3, batch_size=2).to(dev).train()
        x = torch.randn(6, 3, 64, 48, device=dev)
        mask = (torch.rand(6, 1, 64, 48, device=dev) > 0.5).float()
        res = []
        for fuse in (False, True):
            main, shape = copy.deepcopy(main0), copy.deepcopy(shape0)
            seg.enable_relu_fusion(main, fuse)
            seg.enable_relu_fusion(shape, fuse)
            torch.manual_seed(17)
            out = main.update(x, mask, step=0, plot_show=0, two_stage_inputs=x, sp_mask=mask, two_step=True)
            (out[0].mean() + out[3] + out[4]).backward()
            g_main = torch.cat([p.grad.reshape(-1) for p in main.wt_model.parameters()])
            main.zero_grad(set_to_none=True)
            torch.manual_seed(18)
            outs = shape.update(main, x, mask, step=0, plot_show=0, two_stage_inputs=x, two_step=True)
            (outs[0] + outs[1] + outs[4]).backward()
            # (the logvar head of the student feeds only the unused sample: no gradient in either mode)
            g_shape = torch.cat([p.grad.reshape(-1) for p in shape.parameters() if p.grad is not None])
            assert g_shape.numel() > 0.9 * sum(p.numel() for p in shape.parameters())
            res.append(([float(out[3]), float(out[4])] + [float(v) for v in outs], g_main, g_shape))
        assert res[0][0] == res[1][0], (res[0][0], res[1][0])
        for a, b in ((res[0][1], res[1][1]), (res[0][2], res[1][2])):
            err = float((a - b).abs().max() / a.abs().max())
            assert err <= 1e-5, err
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_train_step_with_fused_tail_runs_channels_last_and_captures():
    import numpy as np

    import wtpse_b200 as wb

    dev = torch.device("cuda:0")
    ts = wb.TrainStep(n_per_domain=2, n_domains=3, device=dev, seed=0, fuse_relu=True)
    assert ts.model.wt_model.fused_loss is not None and ts.model_shape.wt_model.fused_loss["fold"] is False
    ref = wb.TrainStep(n_per_domain=2, n_domains=3, device=dev, seed=0, fuse_relu=False)
    image, od, oc = wb.synthetic.fundus_batch(2, 3, 64, 64, dev, seed=4)
    a = ts.step(image.clone(), od, oc)
    b = ref.step(image.clone(), od, oc)
    # sub-step 1's whitening losses depend on the initial weights only (no RNG): identical up to conv layout rounding
    for k in ("ins_wt", "dom_wt"):
        assert abs(float(a[k]) - float(b[k])) <= 2e-4 * max(abs(float(b[k])), 1e-3), (k, float(a[k]), float(b[k]))
    ts.capture(image.clone(), od, oc, warmup=1)
    out = ts.replay(image.clone(), od, oc)
    assert all(np.isfinite(float(v)) for v in out.values())


@pytest.mark.parametrize("B,H,W,n,K", [(6, 64, 64, 2, 3), (8, 256, 256, 2, 3), (5, 37, 29, 2, 2), (3, 30, 30, 1, 3), (30, 48, 40, 10, 3),
                                       (2, 300, 280, 1, 2), (1, 8, 8, 1, 1)])
def test_channels_last_loss_kernels_match_the_nchw_ones(B, H, W, n, K):
    """wtpse_whitening_forward_cl / _backward_cl (plain and fused with the DeepWT-tail ReLU) on channels-last tensors against
    the NCHW kernels on the same values: losses to fp32 rounding (the summation order differs), relu bit-exact, dz to 1e-5."""
    import wtpse_b200 as wb

    dev = torch.device("cuda:0")
    z0 = _z(B, H, W, 21, dev)
    g_relu = torch.randn(B, 16, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    ones = [torch.ones((), device=dev)] * 3

    def run(z, fused):
        z = z.requires_grad_()
        if fused:
            r, off, diag, dom = wb.relu_whitening_terms(z, n, K)
            torch.autograd.backward([off, diag, dom, r], ones + [g_relu])
        else:
            r = None
            off, diag, dom = wb.whitening_terms(z, n, K)
            torch.autograd.backward([off, diag, dom], ones)
        return r, (float(off), float(diag), float(dom)), z.grad

    for fused in (False, True):
        r_a, l_a, dz_a = run(z0.clone(), fused)
        z_cl = z0.clone().contiguous(memory_format=torch.channels_last)
        wb._lib.debug_set("cl_tma_launches", 0)
        r_b, l_b, dz_b = run(z_cl, fused)
        assert wb._lib.debug_get("cl_tma_launches") == 2          # forward + backward took the tensor-map TMA kernels
        assert dz_b.is_contiguous(memory_format=torch.channels_last)
        for a, b in zip(l_a, l_b):
            assert abs(a - b) <= 1e-5 * max(abs(a), 1e-3) or (a != a and b != b), (l_a, l_b)
        assert float((dz_a - dz_b).abs().max()) <= 1e-5 * float(dz_a.abs().max())
        if fused:
            assert torch.equal(r_a, r_b) and r_b.is_contiguous(memory_format=torch.channels_last)
        # the per-thread channels-last kernels (the fallback for unaligned pointers): per pixel the same arithmetic in the
        # same order -> relu and dz bit-identical; the Gram differs in summation order only
        wb._lib.debug_set("cl_tma", 0)
        try:
            r_c, l_c, dz_c = run(z0.clone().contiguous(memory_format=torch.channels_last), fused)
        finally:
            wb._lib.debug_set("cl_tma", 1)
        assert wb._lib.debug_get("cl_tma_launches") == 2
        for a, b in zip(l_b, l_c):
            assert abs(a - b) <= 1e-5 * max(abs(a), 1e-3) or (a != a and b != b), (l_b, l_c)
        assert float((dz_b - dz_c).abs().max()) <= 1e-5 * float(dz_b.abs().max())
        if fused:
            assert torch.equal(r_b, r_c)


def test_train_step_fused_tail_channels_last_matches_unfused():
    import wtpse_b200 as wb

    dev = torch.device("cuda:0")
    a = wb.TrainStep(n_per_domain=2, n_domains=3, device=dev, seed=0, fuse_relu=True)
    b = wb.TrainStep(n_per_domain=2, n_domains=3, device=dev, seed=0, fuse_relu=False)
    assert a.model.wt_model.DoubleConv.double_conv[0].weight.is_contiguous(memory_format=torch.channels_last)
    image, od, oc = wb.synthetic.fundus_batch(2, 3, 64, 64, dev, seed=9)
    oa, ob = a.step(image.clone(), od, oc), b.step(image.clone(), od, oc)
    for k in ("ins_wt", "dom_wt"):          # sub-step 1's whitening losses see only the initial weights (no RNG)
        assert abs(float(oa[k]) - float(ob[k])) <= 2e-5 * max(abs(float(ob[k])), 1e-3), (k, float(oa[k]), float(ob[k]))


@pytest.mark.parametrize("layout", ["nchw", "cl"])
def test_fused_backward_waits_for_a_programmatic_producer_of_grad_relu(layout):
    """Round-1 advisor finding: the fused backward runs as a programmatic dependent, and the upstream ReLU gradient is the output of
    the kernel right in front of it, so the TMA loads of grad_relu must not be issued before griddepcontrol.wait (the z loads may).
    Stress: a producer that signals launch_dependents at once, spins ~1 ms and only then writes grad_relu over a NaN-filled buffer
    (wtpse_debug_pdl_slow_copy), the backward launched straight behind it through the C ABI -- twenty times, bitwise against the
    result with grad_relu in place."""
    import wtpse_b200 as wb
    from wtpse_b200.functional import _forward_outputs, _ptr, _stream_ptr, _workspace

    dev = torch.device("cuda:0")
    lib = wb._lib.load()
    B, H, W, n, K = 12, 96, 96, 4, 3
    P = H * W
    z = _z(B, H, W, 21, dev)
    g = torch.randn(B, 16, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    if layout == "cl":                       # [B][P][16] memory
        z = z.permute(0, 2, 3, 1).contiguous()
        g = g.permute(0, 2, 3, 1).contiguous()
    relu = torch.empty_like(z)
    dz_ref, dz = torch.empty_like(z), torch.empty_like(z)
    ws, ws_bytes = _workspace(lib, B, P, dev)
    losses, (gram, rowstat, domgrad) = _forward_outputs(B, dev)
    one = torch.ones((), device=dev)
    st = _stream_ptr(dev)
    fwd = lib.wtpse_whitening_forward_cl if layout == "cl" else lib.wtpse_whitening_relu_forward
    bwd = lib.wtpse_whitening_backward_cl if layout == "cl" else lib.wtpse_whitening_relu_backward
    wb._lib.check(fwd(_ptr(z), _ptr(relu), B, 16, P, n, K, 0.0, 1e-5, _ptr(losses), _ptr(gram), _ptr(rowstat), _ptr(domgrad), _ptr(ws),
                      ws_bytes, st))

    def backward(grad_relu, out):
        wb._lib.check(bwd(_ptr(z), _ptr(grad_relu), _ptr(gram), _ptr(rowstat), _ptr(domgrad), _ptr(one), _ptr(one), _ptr(one), B, 16, P, n, K,
                          _ptr(out), st))

    backward(g, dz_ref)
    torch.cuda.synchronize()
    assert torch.isfinite(dz_ref).all()
    late = torch.empty_like(g)
    for _ in range(20):
        late.fill_(float("nan"))
        dz.zero_()
        torch.cuda.synchronize()
        wb._lib.check(lib.wtpse_debug_pdl_slow_copy(_ptr(late), _ptr(g), g.numel(), 2_000_000, st))
        backward(late, dz)
        torch.cuda.synchronize()
        assert torch.equal(dz, dz_ref)
