"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors produced by the
unmodified reference, against the oracle on seeded inputs, and through size-independent properties at
BASELINE.json's full sizes.

Tolerances: rel <= 1e-5 (max-abs normalised) on Gram entries, instance terms, KD loss
and all gradients; the domain (MMD) scalar is compared with |delta| <= 1e-5 * max(|ref|, 1) because
Kxx + Kyy - 2Kxy cancels O(1) terms (the reference's own fp32 result is 3.8e-3 away from its fp64 result in
the iid regime).
"""
import numpy as np
import pytest
import torch

from conftest import golden, golden_files, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _close(a, b, tol=TOL, scale=0.0):
    if np.isnan(b):
        return bool(np.isnan(a))
    return abs(a - b) <= tol * max(abs(b), scale, 1e-300)


def _rel_err_dev(a, b):
    """max-abs normalised error computed on the device (full-size tensors stay in HBM)."""
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _synth(B, H, W, seed=1234, offset=True):
    g = torch.Generator().manual_seed(seed)
    z = 0.3 * torch.randn(B, 16, H, W, generator=g)
    if offset:
        z = z + 0.2 * torch.randn(B, 16, 1, 1, generator=g)
    return z


# ---------------------------------------------------------------------------------------------
# golden vectors (reference outputs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fname", golden_files("whitening_"))
def test_whitening_against_reference_golden(fname):
    import wtpse_b200 as wb

    g = golden(fname)
    dev = _dev()
    n, K, margin = int(g["n"]), int(g["K"]), float(g["margin"])
    w = [float(x) for x in g["weights"]]

    # two-value form, WT_PSE.compute_whitening_loss
    z = torch.from_numpy(g["z"]).to(dev).requires_grad_(True)
    ins, dom = wb.whitening_folded(z, n, K, margin, 1e-5)
    (w[0] * ins + w[2] * dom).backward()
    for ref in ("f32_", "f64_"):
        assert _close(float(ins), float(g[ref + "wt_ins"])), (float(ins), float(g[ref + "wt_ins"]))
        assert _close(float(dom), float(g[ref + "wt_dom"]), scale=1.0)
        assert rel_err(z.grad.cpu().numpy(), g[ref + "wt_dz"]) < TOL
    assert rel_err(wb.gram_matrix(z.detach()).cpu().numpy(), g["f32_gram"]) < TOL
    assert rel_err(wb.gram_matrix(z.detach()).cpu().numpy(), g["f64_gram"]) < TOL

    # three-value form, ShapeVariationalDist_x.compute_whitening_loss (3 domains hard-coded)
    z2 = torch.from_numpy(g["z"]).to(dev).requires_grad_(True)
    off, diag, dom3 = wb.whitening_terms(z2, n, 3, margin, 1e-5)
    assert _close(float(off), float(g["f32_sh_off"])) and _close(float(off), float(g["f64_sh_off"]))
    assert _close(float(diag), float(g["f32_sh_diag"])) and _close(float(diag), float(g["f64_sh_diag"]))
    assert _close(float(dom3), float(g["f32_sh_dom"]), scale=1.0)
    if not np.isnan(float(g["f32_sh_dom"])):
        (w[0] * off + w[1] * diag + w[2] * dom3).backward()
        assert rel_err(z2.grad.cpu().numpy(), g["f32_sh_dz"]) < TOL
        assert rel_err(z2.grad.cpu().numpy(), g["f64_sh_dz"]) < TOL


@pytest.mark.parametrize("fname", golden_files("mmd_"))
def test_mmd_against_reference_golden(fname):
    import wtpse_b200 as wb

    g = golden(fname)
    v = torch.from_numpy(g["v"]).to(_dev()).requires_grad_(True)
    loss = wb.mmd_penalty(v, int(g["n"]), int(g["K"]))
    loss.backward()
    assert _close(float(loss), float(g["alg_f64_loss"]), scale=1.0)
    assert _close(float(loss), float(g["sn_f32_loss"]), scale=1.0)
    assert rel_err(v.grad.cpu().numpy(), g["alg_f64_dv"]) < TOL


@pytest.mark.parametrize("fname", golden_files("mse_"))
def test_kd_mse_against_reference_golden(fname):
    import wtpse_b200 as wb

    g = golden(fname)
    a = torch.from_numpy(g["a"]).to(_dev()).requires_grad_(True)
    b = torch.from_numpy(g["b"]).to(_dev()).requires_grad_(True)
    loss = wb.kd_mse(a, b)
    (float(g["gout"]) * loss).backward()
    assert _close(float(loss), float(g["f32_loss"])) and _close(float(loss), float(g["f64_loss"]))
    assert rel_err(a.grad.cpu().numpy(), g["f32_da"]) < TOL
    assert rel_err(b.grad.cpu().numpy(), g["f32_db"]) < TOL


def test_update_level_aggregation_matches_reference():
    """The loss tuples WT_PSE.update / ShapeVariationalDist_x.update returned,
    rebuilt from the CUDA per-embedding losses with the reference's own accumulation statements."""
    import wtpse_b200 as wb

    g = golden("update_b6_16x16.npz")
    dev = _dev()
    n, K = int(g["n"]), int(g["K"])
    # algorithms.py:1257-1267
    ins_acc, dom_acc = 0, 0
    embs = [torch.from_numpy(g[k]).to(dev) for k in ("main_z0", "main_z1")]
    for e in embs:
        i1, d1 = wb.whitening_folded(e, n, K, 0.0, 1e-5)
        ins_acc += i1
        dom_acc += d1
    ins_acc /= 3
    dom_acc /= 3
    assert _close(float(ins_acc), float(g["wt_ins"])) and _close(float(dom_acc), float(g["wt_dom"]), scale=1.0)
    # shape_networks.py:539-554 (including the `instance_wt_loss2 += instance_wt_loss2` overwrite)
    l1, l2, ld = 0, 0, 0
    for k in ("shape_z0", "shape_z1"):
        a, l2, d = wb.whitening_terms(torch.from_numpy(g[k]).to(dev), n, 3, 0.0, 1e-5)
        l1 += a
        l2 = l2 + l2
        ld += d
    l1 /= 3
    l2 = l2 / 3
    ld /= 3
    assert _close(float(l1), float(g["sh_ij"])) and _close(float(l2), float(g["sh_ii"]))
    assert _close(float(l1 + l2), float(g["sh_total"])) and _close(float(ld), float(g["sh_dom"]), scale=1.0)
    kd = wb.kd_mse(torch.from_numpy(g["mu_teacher"]).to(dev), torch.from_numpy(g["mu_student"]).to(dev))
    assert _close(float(kd), float(g["sh_kd"]))


# ---------------------------------------------------------------------------------------------
# oracle on seeded inputs at sizes it finishes in seconds
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,n,K", [(8, 256, 256, 2, 3), (9, 96, 100, 3, 3), (15, 64, 64, 5, 3), (5, 33, 31, 2, 2),
                                       (3, 1024, 1024, 1, 3), (40, 28, 28, 13, 3)])
def test_whitening_against_oracle(B, H, W, n, K):
    import wtpse_b200 as wb
    from oracle import whitening_np as wnp

    z_cpu = _synth(B, H, W, seed=B * 1000 + H)
    z = z_cpu.to(_dev()).requires_grad_(True)
    off, diag, dom = wb.whitening_terms(z, n, K, 0.0, 1e-5)
    (0.9 * off + 1.1 * diag + 1.3 * dom).backward()
    f = wnp.whitening_forward(z_cpu.numpy(), n, K, 0.0, 1e-5)
    dz, _ = wnp.whitening_backward(z_cpu.numpy(), f, n, K, 0.9, 1.1, 1.3)
    assert _close(float(off), float(f["off"])) and _close(float(diag), float(f["diag"]))
    assert _close(float(dom), float(f["dom"]), scale=1.0)
    assert rel_err(wb.gram_matrix(z.detach()).cpu().numpy(), f["gram"]) < TOL
    assert rel_err(z.grad.cpu().numpy(), dz) < TOL


def test_iid_regime_no_worse_than_reference_fp32():
    """Cancellation regime: be at least as close to the float64 truth as the reference's fp32 path."""
    import wtpse_b200 as wb
    from oracle import whitening_np as wnp, whitening_torch as wt

    z_cpu = _synth(12, 64, 64, seed=5, offset=False)
    truth = float(wnp.whitening_forward(z_cpu.numpy(), 4, 3)["dom"])
    _, ref32 = wt.wt_pse_whitening_loss(z_cpu, 4, 3)
    _, got = wb.whitening_folded(z_cpu.to(_dev()), 4, 3)
    assert abs(float(got) - truth) <= max(abs(float(ref32) - truth), 1e-7 * max(abs(truth), 1e-3))


def test_unaligned_and_ragged_inputs_take_the_generic_kernels():
    import wtpse_b200 as wb
    from oracle import whitening_np as wnp

    dev = _dev()
    # P % 4 != 0
    z_cpu = _synth(6, 37, 29, seed=3)
    z = z_cpu.to(dev).requires_grad_(True)
    ins, dom = wb.whitening_folded(z, 2, 3)
    (ins + dom).backward()
    f = wnp.whitening_forward(z_cpu.numpy(), 2, 3)
    dz, _ = wnp.whitening_backward(z_cpu.numpy(), f, 2, 3)
    assert _close(float(ins), float(f["off"] + f["diag"])) and _close(float(dom), float(f["dom"]), scale=1.0)
    assert rel_err(z.grad.cpu().numpy(), dz) < TOL
    # base pointer offset by 4 bytes (contiguous but not 16-byte aligned)
    z_cpu = _synth(6, 32, 32, seed=4)
    buf = torch.empty(z_cpu.numel() + 1, device=dev)
    zu = buf[1:].view_as(z_cpu)
    zu.copy_(z_cpu)
    assert zu.data_ptr() % 16 != 0 and zu.is_contiguous()
    zu.requires_grad_(True)
    ins, dom = wb.whitening_folded(zu, 2, 3)
    (ins + dom).backward()
    f = wnp.whitening_forward(z_cpu.numpy(), 2, 3)
    dz, _ = wnp.whitening_backward(z_cpu.numpy(), f, 2, 3)
    assert _close(float(ins), float(f["off"] + f["diag"]))
    assert rel_err(zu.grad.cpu().numpy(), dz) < TOL
    # non-contiguous (channels-last strides) input is made contiguous like z.contiguous() at algorithms.py:1280
    z_cpu = _synth(6, 16, 16, seed=6)
    zc = z_cpu.to(dev).to(memory_format=torch.channels_last)
    ins2, _ = wb.whitening_folded(zc, 2, 3)
    f = wnp.whitening_forward(z_cpu.numpy(), 2, 3)
    assert _close(float(ins2), float(f["off"] + f["diag"]))


def test_edge_semantics():
    import wtpse_b200 as wb

    dev = _dev()
    z = _synth(6, 16, 16).to(dev)
    # single domain: the reference's penalty stays the python int 0 (algorithms.py:105,115); we return a zero tensor
    _, dom = wb.whitening_folded(z, 6, 1)
    assert float(dom) == 0.0
    # empty chunk -> NaN (mean over an empty slice), e.g. the shape net's literal 3 domains with n=3, B=6
    _, _, dom = wb.whitening_terms(z, 3, 3)
    assert np.isnan(float(dom))
    # NaN / Inf in z propagate (caller's guard is Trainer.py:799-800)
    zn = z.clone()
    zn[1, 3, 2, 2] = float("nan")
    ins, dom = wb.whitening_folded(zn, 2, 3)
    assert np.isnan(float(ins)) and np.isnan(float(dom))
    zi = z.clone()
    zi[0, 0, 0, 0] = float("inf")
    ins, _ = wb.whitening_folded(zi, 2, 3)
    assert not np.isfinite(float(ins))
    # margin large enough to clamp both instance terms: zero loss, zero instance gradient
    zz = z.clone().requires_grad_(True)
    off, diag, dom = wb.whitening_terms(zz, 2, 3, 100.0, 1e-5)
    assert float(off) == 0.0 and float(diag) == 0.0
    (off + diag).backward()
    assert float(zz.grad.abs().max()) == 0.0
    # shape / dtype contract
    with pytest.raises(ValueError):
        wb.whitening_folded(torch.randn(6, 8, 16, 16, device=dev), 2, 3)
    with pytest.raises(TypeError):
        wb.whitening_folded(z.double(), 2, 3)
    with pytest.raises(ValueError):
        wb.mmd_penalty(torch.randn(6, 100, device=dev), 2, 3)
    # zero input: sign(0) == 0 off the diagonal, G_ii = eps < 1 on it
    z0 = torch.zeros(6, 16, 8, 8, device=dev, requires_grad=True)
    off, diag, dom = wb.whitening_terms(z0, 2, 3)
    (off + diag + dom).backward()
    assert float(off) == 0.0 and abs(float(diag) - (1 - 1e-5)) < 1e-6 and float(z0.grad.abs().max()) == 0.0


def test_outputs_behave_like_the_reference_tensors():
    """0-dim, attached to autograd, usable in-place by the caller's accumulation (algorithms.py:1263-1267)."""
    import wtpse_b200 as wb

    z = _synth(6, 16, 16).to(_dev()).requires_grad_(True)
    ins, dom = wb.whitening_folded(z, 2, 3)
    assert ins.dim() == 0 and dom.dim() == 0 and ins.requires_grad and dom.requires_grad
    acc = 0
    acc += ins
    acc += ins
    acc /= 3
    ins *= 2.0          # in-place on an output must not trip autograd's view checks
    (acc + dom).backward()
    assert z.grad is not None and torch.isfinite(z.grad).all()
    # only one of the outputs used
    z2 = _synth(6, 16, 16).to(_dev()).requires_grad_(True)
    off, diag, dom = wb.whitening_terms(z2, 2, 3)
    diag.backward()
    assert z2.grad.abs().max() > 0


def test_run_to_run_bit_reproducible():
    import wtpse_b200 as wb

    z = _synth(9, 128, 128).to(_dev())
    outs = []
    for _ in range(3):
        zz = z.clone().requires_grad_(True)
        ins, dom = wb.whitening_folded(zz, 3, 3)
        (ins + dom).backward()
        outs.append((ins.item(), dom.item(), zz.grad.clone()))
    for o in outs[1:]:
        assert o[0] == outs[0][0] and o[1] == outs[0][1] and torch.equal(o[2], outs[0][2])


def test_in_kernel_tail_is_bitwise_identical_to_the_separate_kernels():
    """Forward tail inside the Gram kernel (last-arriving CTA: slot reduction, f_cor, instance terms, MMD, the backward's
    MMD seed) vs the chain of stand-alone kernels (whitening_epilogue.cu): shared device code, same summation orders ->
    the same bits in every output, and therefore in dz.  Also: the ticket area of the workspace is left zero."""
    import wtpse_b200 as wb
    from wtpse_b200 import functional as wf

    lib = wb._lib.load()
    for B, H, W, n, K in ((9, 96, 96, 3, 3), (32, 128, 128, 10, 3), (6, 32, 20, 2, 3), (8, 64, 64, 2, 3), (3, 300, 300, 1, 3),
                          (7, 40, 40, 3, 2), (5, 24, 24, 5, 1), (1, 64, 64, 1, 3)):
        z = _synth(B, H, W, seed=B).to(_dev())
        res = []
        for fused in (1, 0, 1):
            wb._lib.debug_set("fused_tail", fused)
            try:
                zz = z.clone().requires_grad_(True)
                off, diag, dom = wb.whitening_terms(zz, n, K)
                saved = [t.clone() for t in off.grad_fn.saved_tensors[1:]]          # gram, rowstat, domgrad
                (0.7 * off + 1.9 * diag + 1.3 * dom).backward()
                res.append(([off.item(), diag.item(), dom.item()], saved, zz.grad.clone()))
            finally:
                wb._lib.debug_set("fused_tail", 1)
        M = min(B, n * K) if K > 1 else 0
        for other in res[1:]:
            assert [repr(x) for x in res[0][0]] == [repr(x) for x in other[0]], (B, H, res[0][0], other[0])      # NaN-safe equality
            assert torch.equal(res[0][1][0], other[1][0]) and torch.equal(res[0][1][1], other[1][1])
            assert torch.equal(res[0][1][2][:M], other[1][2][:M])
            assert torch.equal(res[0][2], other[2])
        ws, _ = wf._workspace(lib, B, H * W, z.device)
        nt = lib.wtpse_whitening_ticket_bytes(B)
        assert nt >= 4 * (B + 1) and int(ws[:nt].count_nonzero()) == 0


def test_dirty_tickets_are_the_callers_responsibility_and_clean_ones_stay_clean():
    """The C-ABI contract of wtpse_whitening_ticket_bytes: many back-to-back forwards on one workspace that was zeroed
    once give identical results (every call leaves the tickets zero)."""
    import ctypes

    import wtpse_b200 as wb
    from wtpse_b200.functional import _forward_outputs, _ptr, _stream_ptr

    lib = wb._lib.load()
    dev = _dev()
    B, H, n = 12, 64, 4
    z = _synth(B, H, H, seed=5).to(dev)
    nbytes = lib.wtpse_whitening_workspace_bytes(B, H * H)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev).random_(0, 255)              # garbage everywhere ...
    ws[:lib.wtpse_whitening_ticket_bytes(B)] = 0                                          # ... except the ticket prefix
    outs = []
    for _ in range(5):
        losses, (gram, rowstat, domgrad) = _forward_outputs(B, dev)
        wb._lib.check(lib.wtpse_whitening_forward(_ptr(z), B, 16, H * H, n, 3, 0.0, 1e-5, _ptr(losses), _ptr(gram), _ptr(rowstat),
                                                  _ptr(domgrad), _ptr(ws), nbytes, _stream_ptr(dev)))
        outs.append((losses.clone(), gram.clone(), domgrad.clone()))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))
    ref = wb.whitening_terms(z, n, 3)
    assert [float(x) for x in ref] == [float(x) for x in outs[0][0][:3]]


def test_many_mmd_samples_fall_back_to_the_epilogue_kernel():
    """An MMD over ~100+ samples does not fit the last CTA's pipeline buffers: the forward tail runs as separate kernels
    (the backward is one launch either way)."""
    import wtpse_b200 as wb
    from oracle import whitening_np as wnp

    B, n, K = 70, 23, 3
    z_cpu = _synth(B, 20, 20, seed=70)
    z = z_cpu.to(_dev()).requires_grad_(True)
    ins, dom = wb.whitening_folded(z, n, K)
    (ins + dom).backward()
    f = wnp.whitening_forward(z_cpu.numpy(), n, K)
    dz, _ = wnp.whitening_backward(z_cpu.numpy(), f, n, K)
    assert _close(float(ins), float(f["off"] + f["diag"])) and _close(float(dom), float(f["dom"]), scale=1.0)
    assert rel_err(z.grad.cpu().numpy(), dz) < TOL


def test_dropin_methods_on_reference_shaped_objects():
    """dropin.bind() on stand-ins carrying exactly the attributes the reference's methods read
    (self.margin, self.eps, self.mmd_operator.{batch_size,domain_num}); algorithms.py:1140-1155."""
    import types
    import wtpse_b200 as wb
    from oracle import whitening_np as wnp

    class WT_PSE:  # noqa: N801  (name checked by dropin.bind)
        pass

    class ShapeVariationalDist_x:  # noqa: N801
        pass

    main, shape = WT_PSE(), ShapeVariationalDist_x()
    for o, K in ((main, 2), (shape, 3)):
        o.margin, o.eps = 0, 1e-5
        o.mmd_operator = types.SimpleNamespace(batch_size=3, domain_num=K)
    wb.dropin.bind(main)
    wb.dropin.bind(shape)
    z_cpu = _synth(9, 16, 16)
    z = z_cpu.to(_dev())
    ins, dom = main.compute_whitening_loss(z)
    f = wnp.whitening_forward(z_cpu.numpy(), 3, 2)
    assert _close(float(ins), float(f["off"] + f["diag"])) and _close(float(dom), float(f["dom"]), scale=1.0)
    off, diag, dom = shape.compute_whitening_loss(z)
    f = wnp.whitening_forward(z_cpu.numpy(), 3, 3)
    assert _close(float(off), float(f["off"])) and _close(float(diag), float(f["diag"]))
    a, b = torch.randn(6, 1, 16, 16, device=_dev()), torch.randn(6, 1, 16, 16, device=_dev())
    assert _close(float(shape.wasser_distance(a, b)), float(torch.nn.functional.mse_loss(a, b)))


def test_host_plan_matches_device_path():
    import wtpse_b200 as wb

    z_cpu = _synth(6, 64, 64).pin_memory()
    dz_host = torch.empty_like(z_cpu).pin_memory()
    plan = wb.HostPlan(6, 64, 64)
    off, diag, dom = plan.run(z_cpu, 2, 3, 0.0, 1e-5, (1.0, 1.0, 1.0), dz_host)
    plan.close()
    z = z_cpu.to(_dev()).requires_grad_(True)
    o2, d2, m2 = wb.whitening_terms(z, 2, 3)
    (o2 + d2 + m2).backward()
    assert off == float(o2) and diag == float(d2) and dom == float(m2)
    assert torch.equal(dz_host, z.grad.cpu())


def test_host_plan_pipelined_submissions():
    """Asynchronous submit()/wait(): four steps in flight over two device slots return what run() returns."""
    import wtpse_b200 as wb

    zs = [_synth(6, 32, 32, seed=40 + i).pin_memory() for i in range(4)]
    dzs = [torch.empty_like(z).pin_memory() for z in zs]
    plan = wb.HostPlan(6, 32, 32)
    outs = [plan.submit(z, 2, 3, 0.0, 1e-5, (1.0, 0.5, 2.0), dz) for z, dz in zip(zs, dzs)]
    plan.wait()
    for z, dz, out in zip(zs, dzs, outs):
        ref_dz = torch.empty_like(z)
        ref = plan.run(z, 2, 3, 0.0, 1e-5, (1.0, 0.5, 2.0), ref_dz)
        assert (out[0], out[1], out[2]) == ref and torch.equal(dz, ref_dz)
    plan.close()


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + the op-sequence restatement on the GPU
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,n", [(32, 512, 512, 10), (16, 1024, 1024, 5), (64, 1024, 1024, 21)])
def test_full_size_properties(B, H, W, n):
    import wtpse_b200 as wb
    from oracle import whitening_torch as wt

    dev = _dev()
    z = _synth(B, H, W, seed=42).to(dev)
    zz = z.clone().requires_grad_(True)
    off, diag, dom = wb.whitening_terms(zz, n, 3)
    (off + diag + dom).backward()
    dz = zz.grad
    G = wb.gram_matrix(z, eps=0.0)

    # (1) the reference's operator sequence executed by ATen on the same GPU
    zr = z.clone().requires_grad_(True)
    r_off, r_diag, r_dom, r_G = wt.whitening_terms(zr, n, 3)
    (r_off + r_diag + r_dom).backward()
    assert _close(float(off), float(r_off)) and _close(float(diag), float(r_diag))
    assert _close(float(dom), float(r_dom), scale=1.0)
    assert _rel_err_dev(dz, zr.grad) < TOL
    del zr, r_G
    # (2) homogeneity: G(a z) = a^2 G(z)
    G2 = wb.gram_matrix(2.0 * z, eps=0.0)
    assert _rel_err_dev(G2, 4.0 * G) < 1e-6
    # (3) pixel permutation invariance of the Gram (sum over pixels), exact up to summation order
    perm = torch.randperm(H * W, device=dev)
    Gp = wb.gram_matrix(z.view(B, 16, -1)[:, :, perm].view_as(z).contiguous(), eps=0.0)
    assert _rel_err_dev(Gp, G) < 1e-5
    # (4) adjoint identity: <dz_b, z_b> = sum_ij M_ij G_ij (P-1) with dz = M z  =>  check via a second apply:
    #     the loss is degree-2 homogeneous in z through G, so <dz, z> = 2 * sum_b <dL/dG_b, G_b - eps I>
    zr2 = z.clone().requires_grad_(True)
    o2, d2, m2, G_t = wt.whitening_terms(zr2, n, 3)
    gG = torch.autograd.grad(o2 + d2 + m2, G_t)[0]
    lhs = float((dz.double() * z.double()).sum())
    rhs = float(2.0 * (gG.double() * (G_t.detach().double() - 1e-5 * torch.eye(16, device=dev, dtype=torch.float64))).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(rhs), 1e-12)
