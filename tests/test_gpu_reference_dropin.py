"""The drop-in on the reference's OWN classes, on the GPU (VERDICT r1 "missing #1").

``oracle/_ref`` holds a byte-for-byte copy of the reference's algorithms.py / shape_networks.py (made by
``oracle/build_ref.py`` in the build container, SHA-256 checked against the committed manifest at import).  Here the
real ``algorithms.WT_PSE`` and ``shape_networks.ShapeVariationalDist_x`` -- with their hard-coded ``.cuda()`` sites
(algorithms.py:1162-1164,1296,1305) running as written -- execute ``update()`` (algorithms.py:1216-1275,
shape_networks.py:512-558) and a full ``Trainer.train_epoch``-order iteration (Trainer.py:762-925) twice: stock, and
with the CUDA path installed underneath by ``wtpse_b200.dropin``.

Tolerances: every returned loss within rel 1e-5 (the MMD scalar: 1e-5 * max(|ref|, 1), it cancels
O(1) terms); every parameter gradient within 1e-5 of its tensor's max-abs (floored at 1e-4 of the run's largest gradient:
some tensors carry rounding noise only), on top of the reference's own run-to-run difference (four stock runs with the
same seed: ATen's bilinear-upsampling backward accumulates with float atomics, and at 9x3x256x256 and above that noise --
up to 8e-4 of a tensor's max-abs -- is as large as anything the drop-in changes: tools/dropin_noise_probe.py,
profiles/r2_dropin_noise_probe.txt).  A tensor outside that bar must be within 1e-4 and twice as close to the reference
as the reference's float32 result is to its own float64 one (`_check_grads`).
"""
import contextlib
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

HP = {"whitening": True, "margin": 0, "shape_prior": True, "shape_attention": True, "cat_shape": False,
      "shape_attention_coeffient": 0.3, "shape_start": 0.5, "instance_wt_gm": 1, "domain_wt_gm": 1, "multi-turn": 1}

# (n per domain, H = W): 6x3x64x64, the reference's default 9x3x256x256 (train.py:58-67,89), 15x3x512x512 (BASELINE configs[2])
SHAPES = [(2, 64), (3, 256), (5, 512)]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.fail("oracle/_ref is missing: run __graft_entry__.build() (or python -m oracle.build_ref) in the build container")
    alg, sn, _ = ref_shim.load()
    return alg, sn


@pytest.fixture(autouse=True)
def _fp32_deterministic():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
           torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False          # utils.py:58-65
    yield
    (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
     torch.backends.cudnn.benchmark) = old


def _batch(n, S, dev, seed=3):
    import wtpse_b200 as wb

    return wb.synthetic.fundus_batch(n, 3, S, S, dev, seed=seed)


def _grads(*nets):
    out = {}
    for i, m in enumerate(nets):
        for name, p in m.named_parameters():
            if p.grad is not None:
                out["%d.%s" % (i, name)] = p.grad.detach().clone()
    return out


def _update_pair(main, shape, image, mask):
    """One WT_PSE.update + backward, then one ShapeVariationalDist_x.update + backward (as Trainer.py:779-825 does,
    minus the optimizer steps); returns (10 loss floats, gradients of both networks)."""
    main.zero_grad(set_to_none=True)
    shape.zero_grad(set_to_none=True)
    torch.manual_seed(101)
    out = main.update(image, mask, step=0, plot_show=0, two_stage_inputs=image, sp_mask=mask, two_step=True)
    logits, ins, dom = out[0], out[3], out[4]
    assert ins.dim() == 0 and dom.dim() == 0 and ins.is_cuda and ins.requires_grad and dom.requires_grad
    bce = torch.nn.functional.binary_cross_entropy(torch.sigmoid(logits), mask)
    (bce + ins + dom).backward()
    g_main = _grads(main)
    main.zero_grad(set_to_none=True)
    torch.manual_seed(102)
    kd, tot, ij, ii, dom_s = shape.update(main, image, mask, step=0, plot_show=0, two_stage_inputs=image, two_step=True)
    (kd + tot + dom_s).backward()
    g_shape = _grads(shape, main)            # incl. the (discarded) teacher gradients of shape_networks.py:524
    losses = {"bce": bce, "ins": ins, "dom": dom, "kd": kd, "sh_total": tot, "sh_ij": ij, "sh_ii": ii, "sh_dom": dom_s}
    losses = {k: float(v.detach()) for k, v in losses.items()}
    g_main.update({"s." + k: v for k, v in g_shape.items()})
    return losses, g_main, logits.detach().clone()


def _check_losses(got, want):
    for k, w in want.items():
        tol = 1e-5 * max(abs(w), 1.0) if "dom" in k else 1e-5 * abs(w)
        assert abs(got[k] - w) <= tol, (k, got[k], w)


def _grad_scale(k, want, gmax):
    """What a gradient tensor's error is measured against: its own max-abs, floored at 1e-4 of the run's largest gradient
    (some tensors carry rounding noise only); for a bias also the max-abs of its layer's weight gradient."""
    scale = max(float(want[k].abs().max()), 1e-4 * gmax)
    if k.endswith(".bias") and k[:-4] + "weight" in want:
        # a bias gradient is the plain sum of the same output gradients whose products with the layer input make up
        # the weight gradient: its rounding error scales with the layer's gradient, not with its own (possibly zero) value
        # (the bias of a convolution in front of a BatchNorm has an exactly-zero true gradient: both implementations return
        # the rounding noise of a sum over B*H*W output gradients, which grows with the image size -- 2.5e-9 at 15x3x512x512
        # against 0.87 for the run's largest gradient; floor its scale at 3e-4 of that)
        scale = max(scale, float(want[k[:-4] + "weight"].abs().max()), 3e-4 * gmax)
    return scale


@contextlib.contextmanager
def _float32_noise():
    """`torch.randn_like` (algorithms.py:1072, the reparameterisation noise) draws in float32 whatever the dtype asked for,
    so a float64 run of the reference sees the noise of the float32 run (same seed, same Philox stream)."""
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: orig(t.float(), *a, **k).to(t.dtype)
    try:
        yield
    finally:
        torch.randn_like = orig


def _fp64_truth(main, shape, image, mask):
    """The same update pair by the UNMODIFIED reference classes in float64 (copies of the two networks, same weights, same
    noise): what the float32 gradients are roundings of."""
    m64, s64 = copy.deepcopy(main).double(), copy.deepcopy(shape).double()
    with _float32_noise():
        _, g64, _ = _update_pair(m64, s64, image.double(), mask.double())
    return g64


def _check_grads(got, want, noise, truth=None):
    """Per tensor: max|got - want| <= 1e-5 * scale (`_grad_scale`) on top of the reference's own run-to-run difference
    (`noise`: the gradients of one or more further stock runs with the same seed).

    A tensor that misses this bar may still pass on evidence that the reference's float32 value of it is itself not
    defined to that precision: `truth` (a callable, evaluated only when needed) returns the reference's float64 gradients
    of the same update; the tensor passes if it is within 1e-4 * scale of the reference AND at most half as far from
    the reference's float32 value as that value is from the reference's own float64 one.  (Gradients of the first
    convolutions are sums over B*H*W products with heavy cancellation: a 1e-6 relative change of the loss block's gradient
    -- a different summation order -- moves them by a few 1e-5 of their max-abs, while the reference's own float32 result
    sits 1e-3..1e-2 away from its float64 one; `tools/dropin_noise_probe.py` prints the three distances.)"""
    assert got.keys() == want.keys()
    runs = [want] + (list(noise) if isinstance(noise, (list, tuple)) else [noise])
    gmax = max(float(w.abs().max()) for w in want.values())
    worst = (0.0, None)
    t64 = None
    for k, w in want.items():
        scale = _grad_scale(k, want, gmax)
        err = float((got[k] - w).abs().max())
        # `runs`: repeated stock runs with the same seed.  They differ among themselves (ATen's bilinear-upsampling backward
        # accumulates with float atomics) by as much as the drop-in differs from any of them (tools/dropin_noise_probe.py:
        # up to 8e-4 of a tensor's max-abs at 15x3x512x512, stock against stock); the drop-in's own difference goes through
        # the same amplification, so allow a few times the largest pairwise difference seen
        floor = (8.0 if len(runs) == 2 else 5.0) * max(float((a[k] - b[k]).abs().max()) for i, a in enumerate(runs) for b in runs[:i])
        rel = max(err - floor, 0.0) / scale
        if rel > worst[0]:
            worst = (rel, k)
        if rel > 1e-5 and truth is not None:
            if t64 is None:
                try:
                    t64 = truth()
                except Exception as exc:              # no float64 evidence: the plain bar decides
                    raise AssertionError((k, err, scale, floor, "float64 run of the reference failed: %r" % (exc,)))
            ref_err = float((w.double() - t64[k]).abs().max())
            assert rel <= 1e-4 and err <= 0.5 * ref_err, (k, err, scale, floor, ref_err)
            continue
        assert rel <= 1e-5, (k, err, scale, floor)
    return worst


def _models(alg, sn, n, dev, seed=0):
    torch.manual_seed(seed)
    main = alg.WT_PSE(3, 1, dict(HP), dev, False, per_domain_batch=n, source_domain_num=3).cuda().train()
    shape = sn.ShapeVariationalDist_x(dict(HP), dev, 1, number_source_domain=3, batch_size=n).cuda().train()
    return main, shape


@pytest.mark.parametrize("n,S", SHAPES)
def test_reference_update_stock_vs_dropin_install(ref, n, S):
    """(a) class-level install(): WT_PSE.compute_whitening_loss, ShapeVariationalDist_x.compute_whitening_loss /
    wasser_distance, compute_MMD.forward rebound on the reference's classes."""
    import wtpse_b200 as wb

    alg, sn = ref
    dev = torch.device("cuda:0")
    main, shape = _models(alg, sn, n, dev)
    assert type(main).__module__ == "algorithms" and type(shape).__module__ == "shape_networks"
    image, od, _ = _batch(n, S, dev)
    lib = wb._lib.load()

    stock, g_stock, logits_stock = _update_pair(main, shape, image, od)
    g_again = [_update_pair(main, shape, image, od)[1] for _ in range(3)]                     # the reference's own run-to-run noise
    lib.wtpse_profile_reset()
    saved = wb.dropin.install(alg, sn)
    try:
        assert alg.WT_PSE.compute_whitening_loss is wb.dropin.wt_pse_compute_whitening_loss
        ours, g_ours, logits_ours = _update_pair(main, shape, image, od)
    finally:
        wb.dropin.uninstall(saved)
    assert alg.WT_PSE.compute_whitening_loss is not wb.dropin.wt_pse_compute_whitening_loss
    # the library ran exactly: 4 loss evaluations x (ONE forward launch + ONE backward launch) + KD MSE forward + backward
    assert int(lib.wtpse_profile_launches(-1)) == 4 * 1 + 4 * 1 + 2
    _check_losses(ours, stock)
    _check_grads(g_ours, g_stock, g_again, truth=lambda: _fp64_truth(main, shape, image, od))
    assert torch.equal(logits_ours, logits_stock)          # the drop-in does not touch the backbone or its RNG stream


@pytest.mark.parametrize("n,S", SHAPES[:2])
def test_reference_update_stock_vs_bind_with_fused_tail(ref, n, S):
    """(b) instance-level bind(fuse_relu=True): additionally DeepWT.forward (algorithms.py:1091-1117) takes the fused
    Gram + ReLU pass."""
    import wtpse_b200 as wb

    alg, sn = ref
    dev = torch.device("cuda:0")
    main, shape = _models(alg, sn, n, dev)
    image, od, _ = _batch(n, S, dev, seed=8)
    stock, g_stock, logits_stock = _update_pair(main, shape, image, od)
    g_again = [_update_pair(main, shape, image, od)[1] for _ in range(3)]
    main_b, shape_b = copy.deepcopy(main), copy.deepcopy(shape)
    wb.dropin.bind(main_b, fuse_relu=True)
    wb.dropin.bind(shape_b, fuse_relu=True)
    ours, g_ours, logits_ours = _update_pair(main_b, shape_b, image, od)
    # every term the student's fused tail parked was consumed exactly once; the teacher pass of the shape update
    # (main_b.wt_model, shape_networks.py:516) parks terms nobody asks for -- they are dropped at its next forward
    assert not shape_b.wt_model._pending_terms
    _check_losses(ours, stock)
    _check_grads(g_ours, g_stock, g_again, truth=lambda: _fp64_truth(main, shape, image, od))
    err = float((logits_ours - logits_stock).abs().max() / logits_stock.abs().max())
    assert err <= 1e-6, err                                   # relu(z) of the fused pass is bit-identical to ATen's
    # the class itself is untouched by bind()
    assert "compute_whitening_loss" not in vars(main) and "compute_whitening_loss" in vars(main_b)


@pytest.mark.parametrize("n,S", [(3, 256)])
def test_trainer_order_iteration_stock_vs_dropin(ref, n, S):
    """(c) Trainer.py:779-914 on the real classes (oracle/ref_iteration.py): four update()s, ROI step, host syncs."""
    import wtpse_b200 as wb
    from oracle import ref_iteration as ri

    alg, sn = ref
    dev = torch.device("cuda:0")
    image, od, oc = _batch(n, S, dev, seed=21)

    def run(step_optim, install, double=False):
        nets, optims = ri.build_reference_models(alg, sn, dict(HP), n, 3, dev, seed=0)
        saved = wb.dropin.install(alg, sn) if install else None
        try:
            torch.manual_seed(55)
            if double:                # the reference in float64 (same weights, same noise): evidence for `_check_grads`
                nets = [m.double() for m in nets]
                with _float32_noise():
                    out = ri.trainer_iteration(nets, optims, image.double(), od.double(), oc.double(), dict(HP), step_optim=False)
            else:
                out = ri.trainer_iteration(nets, optims, image.clone(), od, oc, dict(HP), step_optim=step_optim)
        finally:
            if saved:
                wb.dropin.uninstall(saved)
        return out, _grads(*nets), nets

    # gradients of all four networks with the weights frozen
    stock, g_stock, _ = run(False, False)
    g_again = [run(False, False)[1] for _ in range(3)]
    ours, g_ours, _ = run(False, True)
    scalars = [k for k, v in stock.items() if torch.is_tensor(v) and v.dim() == 0]
    assert len(scalars) >= 15
    _check_losses({k: float(ours[k]) for k in scalars}, {k: float(stock[k]) for k in scalars})
    assert torch.equal(ours["od_pred"], stock["od_pred"])
    _check_grads(g_ours, g_stock, g_again, truth=lambda: run(False, False, double=True)[1])

    # and with the four Adam steps taken (Trainer.py:805,825,892,914): later sub-steps see the updated teacher
    stock, _, nets_s = run(True, False)
    ours, _, nets_o = run(True, True)
    for k in scalars:
        w, g = float(stock[k]), float(ours[k])
        assert abs(g - w) <= 1e-4 * max(abs(w), 1.0 if "dom" in k else 0.0), (k, g, w)
    # Adam's first step is lr * g / (|g| + 1e-8): sign-like, so rounding-noise gradients (conv biases in front of a
    # BatchNorm) may move a weight by up to 2 * lr in either run; everything else agrees far below that
    for a, b in zip(nets_s, nets_o):
        for (name, p), q in zip(a.named_parameters(), b.parameters()):
            assert float((p - q).abs().max()) <= 2 * 5e-4 + 1e-6, name


def test_vendored_reference_is_unmodified():
    from oracle import build_ref

    assert build_ref.vendored_present(), "oracle/_ref missing"
    assert build_ref.verify(build_ref.VENDOR_DIR)


# ---- the reference's other hparams branches (VERDICT r1 missing #5) ----------------------------------------------------
# Of the 16 combinations of (whitening, shape_prior, shape_attention, cat_shape) the unmodified reference can run
# exactly six: whitening = shape_prior = shape_attention = True with cat_shape either way, and whitening = shape_prior
# = False (any attention / cat flag; update() then returns python ints, algorithms.py:1270-1275, shape_networks.py:
# 557-558).  The others die inside the reference itself: shape_attention=False leaves `z_posterior_attention_mask`
# unbound (algorithms.py:1271), shape_prior=False with whitening=True leaves `whiting_outputs1` unbound (:1258), and
# whitening=False with shape_prior=True feeds a 4-channel cat into the teacher's 2-channel stem (algorithms.py:1027).
# (Probed by running every combination through oracle/ref_shim.py in the build container.)


def _ours_like(ref_main, ref_shape, hp, n, dev):
    """This repository's classes carrying the reference instances' weights (strict state-dict interchange)."""
    from wtpse_b200 import segmentation as seg

    main = seg.WT_PSE(3, 1, dict(hp), dev, False, per_domain_batch=n, source_domain_num=3).to(dev).train()
    shape = seg.ShapeVariationalDist_x(dict(hp), dev, 1, number_source_domain=3, batch_size=n).to(dev).train()
    main.load_state_dict(ref_main.state_dict(), strict=True)
    shape.load_state_dict(ref_shape.state_dict(), strict=True)
    return main, shape


@pytest.mark.parametrize("n,S", [(2, 64), (3, 256)])
def test_cat_shape_branch_stock_vs_dropin_and_vs_our_classes(ref, n, S):
    """cat_shape=True (algorithms.py:1253: outc sees cat([fused embedding, z_posterior]), 9 channels)."""
    import wtpse_b200 as wb

    alg, sn = ref
    dev = torch.device("cuda:0")
    hp = dict(HP, cat_shape=True)
    torch.manual_seed(0)
    main = alg.WT_PSE(3, 1, dict(hp), dev, False, per_domain_batch=n, source_domain_num=3).cuda().train()
    shape = sn.ShapeVariationalDist_x(dict(hp), dev, 1, number_source_domain=3, batch_size=n).cuda().train()
    assert main.outc[0].weight.shape[1] == 9
    image, od, _ = _batch(n, S, dev, seed=5)
    mine = _ours_like(main, shape, hp, n, dev)              # before any update(): same BatchNorm running statistics

    stock, g_stock, logits_stock = _update_pair(main, shape, image, od)
    g_again = [_update_pair(main, shape, image, od)[1] for _ in range(3)]
    saved = wb.dropin.install(alg, sn)
    try:
        ours, g_ours, logits_ours = _update_pair(main, shape, image, od)
    finally:
        wb.dropin.uninstall(saved)
    _check_losses(ours, stock)
    _check_grads(g_ours, g_stock, g_again, truth=lambda: _fp64_truth(main, shape, image, od))
    assert torch.equal(logits_ours, logits_stock)

    # our own WT_PSE / ShapeVariationalDist_x on the same weights, same RNG stream: the loss scalars do not depend on
    # the backbone's arithmetic path (whitening features only) -> 1e-5; logits go through the re-associated U-Net -> 1e-4
    got, _, logits_mine = _update_pair(mine[0], mine[1], image, od)
    assert mine[0].outc[0].weight.shape[1] == 9
    for k in ("ins", "dom", "sh_ij", "sh_ii", "sh_total", "sh_dom"):
        w = stock[k]
        assert abs(got[k] - w) <= 1e-5 * max(abs(w), 1.0 if "dom" in k else 0.0), (k, got[k], w)
    assert abs(got["kd"] - stock["kd"]) <= 1e-4 * abs(stock["kd"]) and abs(got["bce"] - stock["bce"]) <= 1e-4 * abs(stock["bce"])
    assert float((logits_mine - logits_stock).abs().max()) <= 1e-4 * float(logits_stock.abs().max())


@pytest.mark.parametrize("flags", [dict(shape_attention=True, cat_shape=False), dict(shape_attention=False, cat_shape=True)])
def test_whitening_off_shape_prior_off_returns_python_zeros(ref, flags):
    """hparams['whitening'] == hparams['shape_prior'] == False: WT_PSE.update returns (logits, 0, 0, 0, 0)
    (algorithms.py:1274-1275) and ShapeVariationalDist_x.update returns five python 0s (shape_networks.py:557-558) --
    plain ints, no library launch, with or without the drop-in, and the same from this repository's classes."""
    import wtpse_b200 as wb

    alg, sn = ref
    dev = torch.device("cuda:0")
    n, S = 2, 64
    hp = dict(HP, whitening=False, shape_prior=False, **flags)
    torch.manual_seed(0)
    main = alg.WT_PSE(3, 1, dict(hp), dev, False, per_domain_batch=n, source_domain_num=3).cuda().train()
    shape = sn.ShapeVariationalDist_x(dict(hp), dev, 1, number_source_domain=3, batch_size=n).cuda().train()
    mine = _ours_like(main, shape, hp, n, dev)
    image, od, _ = _batch(n, S, dev, seed=6)
    lib = wb._lib.load()

    def run(m, s):
        m.zero_grad(set_to_none=True)
        out = m.update(image, od, step=0, plot_show=0, two_stage_inputs=image, sp_mask=od, two_step=True)
        assert len(out) == 5 and all(type(v) is int and v == 0 for v in out[1:])
        sh = s.update(m, image, od, step=0, plot_show=0, two_stage_inputs=image, two_step=True)
        assert len(sh) == 5 and all(type(v) is int and v == 0 for v in sh)
        bce = torch.nn.functional.binary_cross_entropy(torch.sigmoid(out[0]), od)
        (bce + 1 * out[3] + 1 * out[4]).backward()               # Trainer.py:802: ints add into the loss
        return out[0].detach().clone(), float(bce), _grads(m)

    logits_stock, bce_stock, g_stock = run(main, shape)
    _, _, g_again = run(main, shape)
    lib.wtpse_profile_reset()
    saved = wb.dropin.install(alg, sn)
    try:
        logits_ours, bce_ours, g_ours = run(main, shape)
    finally:
        wb.dropin.uninstall(saved)
    assert int(lib.wtpse_profile_launches(-1)) == 0            # nothing of the loss path runs in this branch
    assert torch.equal(logits_ours, logits_stock) and bce_ours == bce_stock
    _check_grads(g_ours, g_stock, g_again)

    logits_mine, bce_mine, g_mine = run(*mine)
    assert float((logits_mine - logits_stock).abs().max()) <= 1e-4 * float(logits_stock.abs().max())
    assert abs(bce_mine - bce_stock) <= 1e-5 * bce_stock
    assert g_mine.keys() == g_stock.keys()                       # the same parameters receive gradients


def test_trainer_iteration_multi_turn_2(ref):
    """hparams['multi-turn'] = 2 (Trainer.py:811, 895): each shape network is updated twice per iteration, the second
    time against its own once-stepped weights."""
    import wtpse_b200 as wb
    from oracle import ref_iteration as ri

    alg, sn = ref
    dev = torch.device("cuda:0")
    n, S = 2, 128
    hp = dict(HP)
    hp["multi-turn"] = 2
    image, od, oc = _batch(n, S, dev, seed=22)

    def run(install):
        nets, optims = ri.build_reference_models(alg, sn, dict(hp), n, 3, dev, seed=0)
        lib = wb._lib.load()
        lib.wtpse_profile_reset()
        saved = wb.dropin.install(alg, sn) if install else None
        try:
            torch.manual_seed(56)
            out = ri.trainer_iteration(nets, optims, image.clone(), od, oc, dict(hp), step_optim=True)
        finally:
            if saved:
                wb.dropin.uninstall(saved)
        return out, nets, int(lib.wtpse_profile_launches(-1))

    stock, nets_s, _ = run(False)
    ours, nets_o, launches = run(True)
    # 2 main updates x 2 embeddings + 2 shape networks x 2 turns x 2 embeddings = 12 loss evaluations (fwd + bwd launch
    # each), 4 KD MSEs (fwd + bwd each)
    assert launches == 12 * 2 + 4 * 2, launches
    scalars = [k for k, v in stock.items() if torch.is_tensor(v) and v.dim() == 0]
    for k in scalars:
        w, g = float(stock[k]), float(ours[k])
        assert abs(g - w) <= 1e-4 * max(abs(w), 1.0 if "dom" in k else 0.0), (k, g, w)
    for a, b in zip(nets_s, nets_o):
        for (name, p), q in zip(a.named_parameters(), b.parameters()):
            assert float((p - q).abs().max()) <= 2 * 2 * 5e-4 + 1e-6, name          # two sign-like Adam steps
