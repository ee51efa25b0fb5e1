"""Track W (wavelet) GPU tests.  PARITY UNPINNED: the reference has no wavelet code, so these check the CUDA
kernels against this repository's own float64 specification (oracle/wavelet_np.py) and against the identities
every orthonormal DWT satisfies: perfect reconstruction, Parseval, adjointness."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _maps(shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def _safe_maps(shape, J, seed, wavelet="db2"):
    """Maps whose detail coefficients are bounded away from zero (the inverse transform of coefficients with |c| in
    [0.1, 1]): the sign pattern -- the only discontinuity of the L1 loss -- cannot depend on rounding, so two correct
    implementations with different operation orders (the filter bank and its factored form, csrc/wavelet_db2.cu) agree to
    fp32 rounding.  With plain random maps one coefficient in ~10^6 lies within a few ulps of zero and flips."""
    import wtpse_b200 as wb

    g = torch.Generator().manual_seed(seed)
    c = (0.1 + 0.9 * torch.rand(*shape, generator=g)) * (2.0 * torch.randint(0, 2, shape, generator=g) - 1.0)
    return wb.idwt2d(c.to(_dev()), wavelet, J).contiguous()


@pytest.mark.parametrize("wavelet", ["haar", "db2"])
@pytest.mark.parametrize("shape,J", [((3, 2, 32, 32), 3), ((2, 2, 64, 48), 4), ((1, 1, 16, 80), 2), ((4, 2, 8, 8), 1),
                                     ((2, 2, 128, 64), 4), ((1, 2, 64, 192), 3), ((3, 1, 64, 64), 1), ((1, 1, 256, 128), 2)])
def test_dwt_matches_spec_and_identities(wavelet, shape, J):
    import wtpse_b200 as wb
    from oracle import wavelet_np as wn

    x_cpu = _maps(shape, seed=sum(shape) + J)
    x = x_cpu.to(_dev())
    c = wb.dwt2d(x, wavelet, J)
    assert rel_err(c.cpu().numpy(), wn.dwt2d(x_cpu.numpy(), wavelet, J)) < TOL
    # perfect reconstruction and Parseval
    xr = wb.idwt2d(c, wavelet, J)
    assert rel_err(xr.cpu().numpy(), x_cpu.numpy()) < TOL
    assert abs(float((c.double() ** 2).sum()) - float((x.double() ** 2).sum())) <= TOL * float((x.double() ** 2).sum())
    # adjoint identity <W x, y> = <x, W^T y>
    y = _maps(shape, seed=99).to(_dev())
    lhs = float((c.double() * y.double()).sum())
    rhs = float((x.double() * wb.idwt2d(y, wavelet, J).double()).sum())
    assert abs(lhs - rhs) <= TOL * max(abs(lhs), 1.0)
    # autograd through the transform pair
    xg = x.clone().requires_grad_(True)
    (wb.dwt2d(xg, wavelet, J) * y).sum().backward()
    assert rel_err(xg.grad.cpu().numpy(), wn.idwt2d(y.cpu().numpy(), wavelet, J)) < TOL


def test_dwt_hand_derived_known_answers():
    """Constant, unit impulse and ramp through the CUDA transform, against closed forms derived by hand in
    oracle/wavelet_np.py's header -- no oracle code in the comparison."""
    import wtpse_b200 as wb

    s2, s3 = np.sqrt(2.0), np.sqrt(3.0)
    h = {"haar": np.array([1, 1]) / s2, "db2": np.array([1 + s3, 3 + s3, 3 - s3, 1 - s3]) / (4 * s2)}
    g = {"haar": np.array([1, -1]) / s2, "db2": np.array([1 - s3, -(3 - s3), 3 + s3, -(1 + s3)]) / (4 * s2)}
    N = 16
    for wv in ("haar", "db2"):
        F = len(h[wv])
        # rows = the N unit impulses, repeated down the H axis (constant columns: column pass gives sqrt2 * a, 0)
        x = torch.eye(N).repeat_interleave(N, 0).reshape(N, 1, N, N).to(_dev())        # map m: every row = delta_m
        c = wb.dwt2d(x, wv, 1).cpu().numpy()[:, 0]
        for m in range(N):
            for n in range(N // 2):
                k = (m - 2 * n) % N
                assert abs(c[m, 0, n] / s2 - (h[wv][k] if k < F else 0.0)) < 1e-6            # LL row 0 = sqrt2 * a
                assert abs(c[m, 0, N // 2 + n] / s2 - (g[wv][k] if k < F else 0.0)) < 1e-6   # LH row 0 = sqrt2 * d
            assert np.abs(c[m, N // 2:, :]).max() < 1e-6                                     # HL, HH: zero
        # constant map, 3 levels: LL_3 = 8 c, every detail 0
        c = wb.dwt2d(torch.full((1, 1, N, N), 0.5, device=_dev()), wv, 3).cpu().numpy()[0, 0]
        assert np.abs(c[:2, :2] - 4.0).max() < 1e-5
        c[:2, :2] = 0
        assert np.abs(c).max() < 1e-6
    ramp = torch.arange(N, dtype=torch.float32).repeat(N, 1).reshape(1, 1, N, N).to(_dev())
    c = wb.dwt2d(ramp, "haar", 1).cpu().numpy()[0, 0, 0] / s2
    assert np.allclose(c[:N // 2], (4 * np.arange(N // 2) + 1) / s2, atol=1e-5) and np.allclose(c[N // 2:], -1 / s2, atol=1e-6)
    c = wb.dwt2d(ramp, "db2", 1).cpu().numpy()[0, 0, 0] / s2
    n = np.arange(N // 2 - 1)
    assert np.allclose(c[:N // 2 - 1], 2 * s2 * n + (3 - s3) / s2, atol=1e-5) and np.abs(c[N // 2:N - 1]).max() < 2e-6
    assert abs(c[N // 2 - 1] - (h["db2"][0] * (N - 2) + h["db2"][1] * (N - 1) + h["db2"][3])) < 1e-5
    assert abs(c[N - 1] - (g["db2"][0] * (N - 2) + g["db2"][1] * (N - 1) + g["db2"][3])) < 1e-5


@pytest.mark.parametrize("wavelet,J,weights", [("haar", 3, None), ("db2", 4, None), ("db2", 2, (0.5, 2.0)), ("haar", 1, (3.0,))])
def test_wavelet_shape_loss_forward_backward(wavelet, J, weights):
    import wtpse_b200 as wb
    from oracle import wavelet_np as wn

    # softmax OC/OD-like probability maps: B x 2 x H x W
    logits = _maps((5, 2, 64, 64), seed=J)
    p_cpu = torch.softmax(3.0 * logits, dim=1)
    p = p_cpu.to(_dev()).requires_grad_(True)
    loss = wb.wavelet_shape_loss(p, wavelet, J, weights)
    (1.7 * loss).backward()
    ref_loss, ref_grad = wn.shape_loss(p_cpu.numpy(), wavelet, J, weights)
    assert abs(float(loss) - ref_loss) <= TOL * abs(ref_loss)
    assert rel_err(p.grad.cpu().numpy(), 1.7 * ref_grad) < TOL
    # run-to-run reproducible
    loss2 = wb.wavelet_shape_loss(p.detach(), wavelet, J, weights)
    assert float(loss2) == float(loss)


@pytest.mark.parametrize("wavelet", ["haar", "db2"])
@pytest.mark.parametrize("shape,J", [((40, 2, 64, 64), 3),        # 80 maps: every cluster loops over several maps
                                     ((3, 2, 96, 160), 3),        # non-power-of-two sides
                                     ((2, 1, 256, 256), 5),
                                     ((5, 1, 32, 64), 4),
                                     ((2, 2, 48, 64), 4),         # whole map: one CTA, the halo wraps onto itself
                                     ((3, 1, 64, 128), 1),        # J = 1: the streamed plan has no resident stage
                                     ((2, 1, 64, 72), 2),         # W % 32 != 0: level 1 falls back to per-thread loads
                                     ((1, 2, 512, 512), 4)])
def test_fused_plans_match_per_level_path(wavelet, shape, J):
    """The fused loss + gradient plans (whole map resident in a cluster's distributed shared memory; level 1 streamed +
    low-low band resident) against the per-level kernels: loss, gradient, a second backward through the retained graph,
    the loss-only call."""
    import wtpse_b200 as wb
    from wtpse_b200 import wavelet as wv

    lib = wb._lib.load()
    x = _safe_maps(shape, J, J + shape[0], wavelet)
    weights = tuple(0.5 + 0.25 * j for j in range(J))

    def run():
        xg = x.clone().requires_grad_(True)
        loss = wb.wavelet_shape_loss(xg, wavelet, J, weights)
        (0.3 * loss).backward(retain_graph=True)
        g1 = xg.grad.clone()
        xg.grad = None
        (2.0 * loss).backward()                  # second pass through the retained graph
        return float(loss), g1, xg.grad.clone(), float(wb.wavelet_shape_loss(x, wavelet, J, weights))

    try:
        wb._lib.debug_set("wavelet_resident", 0)
        assert wv.resident_cluster_size(shape[-2], shape[-1], wavelet, J) == 0
        l_p, g_p, g2_p, lo_p = run()
        assert lo_p == l_p
        wb._lib.debug_set("wavelet_resident", 1)
        for split, cmax, tiles in ((0, 8, 1), (1, 8, 1), (1, 2, 1), (1, 8, 0), (-1, 8, 1)):
            wb._lib.debug_set("wavelet_split", split)
            wb._lib.debug_set("wavelet_cluster_max", cmax)
            wb._lib.debug_set("wavelet_tiles", tiles)
            assert wv.resident_cluster_size(shape[-2], shape[-1], wavelet, J) > 0, (split, cmax)
            l_r, g_r, g2_r, lo_r = run()
            assert abs(l_r - l_p) <= 2e-6 * abs(l_p) and lo_r == l_r, (split, cmax)
            assert rel_err(g_r.cpu().numpy(), g_p.cpu().numpy()) < 2e-6, (split, cmax)
            assert rel_err(g2_r.cpu().numpy(), g2_p.cpu().numpy()) < 2e-6, (split, cmax)
            assert rel_err(g2_r.cpu().numpy(), (g_r / 0.3 * 2.0).cpu().numpy()) < 2e-6, (split, cmax)
    finally:
        wb._lib.debug_set("wavelet_resident", 1)
        wb._lib.debug_set("wavelet_split", -1)
        wb._lib.debug_set("wavelet_cluster_max", 8)
        wb._lib.debug_set("wavelet_tiles", 1)


def test_fused_plan_covers_maps_too_large_for_a_cluster():
    """1024 x 1024 (BASELINE configs[4]): 4 MB per map does not fit a cluster.  The streamed plan peels levels two at a time
    (1024 -> 256 -> 64, level 5 resident in one CTA per band); with one peeled level the 512 x 512 bands go resident in
    clusters of 8.  Both against the per-level kernels."""
    import wtpse_b200 as wb
    from wtpse_b200 import wavelet as wv

    lib = wb._lib.load()
    x = _safe_maps((2, 2, 1024, 1024), 5, 11)
    res = []
    try:
        for resident, peel, cs in ((1, 8, 1), (1, 1, 8), (0, 8, 0)):
            wb._lib.debug_set("wavelet_resident", resident)
            wb._lib.debug_set("wavelet_peel_max", peel)
            assert wv.resident_cluster_size(1024, 1024, "db2", 5) == cs
            xg = x.clone().requires_grad_(True)
            loss = wb.wavelet_shape_loss(xg, "db2", 5)
            loss.backward()
            res.append((float(loss), xg.grad.clone()))
    finally:
        wb._lib.debug_set("wavelet_resident", 1)
        wb._lib.debug_set("wavelet_peel_max", 8)
    for l, g in res[:2]:
        assert abs(l - res[2][0]) <= 2e-6 * abs(res[2][0])
        assert rel_err(g.cpu().numpy(), res[2][1].cpu().numpy()) < 2e-6


@pytest.mark.parametrize("wavelet", ["haar", "db2"])
def test_streamed_plan_with_every_level_streamed(wavelet):
    """1024 x 1024, J = 2: the 512 x 512 low-low band would need a cluster of 8, level 2 can stream -> nothing goes
    resident (reported cluster size 1).  Shapes with no fused plan at all report 0 and take the per-level path."""
    import wtpse_b200 as wb
    from wtpse_b200 import wavelet as wv

    lib = wb._lib.load()
    assert wv.resident_cluster_size(64, 96, wavelet, 5) == 0        # 96 % 2^(J+1) != 0 and 48, 24 .. are not tileable
    x = _safe_maps((1, 2, 1024, 1024), 2, 5, wavelet)
    res = []
    try:
        wb._lib.debug_set("wavelet_split", 1)            # Haar would otherwise take the single band kernel here
        for resident, peel, cs in ((1, 8, 1), (1, 1, 8), (0, 8, 0)):
            wb._lib.debug_set("wavelet_resident", resident)
            wb._lib.debug_set("wavelet_peel_max", peel)
            # Haar row bands are independent work items: the resident stage never needs a cluster
            assert wv.resident_cluster_size(1024, 1024, wavelet, 2) == (cs if wavelet == "db2" or cs == 0 else 1)
            xg = x.clone().requires_grad_(True)
            loss = wb.wavelet_shape_loss(xg, wavelet, 2, (1.0, 2.0))
            (0.5 * loss).backward()
            res.append((float(loss), xg.grad.clone()))
    finally:
        wb._lib.debug_set("wavelet_resident", 1)
        wb._lib.debug_set("wavelet_peel_max", 8)
        wb._lib.debug_set("wavelet_split", -1)
    for l, g in res[:2]:
        assert abs(l - res[2][0]) <= 2e-6 * abs(res[2][0])
        assert rel_err(g.cpu().numpy(), res[2][1].cpu().numpy()) < 2e-6


def test_fused_plans_random_shapes_and_determinism():
    """Seeded sweep over map sizes / level counts / batch sizes: whatever plan the planner picks must agree with the
    per-level kernels, and five repeated launches must be bit-identical (a race between the TMA ring, the cluster
    barriers and the DSMEM halo copies would show up as run-to-run differences)."""
    import random

    import wtpse_b200 as wb
    from wtpse_b200 import wavelet as wv

    lib = wb._lib.load()
    rng = random.Random(20261018)
    cases = 0
    while cases < 28:
        H = rng.choice([32, 48, 64, 96, 128, 160, 192, 256, 320, 384, 512])
        W = rng.choice([32, 64, 96, 128, 160, 256, 384, 512, 640])
        wavelet = rng.choice(["haar", "db2"])
        J = rng.randint(1, 5)
        if H % (1 << J) or W % (1 << J) or (wavelet == "db2" and min(H, W) >> (J - 1) < 4):
            continue
        nmaps = rng.choice([1, 2, 3, 7, 16, 37, 70])
        if nmaps * H * W > 12 * 1024 * 1024:
            continue
        cases += 1
        x = _safe_maps((nmaps, 1, H, W), J, cases, wavelet)
        weights = tuple(rng.choice([0.25, 1.0, 3.0]) for _ in range(J))

        def run():
            xg = x.clone().requires_grad_(True)
            loss = wb.wavelet_shape_loss(xg, wavelet, J, weights)
            loss.backward()
            return loss.detach().clone(), xg.grad.clone()

        fused_available = wv.resident_cluster_size(H, W, wavelet, J) > 0
        l0, g0 = run()
        for _ in range(4):
            l1, g1 = run()
            assert torch.equal(l0, l1) and torch.equal(g0, g1), (H, W, wavelet, J, nmaps)
        if fused_available:
            wb._lib.debug_set("wavelet_resident", 0)
            try:
                lp, gp = run()
            finally:
                wb._lib.debug_set("wavelet_resident", 1)
            assert abs(float(l0) - float(lp)) <= 2e-6 * abs(float(lp)), (H, W, wavelet, J, nmaps)
            assert rel_err(g0.cpu().numpy(), gp.cpu().numpy()) < 2e-6, (H, W, wavelet, J, nmaps)


def test_resident_path_reproducible_and_unit_upstream():
    import wtpse_b200 as wb

    x = torch.rand(32, 2, 128, 128, device=_dev())
    outs = []
    for _ in range(2):
        xg = x.clone().requires_grad_(True)
        loss = wb.wavelet_shape_loss(xg, "db2", 4)
        loss.backward()                              # upstream gradient exactly 1: the scale kernel exits early
        outs.append((float(loss), xg.grad.clone()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])
    assert abs(float((outs[0][1].double() * x.double()).sum()) - outs[0][0]) <= 1e-4 * outs[0][0]


def test_wavelet_contract_errors():
    import wtpse_b200 as wb

    x = torch.randn(2, 2, 24, 24, device=_dev())
    with pytest.raises(wb._lib.WtpseError):
        wb.dwt2d(x, "haar", 4)                      # 24 not divisible by 16
    with pytest.raises(ValueError):
        wb.dwt2d(x, "sym8", 1)
    with pytest.raises(RuntimeError):
        wb.dwt2d(x.cpu(), "haar", 1)


@pytest.mark.parametrize("shape,J", [((3, 1, 64, 128), 1), ((2, 1, 64, 128), 3), ((5, 1, 32, 256), 2), ((3, 2, 256, 256), 4),
                                     ((1, 2, 512, 512), 4), ((2, 2, 512, 512), 2), ((7, 1, 128, 512), 3), ((1, 1, 1024, 1024), 1),
                                     ((1, 2, 1024, 1024), 2), ((2, 1, 1024, 1024), 5), ((37, 1, 96, 256), 2), ((2, 1, 16, 256), 2),
                                     ((70, 1, 64, 256), 3)])
def test_db2_factored_passes(shape, J):
    """csrc/wavelet_db2.cu (factored db2, one or two levels per pass, level groups on separate warps) forced into the plan
    (wavelet_split = 1) against the per-level kernels: loss, gradient with a non-unit upstream gradient, run-to-run bits;
    one level per pass, two levels per pass and the round-1 level kernels must all agree.  The small cases also against the
    float64 specification."""
    import wtpse_b200 as wb
    from oracle import wavelet_np as wn

    x = _safe_maps(shape, J, seed=J + shape[0])
    weights = tuple(0.5 + 0.25 * j for j in range(J))

    def run():
        xg = x.clone().requires_grad_(True)
        loss = wb.wavelet_shape_loss(xg, "db2", J, weights)
        (0.7 * loss).backward()
        return float(loss.detach()), xg.grad.clone()

    try:
        wb._lib.debug_set("wavelet_resident", 0)
        lp, gp = run()
        wb._lib.debug_set("wavelet_resident", 1)
        wb._lib.debug_set("wavelet_split", 1)
        for db2, two in ((0, 0), (1, 0), (1, 1)):
            wb._lib.debug_set("wavelet_db2", db2)
            wb._lib.debug_set("wavelet_db2_two", two)
            l, g = run()
            l2, g2 = run()
            assert l2 == l and torch.equal(g, g2), (db2, two)
            assert abs(l - lp) <= 2e-6 * abs(lp), (db2, two)
            assert rel_err(g.cpu().numpy(), gp.cpu().numpy()) < 2e-6, (db2, two)
        if x.numel() <= 1 << 18:
            ref_loss, ref_grad = wn.shape_loss(x.cpu().numpy(), "db2", J, weights)
            assert abs(l - ref_loss) <= TOL * abs(ref_loss)
            assert rel_err(g.cpu().numpy(), 0.7 * ref_grad) < TOL
    finally:
        wb._lib.debug_set("wavelet_resident", 1)
        wb._lib.debug_set("wavelet_split", -1)
        wb._lib.debug_set("wavelet_db2", 1)
        wb._lib.debug_set("wavelet_db2_two", 1)


def test_db2_factored_zero_and_flat_maps():
    """sign(0) = 0: an all-zero map has zero loss and an exactly zero gradient; a constant map has (numerically) zero
    detail coefficients, so its loss is at rounding level."""
    import wtpse_b200 as wb

    try:
        wb._lib.debug_set("wavelet_split", 1)
        x = torch.zeros(2, 1, 64, 256, device=_dev(), requires_grad=True)
        loss = wb.wavelet_shape_loss(x, "db2", 2)
        loss.backward()
        assert float(loss) == 0.0 and float(x.grad.abs().max()) == 0.0
        y = torch.full((2, 1, 64, 256), 0.75, device=_dev())
        assert float(wb.wavelet_shape_loss(y, "db2", 2)) < 1e-6
    finally:
        wb._lib.debug_set("wavelet_split", -1)


def test_wavelet_full_size_properties():
    """BASELINE configs[1] as literally written: 32 x 2 x 512 x 512, db2, J = 4."""
    import wtpse_b200 as wb

    x = torch.rand(32, 2, 512, 512, device=_dev())
    c = wb.dwt2d(x, "db2", 4)
    assert rel_err(wb.idwt2d(c, "db2", 4).cpu().numpy(), x.cpu().numpy()) < TOL
    e0, e1 = float((x.double() ** 2).sum()), float((c.double() ** 2).sum())
    assert abs(e0 - e1) <= TOL * e0
    xg = x.clone().requires_grad_(True)
    loss = wb.wavelet_shape_loss(xg, "db2", 4)
    loss.backward()
    # the loss is 1-homogeneous in x: <grad, x> == loss (Euler), a size-independent gradient check
    assert abs(float((xg.grad.double() * x.double()).sum()) - float(loss)) <= 1e-4 * float(loss)


@pytest.mark.parametrize("wavelet,shape,J", [("db2", (3, 1, 64, 128), 1), ("db2", (5, 1, 32, 256), 2), ("db2", (3, 1, 512, 512), 4),
                                             ("db2", (1, 1, 1024, 1024), 2), ("db2", (7, 1, 96, 256), 3), ("haar", (3, 1, 512, 512), 3)])
def test_fused_plans_stay_inside_their_buffers(wavelet, shape, J):
    """Guard bands (compute-sanitizer is not available on the GPU pool): input, gradient and workspace sit inside larger
    buffers filled with a byte pattern; after the fused call through the C ABI every guard byte is untouched, the input is
    unchanged and the result equals the ordinary call's -- for the streamed db2 passes (wavelet_split = 1) and the default plan."""
    import wtpse_b200 as wb
    from wtpse_b200.functional import _ptr, _stream_ptr
    from wtpse_b200.wavelet import _weights

    dev = _dev()
    lib = wb._lib.load()
    nmaps, H, W = shape[0] * shape[1], shape[2], shape[3]
    wid = {"haar": 0, "db2": 1}[wavelet]
    x = _safe_maps(shape, J, 3, wavelet)
    n = x.numel()
    G = 4096                                                     # guard floats / bytes on each side
    try:
        for split in (1, -1):
            wb._lib.debug_set("wavelet_split", split)
            if lib.wtpse_wavelet_resident_cluster(H, W, wid, J) == 0:
                continue
            xg = x.clone().requires_grad_(True)
            ref_loss = wb.wavelet_shape_loss(xg, wavelet, J)
            ref_loss.backward()
            nbytes = lib.wtpse_wavelet_workspace_bytes(nmaps, H, W, J)
            xin = torch.full((n + 2 * G,), 7.0, device=dev)
            xin[G:G + n] = x.reshape(-1)
            gbuf = torch.full((n + 2 * G,), -3.0, device=dev)
            wbuf = torch.full((nbytes + 2 * G,), 0x5A, dtype=torch.uint8, device=dev)
            loss = torch.zeros(2, device=dev)
            wb._lib.check(lib.wtpse_wavelet_loss_resident(_ptr(xin[G:]), nmaps, H, W, wid, J, _weights(None, J), None, _ptr(loss),
                                                          _ptr(gbuf[G:]), _ptr(wbuf[G:]), nbytes, _stream_ptr(dev)))
            torch.cuda.synchronize()
            assert bool((xin[:G] == 7.0).all()) and bool((xin[G + n:] == 7.0).all()) and torch.equal(xin[G:G + n], x.reshape(-1))
            assert bool((gbuf[:G] == -3.0).all()) and bool((gbuf[G + n:] == -3.0).all())
            assert bool((wbuf[:G] == 0x5A).all()) and bool((wbuf[G + nbytes:] == 0x5A).all())
            assert float(loss[1]) == 0.0 and float(loss[0]) == float(ref_loss)
            assert torch.equal(gbuf[G:G + n].view_as(x), xg.grad)
    finally:
        wb._lib.debug_set("wavelet_split", -1)


@pytest.mark.parametrize("shape,J", [((3, 1, 64, 128), 1), ((5, 1, 32, 256), 2), ((3, 2, 256, 256), 4), ((2, 1, 512, 512), 3),
                                     ((7, 1, 128, 512), 2), ((1, 2, 1024, 1024), 5), ((37, 1, 96, 256), 3), ((2, 1, 16, 256), 2)])
def test_haar_through_the_pass_kernels(shape, J):
    """Haar levels through the one- / two-level pass kernels of csrc/wavelet_db2.cu (no overlap: no halo rows, no neighbour
    exchange, a 2 x 2 Hadamard butterfly per site in the synthesis) against the per-level kernels and, for the small cases,
    the float64 specification; the band kernel (wavelet_haar_passes = 0) must agree too."""
    import wtpse_b200 as wb
    from oracle import wavelet_np as wn

    x = _safe_maps(shape, J, seed=2 * J + shape[0], wavelet="haar")
    weights = tuple(1.0 - 0.15 * j for j in range(J))

    def run():
        xg = x.clone().requires_grad_(True)
        loss = wb.wavelet_shape_loss(xg, "haar", J, weights)
        (1.3 * loss).backward()
        return float(loss.detach()), xg.grad.clone()

    try:
        wb._lib.debug_set("wavelet_resident", 0)
        lp, gp = run()
        wb._lib.debug_set("wavelet_resident", 1)
        wb._lib.debug_set("wavelet_haar_min_log2px", 0)          # every size that has a pass plan takes it
        for passes, split in ((1, -1), (1, 1), (0, -1)):
            wb._lib.debug_set("wavelet_haar_passes", passes)
            wb._lib.debug_set("wavelet_split", split)
            l, g = run()
            l2, g2 = run()
            assert l2 == l and torch.equal(g, g2), (passes, split)
            assert abs(l - lp) <= 2e-6 * abs(lp), (passes, split)
            assert rel_err(g.cpu().numpy(), gp.cpu().numpy()) < 2e-6, (passes, split)
        if x.numel() <= 1 << 18:
            ref_loss, ref_grad = wn.shape_loss(x.cpu().numpy(), "haar", J, weights)
            assert abs(l - ref_loss) <= TOL * abs(ref_loss)
            assert rel_err(g.cpu().numpy(), 1.3 * ref_grad) < TOL
    finally:
        wb._lib.debug_set("wavelet_resident", 1)
        wb._lib.debug_set("wavelet_split", -1)
        wb._lib.debug_set("wavelet_haar_passes", 1)
        wb._lib.debug_set("wavelet_haar_min_log2px", 16)
