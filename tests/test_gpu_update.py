"""GPU parity of the reference-facing entry points: WT_PSE.update / ShapeVariationalDist_x.update against the
loss tuples the unmodified reference returned for the same seed-0 weights and inputs, and the train-step harness."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

HP = {"whitening": True, "margin": 0, "shape_prior": True, "shape_attention": True, "cat_shape": False,
      "shape_attention_coeffient": 0.3, "shape_start": 0.5, "instance_wt_gm": 1, "domain_wt_gm": 1, "multi-turn": 1}


@pytest.fixture(autouse=True)
def _fp32_backbone():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _close(a, b, tol):
    return abs(a - b) <= tol * max(abs(b), 1e-12)


def test_update_entry_points_match_reference_golden():
    from wtpse_b200 import segmentation as seg

    g = golden("update_b6_16x16.npz")
    n, K = int(g["n"]), int(g["K"])
    dev = torch.device("cuda:0")
    torch.manual_seed(0)                        # same construction order as the reference -> same weights (CPU RNG)
    main = seg.WT_PSE(3, 1, dict(HP), dev, False, per_domain_batch=n, source_domain_num=K)
    shape = seg.ShapeVariationalDist_x(dict(HP), dev, 1, number_source_domain=K, batch_size=n)
    main.to(dev).train(); shape.to(dev).train()
    image = torch.from_numpy(g["image"]).to(dev)
    mask = torch.from_numpy(g["mask"]).to(dev)

    logits, att1, att2, ins, dom = main.update(image, mask, step=0, plot_show=0, two_stage_inputs=image, sp_mask=mask,
                                               two_step=True)
    assert tuple(logits.shape) == tuple(g["logits_shape"]) and att1 is att2 and att1.shape == mask.shape
    assert ins.dim() == 0 and dom.dim() == 0 and ins.requires_grad and dom.requires_grad
    # loss scalars are RNG-free (they see only the whitening features); conv arithmetic differs CPU vs cuDNN
    assert _close(float(ins), float(g["wt_ins"]), 1e-4), (float(ins), float(g["wt_ins"]))
    assert abs(float(dom) - float(g["wt_dom"])) <= 1e-5

    kd, tot, ij, ii, dom_s = shape.update(main, image, mask, step=0, plot_show=0, two_stage_inputs=image, two_step=True)
    assert _close(float(kd), float(g["sh_kd"]), 1e-4)
    assert _close(float(ij), float(g["sh_ij"]), 1e-4) and _close(float(ii), float(g["sh_ii"]), 1e-4)
    assert _close(float(tot), float(g["sh_total"]), 1e-4) and abs(float(dom_s) - float(g["sh_dom"])) <= 1e-5

    # the whole thing is differentiable down to the conv weights, and only through real numbers
    (torch.nn.functional.binary_cross_entropy(torch.sigmoid(logits), mask) + ins + dom).backward()
    (kd + tot + dom_s).backward()
    for m in (main, shape):
        grads = [p.grad for p in m.parameters() if p.grad is not None]
        assert grads and all(torch.isfinite(x).all() for x in grads)
    assert main.wt_model.DoubleConv.double_conv[0].weight.grad.abs().max() > 0
    assert shape.wt_model.DoubleConv2.double_conv[2].weight.grad.abs().max() > 0


def test_update_gradients_match_autograd_through_the_reference_operator_sequence():
    """Same model, same batch: d(loss)/d(conv weights) with the CUDA loss path vs the op-sequence restatement."""
    from oracle import whitening_torch as wt
    from wtpse_b200 import segmentation as seg

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    net = seg.DeepWT(3, 16).to(dev)
    x = torch.randn(6, 3, 48, 40, device=dev)
    grads = []
    for use_cuda_path in (True, False):
        net.zero_grad(set_to_none=True)
        f = net(x)
        if use_cuda_path:
            import wtpse_b200 as wb
            losses = [wb.whitening_folded(t, 2, 3) for t in f[:2]]
        else:
            losses = [wt.wt_pse_whitening_loss(t, 2, 3) for t in f[:2]]
        sum(a + b for a, b in losses).backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone())
    err = (grads[0] - grads[1]).abs().max() / grads[1].abs().max()
    assert err < 1e-4, float(err)


def test_reparameterisation_equals_torch_normal():
    """The capturable spelling of `torch.normal(mu, std) * std + mu` draws the same numbers and has the same gradients."""
    from wtpse_b200 import segmentation as seg

    dev = torch.device("cuda:0")
    net = seg.ShapeVariationalDist_x(dict(HP), dev, 1, number_source_domain=3, batch_size=2)
    mu = torch.randn(6, 1, 16, 16, device=dev, requires_grad=True)
    logvar = torch.randn(6, 1, 16, 16, device=dev, requires_grad=True)
    torch.manual_seed(11)
    ours = net.reparameterization(mu, logvar)
    ours.sum().backward()
    g_ours = (mu.grad.clone(), logvar.grad.clone())
    mu.grad = logvar.grad = None
    torch.manual_seed(11)
    std = torch.exp(logvar / 2)
    ref = torch.normal(mu, std) * std + mu               # shape_networks.py:507-509 verbatim
    ref.sum().backward()
    assert torch.equal(ours, ref) and torch.equal(g_ours[0], mu.grad) and torch.allclose(g_ours[1], logvar.grad, rtol=1e-6, atol=0)


def test_skipping_the_dead_teacher_backward_changes_nothing_the_trainer_keeps():
    """ShapeVariationalDist_x.teacher_grad=False (what TrainStep uses): same five losses, same student gradients, same
    teacher BatchNorm statistics; only the teacher gradients the trainer zeroes anyway (Trainer.py:768) are not produced."""
    import copy

    from wtpse_b200 import segmentation as seg

    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    main0 = seg.WT_PSE(3, 1, HP, dev, True, per_domain_batch=2, source_domain_num=3).to(dev).train()
    shape0 = seg.ShapeVariationalDist_x(HP, dev, 1, number_source_domain=3, batch_size=2).to(dev).train()
    assert shape0.teacher_grad is True                                  # the class default is the reference's behaviour
    x = torch.randn(6, 3, 64, 48, device=dev)
    mask = (torch.rand(6, 1, 64, 48, device=dev) > 0.5).float()
    res = []
    for teacher_grad in (True, False):
        main, shape = copy.deepcopy(main0), copy.deepcopy(shape0)
        shape.teacher_grad = teacher_grad
        torch.manual_seed(17)
        out = shape.update(main, x, mask, step=0, plot_show=0, two_stage_inputs=x, two_step=True)
        (out[0] + out[1] + out[4]).backward()
        # (the student's logvar head feeds only the unused sample: it has no gradient in either mode)
        res.append(([float(v) for v in out], torch.cat([p.grad.reshape(-1) for p in shape.parameters() if p.grad is not None]),
                    main.prior_dist.down1.bn1.running_mean.clone(),
                    [p.grad for p in main.parameters()]))
    assert res[0][0] == res[1][0]
    err = float((res[0][1] - res[1][1]).abs().max() / res[0][1].abs().max())     # cuDNN's weight-gradient reductions are
    assert err <= 2e-5, err                                                      # not bit-reproducible run to run
    assert torch.equal(res[0][2], res[1][2]) and not torch.equal(res[0][2], main0.prior_dist.down1.bn1.running_mean)
    assert any(g is not None for g in res[0][3]) and all(g is None for g in res[1][3])
    import wtpse_b200 as wb
    ts = wb.TrainStep(n_per_domain=2, n_domains=3, device=dev, seed=0)
    assert ts.model_shape.teacher_grad is False and ts.model_shape_oc.teacher_grad is False
    assert wb.TrainStep(n_per_domain=2, device=dev, teacher_backward=True).model_shape.teacher_grad is True


def test_train_step_runs_and_learns():
    import wtpse_b200 as wb

    dev = torch.device("cuda:0")
    ts = wb.TrainStep(n_per_domain=2, n_domains=3, device=dev, seed=0)
    before = [p.detach().clone() for p in ts.model_shape_oc.parameters()][:4]
    first = None
    for it in range(3):
        image, od, oc = wb.synthetic.fundus_batch(2, 3, 64, 64, dev, seed=it)
        assert image.min() >= -1.0 and image.max() <= 1.0 and set(od.unique().tolist()) <= {0.0, 1.0}
        assert float((oc * (1 - od)).sum()) == 0.0               # the cup lies inside the disc
        img_in = image.clone()
        out = ts.step(image, od, oc)
        assert torch.equal(image, img_in + 1)                     # Trainer.py:850 mutates the batch in place
        vals = {k: float(v) for k, v in out.items()}
        assert all(np.isfinite(v) for v in vals.values()), vals
        first = first or vals
    after = [p.detach() for p in ts.model_shape_oc.parameters()][:4]
    assert any(not torch.equal(a, b) for a, b in zip(before, after))
    assert set(first) >= {"loss_seg", "ins_wt", "dom_wt", "kd", "ins_ii", "ins_ij", "loss_seg_oc", "kd_oc"}


def test_train_step_cuda_graph_replay_matches_eager():
    """One iteration replayed from a CUDA graph gives the losses of the eager iteration on the same weights/batch
    (the RNG only feeds the re-parameterisation, which no loss scalar depends on)."""
    import wtpse_b200 as wb

    dev = torch.device("cuda:0")
    batches = [wb.synthetic.fundus_batch(2, 3, 64, 64, dev, seed=50 + i) for i in range(3)]
    eager = wb.TrainStep(n_per_domain=2, n_domains=3, device=dev, seed=1)
    graph = wb.TrainStep(n_per_domain=2, n_domains=3, device=dev, seed=1)
    graph.capture(*[t.clone() for t in batches[0]], warmup=0)
    # capture itself ran no real step (warmup=0 and graph capture only records): both start from the same weights
    for it, (img, od, oc) in enumerate(batches):
        a = eager.step(img.clone(), od, oc)
        b = graph.replay(img.clone(), od, oc)
        vals = {k: float(v) for k, v in b.items()}
        assert all(np.isfinite(v) for v in vals.values()), vals
        if it == 0:
            # same weights, same batch: every loss scalar that does not depend on the re-parameterisation noise agrees.
            # (Later iterations diverge legitimately: the logits see z_posterior, i.e. the RNG, and the two runs
            # draw different noise, so their weights differ after the first optimizer step.)
            # Sub-step 1's whitening losses see only the initial weights; everything after the first optimizer step has
            # already absorbed the (different) noise through the segmentation loss, so it only has to be close.
            for k in ("ins_wt", "dom_wt"):
                assert abs(float(a[k]) - float(b[k])) <= 2e-4 * max(abs(float(a[k])), 1e-3), (k, float(a[k]), float(b[k]))
            for k in ("kd", "ins_ii", "ins_ij", "ins_wt_oc"):
                assert abs(float(a[k]) - float(b[k])) <= 0.05 * max(abs(float(a[k])), 1e-3), (k, float(a[k]), float(b[k]))
    first_kd = float(graph.replay(*batches[0])["kd"])          # replay() returns the graph's static output tensors
    assert float(graph.replay(*batches[0])["kd"]) != first_kd  # same batch again: the weights moved, the graph keeps training
