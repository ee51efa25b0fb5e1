"""One training iteration of WT-PSE: the body of ``Trainer.train_epoch`` (Trainer.py:762-925) as a harness
that takes a device batch and returns device scalars (no per-loss ``.item()`` host syncs).

Four sub-steps, exactly in the reference's order and with its loss definitions:
  1. OD segmentation net   BCE(sigmoid(out), od) + w_i * ins + w_d * dom          Trainer.py:779-805
  2. OD shape net          kd + w_i * ins_total + w_d * dom                        Trainer.py:811-825
  -- coarse-to-fine: od_pred = sigmoid(out) > 0.75 ; image += 1 ; roi = image * od_pred - 1   :842-853 --
  3. OC segmentation net   BCE-with-logits(out * od_pred, oc, pos_weight) + ...    Trainer.py:856-892
  4. OC shape net                                                                  Trainer.py:896-914
Adam(lr 5e-4, betas (0.9, 0.99)) per network (train.py:120-138).  With a process group, each backward is
followed by one bucketed all-reduce of that network's gradients (dp.FlatGradBucket); the gradients the
shape update deposits into the segmentation network (teacher not detached) are never reduced.
"""
import torch
import torch.nn.functional as F

from . import segmentation as seg
from .dp import FlatGradBucket
from .elementwise import od_roi

DEFAULT_HPARAMS = {   # hparams_registry.py:75-93
    "whitening": True, "margin": 0, "shape_prior": True, "shape_attention": True, "cat_shape": False,
    "shape_attention_coeffient": 0.3, "shape_start": 0.5, "instance_wt_gm": 1, "domain_wt_gm": 1, "multi-turn": 1,
}


class TrainStep:
    def __init__(self, n_per_domain, n_domains=3, device="cuda", hparams=None, lr=5e-4, seed=0, process_group=None,
                 channels_last=True, fused_adam=True, teacher_backward=False, fuse_relu=True,
                 fold_conv_bias=True, cuda_upsample=True, fast_bias=True, cuda_batchnorm=True,
                 cuda_pool=True, grad_segments=3, grad_allreduce=True):
        self.hp = dict(DEFAULT_HPARAMS if hparams is None else hparams)
        self.device = torch.device(device)
        torch.manual_seed(seed)                                   # identical initial weights on every rank
        mk = lambda two_step: seg.WT_PSE(3, 1, self.hp, self.device, two_step, per_domain_batch=n_per_domain,
                                         source_domain_num=n_domains)
        sh = lambda: seg.ShapeVariationalDist_x(self.hp, self.device, 1, number_source_domain=n_domains,
                                                batch_size=n_per_domain)
        self.model, self.model_shape = mk(False).to(self.device), sh().to(self.device)
        self.model_oc, self.model_shape_oc = mk(True).to(self.device), sh().to(self.device)
        self.nets = (self.model, self.model_shape, self.model_oc, self.model_shape_oc)
        # The reference back-propagates the KD loss into the teacher and discards the result at the next zero_grad
        # (SURVEY appendix A.3 item 6).  teacher_backward=False skips that dead backward pass; weights and losses of
        # all four networks are unchanged (tests/test_gpu_update.py).  True reproduces the reference's work exactly.
        self.model_shape.teacher_grad = self.model_shape_oc.teacher_grad = bool(teacher_backward)
        for m in self.nets:
            m.train()
            # conv -> BatchNorm stages skip ATen's separate bias-add pass (segmentation._conv_bn; same outputs, same
            # gradients, same running statistics up to rounding)
            seg.set_conv_bias_folding(m, fold_conv_bias)
            seg.set_cuda_upsample(m, cuda_upsample and channels_last)     # ATen's NHWC bilinear kernel: 26 ms of 242
            seg.set_fast_bias(m, fast_bias and channels_last)             # bias (+ ReLU) of the convs without BatchNorm
            seg.set_cuda_batchnorm(m, cuda_batchnorm and channels_last)   # BatchNorm + ReLU, forward and backward
            seg.set_cuda_pool(m, cuda_pool and channels_last)             # 2x2 max pooling with a one-byte argmax
            if channels_last:
                # cuDNN's NCHW BatchNorm runs one CTA per channel (16-256 CTAs on 148 SMs: 55 % of the step at
                # 512x512); with channels-last weights every conv/BN of the backbone takes the NHWC kernels.
                # The loss kernels keep the reference's NCHW layout (`z.contiguous()`, algorithms.py:1280).
                m.to(memory_format=torch.channels_last)
            if fuse_relu:
                # Gram + ReLU in one pass over each embedding (the *_cl kernels read the channels-last
                # embeddings in place, so there is no layout conversion around the loss either way)
                seg.enable_relu_fusion(m, True)
        # each network's gradients are reduced in `grad_segments` pieces, started from autograd hooks while the rest of the
        # backward pass is still running (dp.FlatGradBucket); 1 = one collective after the backward pass
        self.buckets = [FlatGradBucket(m, process_group, segments=grad_segments) for m in self.nets]
        fused = bool(fused_adam) and self.device.type == "cuda"       # one multi-tensor kernel per optimizer step
        self.optims = [torch.optim.Adam(m.parameters(), lr=lr, betas=(0.9, 0.99), fused=fused, capturable=fused)
                       for m in self.nets]
        # identical weights, DIFFERENT noise: the posterior's re-parameterisation noise (torch.randn in the shape networks) must
        # not be the same on every data-parallel rank, so the generator is re-seeded by rank once the weights exist
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank = torch.distributed.get_rank(process_group)
            torch.manual_seed(seed + 7919 * (rank + 1))
        self._graph = None
        self.iteration = 0
        self.grad_allreduce = bool(grad_allreduce)     # False: independent replicas (bench.py's scaling diagnostic)

    def _finish(self, idx, loss):
        if self.grad_allreduce:
            self.buckets[idx].arm()
        loss.backward()
        if self.grad_allreduce:
            self.buckets[idx].finish()
        self.optims[idx].step()

    def step(self, image, target_od, target_oc):
        """image B x 3 x H x W in [-1, 1] (MUTATED in place: += 1, as Trainer.py:850 does); targets B x 1 x H x W.
        Returns a dict of 0-dim device tensors."""
        hp = self.hp
        wi, wd = hp["instance_wt_gm"], hp["domain_wt_gm"]
        out = {}
        # ---- 1. OD segmentation ---------------------------------------------------------------------
        self.buckets[0].zero()
        output, _, _, ins, dom = self.model.update(image, target_od, step=self.iteration, plot_show=0,
                                                   two_stage_inputs=image, sp_mask=target_od, two_step=True)
        loss_seg = F.binary_cross_entropy(torch.sigmoid(output), target_od)
        self._finish(0, loss_seg + wi * ins + wd * dom)
        out.update(loss_seg=loss_seg.detach(), ins_wt=ins.detach(), dom_wt=dom.detach())
        # ---- 2. OD shape network --------------------------------------------------------------------
        for _ in range(hp["multi-turn"]):
            self.buckets[1].zero()
            kd, ins_s, ins_ij, ins_ii, dom_s = self.model_shape.update(self.model, image, target_od, step=self.iteration,
                                                                      plot_show=0, two_stage_inputs=image, two_step=True)
            self._finish(1, kd + wi * ins_s + wd * dom_s)
        out.update(kd=kd.detach(), ins_wt_shape=ins_s.detach(), ins_ij=ins_ij.detach(), ins_ii=ins_ii.detach(),
                   dom_wt_shape=dom_s.detach())
        # ---- coarse-to-fine ROI (one kernel; also yields the BCE pos_weight without a host sync) ------
        od_pred, image_roi, sums = od_roi(output, image, target_oc, 0.75)
        pos_weight = sums[2]
        # ---- 3. OC segmentation ---------------------------------------------------------------------
        self.buckets[2].zero()
        output_oc, _, _, ins_oc, dom_oc = self.model_oc.update(image_roi, target_oc, step=self.iteration, plot_show=0,
                                                               two_stage_inputs=image_roi, two_step=True)
        loss_seg_oc = F.binary_cross_entropy_with_logits(output_oc * od_pred, target_oc, pos_weight=pos_weight)
        self._finish(2, loss_seg_oc + wi * ins_oc + wd * dom_oc)
        out.update(loss_seg_oc=loss_seg_oc.detach(), ins_wt_oc=ins_oc.detach(), dom_wt_oc=dom_oc.detach())
        # ---- 4. OC shape network --------------------------------------------------------------------
        for _ in range(hp["multi-turn"]):
            self.buckets[3].zero()
            kd_oc, ins_s_oc, _, _, dom_s_oc = self.model_shape_oc.update(self.model_oc, image_roi, target_oc,
                                                                        step=self.iteration, plot_show=0,
                                                                        two_stage_inputs=image_roi, two_step=True)
            self._finish(3, kd_oc + wi * ins_s_oc + wd * dom_s_oc)
        out.update(kd_oc=kd_oc.detach(), ins_wt_shape_oc=ins_s_oc.detach(), dom_wt_shape_oc=dom_s_oc.detach())
        self.iteration += 1
        return out

    # ---- CUDA-graph replay of the whole iteration --------------------------------------
    def capture(self, image, target_od, target_oc, warmup=3):
        """Record one full iteration (all four sub-steps, ~5 000 kernel launches, the NCCL all-reduces included)
        into a CUDA graph on static input buffers.  Possible because nothing on the path synchronises the host:
        the losses stay on the device, the ROI pos-weight is computed by a kernel, the NaN scrubs are `where`s.
        Call once with a representative batch; afterwards `replay()` copies a new batch in and launches the graph."""
        self._static = [image.clone(), target_od.clone(), target_oc.clone()]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                                   # allocator / cuDNN / optimizer state warm-up
                self._static[0].copy_(image)
                self.step(*self._static)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._static[0].copy_(image)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_out = self.step(*self._static)
        return self

    def replay(self, image, target_od, target_oc):
        if self._graph is None:
            raise RuntimeError("call capture() first")
        self._static[0].copy_(image)
        self._static[1].copy_(target_od)
        self._static[2].copy_(target_oc)
        self._graph.replay()
        return self._static_out
