"""Data-parallel plumbing for the train step: one process per GPU, NCCL all-reduce of
the fp32 gradients once per backward pass, nothing inside the loss kernels.

The reference has no distributed code at all; this is the standard DDP objective (mean over ranks of the
per-rank reference objective).  Each rank draws whole ``[K domains x n]`` batches, because both BatchNorm
and the MMD couple the samples of one batch (BN statistics stay local, as in the reference -- SyncBN
would change semantics).

``FlatGradBucket`` keeps every parameter's ``.grad`` as a view into one contiguous buffer (25.5 MB for WT_PSE, 12.8 MB
for the shape network), cut into a few contiguous SEGMENTS.  A segment's all-reduce is started from autograd's
post-accumulate-grad hooks the moment its last gradient has been written -- decoder and head segments first, while the
encoder's backward is still running -- as an asynchronous NCCL collective (its own stream); ``finish()`` starts whatever
has not fired and makes the current stream wait for all of them right before the optimizer step.  On NVSwitch the
collectives are latency-, not bandwidth-bound: what the overlap hides is their launch and wire time, not the time the
ranks differ by when they arrive.  Everything is stream-ordered, so a whole iteration -- collectives included -- can
be captured into one CUDA graph (train_step.TrainStep.capture).
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """segments: how many contiguous pieces the flat buffer is reduced in (1 = one collective after the backward pass).
    Segment boundaries follow parameter boundaries; segment 0 holds the LAST-registered parameters (the ones whose
    gradients autograd finishes first) and is the largest, the last segment (first layers, ready only when the backward
    ends, so its collective is the exposed one) the smallest."""

    # shares of the buffer, in firing order
    SPLITS = {1: (1.0,), 2: (0.7, 0.3), 3: (0.5, 0.35, 0.15), 4: (0.4, 0.3, 0.2, 0.1)}

    def __init__(self, module, process_group=None, segments=1):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self._rebind()
        # ---- segments (contiguous parameter ranges, walking the registration order backwards) ----
        shares = self.SPLITS[max(1, min(int(segments), 4))]
        ends = []                                        # cumulative element targets, from the end of the buffer
        acc = 0.0
        for sh in shares[:-1]:
            acc += sh
            ends.append(acc * total)
        self.segments = []                               # (lo, hi) element ranges, in firing order
        self._seg_of = {}                                # id(param) -> segment index
        hi, seg, taken = total, 0, 0
        off_after = total
        cur = []
        for p in reversed(self.params):
            cur.append(p)
            taken += p.numel()
            off_after -= p.numel()
            if seg < len(ends) and taken >= ends[seg]:
                self.segments.append((off_after, hi))
                for q in cur:
                    self._seg_of[id(q)] = seg
                hi, cur, seg = off_after, [], seg + 1
        if cur:
            self.segments.append((0, hi))
            for q in cur:
                self._seg_of[id(q)] = len(self.segments) - 1
        self._seg_params = [sum(1 for p in self.params if self._seg_of[id(p)] == k) for k in range(len(self.segments))]
        self._pending = None                             # per-segment countdown while armed
        self._fired = None
        self._works = []
        self._hooks = []
        if len(self.segments) > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ---- gradient storage ------------------------------------------------------------------------------------------
    def zero(self):
        """Replaces optimizer.zero_grad()/module.zero_grad() (Trainer.py:767-768): one memset, views stay bound."""
        self.flat.zero_()
        for p in self.params:                                        # re-bind if someone set grads to None
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * self.flat.element_size():
                self._rebind()
                break

    def _rebind(self):
        off = 0
        for p in self.params:
            # a view with the parameter's own strides (channels-last weights keep their layout, as the
            # gradient-layout contract and the fused optimizers require); autograd accumulates in place into it
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), storage_offset=off)
            off += p.numel()

    # ---- collectives ---------------------------------------------------------------------------------------------
    def _world(self):
        if not dist.is_available() or not dist.is_initialized():
            return 1
        return dist.get_world_size(self.group)

    def _reduce(self, lo, hi, async_op):
        view = self.flat[lo:hi]
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op), None
        # gloo (CPU tests) has no AVG: sum, then divide once the collective has finished
        return dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op), view

    def arm(self):
        """Call right before the backward pass whose gradients this bucket owns: from now on every segment starts its
        all-reduce as soon as its last gradient has been accumulated.  Backward passes that merely deposit gradients here
        (the shape update's dead teacher gradients, shape_networks.py:524) run un-armed and trigger nothing."""
        if len(self.segments) > 1 and self._world() > 1:
            self._pending = list(self._seg_params)
            self._fired = [False] * len(self.segments)
            self._works = []

    def _on_grad(self, p):
        if self._pending is None:
            return
        k = self._seg_of[id(p)]
        self._pending[k] -= 1
        if self._pending[k] == 0 and not self._fired[k]:
            self._fire(k)

    def _fire(self, k):
        self._fired[k] = True
        lo, hi = self.segments[k]
        self._works.append(self._reduce(lo, hi, True))

    def finish(self):
        """After the backward pass: start the segments that have not fired (a parameter without a gradient in this pass
        holds its segment back), then make the current stream wait for every collective.  Un-armed (or one segment):
        one collective over the whole buffer."""
        world = self._world()
        if world == 1:
            self._pending = None
            return
        if self._pending is None:
            work, view = self._reduce(0, self.flat.numel(), False)
            if view is not None:
                view.div_(world)
            return
        for k in range(len(self.segments)):
            if not self._fired[k]:
                self._fire(k)
        self._pending = None
        for work, view in self._works:
            work.wait()
            if view is not None:
                view.div_(world)
        self._works = []

    def allreduce_mean(self):
        """Average the bucket over the ranks (no-op for a single process); with segments, finish() of an armed pass."""
        self.finish()


def rank_batch_seed(base_seed, rank, iteration):
    """Rank-distinct, iteration-distinct seed for the synthetic loader: no two ranks see the same batch."""
    return (base_seed * 1000003 + iteration) * 4099 + rank


def per_rank_batch(global_batch, world_size, n_domains):
    """Reference batch arithmetic (Trainer.py:1013, train.py:89) per rank: the nominal per-rank batch is
    global // world, the batch actually used is n_domains * (per_rank // n_domains) whole domain groups."""
    per_rank = global_batch // world_size
    n_per_domain = per_rank // n_domains
    return n_per_domain, n_per_domain * n_domains
