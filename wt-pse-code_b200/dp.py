"""Data-parallel plumbing for the train step: one process per GPU, NCCL all-reduce of
the fp32 gradients once per backward pass, nothing inside the loss kernels.

The reference has no distributed code at all; this is the standard DDP objective (mean over ranks of the
per-rank reference objective).  Each rank draws whole ``[K domains x n]`` batches, because both BatchNorm
and the MMD couple the samples of one batch (BN statistics stay local, as in the reference -- SyncBN
would change semantics).

``FlatGradBucket`` keeps every parameter's ``.grad`` as a view into one contiguous buffer, so a backward
pass is followed by exactly ONE collective per model (25.5 MB for WT_PSE, 12.8 MB for the shape network)
instead of one per tensor; on NVSwitch that is latency-, not bandwidth-bound.
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self._rebind()

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

    def allreduce_mean(self):
        """Average the bucket over the ranks (no-op for a single process)."""
        if not dist.is_available() or not dist.is_initialized():
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:                                                        # gloo (CPU tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(world)


def rank_batch_seed(base_seed, rank, iteration):
    """Rank-distinct, iteration-distinct seed for the synthetic loader: no two ranks see the same batch."""
    return (base_seed * 1000003 + iteration) * 4099 + rank


def per_rank_batch(global_batch, world_size, n_domains):
    """Reference batch arithmetic (Trainer.py:1013, train.py:89) per rank: the nominal per-rank batch is
    global // world, the batch actually used is n_domains * (per_rank // n_domains) whole domain groups."""
    per_rank = global_batch // world_size
    n_per_domain = per_rank // n_domains
    return n_per_domain, n_per_domain * n_domains
