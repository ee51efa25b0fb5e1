"""Data-parallel host logic on CPU: world_size-2 gloo run of the bucketed gradient all-reduce
(allreduced grads == mean of the per-rank grads), and the reference's batch arithmetic per rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(0)          # small first layer, most parameters at the end: cuts into 3 segments
    return torch.nn.Sequential(torch.nn.Conv2d(3, 2, 1), torch.nn.BatchNorm2d(2), torch.nn.ReLU(),
                               torch.nn.Conv2d(2, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 3, padding=1),
                               torch.nn.ReLU(), torch.nn.Conv2d(4, 1, 1))


def _local_grads(rank):
    from wtpse_b200.dp import rank_batch_seed

    m = _model()
    g = torch.Generator().manual_seed(rank_batch_seed(7, rank, 0))
    x = torch.randn(6, 3, 8, 8, generator=g)
    m(x).square().mean().backward()
    return m, torch.cat([p.grad.reshape(-1) for p in m.parameters()])


def _worker(rank, world, port, out, segments=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wtpse_b200.dp import FlatGradBucket, rank_batch_seed

        m = _model()
        bucket = FlatGradBucket(m, segments=segments)
        if segments > 1:
            # contiguous, disjoint, covering; the segment that fires first holds the last-registered parameters
            assert len(bucket.segments) == segments
            assert sorted(bucket.segments)[0][0] == 0 and max(h for _, h in bucket.segments) == bucket.flat.numel()
            assert sum(h - l for l, h in bucket.segments) == bucket.flat.numel()
            assert bucket.segments[0][1] == bucket.flat.numel()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        for it in range(2):                      # second round checks zero() keeps the views bound
            bucket.zero()
            g = torch.Generator().manual_seed(rank_batch_seed(7, rank, 0))
            x = torch.randn(6, 3, 8, 8, generator=g)
            if it == 0:
                bucket.arm()                     # segments > 1: collectives start from the autograd hooks
            m(x).square().mean().backward()
            for p in m.parameters():             # grads are views of the flat buffer
                assert p.grad.data_ptr() >= bucket.flat.data_ptr()
            if it == 0:
                if segments > 1:
                    assert all(bucket._fired), bucket._fired      # every segment fired during the backward pass
                bucket.finish()
                expect = sum(_local_grads(r)[1] for r in range(world)) / world
                assert torch.allclose(bucket.flat, expect, atol=1e-7), (bucket.flat - expect).abs().max()
            opt.step()
        # set_to_none elsewhere must not break the bucket
        m.zero_grad(set_to_none=True)
        bucket.zero()
        assert all(p.grad is not None for p in m.parameters())
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_is_mean_of_rank_grads():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_segmented_allreduce_fired_from_autograd_hooks():
    """Overlapped variant: the buffer is reduced in 3 segments, each started by the post-accumulate-grad hook of its
    last gradient; an un-armed backward (the shape update's dead teacher gradients) triggers nothing."""
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, 3), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_unarmed_backward_starts_no_collective():
    from wtpse_b200.dp import FlatGradBucket

    m = _model()
    bucket = FlatGradBucket(m, segments=3)
    bucket.zero()
    m(torch.randn(2, 3, 8, 8)).square().mean().backward()        # no process group, never armed
    assert bucket._pending is None and bucket._works == []
    bucket.finish()                                               # single process: no-op
    assert bucket.flat.abs().sum() > 0


def test_per_rank_batch_arithmetic_and_seeds():
    from wtpse_b200.dp import per_rank_batch, rank_batch_seed

    # BASELINE configs[3]: nominal global batch 128 on 8 ranks -> 16 per rank -> 3 x 5 = 15 used (Trainer.py:1013)
    assert per_rank_batch(128, 8, 3) == (5, 15)
    assert per_rank_batch(16, 1, 3) == (5, 15)
    assert per_rank_batch(9, 1, 3) == (3, 9)                     # the reference's default --batch-size 9
    seeds = {rank_batch_seed(1, r, it) for r in range(8) for it in range(50)}
    assert len(seeds) == 400
