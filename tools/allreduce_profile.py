"""NCCL all-reduce time per train iteration, overlapped segments vs one collective per backward (CUPTI through torch.profiler; launch
under torchrun, rank 0 prints).  For each setting: the NCCL kernels' summed duration per iteration, how much of it runs while
another kernel of this rank is executing (hidden) and how much is exposed, and the iteration time.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import wtpse_b200 as wb

rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True
S, ITERS = 512, 4


def measure(segments):
    ts = wb.TrainStep(n_per_domain=5, n_domains=3, device=dev, seed=0, grad_segments=segments)
    batches = [wb.synthetic.fundus_batch(5, 3, S, S, dev, seed=100 * rank + i) for i in range(2)]
    for i in range(3):
        ts.step(*batches[i % 2])
    torch.cuda.synchronize()
    dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(ITERS):
            ts.step(*batches[i % 2])
        torch.cuda.synchronize()
    ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    nccl = [(a, b) for a, b, n in ev if "nccl" in n.lower()]
    other = sorted((a, b) for a, b, n in ev if "nccl" not in n.lower() and "memcpy" not in n.lower() and "memset" not in n.lower())
    # merge the compute intervals, then measure how much of every NCCL kernel they cover
    merged = []
    for a, b in other:
        if merged and a <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], b)
        else:
            merged.append([a, b])
    hidden = 0.0
    for a, b in nccl:
        for c, d in merged:
            if d <= a:
                continue
            if c >= b:
                break
            hidden += min(b, d) - max(a, c)
    total = sum(b - a for a, b in nccl)
    span = (max(b for _, b, _ in ev) - min(a for a, _, _ in ev)) / ITERS
    return len(nccl) / ITERS, total / ITERS, hidden / ITERS, span


for seg in (3, 1):
    n, tot, hid, span = measure(seg)
    if rank == 0:
        print("segments=%d  world=%d  NCCL kernels/iteration %.1f  NCCL time/iteration %.3f ms  hidden under this rank's kernels %.3f ms  "
              "exposed %.3f ms  iteration (eager, profiled) %.1f ms" % (seg, dist.get_world_size(), n, tot / 1e3, hid / 1e3, (tot - hid) / 1e3, span / 1e3),
              flush=True)
dist.destroy_process_group()
