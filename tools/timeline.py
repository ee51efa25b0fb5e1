"""Kernel start/end times (CUPTI, through torch.profiler) of the device-resident loops of tools/step_breakdown.py: forward
only, backward only, and the pair -- where the pair's extra microseconds sit (gaps and overlaps between consecutive launches).
Usage: python tools/timeline.py [B] [H] [n]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wtpse_b200 as wb  # noqa: E402
from wtpse_b200 import functional as wf  # noqa: E402

B, H, n = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 32), (2, 512), (3, 10)))
lib = wb._lib.load()
dev = torch.device("cuda:0")
P = H * H
zs = [0.3 * torch.randn(B, 16, H, H, device=dev) + 0.2 * torch.randn(B, 16, 1, 1, device=dev) for _ in range(2)]
dz = torch.empty_like(zs[0])
ws, ws_bytes = wf._workspace(lib, B, P, dev)
losses, (gram, rowstat, domgrad) = wf._forward_outputs(B, dev)
one = torch.ones((), device=dev)
st = wf._stream_ptr(dev)
p = wf._ptr


def fwd(z):
    wb._lib.check(lib.wtpse_whitening_forward(p(z), B, 16, P, n, 3, 0.0, 1e-5, p(losses), p(gram), p(rowstat), p(domgrad), p(ws), ws_bytes, st))


def bwd(z):
    wb._lib.check(lib.wtpse_whitening_backward(p(z), p(gram), p(rowstat), p(domgrad), p(one), p(one), p(one), B, 16, P, n, 3, p(dz), st))


def trace(name, fn, iters=12):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(iters):
            fn(i)
        torch.cuda.synchronize()
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "wtpse" in e.name),
                key=lambda e: e.time_range.start)
    print("== %s: %d kernels" % (name, len(ev)))
    prev_end = None
    t0 = ev[0].time_range.start
    rows = []
    for e in ev[4:]:
        s, t = e.time_range.start, e.time_range.end
        short = "gram " if "gram" in e.name else "apply"
        rows.append((short, s - t0, t - s, (s - prev_end) if prev_end is not None else 0.0))
        prev_end = t
    prev_end = ev[3].time_range.end
    for short, s, d, _ in rows[:8]:
        pass
    prev = ev[3]
    for e in ev[4:12]:
        s, t = e.time_range.start, e.time_range.end
        print("  %s start %8.1f us  dur %7.1f us  start - previous end %+6.1f us" % ("gram " if "gram" in e.name else "apply", s - t0, t - s,
                                                                                      s - prev.time_range.end))
        prev = e
    span = (ev[-1].time_range.end - ev[4].time_range.start) / max(1, (len(ev) - 4))
    print("  per kernel over the window: %.1f us" % span)


trace("forward only", lambda i: fwd(zs[i & 1]))
trace("backward only", lambda i: bwd(zs[i & 1]))
trace("pair", lambda i: (fwd(zs[i & 1]), bwd(zs[i & 1])))
