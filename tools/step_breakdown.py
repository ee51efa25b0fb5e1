"""Where a device-resident step goes: forward alone, backward alone and the pair, each as a back-to-back loop through the
C ABI (programmatic dependent launch in effect, no events inside the loops).  Usage: python tools/step_breakdown.py [B] [H] [n] [iters]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wtpse_b200 as wb  # noqa: E402
from wtpse_b200 import functional as wf  # noqa: E402

B, H, n, iters = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 32), (2, 512), (3, 10), (4, 50)))
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


def timed(fn):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


t_f = timed(lambda i: fwd(zs[i & 1]))
t_b = timed(lambda i: bwd(zs[i & 1]))
t_p = timed(lambda i: (fwd(zs[i & 1]), bwd(zs[i & 1])))
bytes_f, bytes_b = 64.0 * B * P, 128.0 * B * P
print("B=%d H=%d n=%d  forward %.1f us (%.3f of 6548.2 GB/s)  backward %.1f us (%.3f)  pair %.1f us (%.3f)  pair - (fwd + bwd) = %+.1f us"
      % (B, H, n, t_f, bytes_f / t_f / 6548.2e3, t_b, bytes_b / t_b / 6548.2e3, t_p, (bytes_f + bytes_b) / t_p / 6548.2e3, t_p - t_f - t_b))
if os.environ.get("SWEEP_SCHEDULE"):
    for rr in (1, 0):
        wb._lib.debug_set("apply_round_robin", rr)
        t_b2 = timed(lambda i: bwd(zs[i & 1]))
        t_p2 = timed(lambda i: (fwd(zs[i & 1]), bwd(zs[i & 1])))
        print("apply_round_robin=%d  backward alone %.1f us (%.3f)  pair %.1f us (%.3f)" % (rr, t_b2, bytes_b / t_b2 / 6548.2e3, t_p2, (bytes_f + bytes_b) / t_p2 / 6548.2e3))
    wb._lib.debug_set("apply_round_robin", 1)
