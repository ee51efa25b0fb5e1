"""Runs the four loss variants (NCHW / channels-last x plain / fused with the DeepWT-tail ReLU), forward + backward, at the
bench size -- the target of the ncu launch lists and --set full captures under profiles/.
Usage: python tools/loss_probe.py [iters] [variants, e.g. nchw,nchw_relu,cl,cl_relu]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import wtpse_b200 as wb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["nchw", "nchw_relu", "cl", "cl_relu"]
dev = torch.device("cuda:0")
torch.manual_seed(1234)
z0 = 0.3 * torch.randn(32, 16, 512, 512, device=dev) + 0.2 * torch.randn(32, 16, 1, 1, device=dev)
g0 = torch.randn_like(z0)
one = torch.ones((), device=dev)
for name in which:
    cl = name.startswith("cl")
    z = (z0.contiguous(memory_format=torch.channels_last) if cl else z0.clone()).requires_grad_(True)
    g = g0.contiguous(memory_format=torch.channels_last) if cl else g0
    for _ in range(iters):
        z.grad = None
        if name.endswith("relu"):
            r, ins, dom = wb.relu_whitening_folded(z, 10, 3)
            torch.autograd.backward([r, ins, dom], [g, one, one])
        else:
            ins, dom = wb.whitening_folded(z, 10, 3)
            torch.autograd.backward([ins, dom], [one, one])
    torch.cuda.synchronize()
    if os.environ.get("PROBE_TIMES"):
        import ctypes
        lib = wb._lib.load()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            z.grad = None
            if name.endswith("relu"):
                r, ins, dom = wb.relu_whitening_folded(z, 10, 3)
                torch.autograd.backward([r, ins, dom], [g, one, one])
            else:
                ins, dom = wb.whitening_folded(z, 10, 3)
                torch.autograd.backward([ins, dom], [one, one])
        ev1.record()
        torch.cuda.synchronize()
        lib.wtpse_profile_reset()
        lib.wtpse_profile_enable(1)
        for _ in range(3):
            z.grad = None
            if name.endswith("relu"):
                r, ins, dom = wb.relu_whitening_folded(z, 10, 3)
                torch.autograd.backward([r, ins, dom], [g, one, one])
            else:
                ins, dom = wb.whitening_folded(z, 10, 3)
                torch.autograd.backward([ins, dom], [one, one])
        lib.wtpse_profile_enable(0)
        torch.cuda.synchronize()
        kern = {}
        for kid in range(lib.wtpse_profile_kernel_count()):
            cnt, ms = ctypes.c_longlong(0), ctypes.c_double(0.0)
            lib.wtpse_profile_read(kid, ctypes.byref(cnt), ctypes.byref(ms))
            if cnt.value:
                kern[lib.wtpse_profile_kernel_name(kid).decode()] = round(ms.value / cnt.value * 1e3, 1)
        print(name, "pair %.1f us" % (ev0.elapsed_time(ev1) * 100), kern)
    print(name, "ok", float(ins), float(dom))
    del z
