"""Track W, factored db2 passes (csrc/wavelet_db2.cu) against the per-level kernels, on inputs whose detail coefficients are
bounded away from zero (x = inverse transform of coefficients with |c| in [0.1, 1]): the sign pattern -- the only discontinuity of
the L1 loss -- then does not depend on rounding, so two correct implementations must agree to fp32 rounding.  Also times the plans.
Usage: python tools/wavelet_db2_check.py [quick]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import wtpse_b200 as wb
from wtpse_b200 import wavelet as wv

dev = torch.device("cuda:0")


def safe_maps(shape, J, seed):
    g = torch.Generator().manual_seed(seed)
    c = (0.1 + 0.9 * torch.rand(*shape, generator=g)) * (2.0 * torch.randint(0, 2, shape, generator=g) - 1.0)
    return wb.idwt2d(c.to(dev), "db2", J).contiguous()


def run(x, J, weights):
    xg = x.clone().requires_grad_(True)
    loss = wb.wavelet_shape_loss(xg, "db2", J, weights)
    (0.7 * loss).backward()
    return float(loss), xg.grad.clone()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


cases = [((3, 1, 64, 128), 1), ((2, 1, 64, 128), 3), ((5, 1, 32, 256), 2), ((3, 2, 256, 256), 4), ((2, 1, 256, 256), 2), ((1, 2, 512, 512), 4),
         ((2, 2, 512, 512), 2), ((7, 1, 128, 512), 3), ((1, 1, 1024, 1024), 1), ((1, 2, 1024, 1024), 2), ((2, 1, 1024, 1024), 5),
         ((37, 1, 96, 256), 2), ((3, 1, 48, 128), 2), ((2, 1, 16, 256), 2), ((70, 1, 64, 256), 3)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    cases = cases[:6]
bad = 0
for shape, J in cases:
    x = safe_maps(shape, J, seed=J + shape[0])
    weights = tuple(0.5 + 0.25 * j for j in range(J))
    wb._lib.debug_set("wavelet_resident", 0)
    lp, gp = run(x, J, weights)
    wb._lib.debug_set("wavelet_resident", 1)
    out = []
    for db2, two in ((0, 0), (1, 0), (1, 1)):
        wb._lib.debug_set("wavelet_db2", db2)
        wb._lib.debug_set("wavelet_db2_two", two)
        cs = wv.resident_cluster_size(shape[-2], shape[-1], "db2", J)
        l, g = run(x, J, weights)
        l2, g2 = run(x, J, weights)
        ok = abs(l - lp) <= 2e-6 * abs(lp) and rel(g, gp) < 2e-6 and l2 == l and torch.equal(g, g2)
        bad += 0 if ok else 1
        out.append("db2=%d two=%d cs=%d dl=%.1e dg=%.1e%s" % (db2, two, cs, abs(l - lp) / abs(lp), rel(g, gp), "" if ok else "  <-- FAIL"))
    wb._lib.debug_set("wavelet_db2", 1)
    wb._lib.debug_set("wavelet_db2_two", 1)
    print(shape, "J=%d" % J, " | ".join(out), flush=True)
print("failures:", bad)

# timing through the C ABI is bench.py's job; here: autograd calls, relative numbers only
if not (len(sys.argv) > 1 and sys.argv[1] == "quick"):
    for shape, J in (((32, 2, 512, 512), 4), ((64, 2, 1024, 1024), 1), ((64, 2, 1024, 1024), 2)):
        x = torch.rand(*shape, device=dev)
        for db2, two in ((0, 0), (1, 0), (1, 1)):
            wb._lib.debug_set("wavelet_db2", db2)
            wb._lib.debug_set("wavelet_db2_two", two)
            for _ in range(3):
                run(x, J, None)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                run(x, J, None)
            torch.cuda.synchronize()
            print(shape, J, "db2=%d two=%d  %.1f us per fwd+bwd (autograd path)" % (db2, two, (time.perf_counter() - t0) * 1e5))
    wb._lib.debug_set("wavelet_db2", 1)
    wb._lib.debug_set("wavelet_db2_two", 1)
sys.exit(1 if bad else 0)
