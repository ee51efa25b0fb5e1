"""Phase clocks of the in-kernel forward tail (whitening_tail.cuh): the critical path between the last pixel tile and the end
of the Gram kernel.  Usage: python tools/tail_phases.py [B] [H] [n]   (GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wtpse_b200 as wb  # noqa: E402
from wtpse_b200 import functional as wf  # noqa: E402

B, H, n = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 32), (2, 512), (3, 10)))
lib = wb._lib.load()
dev = torch.device("cuda:0")
z = (0.3 * torch.randn(B, 16, H, H, device=dev) + 0.2 * torch.randn(B, 16, 1, 1, device=dev))
wb._lib.debug_set("tail_stamps", 1)
names = ["flush begin", "flush end", "ticket(sample)", "reduce", "ticket(batch)", "(unused)", "final begin", "A stage v/stat",
         "B pairwise", "C block sums", "D domgrad (warps 1..; warp 0: scalars)"]
rows = []
for it in range(6):
    wb.whitening_terms(z, n, 3)
    torch.cuda.synchronize()
    ws, nbytes = wf._workspace(lib, B, H * H, dev)
    st = ws[nbytes - 128:nbytes].view(torch.int64).cpu().tolist()
    rows.append(st)
wb._lib.debug_set("tail_stamps", 0)
clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
for st in rows[2:]:
    t0 = st[0]
    print("  ".join("%s +%.2fus" % (names[i], (st[i] - t0) / (clk * 1e-3 if clk > 10000 else clk)) for i in (1, 2, 3, 4, 6, 7, 8, 9, 10)))
