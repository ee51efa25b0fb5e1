"""Runs the fused DeepWT-tail pair (wtpse_whitening_relu_forward/backward) a few times at the bench size; the target of
the ncu captures under profiles/ (r1_ncu_fusion_*).  Usage: python tools/fusion_probe.py [iters] [cl]
("cl": channels-last tensors -> the wtpse_whitening_*_cl kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import wtpse_b200 as wb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
torch.manual_seed(1234)
z = (0.3 * torch.randn(32, 16, 512, 512, device=dev) + 0.2 * torch.randn(32, 16, 1, 1, device=dev)).requires_grad_(True)
if len(sys.argv) > 2 and sys.argv[2] == "cl":
    z = z.detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)
g = torch.randn_like(z)
one = torch.ones((), device=dev)
for _ in range(iters):
    z.grad = None
    r, ins, dom = wb.relu_whitening_folded(z, 10, 3)
    torch.autograd.backward([r, ins, dom], [g, one, one])
torch.cuda.synchronize()
print("ok", float(ins), float(dom))
