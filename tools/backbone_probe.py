"""Runs the channels-last backbone kernels (BatchNorm + ReLU fwd/bwd, channel sum, bias + ReLU, x2 up-sampling) once per
U-Net stage shape of the 15 x 512 x 512 train step; the target of the ncu launch list profiles/r1_launches_backbone.csv."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import wtpse_b200 as wb
from wtpse_b200.elementwise import batch_norm_act, channel_sum, max_pool2

dev = torch.device("cuda:0")
torch.manual_seed(0)
for C, S in ((16, 512), (32, 512), (32, 256), (64, 128), (128, 64), (256, 32)):
    x = torch.randn(15, C, S, S, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    g = torch.randn(15, C, S, S, device=dev).contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(C).to(dev).train()
    for _ in range(2):
        x.grad = None
        y = batch_norm_act(x, bn, True)
        y.backward(g)
        channel_sum(g)
    if S < 512:
        u = wb.upsample2x(x)
        u.backward(torch.randn_like(u))
    if S > 32:
        m = max_pool2(x)
        m.backward(torch.randn_like(m))
torch.cuda.synchronize()
print("ok")
