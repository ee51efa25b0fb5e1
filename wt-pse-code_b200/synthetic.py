"""Synthetic fundus-shaped batches generated on the device.

Replaces the reference's in-memory PIL pools (fundus_dataloader.py:86-99) for benchmarks and tests: a
smooth reddish background with a bright disc and a darker cup, plus the raw uint8 mask in the dataset's
encoding (0 inside the cup, 128 in the disc rim, 255 background -- custom_transforms.py:473-497).  The
batch is ordered [domain A x n | domain B x n | ...] like Trainer.get_multi_batch (Trainer.py:45-55), with
a per-domain colour/intensity shift so the MMD sees distinct domains.  The uint8 -> fp32 / label
conversion is the bit-exact ``prepare_batch`` kernel, i.e. the same integer path real data takes.
"""
import torch

from .elementwise import prepare_batch


def raw_fundus_batch(n_per_domain, n_domains, H, W, device, seed=0):
    """Returns (img uint8 B x H x W x 3, raw_mask uint8 B x H x W), B = n_per_domain * n_domains."""
    B = n_per_domain * n_domains
    g = torch.Generator(device="cpu").manual_seed(seed)
    # per-sample geometry on the host (B numbers), rasterised on the device
    cx = (0.5 + 0.12 * (torch.rand(B, generator=g) - 0.5)).to(device).view(B, 1, 1)
    cy = (0.5 + 0.12 * (torch.rand(B, generator=g) - 0.5)).to(device).view(B, 1, 1)
    rd = (0.22 + 0.06 * torch.rand(B, generator=g)).to(device).view(B, 1, 1)          # disc radius
    ecc = (0.85 + 0.3 * torch.rand(B, generator=g)).to(device).view(B, 1, 1)          # ellipse aspect
    cup = (0.35 + 0.3 * torch.rand(B, generator=g)).to(device).view(B, 1, 1)          # cup/disc ratio
    noise_seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g))
    ys = torch.linspace(0, 1, H, device=device).view(1, H, 1)
    xs = torch.linspace(0, 1, W, device=device).view(1, 1, W)
    r = torch.sqrt(((xs - cx) / ecc) ** 2 + (ys - cy) ** 2)
    raw = torch.full((B, H, W), 255, dtype=torch.uint8, device=device)
    raw[r <= rd] = 128
    raw[r <= rd * cup] = 0
    # image: background falloff + bright disc + brighter cup, per-domain colour shift, mild noise
    domain = (torch.arange(B, device=device) // n_per_domain).view(B, 1, 1, 1).float()
    base = torch.tensor([150.0, 70.0, 40.0], device=device).view(1, 1, 1, 3)
    shift = torch.tensor([18.0, -9.0, 12.0], device=device).view(1, 1, 1, 3) * (domain - (n_domains - 1) / 2)
    vign = (1.0 - 0.9 * ((xs - 0.5) ** 2 + (ys - 0.5) ** 2)).unsqueeze(-1)
    disc = torch.sigmoid((rd - r) * 60.0).unsqueeze(-1)
    cupm = torch.sigmoid((rd * cup - r) * 60.0).unsqueeze(-1)
    gd = torch.Generator(device=device).manual_seed(noise_seed)
    img = (base + shift) * vign + 70.0 * disc + 35.0 * cupm + 4.0 * torch.randn(B, H, W, 3, device=device, generator=gd)
    return img.clamp_(0, 255).to(torch.uint8), raw


def fundus_batch(n_per_domain, n_domains, H, W, device, seed=0):
    """(image B x 3 x H x W in [-1, 1], label_od, label_oc) through the integer label path."""
    img, raw = raw_fundus_batch(n_per_domain, n_domains, H, W, device, seed)
    return prepare_batch(raw, img)
