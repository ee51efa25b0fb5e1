"""How far do parameter gradients move when the drop-in replaces the reference's loss block, against (i) the reference's
own run-to-run difference and (ii) the reference's own fp32-vs-fp64 difference?  Evidence for the tolerances of
tests/test_gpu_reference_dropin.py (``_check_grads``).  TEST INFRASTRUCTURE: imports oracle/ (the unmodified reference).

    python tools/dropin_noise_probe.py [seeds]      # prints one line per (shape, seed, mode)
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_reference_dropin as T  # noqa: E402


def main():
    import wtpse_b200 as wb
    from oracle import ref_shim

    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    alg, sn, _ = ref_shim.load()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    for n, S in T.SHAPES:
        for seed in range(seeds if S < 512 else 1):
            main_, shape = T._models(alg, sn, n, dev, seed=seed)
            image, od, _ = T._batch(n, S, dev, seed=10 + seed)
            _, g_stock, _ = T._update_pair(main_, shape, image, od)
            _, g_again, _ = T._update_pair(main_, shape, image, od)
            truth = T._fp64_truth(main_, shape, image, od)
            for mode in ("install", "bind_fused"):
                if mode == "install":
                    saved = wb.dropin.install(alg, sn)
                    try:
                        _, g_ours, _ = T._update_pair(main_, shape, image, od)
                    finally:
                        wb.dropin.uninstall(saved)
                else:
                    mb, sb = copy.deepcopy(main_), copy.deepcopy(shape)
                    wb.dropin.bind(mb, fuse_relu=True)
                    wb.dropin.bind(sb, fuse_relu=True)
                    _, g_ours, _ = T._update_pair(mb, sb, image, od)
                rows = []
                gmax = max(float(w.abs().max()) for w in g_stock.values())
                for k, w in g_stock.items():
                    scale = T._grad_scale(k, g_stock, gmax)
                    e = float((g_ours[k] - w).abs().max())
                    nz = float((g_again[k] - w).abs().max())
                    t = float((w.double() - truth[k]).abs().max())
                    rows.append((e / scale, nz / scale, t / scale, k))
                rows.sort(reverse=True)
                r = rows[0]
                print("n=%d S=%d seed=%d %-10s worst ours-vs-stock %.2e (stock-vs-stock %.2e, stock-vs-fp64 %.2e) %s; tensors above 1e-5: %d of %d"
                      % (n, S, seed, mode, r[0], r[1], r[2], r[3], sum(x[0] > 1e-5 for x in rows), len(rows)), flush=True)


if __name__ == "__main__":
    main()
