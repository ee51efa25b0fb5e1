"""Pure-write, pure-read and copy bandwidth of this GPU with library kernels (cudaMemset, a reduction, cudaMemcpy D2D) on buffers far larger
than L2 -- the ceilings the write-dominated (wavelet synthesis) and read-dominated (analysis, Gram) kernels should be read against.
MEASURED_PEAKS.json's hbm_gbs is a COPY figure (read + write)."""
import torch

dev = torch.device("cuda:0")
n = 1 << 29                                   # 2 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


t_w = timed(lambda: a.zero_())
t_r = timed(lambda: a.sum())
t_c = timed(lambda: b.copy_(a))
gb = n * 4 / 1e9
print("write (memset)   %.0f GB/s   (zeros: the memory system may compress them -- an upper bound, not a target)" % (gb / t_w))
print("read  (sum)      %.0f GB/s" % (gb / t_r))
print("copy  (D2D)      %.0f GB/s read + write" % (2 * gb / t_c))
