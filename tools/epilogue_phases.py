"""Print clock64 deltas between the phases of the single-CTA epilogue kernels (diagnostic; needs a GPU).
Forces the one-kernel forward epilogue and the single-CTA backward epilogue (debug modes), which carry the stamps."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wtpse_b200 as wb

B, H, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (32, 512, 10)
lib = wb._lib.load()
lib.wtpse_debug_set_two_stage_epilogue(0)
lib.wtpse_debug_set_backward_mode(2)
lib.wtpse_debug_set_epilogue_repeat(int(os.environ.get("REPEAT", "1")))
dev = torch.device("cuda:0")
z = (0.3 * torch.randn(B, 16, H, H, device=dev) + 0.2 * torch.randn(B, 16, 1, 1, device=dev)).requires_grad_(True)
stamps = torch.zeros(16, dtype=torch.int64, device=dev)
one = torch.ones((), device=dev)
for it in range(4):
    if it == 3:
        lib.wtpse_debug_set_stamp_buffer(ctypes.c_void_p(stamps.data_ptr()))
    z.grad = None
    ins, dom = wb.whitening_folded(z, n, 3)
    torch.autograd.backward([ins, dom], [one, one])
torch.cuda.synchronize()
lib.wtpse_debug_set_stamp_buffer(None)
s = stamps.cpu().tolist()
rep = "x%d" % int(os.environ.get("REPEAT", "1"))
print(rep, "fwd phases (cycles): A=%d B=%d C=%d D=%d total=%d" % (s[1]-s[0], s[2]-s[1], s[3]-s[2], s[4]-s[3], s[4]-s[0]))
print("bwd phases (cycles): stage_v=%d pairwise=%d mmat=%d total=%d" % (s[9]-s[8], s[10]-s[9], s[11]-s[10], s[11]-s[8]))
