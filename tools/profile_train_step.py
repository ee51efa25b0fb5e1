"""Top CUDA kernels of one full train step (diagnostic; needs a GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wtpse_b200 as wb
from torch.profiler import profile, ProfilerActivity

if os.environ.get("CUDNN_BENCHMARK", "0") == "1":
    torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
S = int(os.environ.get("SIZE", "512"))
ts = wb.TrainStep(n_per_domain=5, n_domains=3, device=dev, seed=0, channels_last=os.environ.get('CHANNELS_LAST', '0') == '1')
def one(it):
    image, od, oc = wb.synthetic.fundus_batch(5, 3, S, S, dev, seed=it)
    return ts.step(image, od, oc)
for it in range(3):
    one(it)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for it in range(3):
    one(10 + it)
torch.cuda.synchronize()
print("ms/step %.1f" % ((time.perf_counter() - t0) / 3 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    one(20)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
