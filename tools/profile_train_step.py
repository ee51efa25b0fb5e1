"""Top CUDA kernels of one full train step, and the element-wise / copy operators grouped by input shape and strides
(diagnostic; needs a GPU).  Env: SIZE (512), CHANNELS_LAST (1), CUDNN_BENCHMARK (1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import wtpse_b200 as wb

if os.environ.get("CUDNN_BENCHMARK", "1") == "1":
    torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
S = int(os.environ.get("SIZE", "512"))
ts = wb.TrainStep(n_per_domain=5, n_domains=3, device=dev, seed=0, channels_last=os.environ.get("CHANNELS_LAST", "1") == "1")


def one(it):
    image, od, oc = wb.synthetic.fundus_batch(5, 3, S, S, dev, seed=it)
    return ts.step(image, od, oc)


for it in range(3):
    one(it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    one(20)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
print("==== by input shape ====")
rows = [e for e in prof.key_averages(group_by_input_shape=True)
        if e.key.startswith(("aten::add", "aten::copy_", "aten::mul", "aten::sum", "aten::clone", "aten::contiguous", "aten::cat",
                             "aten::threshold_backward", "aten::relu", "aten::upsample", "aten::max_pool", "aten::div",
                             "aten::sigmoid", "aten::fill_", "aten::zero_", "aten::where", "aten::native_batch_norm",
                             "aten::cudnn_batch_norm"))]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:45]:
    print("%-34s n=%-4d self_cuda=%8.2f ms  %s" % (e.key, e.count, e.self_device_time_total / 1e3, str(e.input_shapes)[:150]))
