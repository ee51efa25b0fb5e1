#!/usr/bin/env python
"""bench.py -- shape-loss fwd+bwd throughput (Mpix/s) on B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one forward + backward of the whitening/MMD shape-regularization loss
(WT_PSE.compute_whitening_loss, algorithms.py:1277-1309) over one synthetic batch.  At N=1 the workload is
BASELINE.json configs[1] mapped onto what the reference's entry point accepts:
z = 32 x 16 x 512 x 512 fp32 feature maps, 3 domains x 10 samples.  1 pix = one (b,h,w) site, all 16 channels.

  value   device-resident throughput (inputs already in HBM), whole job over all ranks
  e2e     same metric through the host-buffer C-ABI entry point (wtpse_host_plan_run): pinned host z in,
          H2D, forward, backward, D2H of dz and of the losses, every step
  roofline  the dominant kernel (apply_tma_kernel, backward) -- algorithmic bytes / its CUDA-event duration
            measured inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU path (the unmodified algorithms.py from oracle/_ref, `kind: "reference"`; the
          oracle's operator-sequence port only if that copy is missing) on this box's host cores, SAME workload
  configs   the other BASELINE configs mapped onto Track R (8x16x256x256, 64x16x1024x1024) -- device-resident fwd+bwd
  wavelet   BASELINE configs as literally written (Track W, parity unpinned): Haar J=3 8x2x256^2, db2 J=4 32x2x512^2,
          Haar + db2 J=1..5 at 64x2x1024^2

N > 1 (torchrun, one rank per GPU): the path shards over whole [K domains x n] batches with no data-path
collective; every rank processes its own batch -> "scaling": "weak".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(B=32, C=16, H=512, W=512, n_per_domain=10, n_domains=3, margin=0.0, eps=1e-5)
ALGO_BYTES_PER_PIX = {"gram_tma_kernel": 64, "apply_tma_kernel": 128}   # fwd read z; bwd read z + write dz
FALLBACK_PEAK_GBS = 6650.0                                               # B200_PROFILING.md fallback
METRIC = "shape-loss fwd+bwd Mpix/s"


def workload_config(B, H, W, n, K):
    """`config` of the JSON line -- built by ONE function so that both arms (ours / --impl reference) print the same dict."""
    return {"workload": "whitening+MMD shape loss fwd+bwd (WT_PSE.compute_whitening_loss), z=%dx16x%dx%d fp32, "
                        "n=%d K=%d per GPU [BASELINE configs[1] mapped per SURVEY 8(d)]" % (B, H, W, n, K),
            "per_gpu_batch": B,
            "l2": "inputs (%.0f MB, two alternating buffers) larger than L2" % (B * 16 * H * W * 4 / 1e6),
            "sharding": "independent [K x n] batches per rank, no data-path collective"}


def workload_dims(args):
    B, H = args.batch, args.size
    n = max(1, B // 3) if B != WORKLOAD["B"] else WORKLOAD["n_per_domain"]
    return B, H, H, n, WORKLOAD["n_domains"]


# the CUDA sources each traffic entry of profiles/traffic.json depends on (kernels + the host code that picks and shapes them)
TRAFFIC_SOURCES = {
    "whitening": ("common.cuh", "mmd_device.cuh", "whitening_tail.cuh", "whitening_matrix.cuh", "whitening_gram.cu", "whitening_apply.cu"),
    "wavelet": ("common.cuh", "wavelet_db2.cu", "wavelet_resident.cu", "wavelet_stream.cu", "wavelet_tiles.cu", "wavelet_level.cuh",
                "wavelet_bank.cuh"),
}


def kernel_source_hash(group="whitening"):
    """sha256 over the CUDA sources a traffic entry was captured with (the GPU box has no .git): profiles/traffic.json records the hash
    per group, so a stale `roofline.traffic` is flagged instead of silently reused -- and a change to an unrelated kernel does not."""
    import hashlib

    csrc = os.path.join(ROOT, "wt-pse-code_b200", "csrc")
    h = hashlib.sha256()
    for name in TRAFFIC_SOURCES[group]:
        h.update(name.encode())
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def traffic_entry(key, group="whitening"):
    """(bytes per launch or None, stale flag, capture hash) from profiles/traffic.json."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, None, None
    t = json.load(open(tpath))
    cap = t.get("kernel_source_hash" if group == "whitening" else group + "_source_hash")
    return t.get(key), (cap != kernel_source_hash(group)), cap


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-reference-work", type=int, default=1,
                    help="also time the iteration with the reference's dead teacher backward kept (train_step.with_teacher_backward)")
    ap.add_argument("--train-fuse-relu", type=int, default=1, help="train step with the fused DeepWT tail (SURVEY 8(f).1)")
    ap.add_argument("--train-fuse-compare", type=int, default=1, help="also time the train step with the other --train-fuse-relu setting")
    ap.add_argument("--train-cudnn-benchmark", type=int, default=1, help="torch.backends.cudnn.benchmark for the train-step backbone")
    ap.add_argument("--wavelet-name", default="db2", choices=["haar", "db2"])
    ap.add_argument("--wavelet-levels", type=int, default=4)
    ap.add_argument("--wavelet-batch", type=int, default=32)
    ap.add_argument("--wavelet-split", type=int, default=-1, help="Track W fused plan: -1 auto, 0 whole map resident, 1 level 1 streamed")
    ap.add_argument("--wavelet-peel-max", type=int, default=8, help="Track W: most levels streamed before the resident stage")
    ap.add_argument("--wavelet-tiles", type=int, default=1, help="Track W: level 1 of the streamed plan as TMA pipelines (1) or per-thread loads (0)")
    ap.add_argument("--wavelet-cluster-max", type=int, default=8, help="Track W: largest cluster size of the resident stage")
    ap.add_argument("--wavelet-resident", type=int, default=1, help="Track W: 0 = per-level kernels instead of the cluster-resident kernel")
    ap.add_argument("--track", default="whitening", choices=["whitening", "wavelet"],
                    help="wavelet = BASELINE configs[1] as literally written (Track W, parity unpinned); not the default")
    ap.add_argument("--e2e-steps", type=int, default=20, help="steps of the host-buffer (PCIe-bound) loop")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget for the cpu_baseline leg")
    ap.add_argument("--batch", type=int, default=WORKLOAD["B"])
    ap.add_argument("--size", type=int, default=WORKLOAD["H"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-steps", type=int, default=20, help="timed iterations of the full WT-PSE train step (0 = skip)")
    ap.add_argument("--train-grad-segments", type=int, default=3,
                    help="pieces each network's gradient buffer is all-reduced in, started from autograd hooks during the backward pass (1 = one collective after it)")
    ap.add_argument("--train-segments-compare", type=int, default=1, help="N > 1: also time the train step with one un-overlapped all-reduce per backward")
    ap.add_argument("--train-size", type=int, default=512)
    ap.add_argument("--train-graph", type=int, default=1, help="replay the train iteration as a CUDA graph")
    ap.add_argument("--train-batch", type=int, default=16, help="nominal per-GPU batch (the reference uses 3 * (batch // 3))")
    ap.add_argument("--no-kernel-events", action="store_true", help="diagnostic: timed region without per-kernel CUDA events")
    ap.add_argument("--debug", action="append", default=[], metavar="NAME=VALUE",
                    help="diagnostic switch of include/wtpse_b200_debug.h, e.g. --debug fused_tail=0 (repeatable)")
    ap.add_argument("--extra-configs", type=int, default=1, help="also time the other BASELINE configs (Track R 256^2 / 1024^2, Track W) at N=1")
    ap.add_argument("--numa-affinity", type=int, default=1, help="bind each rank to the CPUs NVML reports as local to its GPU before allocating pinned buffers")
    ap.add_argument("--train-reference-eager", type=int, default=1, help="also time the UNMODIFIED reference classes' iteration (oracle/_ref) on the same GPU")
    ap.add_argument("--event-stride", type=int, default=10, help="bracket kernels with CUDA events on every n-th timed step")
    ap.add_argument("--launch-gate-ms", type=float, default=2.0,
                    help="device-side spin in front of the timed region so the host has steps queued when it starts (0 = none)")
    return ap.parse_args()


def synth_batch(B, H, W, seed, device=None, pin=False):
    """SURVEY.md 8(d) synthetic input: 0.3*randn + per-sample-per-channel offset 0.2*randn (domains differ)."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    off = 0.2 * torch.randn(B, 16, 1, 1, generator=g)
    if device is not None:
        gd = torch.Generator(device=device).manual_seed(seed)
        z = 0.3 * torch.randn(B, 16, H, W, generator=gd, device=device) + off.to(device)
        return z
    z = torch.empty(B, 16, H, W, pin_memory=pin)
    torch.randn(B, 16, H, W, generator=g, out=z)
    z.mul_(0.3).add_(off)
    return z


# -------------------------------------------------------------------------------------------------
# clocks sampler (NVML, background thread)
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def mark(self):
        """Forget what was sampled so far (warm-up): the report covers the region that starts here."""
        self.samples, self.reasons = [], set()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# -------------------------------------------------------------------------------------------------
# CPU legs (oracle port, or the real reference when its tree is on this box)
# -------------------------------------------------------------------------------------------------
def cpu_step_fn(n, K):
    """Returns (callable(z) -> None doing one fwd+bwd on HOST tensors, kind, context manager to run it under)."""
    import contextlib

    import torch
    from oracle import ref_shim

    if ref_shim.available():
        # the unmodified reference (oracle/_ref: byte-for-byte copy, sha256-checked; /root/reference in the build
        # container); its hard-coded .cuda() calls are neutralised for the CPU leg (ref_shim.cpu_only)
        alg, _, _ = ref_shim.load()
        hp = dict(ref_shim.DEFAULT_HPARAMS)
        with ref_shim.cpu_only():
            torch.manual_seed(0)
            model = alg.WT_PSE(3, 1, hp, "cpu", False, per_domain_batch=n, source_domain_num=K)

        def step(z):
            z = z.detach().requires_grad_(True)
            ins, dom = model.compute_whitening_loss(z)          # algorithms.py:1277-1309, as published
            (ins + dom).backward()
            return float(ins.detach()), float(dom.detach())
        return step, "reference", ref_shim.cpu_only
    from oracle import whitening_torch as wt

    def step(z):
        ins, dom, _ = wt.fwd_bwd(z, n, K)
        return float(ins.detach()), float(dom.detach())
    return step, "port", contextlib.nullcontext


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def time_cpu(B, H, W, n, K, budget_s, max_iters, z=None):
    """cpu_baseline: the SAME workload as the GPU arm on this box's host cores; bounded by time, not by shrinking it.
    z: host copy of the very batch the GPU arm's reported losses come from (else a seeded host batch)."""
    import torch

    old = torch.get_num_threads()
    torch.set_num_threads(host_threads())          # torchrun exports OMP_NUM_THREADS=1
    step, kind, ctx = cpu_step_fn(n, K)
    if z is None:
        z = synth_batch(B, H, W, seed=7)
    with ctx():
        step(z)                                     # warm-up
        times = []
        t_start = time.perf_counter()
        while len(times) < max_iters and (len(times) < 2 or (time.perf_counter() - t_start) < budget_s):
            t0 = time.perf_counter()
            losses = step(z)
            times.append(time.perf_counter() - t0)
    pix = B * H * W
    out = {"value": pix / min(times) / 1e6, "mean_value": pix / (sum(times) / len(times)) / 1e6, "unit": "Mpix/s",
           "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": kind, "iters": len(times),
           "losses": list(losses),
           "sample": "%dx16x%dx%d fp32 (n=%d,K=%d) fwd+bwd -- the full workload, best of %d after 1 warm-up" % (B, H, W, n, K, len(times))}
    torch.set_num_threads(old)
    return out


def time_gpu_eager(z, n, K, iters=10):
    """The reference's own eager-CUDA path on the same GPU and the same resident input: the unmodified
    `algorithms.WT_PSE.compute_whitening_loss` (its hard-coded .cuda() copies and .item() sync included) when oracle/_ref
    is present, else the oracle's operator-sequence port.  A reported baseline -- test infrastructure, never the product path."""
    import torch
    from oracle import ref_shim

    if ref_shim.available():
        alg, _, _ = ref_shim.load()
        torch.manual_seed(0)
        model = alg.WT_PSE(3, 1, dict(ref_shim.DEFAULT_HPARAMS), z.device, False, per_domain_batch=n, source_domain_num=K)
        kind = "reference: unmodified algorithms.WT_PSE.compute_whitening_loss in PyTorch eager (CUDA), same GPU, same device-resident input"

        def fwd_bwd():
            zz = z.detach().requires_grad_(True)
            ins, dom = model.compute_whitening_loss(zz)
            (ins + dom).backward()
            return ins, dom
    else:
        from oracle import whitening_torch as wt
        kind = "port: reference operator sequence in PyTorch eager (CUDA), same GPU, same device-resident input"

        def fwd_bwd():
            ins, dom, _ = wt.fwd_bwd(z, n, K)
            return ins, dom
    for _ in range(2):
        fwd_bwd()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(iters):
        ins, dom = fwd_bwd()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    return {"value": z.shape[0] * z.shape[2] * z.shape[3] / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": ms, "iters": iters,
            "kind": kind, "losses": [float(ins.detach()), float(dom.detach())]}


def time_relu_fusion(z, n, K, peak, iters=10):
    """embedding + the ReLU that follows it (DeepWT tail), forward and backward with an upstream gradient
    on relu(z).  `unfused` = ATen relu / threshold_backward / gradient add around the plain loss kernels; `fused` =
    wtpse_whitening_relu_forward/backward.  Device-resident, same input as the headline step."""
    import torch

    import wtpse_b200 as wb

    zz = z.detach().clone().requires_grad_(True)
    g = torch.randn_like(zz)
    one = torch.ones((), device=z.device)
    pix = z.shape[0] * z.shape[2] * z.shape[3]

    def unfused():
        zz.grad = None
        ins, dom = wb.whitening_folded(zz, n, K)
        r = torch.relu(zz)
        torch.autograd.backward([ins, dom, r], [one, one, g])

    def fused():
        zz.grad = None
        r, ins, dom = wb.relu_whitening_folded(zz, n, K)
        torch.autograd.backward([r, ins, dom], [g, one, one])

    zc = z.detach().clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gc_ = g.contiguous(memory_format=torch.channels_last)

    def fused_cl():                                   # the same pair on channels-last tensors (wtpse_whitening_*_cl)
        zc.grad = None
        r, ins, dom = wb.relu_whitening_folded(zc, n, K)
        torch.autograd.backward([r, ins, dom], [gc_, one, one])

    def plain_cl():                                   # plain loss fwd+bwd on a channels-last z: 192 B/pix, no conversion
        zc.grad = None
        ins, dom = wb.whitening_folded(zc, n, K)
        torch.autograd.backward([ins, dom], [one, one])

    out = {}
    for name, fn, bytes_per_pix in (("unfused", unfused, 704), ("fused", fused, 320), ("fused_channels_last", fused_cl, 320),
                                    ("plain_loss_channels_last", plain_cl, 192)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(iters):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) / iters * 1e3
        out[name] = {"us_per_step": us, "algorithmic_bytes_per_pix": bytes_per_pix,
                     "hbm_frac": bytes_per_pix * pix / (us * 1e-6) / 1e9 / peak}
        # per-kernel durations (CUDA events around each launch; they serialise the launches, so these are measured
        # outside the timed loop above)
        lib = wb._lib.load()
        lib.wtpse_profile_reset()
        lib.wtpse_profile_enable(1)
        for _ in range(3):
            fn()
        lib.wtpse_profile_enable(0)
        torch.cuda.synchronize()
        import ctypes
        kern = {}
        for kid in range(lib.wtpse_profile_kernel_count()):
            cnt, ms = ctypes.c_longlong(0), ctypes.c_double(0.0)
            wb._lib.check(lib.wtpse_profile_read(kid, ctypes.byref(cnt), ctypes.byref(ms)))
            if cnt.value:
                kern[lib.wtpse_profile_kernel_name(kid).decode()] = round(ms.value / cnt.value * 1e3, 1)
        out[name]["kernel_us"] = kern
    out["what"] = ("relu(z) + whitening loss fwd, and bwd with an upstream gradient on relu(z); unfused: 64 + 128 fwd, "
                   "128 + 192 + 192 bwd B/pix; fused: 128 fwd, 192 bwd")
    out["speedup"] = out["unfused"]["us_per_step"] / out["fused"]["us_per_step"]
    return out


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref, unmodified) on all host cores,
    SAME metric, SAME config, SAME workload as the GPU arm (32x16x512x512 is ~0.5 s per fwd+bwd on 8 cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    torch.set_num_threads(host_threads())         # torchrun exports OMP_NUM_THREADS=1
    B, H, W, n, K = workload_dims(args)
    step, kind, ctx = cpu_step_fn(n, K)
    z = synth_batch(B, H, W, seed=7)
    with ctx():
        for _ in range(max(1, min(args.warmup, 2))):
            step(z)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            losses = step(z)
        dt = time.perf_counter() - t0
    pix = B * H * W
    val = pix * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, H, W, n, K),
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": "%dx16x%dx%d per step (the full workload), %d steps" % (B, H, W, args.steps)},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "losses": list(losses),
    }
    print(json.dumps(line))


def pin_to_gpu_numa_node(gpu_index):
    """Bind this rank's threads (and therefore the first-touch placement of the pinned host buffers it allocates next)
    to the CPUs NVML reports as local to its GPU.  Returns a short description, or None when NVML cannot tell."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "%d CPUs local to GPU %d (NVML)" % (len(cpus), gpu_index)
    except Exception:
        return None


def time_link(dev, h_src, h_dst, world, barrier, reps=4):
    """Plain pinned cudaMemcpyAsync of one step's bytes: H2D alone, D2H alone, and both at once on two streams (what the
    e2e pipeline can reach at best).  GB/s per rank; at N > 1 all ranks copy at the same time (min over ranks)."""
    import torch
    import torch.distributed as dist

    nbytes = h_src.numel() * h_src.element_size()
    d_in = torch.empty(h_src.shape, device=dev)
    d_out = torch.randn(h_dst.shape, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_src, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    h_dst.copy_(d_out, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()
        dt = time.perf_counter() - t0
        return (int(h2d) + int(d2h)) * nbytes * reps / dt / 1e9

    run(True, True)                                   # warm-up
    res = {"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "bidir_gbs": run(True, True)}
    if world > 1:
        t = torch.tensor([res["h2d_gbs"], res["d2h_gbs"], res["bidir_gbs"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        res = {"h2d_gbs": float(t[0]), "d2h_gbs": float(t[1]), "bidir_gbs": float(t[2]),
               "note": "all %d ranks copying at the same time, min over ranks" % world}
    res["bytes_per_direction"] = nbytes
    return res


def time_track_r_configs(dev, peak, steps=20):
    """BASELINE configs[0] and configs[4] mapped onto Track R: device-resident fwd+bwd of the whitening/MMD
    loss at 8x16x256x256 (n=2, K=3: launch-bound, 100 MB -- no roofline claim) and 64x16x1024x1024 (n=21, K=3)."""
    import torch

    import wtpse_b200 as wb

    out = {}
    one = torch.ones((), device=dev)
    for name, B, S, n in (("8x16x256x256", 8, 256, 2), ("64x16x1024x1024", 64, 1024, 21)):
        nbuf = 2 if B * 16 * S * S * 4 > 130e6 else 8           # rotate through more than the 126 MB L2
        zs = [synth_batch(B, S, S, seed=500 + i, device=dev).requires_grad_(True) for i in range(nbuf)]

        def step(i):
            z = zs[i % nbuf]
            z.grad = None
            ins, dom = wb.whitening_folded(z, n, 3, 0.0, 1e-5)
            torch.autograd.backward([ins, dom], [one, one])
            return ins, dom
        for i in range(nbuf + 4):                                # every buffer's gradient block comes from the allocator cache
            step(i)
        torch.cuda.synchronize()

        def timed(fn):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for i in range(steps):
                r = fn(i)
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) / steps, r

        ms, (ins, dom) = timed(step)
        pix = B * S * S
        entry = {"ms_per_step": ms, "value": pix / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "steps": steps, "n_per_domain": n,
                 "launch_mode": "eager (two C-ABI calls per step from Python)",
                 "step_frac": 192.0 * pix / (ms * 1e-3) / 1e9 / peak if S >= 512 else None,
                 "note": None if S >= 512 else "100.7 MB per step = 15 us at the HBM peak: bound by launch and forward-tail latency (the single-SM tail alone is 10 us), not by HBM",
                 "l2": "%d alternating inputs of %.0f MB" % (nbuf, B * 16 * S * S * 4 / 1e6),
                 "losses": [float(ins.detach()), float(dom.detach())]}
        if S < 512:
            # a step this short is bound by the host's launch path: replay it as a CUDA graph (one graph per input buffer;
            # the C ABI neither allocates nor synchronises, so forward + backward capture as they are)
            try:
                # the captured step asks autograd for dz directly (torch.autograd.grad): the eager steps above left z's
                # AccumulateGrad node bound to the default stream, and running THAT node inside a capture on another stream
                # makes the engine synchronise with the legacy stream, which a capture forbids
                def gstep(i):
                    z = zs[i % nbuf]
                    ins, dom = wb.whitening_folded(z, n, 3, 0.0, 1e-5)
                    (dz,) = torch.autograd.grad([ins, dom], [z], [one, one])
                    return ins, dom, dz
                del ins, dom
                graphs, outs = [], []
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for i in range(nbuf):
                        gstep(i)
                torch.cuda.current_stream(dev).wait_stream(side)
                for i in range(nbuf):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        outs.append(gstep(i))
                    graphs.append(g)

                def replay(i):
                    graphs[i % nbuf].replay()
                    return outs[i % nbuf][:2]
                for i in range(nbuf):
                    replay(i)
                torch.cuda.synchronize()
                ms_g, (ins_g, dom_g) = timed(replay)
                entry.update({"eager_ms_per_step": ms, "ms_per_step": ms_g, "value": pix / (ms_g * 1e-3) / 1e6,
                              "step_frac": 192.0 * pix / (ms_g * 1e-3) / 1e9 / peak,
                              "launch_mode": "cuda-graph replay of forward + backward (eager: eager_ms_per_step)"})
                del graphs, outs
            except Exception as exc:
                entry["graph_capture_failed"] = str(exc).splitlines()[0][:160]
        out[name] = entry
        del zs
        torch.cuda.empty_cache()
    return out


def wavelet_case(dev, B, S, wv, J, steps=20, warmup=5, with_transform=False, debug=None):
    """One Track-W measurement: wavelet shape loss fwd+bwd on B x 2 x S x S softmax maps through the C ABI with preallocated
    outputs (the launches the autograd Function makes, without its Python bookkeeping).  PARITY UNPINNED."""
    import torch

    import wtpse_b200 as wb
    from wtpse_b200 import wavelet as wvm
    from wtpse_b200.functional import _ptr, _stream_ptr

    lib = wb._lib.load()
    C, H, W = 2, S, S
    wid = {"haar": 0, "db2": 1}[wv]
    cs = wvm.resident_cluster_size(H, W, wv, J)
    nin = max(4, int(300e6 // (B * C * H * W * 4)) + 1)          # the rotation is larger than the 126 MB L2
    nin = min(nin, 16)
    gen = torch.Generator(device=dev).manual_seed(4321)
    xs = [torch.softmax(3 * torch.randn(B, C, H, W, device=dev, generator=gen), 1) for _ in range(nin)]
    one = torch.ones((), device=dev)
    nmaps = B * C
    nbytes = lib.wtpse_wavelet_workspace_bytes(nmaps, H, W, J)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    loss_abi = torch.zeros((), device=dev)
    grad = torch.empty(B, C, H, W, device=dev)
    gcoef = None if cs else torch.empty(B, C, H, W, device=dev)
    st = _stream_ptr(dev)

    def step_abi(i):
        x = xs[i % nin]
        if cs:
            wb._lib.check(lib.wtpse_wavelet_loss_resident(_ptr(x), nmaps, H, W, wid, J, None, None, _ptr(loss_abi), _ptr(grad),
                                                          _ptr(ws), nbytes, st))
            wb._lib.check(lib.wtpse_scale_unless_one(_ptr(grad), grad.numel(), _ptr(one), st))
        else:
            wb._lib.check(lib.wtpse_wavelet_loss_forward(_ptr(x), nmaps, H, W, wid, J, None, _ptr(loss_abi), _ptr(gcoef), _ptr(ws),
                                                         nbytes, st))
            wb._lib.check(lib.wtpse_dwt2d_inverse(_ptr(gcoef), nmaps, H, W, wid, J, _ptr(grad), _ptr(one), _ptr(ws), nbytes, st))
        return loss_abi

    def timed(step):
        for i in range(max(warmup, 3)):
            step(i)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            out = step(i)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / steps, out

    ms, loss = timed(step_abi)
    elems = B * C * H * W
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else FALLBACK_PEAK_GBS
    res = {"us_per_step": ms * 1e3, "value": B * H * W / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "steps": steps,
           "frac": 8.0 * elems / (ms * 1e-3) / 1e9 / peak, "loss": float(loss),
           "path": ("fused plan, resident stage in clusters of %d CTAs" % cs) if cs else "one kernel per level",
           "l2": "%d alternating inputs of %.0f MB" % (nin, elems * 4 / 1e6)}
    if with_transform:
        coef = torch.empty(B, C, H, W, device=dev)
        back = torch.empty(B, C, H, W, device=dev)

        def step_fwd(i):
            wb._lib.check(lib.wtpse_dwt2d_forward(_ptr(xs[i % nin]), nmaps, H, W, wid, J, _ptr(coef), _ptr(ws), nbytes, st))
            return coef

        def step_inv(i):
            wb._lib.check(lib.wtpse_dwt2d_inverse(_ptr(coef), nmaps, H, W, wid, J, _ptr(back), None, _ptr(ws), nbytes, st))
            return back
        ms_fwd, _ = timed(step_fwd)
        ms_inv, _ = timed(step_inv)
        res["transform"] = {"dwt2d_us": ms_fwd * 1e3, "idwt2d_us": ms_inv * 1e3,
                            "max_abs_reconstruction_error": float((back - xs[(steps - 1) % nin]).abs().max()),
                            "frac_of_8B_per_element_roofline": [8.0 * elems / (t * 1e-3) / 1e9 / peak for t in (ms_fwd, ms_inv)]}
    return res


def time_wavelet_configs(dev, peak):
    """BASELINE configs[0], [1], [4] as literally written (Track W): the reference has no wavelet code, so every entry is
    parity UNPINNED -- checked against this repository's own float64 specification only (oracle/wavelet_np.py)."""
    out = {"parity": "unpinned", "metric": "wavelet shape-loss fwd+bwd, frac = 8 B/element (4N read + 4N gradient written) / time / HBM peak",
           "peak_gbs": peak}
    out["haar_J3_8x2x256x256"] = wavelet_case(dev, 8, 256, "haar", 3)
    out["haar_J3_8x2x256x256"]["note"] = "4.2 MB per step: launch-bound, no roofline claim"
    out["db2_J4_32x2x512x512"] = wavelet_case(dev, 32, 512, "db2", 4, with_transform=True)
    sweep = {}
    for wv in ("haar", "db2"):
        for J in range(1, 6):
            r = wavelet_case(dev, 64, 1024, wv, J, steps=10, warmup=3)
            sweep["%s_J%d" % (wv, J)] = {"us_per_step": r["us_per_step"], "frac": r["frac"], "value": r["value"], "path": r["path"]}
    out["sweep_64x2x1024x1024"] = sweep
    return out


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import wtpse_b200 as wb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, H, W, n, K = workload_dims(args)
    margin, eps = WORKLOAD["margin"], WORKLOAD["eps"]
    pix = B * H * W
    lib = wb._lib.load()
    affinity = pin_to_gpu_numa_node(physical_gpu_index(local_rank)) if args.numa_affinity else None
    for item in args.debug:
        name, _, val = item.partition("=")
        wb._lib.debug_set(name, int(val))

    # two resident input batches, alternated, each 4x the 126 MB L2 -> no timed step finds its input in L2
    zs = [synth_batch(B, H, W, seed=1234 + 17 * rank + i, device=dev).requires_grad_(True) for i in range(2)]

    one = torch.ones((), device=dev)

    def step(i):
        z = zs[i & 1]
        z.grad = None
        ins, dom = wb.whitening_folded(z, n, K, margin, eps)
        torch.autograd.backward([ins, dom], [one, one])       # unit upstream gradients, no extra kernels
        return ins, dom

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # NVML is initialised and the sampler thread started BEFORE the warm-up steps (both take tens of milliseconds, during which
    # an idle GPU drops its clocks): the warm-up then runs right up to the barrier that opens the timed region.
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    if args.launch_gate_ms > 0:
        try:
            torch.cuda._sleep(1000)          # loads the spin kernel of the launch gate (below) now, not between barrier and ev0
        except Exception:
            pass
    for i in range(max(args.warmup, 3)):
        step(i)
    if world > 1:
        # the first collective of a process group sets up NCCL's channels (hundreds of milliseconds, GPU idle): have it
        # here, then warm up again, so that the barrier that opens the timed region is a settled one on a busy GPU
        barrier()
        for i in range(3):
            step(i)
    # Per-kernel CUDA events are recorded inside the timed region on every `stride`-th step only: each
    # bracketed launch costs ~5 us of stream serialisation (measured: 318 us/step with events on every
    # launch vs 302 us/step with none), so sampling keeps the kernel timings "live" without taxing `value`.
    # The bracketed steps sit in the MIDDLE of each stride (never step 0: the first step after the barrier starts on an idle
    # GPU and is not what the other K - 1 look like), and a short run (the driver's K = 20) still brackets three of them.
    stride = 0 if args.no_kernel_events else max(1, args.event_stride)
    if stride and args.steps < 3 * stride:
        stride = max(1, args.steps // 3)
    phase = stride // 2
    lib.wtpse_profile_reset()
    lib.wtpse_profile_enable(0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()                    # the clocks reported are those sampled from here on (the timed region)
    # Device-side gate: a spin kernel of ~2 ms in front of ev0, so that the host has the first steps queued before the device
    # starts the timed K and runs them back to back.  A step costs the host ~70 us against 270 us on the device, so the
    # region is never host-bound in steady state -- but with K = 20 it is only 5.6 ms long, and one 0.7 ms hiccup of one rank's
    # host thread right after the barrier (scheduler, garbage collector; seen on one of eight ranks: 0.314 ms per step against
    # 0.278-0.282 on the other seven) became the max-over-ranks result.  The collector is off for the same reason.
    import gc
    gc_was_on = gc.isenabled()
    gc.disable()
    gate_ms = 0.0
    if args.launch_gate_ms > 0:
        try:
            torch.cuda._sleep(int(args.launch_gate_ms * 1e-3 * (sampler.max_mhz or 1900.0) * 1e6))
            gate_ms = float(args.launch_gate_ms)
        except Exception:
            gate_ms = 0.0
    ev0.record()
    for i in range(args.steps):
        if stride and i % stride == phase:
            lib.wtpse_profile_enable(1)
            ins, dom = step(i)
            lib.wtpse_profile_enable(0)
        else:
            ins, dom = step(i)
    ev1.record()
    if gc_was_on:
        gc.enable()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    losses = (float(ins.detach()), float(dom.detach()))

    launches = int(lib.wtpse_profile_launches(-1))
    import ctypes
    kern = {}
    for kid in range(lib.wtpse_profile_kernel_count()):
        cnt, ms = ctypes.c_longlong(0), ctypes.c_double(0.0)
        wb._lib.check(lib.wtpse_profile_read(kid, ctypes.byref(cnt), ctypes.byref(ms)))
        if cnt.value:
            kern[lib.wtpse_profile_kernel_name(kid).decode()] = {"launches": cnt.value, "avg_us": ms.value / cnt.value * 1e3}

    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    per_rank_ms = [ms_total / args.steps]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x.item()) / args.steps for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * pix / (ms_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI entry point ------------------------------------------------
    # Every step: its own pinned z goes H2D, forward, backward, dz + losses come back D2H.  Steps are submitted
    # back to back (wtpse_host_plan_submit); the plan's two device slots let one step's D2H overlap the next
    # step's H2D on the full-duplex link.  Two host buffer sets alternate (a step's buffers are reused only after
    # the plan has drained that slot).
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    z_host = [synth_batch(B, H, W, seed=99 + rank + i, pin=True) for i in range(2)]
    dz_host = [torch.empty(B, 16, H, W, pin_memory=True) for _ in range(2)]
    plan = wb.HostPlan(B, H, W)
    plan.run(z_host[0], n, K, margin, eps, (1.0, 1.0, 1.0), dz_host[0])       # warm-up
    barrier()
    t0 = time.perf_counter()
    outs = []
    for i in range(e2e_steps):
        outs.append(plan.submit(z_host[i & 1], n, K, margin, eps, (1.0, 1.0, 1.0), dz_host[i & 1]))
    plan.wait()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_losses = [float(outs[-1][3]), float(outs[-1][2])]
    plan.close()
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    nbytes = B * 16 * H * W * 4
    e2e = {"value": world * pix * e2e_steps / e2e_s / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": nbytes,
           "d2h_bytes_per_step": nbytes + 16, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "api": "wtpse_host_plan_submit/_wait (pinned host z -> H2D -> fwd -> bwd -> D2H dz + losses every step; "
                  "consecutive steps overlap their PCIe transfers)", "losses": e2e_losses}
    # the e2e leg's own roofline: plain pinned cudaMemcpyAsync of the same bytes, H2D and D2H at once on two streams
    link = time_link(dev, z_host[0], dz_host[0], world, barrier)
    e2e_rate = (2 * nbytes + 16) / (e2e_s / e2e_steps) / 1e9               # GB/s per rank, both directions
    e2e.update({"link_gbs": link["bidir_gbs"], "link_frac": e2e_rate / link["bidir_gbs"], "achieved_gbs": e2e_rate,
                "link": link, "cpu_affinity": affinity,
                "bound": "host link (PCIe): %.2f GB per step against %.2f ms of kernels" % ((2 * nbytes) / 1e9, ms_step)})
    del z_host, dz_host

    train = None
    if args.train_steps > 0:
        train = time_train_step(args, dev, rank, world, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline for the dominant kernel --------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md)"
    dom_name = max((k for k in kern if k in ALGO_BYTES_PER_PIX), key=lambda k: kern[k]["avg_us"], default=None)
    roofline = None
    if dom_name:
        achieved = ALGO_BYTES_PER_PIX[dom_name] * pix / (kern[dom_name]["avg_us"] * 1e-6) / 1e9
        traffic, stale, cap_hash = traffic_entry(dom_name)
        roofline = {"bound": "hbm", "kernel": dom_name, "event_stride": stride, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_stale": stale,
                    "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch), "
                                      "captured at kernel-source hash %s; this run's sources hash to %s" % (cap_hash, kernel_source_hash()),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PIX[dom_name] * pix,
                    "kernel_avg_us": kern[dom_name]["avg_us"]}
        other = "gram_tma_kernel" if dom_name == "apply_tma_kernel" else "apply_tma_kernel"
        if other in kern:
            a2 = ALGO_BYTES_PER_PIX[other] * pix / (kern[other]["avg_us"] * 1e-6) / 1e9
            roofline["also"] = {"kernel": other, "achieved": a2, "frac": a2 / peak, "kernel_avg_us": kern[other]["avg_us"]}
        roofline["step_frac"] = 192.0 * pix / (ms_step * 1e-3) / 1e9 / peak    # whole fwd+bwd step vs 192 B/pix

    cpu = None
    gpu_eager = None
    if not args.no_cpu_baseline and world == 1:
        cpu = time_cpu(B, H, W, n, K, args.cpu_seconds, 40, z=zs[(args.steps - 1) & 1].detach().cpu())
        cpu["losses_match_ours"] = bool(abs(cpu["losses"][0] - losses[0]) <= 1e-5 * abs(losses[0]) and abs(cpu["losses"][1] - losses[1]) <= 1e-5)
        try:
            gpu_eager = time_gpu_eager(zs[(args.steps - 1) & 1], n, K)         # the input whose losses the line reports
            gpu_eager["losses_match_ours"] = bool(abs(gpu_eager["losses"][0] - losses[0]) <= 1e-5 * abs(losses[0]) and
                                                  abs(gpu_eager["losses"][1] - losses[1]) <= 1e-4)
        except Exception as exc:                          # e.g. out of memory on a smaller device: report, do not hide
            gpu_eager = {"unavailable": str(exc).splitlines()[0][:160]}

    relu_fusion = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            relu_fusion = time_relu_fusion(zs[0], n, K, peak)
        except Exception as exc:
            relu_fusion = {"unavailable": str(exc).splitlines()[0][:160]}

    configs = wavelet = None
    if world == 1 and args.extra_configs:
        del zs
        torch.cuda.empty_cache()
        try:
            configs = time_track_r_configs(dev, peak)
        except Exception as exc:
            configs = {"unavailable": str(exc).splitlines()[0][:160]}
        torch.cuda.empty_cache()
        try:
            wavelet = time_wavelet_configs(dev, peak)
        except Exception as exc:
            wavelet = {"unavailable": str(exc).splitlines()[0][:160]}

    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, H, W, n, K),
        "ms_per_step_per_rank": per_rank_ms,          # ms_per_step is their maximum (every rank times its own K steps on its device)
        "launch_gate_ms": gate_ms,                    # device-side spin between the opening barrier and the first CUDA event (not timed)
        "e2e": e2e, "gpu_launches": launches, "kernels": kern, "roofline": roofline, "cpu_baseline": cpu,
        "gpu_eager_baseline": gpu_eager, "clocks": clocks, "losses": losses, "train_step": train,
        "relu_fusion": relu_fusion, "configs": configs, "wavelet": wavelet,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def time_train_step(args, dev, rank, world, barrier):
    """BASELINE configs[2]/[3]: full WT-PSE iteration (4 network updates: reference U-Net backbone in PyTorch/cuDNN
    + the CUDA shape-loss path) on synthetic fundus-shaped batches, data-parallel over the ranks with one bucketed
    NCCL all-reduce per backward.  Reported as images/s over all ranks; loader time (device-side generator +
    bit-exact label kernel) is inside the timed region.

    `value` skips the reference's dead teacher backward in the two shape updates (its gradients are zeroed unread,
    Trainer.py:768; weights and losses unchanged, tests/test_gpu_update.py); `with_teacher_backward` is the same
    iteration doing exactly the reference's work."""
    import gc

    import torch

    res = _time_train_variant(args, dev, rank, world, barrier, teacher_backward=False, steps=args.train_steps,
                              fuse_relu=bool(args.train_fuse_relu))
    res["fused_deepwt_tail"] = bool(args.train_fuse_relu)
    res["dead_teacher_backward"] = "skipped (SURVEY 8(f).3); weights, losses and BatchNorm statistics unchanged"
    if args.train_reference_work:
        gc.collect()
        torch.cuda.empty_cache()
        ref = _time_train_variant(args, dev, rank, world, barrier, teacher_backward=True, steps=max(2, args.train_steps // 2))
        res["with_teacher_backward"] = {k: ref[k] for k in ("value", "unit", "ms_per_step", "steps", "launch_mode")}
    if world > 1 and args.train_segments_compare and args.train_grad_segments > 1:
        gc.collect()
        torch.cuda.empty_cache()
        alt = _time_train_variant(args, dev, rank, world, barrier, teacher_backward=False, steps=max(2, args.train_steps // 2),
                                  fuse_relu=bool(args.train_fuse_relu), grad_segments=1)
        res["one_allreduce_after_backward"] = {k: alt[k] for k in ("value", "unit", "ms_per_step", "steps", "launch_mode", "grad_allreduce")}
        gc.collect()
        torch.cuda.empty_cache()
        # scaling diagnostic: the same N processes WITHOUT any gradient exchange.  max - min over ranks is how much the
        # GPUs / processes differ on their own (clocks, cuDNN's per-process algorithm choice); what the all-reduce variants
        # lose beyond that is the collectives' cost.  Not a valid training configuration, never the reported value.
        alt = _time_train_variant(args, dev, rank, world, barrier, teacher_backward=False, steps=max(2, args.train_steps // 2),
                                  fuse_relu=bool(args.train_fuse_relu), grad_allreduce=False)
        res["replicas_without_allreduce"] = {"ms_per_step_slowest_rank": alt["ms_per_step"],
                                             "ms_per_step_fastest_rank": alt["ms_per_step_fastest_rank"], "steps": alt["steps"],
                                             "what": "diagnostic: N independent replicas, no gradient exchange"}
    if args.train_fuse_compare:
        gc.collect()
        torch.cuda.empty_cache()
        alt = _time_train_variant(args, dev, rank, world, barrier, teacher_backward=False, steps=max(2, args.train_steps // 2),
                                  fuse_relu=not args.train_fuse_relu)
        key = "with_fused_deepwt_tail" if not args.train_fuse_relu else "without_fused_deepwt_tail"
        res[key] = {k: alt[k] for k in ("value", "unit", "ms_per_step", "steps", "launch_mode", "our_kernels_per_iteration")}
    if args.train_reference_eager and world == 1:
        gc.collect()
        torch.cuda.empty_cache()
        try:
            res["reference_eager"] = time_reference_train(args, dev, rank)
            res["speedup_vs_reference_eager"] = res["value"] / res["reference_eager"]["value"]
            if "dropin_installed" in res["reference_eager"]:
                res["reference_eager"]["dropin_installed"]["speedup"] = (res["reference_eager"]["dropin_installed"]["value"] /
                                                                        res["reference_eager"]["value"])
        except Exception as exc:
            res["reference_eager"] = {"unavailable": str(exc).splitlines()[0][:200]}
    return res


def time_reference_train(args, dev, rank):
    """The reference's OWN training iteration on the same GPU: the unmodified algorithms.WT_PSE / shape_networks.
    ShapeVariationalDist_x classes (oracle/_ref) driven in Trainer.train_epoch's order with its per-loss .item() host
    syncs (oracle/ref_iteration.py restates Trainer.py:762-925; Trainer.py itself needs packages this image lacks),
    PyTorch eager, torch defaults (NCHW, cuDNN with TF32 convolutions allowed, cudnn.deterministic=True / benchmark=False
    as utils.seed_initialization sets them), same synthetic batches, same batch arithmetic (15 of 16 images used).
    Also timed: the same reference classes with wtpse_b200.dropin.install() underneath -- the drop-in as a user gets it."""
    import torch

    import wtpse_b200 as wb
    from oracle import ref_iteration as ri
    from oracle import ref_shim

    if not ref_shim.available():
        return {"unavailable": "oracle/_ref missing (run __graft_entry__.build() where /root/reference exists)"}
    alg, sn, _ = ref_shim.load()
    hp = dict(ref_shim.DEFAULT_HPARAMS)
    n_per_domain, used = wb.dp.per_rank_batch(args.train_batch, 1, 3)
    S = args.train_size
    old = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False          # utils.py:58-65
    steps = max(2, args.train_steps // 2)

    def batch(it):
        return wb.synthetic.fundus_batch(n_per_domain, 3, S, S, dev, seed=wb.dp.rank_batch_seed(1, rank, it))

    def run(install):
        nets, optims = ri.build_reference_models(alg, sn, hp, n_per_domain, 3, dev, seed=0)
        saved = wb.dropin.install(alg, sn) if install else None
        try:
            for it in range(2):
                ri.trainer_iteration(nets, optims, *batch(it), hp, epoch=0, host_syncs=True)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for it in range(steps):
                out = ri.trainer_iteration(nets, optims, *batch(3 + it), hp, epoch=0, host_syncs=True)
            ev1.record()
            torch.cuda.synchronize()
        finally:
            if saved:
                wb.dropin.uninstall(saved)
        ms = ev0.elapsed_time(ev1) / steps
        del nets, optims
        return {"value": used / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "steps": steps,
                "losses": {k: float(v) for k, v in out.items() if torch.is_tensor(v) and v.dim() == 0}}

    try:
        res = run(False)
        res["kind"] = ("reference: unmodified algorithms.WT_PSE / shape_networks.ShapeVariationalDist_x (oracle/_ref), "
                       "Trainer.train_epoch order with its .item() syncs, PyTorch eager CUDA, same GPU")
        res["image_size"], res["per_gpu_batch_used"] = S, used
        torch.cuda.empty_cache()
        res["dropin_installed"] = run(True)
        res["dropin_installed"]["what"] = "same reference classes and loop with wtpse_b200.dropin.install(algorithms, shape_networks)"
    finally:
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = old
    return res


def _time_train_variant(args, dev, rank, world, barrier, teacher_backward, steps, fuse_relu=False, grad_segments=None,
                        grad_allreduce=True):
    import torch
    import torch.distributed as dist

    import wtpse_b200 as wb

    # fixed shapes: let cuDNN time its algorithms once (during the eager warm-up iterations, before the graph capture)
    torch.backends.cudnn.benchmark = bool(args.train_cudnn_benchmark)
    n_per_domain, used = wb.dp.per_rank_batch(args.train_batch * world, world, 3)
    S = args.train_size
    grad_segments = args.train_grad_segments if grad_segments is None else grad_segments
    ts = wb.TrainStep(n_per_domain=n_per_domain, n_domains=3, device=dev, seed=0, teacher_backward=teacher_backward,
                      fuse_relu=fuse_relu, grad_segments=grad_segments, grad_allreduce=grad_allreduce)
    lib = wb._lib.load()

    mode = "eager"

    def batch(it):
        return wb.synthetic.fundus_batch(n_per_domain, 3, S, S, dev, seed=wb.dp.rank_batch_seed(1, rank, it))

    def one(it):
        image, od, oc = batch(it)
        return ts.replay(image, od, oc) if mode == "cuda-graph" else ts.step(image, od, oc)

    lib.wtpse_profile_reset()
    for it in range(3):
        one(it)
    kernels_per_iteration = int(lib.wtpse_profile_launches(-1)) // 3          # counted on the eager warm-up iterations
    if args.train_graph:
        try:                                              # the whole iteration as one CUDA graph (no host syncs on the path)
            ts.capture(*batch(100))
            mode = "cuda-graph"
            one(101)
        except Exception as exc:                          # report, do not hide: the eager path is still the product
            mode = "eager (graph capture failed: %s)" % str(exc).splitlines()[0][:120]
    barrier()
    lib.wtpse_profile_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for it in range(steps):
        out = one(3 + it)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tmin = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    ms = float(t.item()) / steps
    return {"metric": "train images/s", "value": world * used / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
            "ms_per_step_fastest_rank": float(tmin.item()) / steps,
            "steps": steps, "launch_mode": mode, "image_size": S, "per_gpu_batch_nominal": args.train_batch, "per_gpu_batch_used": used,
            "global_batch_used": world * used, "our_kernels_per_iteration": kernels_per_iteration,
            "backbone": "PyTorch/cuDNN (benchmark=%d), channels-last weights, fp32 storage (torch-default TF32 convs), fused Adam" % int(args.train_cudnn_benchmark), "grad_allreduce": ("NCCL all-reduce(AVG) of each network's flat gradient buffer in %d segment(s)%s" % (
                max(len(b.segments) for b in ts.buckets), ", started from autograd hooks during the backward pass (async, own stream)"
                if grad_segments > 1 else " after the backward pass")) if world > 1 else None,
            "losses": {k: float(v) for k, v in out.items()}}


def run_wavelet(args):
    """Track W side bench: db2 J=4 L1 detail loss fwd+bwd on 32x2x512x512 maps (BASELINE configs[1] verbatim).
    PARITY UNPINNED (no wavelet code in the reference); reported for completeness, never the headline."""
    import torch

    import wtpse_b200 as wb

    import torch.distributed as dist

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:                           # independent maps: every rank its own batch, no data-path collective
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)
    B, C, H, W, J, wv = args.wavelet_batch, 2, args.size, args.size, args.wavelet_levels, args.wavelet_name
    wid = {"haar": 0, "db2": 1}[wv]
    from wtpse_b200 import wavelet as wvm
    from wtpse_b200.functional import _ptr, _stream_ptr

    lib = wb._lib.load()
    wb._lib.debug_set("wavelet_resident", int(args.wavelet_resident))
    wb._lib.debug_set("wavelet_split", int(args.wavelet_split))
    wb._lib.debug_set("wavelet_cluster_max", int(args.wavelet_cluster_max))
    wb._lib.debug_set("wavelet_tiles", int(args.wavelet_tiles))
    wb._lib.debug_set("wavelet_peel_max", int(args.wavelet_peel_max))
    for item in args.debug:
        name, _, val = item.partition("=")
        wb._lib.debug_set(name, int(val))
    cs = wvm.resident_cluster_size(H, W, wv, J)
    xs = [torch.softmax(3 * torch.randn(B, C, H, W, device=dev), 1).requires_grad_(True) for _ in range(4)]
    one = torch.ones((), device=dev)

    def step_autograd(i):                   # the public API: autograd Function, fresh output tensors every call
        x = xs[i % 4]
        x.grad = None
        loss = wb.wavelet_shape_loss(x, wv, J)
        loss.backward(one)
        return loss

    # the same launches through the C ABI with preallocated outputs (no Python autograd bookkeeping between them)
    nmaps = B * C
    nbytes = lib.wtpse_wavelet_workspace_bytes(nmaps, H, W, J)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    loss_abi = torch.zeros((), device=dev)
    grad = torch.empty(B, C, H, W, device=dev)
    gcoef = torch.empty(B, C, H, W, device=dev)
    st = _stream_ptr(dev)

    def step_abi(i):
        x = xs[i % 4]
        if cs:
            wb._lib.check(lib.wtpse_wavelet_loss_resident(_ptr(x), nmaps, H, W, wid, J, None, None, _ptr(loss_abi), _ptr(grad),
                                                          _ptr(ws), nbytes, st))
            wb._lib.check(lib.wtpse_scale_unless_one(_ptr(grad), grad.numel(), _ptr(one), st))
        else:
            wb._lib.check(lib.wtpse_wavelet_loss_forward(_ptr(x), nmaps, H, W, wid, J, None, _ptr(loss_abi), _ptr(gcoef), _ptr(ws),
                                                         nbytes, st))
            wb._lib.check(lib.wtpse_dwt2d_inverse(_ptr(gcoef), nmaps, H, W, wid, J, _ptr(grad), _ptr(one), _ptr(ws), nbytes, st))
        return loss_abi

    def timed(step):
        for i in range(max(args.warmup, 3)):
            step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            out = step(i)
        ev1.record()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)        # max over ranks, on the device
        return float(t), out

    ms_autograd, loss_ag = timed(step_autograd)
    ms, loss = timed(step_abi)

    # the plain transform pair (W1 / W2): coefficients in Mallat layout and back, through the C ABI
    coef = torch.empty(B, C, H, W, device=dev)
    back = torch.empty(B, C, H, W, device=dev)

    def step_fwd(i):
        wb._lib.check(lib.wtpse_dwt2d_forward(_ptr(xs[i % 4]), nmaps, H, W, wid, J, _ptr(coef), _ptr(ws), nbytes, st))
        return coef

    def step_inv(i):
        wb._lib.check(lib.wtpse_dwt2d_inverse(_ptr(coef), nmaps, H, W, wid, J, _ptr(back), None, _ptr(ws), nbytes, st))
        return back

    ms_fwd, _ = timed(step_fwd)
    ms_inv, _ = timed(step_inv)
    recon_err = float((back - xs[(args.steps - 1) % 4].detach()).abs().max())
    assert abs(float(loss) - float(loss_ag)) <= 1e-6 * abs(float(loss_ag)), (float(loss), float(loss_ag))
    elems = B * C * H * W
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else FALLBACK_PEAK_GBS
    algo = 8.0 * elems                      # SURVEY 8(d) Track W: 4N read forward + 4N written backward
    # bytes the kernels actually move: resident = one read + one write; per-level = (read + write) * 4/3 per pass, two passes
    moved = algo if cs else 4.0 * elems * (8.0 / 3.0) * 2
    traffic, traffic_stale = None, None
    if cs and (B, C, H, W, J, wv) == (32, 2, 512, 512, 4, "db2"):
        traffic, traffic_stale, _ = traffic_entry("wavelet_fused_step_32x2x512x512_db2_J4", "wavelet")
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    print(json.dumps({
        "metric": "wavelet shape-loss fwd+bwd Mpix/s (Track W, parity unpinned)", "value": world * B * H * W / (ms * 1e-3) / 1e6,
        "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s DWT J=%d L1 detail-coefficient loss fwd+bwd, %dx%dx%dx%d softmax maps" % (wv, J, B, C, H, W),
                   "l2": "four alternating 67 MB inputs (each below the 126 MB L2: the 268 MB rotation is not)",
                   "path": ("fused plan (wavelet-split=%d), resident stage in clusters of %d CTAs" % (args.wavelet_split, cs)) if cs else "one kernel per level",
                   "timed_through": "C ABI, preallocated outputs"},
        "autograd_ms_per_step": ms_autograd,
        "transform": {"dwt2d_ms": ms_fwd, "idwt2d_ms": ms_inv, "max_abs_reconstruction_error": recon_err,
                      "frac_of_8B_per_element_roofline": [8.0 * B * C * H * W / (t * 1e-3) / 1e9 / (float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else FALLBACK_PEAK_GBS) for t in (ms_fwd, ms_inv)]},
        "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "per": "GPU",
                     "frac": algo / (ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_stale": traffic_stale,
                     "moved_estimate_frac": moved / (ms * 1e-3) / 1e9 / peak},
        "loss": float(loss.detach())}))


def main():
    args = parse_args()
    if args.track == "wavelet" and args.impl != "reference":
        return run_wavelet(args)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
