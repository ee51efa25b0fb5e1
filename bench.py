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
  cpu_baseline  the oracle's PyTorch-CPU port of the reference path on this box's host cores (bounded sample)

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
    ap.add_argument("--train-steps", type=int, default=6, help="timed iterations of the full WT-PSE train step (0 = skip)")
    ap.add_argument("--train-size", type=int, default=512)
    ap.add_argument("--train-graph", type=int, default=1, help="replay the train iteration as a CUDA graph")
    ap.add_argument("--train-batch", type=int, default=16, help="nominal per-GPU batch (the reference uses 3 * (batch // 3))")
    ap.add_argument("--no-kernel-events", action="store_true", help="diagnostic: timed region without per-kernel CUDA events")
    ap.add_argument("--debug-backward-mode", type=int, default=0)
    ap.add_argument("--debug-round-robin", type=int, default=1)
    ap.add_argument("--debug-gram-group", type=int, default=-1)
    ap.add_argument("--debug-l2-hint", type=int, default=-1)
    ap.add_argument("--debug-gram-variant", type=int, default=-1)
    ap.add_argument("--debug-two-stage-epilogue", type=int, default=1)
    ap.add_argument("--event-stride", type=int, default=8, help="bracket kernels with CUDA events on every n-th timed step")
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
    """Returns (callable(z) -> None doing one fwd+bwd, kind)."""
    import torch
    from oracle import ref_shim

    if ref_shim.available() and not torch.cuda.is_available():
        # build container: the unmodified reference through the import shim
        alg, _, _ = ref_shim.load()
        hp = dict(ref_shim.DEFAULT_HPARAMS)
        torch.manual_seed(0)
        model = alg.WT_PSE(3, 1, hp, "cpu", False, per_domain_batch=n, source_domain_num=K)

        def step(z):
            z = z.detach().requires_grad_(True)
            ins, dom = model.compute_whitening_loss(z)
            (ins + dom).backward()
        return step, "reference"
    from oracle import whitening_torch as wt

    def step(z):
        wt.fwd_bwd(z, n, K)
    return step, "port"


def time_cpu(B, H, W, n, K, budget_s, max_iters):
    import torch

    step, kind = cpu_step_fn(n, K)
    z = synth_batch(B, H, W, seed=7)
    step(z)                                     # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_start) < budget_s:
        t0 = time.perf_counter()
        step(z)
        times.append(time.perf_counter() - t0)
    pix = B * H * W
    return {"value": pix / min(times) / 1e6, "mean_value": pix / (sum(times) / len(times)) / 1e6, "unit": "Mpix/s",
            "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": kind, "iters": len(times),
            "sample": "%dx16x%dx%d fp32 (n=%d,K=%d) fwd+bwd, best of %d after 1 warm-up" % (B, H, W, n, K, len(times))}


def time_gpu_eager(z, n, K, iters=10):
    """The reference's operator sequence (oracle/whitening_torch.py: bmm + ATen element-wise + the O(K^2) MMD loop, i.e.
    what the unmodified reference launches on a GPU) in PyTorch eager on the same device and the same resident input.
    A reported baseline beside cpu_baseline -- test infrastructure, never the product path."""
    import torch
    from oracle import whitening_torch as wt

    for _ in range(2):
        wt.fwd_bwd(z, n, K)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(iters):
        ins, dom, _ = wt.fwd_bwd(z, n, K)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    return {"value": z.shape[0] * z.shape[2] * z.shape[3] / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": ms, "iters": iters,
            "kind": "port: reference operator sequence in PyTorch eager (CUDA), same GPU, same device-resident input",
            "losses": [float(ins), float(dom)]}


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
    out["what"] = ("relu(z) + whitening loss fwd, and bwd with an upstream gradient on relu(z); unfused: 64 + 128 fwd, "
                   "128 + 192 + 192 bwd B/pix; fused: 128 fwd, 192 bwd")
    out["speedup"] = out["unfused"]["us_per_step"] / out["fused"]["us_per_step"]
    return out


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread this process may run on
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    H = args.size
    n, K = 2, 3
    B = n * K                                   # bounded sample: 6 of the workload's samples per step
    step, kind = cpu_step_fn(n, K)
    z = synth_batch(B, H, H, seed=7)
    for _ in range(max(1, min(args.warmup, 3))):
        step(z)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(z)
    dt = time.perf_counter() - t0
    pix = B * H * H
    val = pix * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": "shape-loss fwd+bwd Mpix/s", "value": val, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "whitening+MMD shape loss fwd+bwd, %dx16x%dx%d fp32, n=%d K=%d (bounded sample of the "
                               "32x16x512x512 workload, CPU)" % (B, H, H, n, K)},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": "%dx16x%dx%d per step, %d steps" % (B, H, H, args.steps)},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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

    B, H, W = args.batch, args.size, args.size
    n = max(1, B // 3) if B != WORKLOAD["B"] else WORKLOAD["n_per_domain"]
    K = WORKLOAD["n_domains"]
    margin, eps = WORKLOAD["margin"], WORKLOAD["eps"]
    pix = B * H * W
    lib = wb._lib.load()
    lib.wtpse_debug_set_backward_mode(args.debug_backward_mode)
    lib.wtpse_debug_set_apply_round_robin(args.debug_round_robin)
    if args.debug_gram_variant >= 0:
        lib.wtpse_debug_set_gram_variant(args.debug_gram_variant)
    if args.debug_l2_hint >= 0:
        lib.wtpse_debug_set_l2_hint(args.debug_l2_hint)
    if args.debug_gram_group >= 0:
        lib.wtpse_debug_set_gram_group(args.debug_gram_group)
    lib.wtpse_debug_set_two_stage_epilogue(args.debug_two_stage_epilogue)

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

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    # Per-kernel CUDA events are recorded inside the timed region on every `stride`-th step only: each
    # bracketed launch costs ~5 us of stream serialisation (measured: 318 us/step with events on every
    # launch vs 302 us/step with none), so sampling keeps the kernel timings "live" without taxing `value`.
    stride = 0 if args.no_kernel_events else max(1, args.event_stride)
    lib.wtpse_profile_reset()
    lib.wtpse_profile_enable(0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record()
    for i in range(args.steps):
        if stride and i % stride == 0:
            lib.wtpse_profile_enable(1)
            ins, dom = step(i)
            lib.wtpse_profile_enable(0)
        else:
            ins, dom = step(i)
    ev1.record()
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
    if world > 1:
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
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom_name)
        roofline = {"bound": "hbm", "kernel": dom_name, "event_stride": stride, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
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
        cpu = time_cpu(6, H, W, 2, 3, args.cpu_seconds, 40)
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

    line = {
        "metric": "shape-loss fwd+bwd Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "whitening+MMD shape loss fwd+bwd (WT_PSE.compute_whitening_loss), z=%dx16x%dx%d fp32, "
                               "n=%d K=%d per GPU [BASELINE configs[1] mapped per SURVEY 8(d)]" % (B, H, W, n, K),
                   "per_gpu_batch": B, "l2": "inputs (%.0f MB, two alternating buffers) larger than L2" % (nbytes / 1e6),
                   "sharding": "independent [K x n] batches per rank, no data-path collective"},
        "e2e": e2e, "gpu_launches": launches, "kernels": kern, "roofline": roofline, "cpu_baseline": cpu,
        "gpu_eager_baseline": gpu_eager, "clocks": clocks, "losses": losses, "train_step": train,
        "relu_fusion": relu_fusion,
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
    if args.train_fuse_compare:
        gc.collect()
        torch.cuda.empty_cache()
        alt = _time_train_variant(args, dev, rank, world, barrier, teacher_backward=False, steps=max(2, args.train_steps // 2),
                                  fuse_relu=not args.train_fuse_relu)
        key = "with_fused_deepwt_tail" if not args.train_fuse_relu else "without_fused_deepwt_tail"
        res[key] = {k: alt[k] for k in ("value", "unit", "ms_per_step", "steps", "launch_mode", "our_kernels_per_iteration")}
    return res


def _time_train_variant(args, dev, rank, world, barrier, teacher_backward, steps, fuse_relu=False):
    import torch
    import torch.distributed as dist

    import wtpse_b200 as wb

    # fixed shapes: let cuDNN time its algorithms once (during the eager warm-up iterations, before the graph capture)
    torch.backends.cudnn.benchmark = bool(args.train_cudnn_benchmark)
    n_per_domain, used = wb.dp.per_rank_batch(args.train_batch * world, world, 3)
    S = args.train_size
    ts = wb.TrainStep(n_per_domain=n_per_domain, n_domains=3, device=dev, seed=0, teacher_backward=teacher_backward,
                      fuse_relu=fuse_relu)
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
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"metric": "train images/s", "value": world * used / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
            "steps": steps, "launch_mode": mode, "image_size": S, "per_gpu_batch_nominal": args.train_batch, "per_gpu_batch_used": used,
            "global_batch_used": world * used, "our_kernels_per_iteration": kernels_per_iteration,
            "backbone": "PyTorch/cuDNN (benchmark=%d), channels-last weights, fp32 storage (torch-default TF32 convs), fused Adam" % int(args.train_cudnn_benchmark), "grad_allreduce": "NCCL, 1 bucket per backward" if world > 1 else None,
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
    lib.wtpse_debug_set_wavelet_resident(int(args.wavelet_resident))
    lib.wtpse_debug_set_wavelet_split(int(args.wavelet_split))
    lib.wtpse_debug_set_wavelet_cluster_max(int(args.wavelet_cluster_max))
    lib.wtpse_debug_set_wavelet_tiles(int(args.wavelet_tiles))
    lib.wtpse_debug_set_wavelet_peel_max(int(args.wavelet_peel_max))
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
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if cs and os.path.exists(tpath) and (B, C, H, W, J, wv) == (32, 2, 512, 512, 4, "db2"):
        traffic = json.load(open(tpath)).get("wavelet_fused_step_32x2x512x512_db2_J4")
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
                     "frac": algo / (ms * 1e-3) / 1e9 / peak, "traffic": traffic,
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
