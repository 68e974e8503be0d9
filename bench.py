#!/usr/bin/env python
"""Headline benchmark: SR train images/s (forward + loss + backward + Adam) on synthetic
Food101-shaped crops, BASELINE.json config C2 (ResNet-SR 16 blocks x 64 ch, x4, 64->256, batch 64 per
GPU, NLPD loss), one process per GPU, weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl srk|reference] [--dtype bf16|fp32]

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement").  `--impl reference` times the CPU port of
the reference path (oracle/sr_oracle.py: the same ATen calls the reference modules make) on the host
cores with all threads, on a bounded sample of the same workload."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# headline workload = BASELINE config C2; C1 / C3 are selectable for secondary measurements (--config)
CONFIGS = {
    "C2": dict(arch="RESNET", loss="nlpd", lr_hw=64, scale=4, batch=64, gflop=54.39,
               name="C2: ResNet-SR 16x64ch x4, 64->256 synthetic crops, batch 64/GPU, NLPD loss, fwd+bwd+Adam"),
    "C3": dict(arch="AttentionSR", loss="mae", lr_hw=64, scale=4, batch=32, gflop=158.9,
               name="C3: AttentionSR 32x96ch x4, 64->256 synthetic crops, batch 32/GPU, MAE loss, fwd+bwd+Adam"),
    "C1": dict(arch="SRCNN", loss="nlpd", lr_hw=128, scale=2, batch=16, gflop=7.57,
               name="C1: SRCNN 9-1-5 x2, 128->256 synthetic crops, batch 16/GPU, NLPD loss, fwd+bwd+Adam"),
}
ARCH = "RESNET"
LOSS = "nlpd"
LR_HW = 64
SCALE = 4
BATCH_PER_GPU = 64
FWD_BWD_GFLOP_PER_IMG = 54.39  # 2*MAC over convs, fwd + dgrad + wgrad (no input dgrad)
WORKLOAD = CONFIGS["C2"]["name"]


def select_config(name):
    global ARCH, LOSS, LR_HW, SCALE, BATCH_PER_GPU, FWD_BWD_GFLOP_PER_IMG, WORKLOAD
    c = CONFIGS[name]
    ARCH, LOSS, LR_HW, SCALE, BATCH_PER_GPU = c["arch"], c["loss"], c["lr_hw"], c["scale"], c["batch"]
    FWD_BWD_GFLOP_PER_IMG, WORKLOAD = c["gflop"], c["name"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx, pw = [], set(), None, []
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                pw.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


def shutdown_distributed(*holders):
    """Captured CUDA graphs that contain NCCL kernels keep the communicator busy: destroy_process_group() then waits
    for ever (observed with torch 2.11 / NCCL 2.28).  Drop the graphs first, and never let the teardown outlive the
    measurement: after 20 s the process exits on its own (the JSON line has been printed by then)."""
    import gc
    import threading
    import torch.distributed as dist
    for h in holders:
        if h is not None and hasattr(h, "_graphs"):
            h._graphs.clear()
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    t = threading.Timer(20.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------------------------------
class _CpuReference:
    """The reference's own CPU training step: its unmodified nn.Modules (oracle/_ref: byte copies of the reference's
    src/models.py and src/loss.py, see oracle/build_ref.py) driven the way reference train.py:53-56,116-120 drives
    them - get_model, get_loss_function, optim.Adam(betas=(0.5, 0.999)), zero_grad / forward / loss / backward / step.
    Falls back to the pinned port (oracle/sr_oracle.py, the same ATen calls) when the copy is not in the tree."""

    def __init__(self, batch):
        from oracle import ref_modules
        from oracle import sr_oracle as O
        self.O, self.batch = O, batch
        torch.manual_seed(0)
        self.lr, self.hr = O.synthetic_pair(batch, LR_HW, LR_HW, SCALE)
        if ref_modules.available():
            ref = ref_modules.load()
            self.kind = "reference"
            self.model = ref.models.get_model(ARCH, scale_factor=SCALE, device="cpu").train()
            self.crit = ref.loss.get_loss_function(LOSS, "cpu")
            self.opt = torch.optim.Adam(self.model.parameters(), lr=4e-4, betas=(0.5, 0.999))
        else:
            self.kind = "port"
            from src.models import get_model   # constructors only (seeded init); no libsrk kernel runs on CPU tensors
            sd = {k: v.clone() for k, v in get_model(ARCH, SCALE, "cpu").state_dict().items()}
            self.params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
                           if v.is_floating_point() and "running_" not in k}
            self.work = dict(sd)
            self.work.update(self.params)
            self.opt = torch.optim.Adam(list(self.params.values()), lr=4e-4, betas=(0.5, 0.999))

    def step(self):
        self.opt.zero_grad()
        if self.kind == "reference":
            loss = self.crit(self.model(self.lr), self.hr)
        else:
            out = self.O.model_forward(ARCH, self.work, self.lr, training=True, scale_factor=SCALE)
            loss = self.O.loss_fn(LOSS)(out, self.hr)
        loss.backward()
        self.opt.step()
        return loss.item()


def _cpu_probe_rate(threads):
    """images/s of the CPU reference step at a small batch (one warm-up + one timed step)."""
    ref = _CpuReference(min(4, BATCH_PER_GPU))
    ref.step()
    t0 = time.perf_counter()
    ref.step()
    return ref.batch / (time.perf_counter() - t0)


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation on the host cores (rank 0 only), all threads, same
    workload definition; the per-step batch is the config's own when (steps + warmup) of it fit ~150 s of CPU time,
    else the largest batch that does (stated in `sample`)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = host_threads()
    rate = _cpu_probe_rate(threads)
    budget_s = float(os.environ.get("SRK_REF_BUDGET_S", "150"))
    fit = int(rate * budget_s / max(args.steps + args.warmup, 1))
    sample_batch = args.cpu_batch if args.cpu_batch > 0 else max(1, min(args.batch, fit))
    ref = _CpuReference(sample_batch)
    for _ in range(args.warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    dt = (time.perf_counter() - t0) / args.steps
    v = sample_batch / dt
    what = "unmodified reference modules (oracle/_ref)" if ref.kind == "reference" else "port oracle/sr_oracle.py"
    sample = "batch %d of %d per step, %s, fp32, torch CPU ATen ops, %d threads" % (sample_batch, args.batch, what, threads)
    line = {"impl": "reference", "metric": "sr_train_images_per_sec", "value": round(v, 3), "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample, "sample_batch": sample_batch},
            "cpu_baseline": {"value": round(v, 3), "unit": "images/s", "cores": threads, "kind": ref.kind, "sample": sample},
            "e2e": {"value": round(v, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def host_threads():
    """Threads of the CPU legs.  torch's own default (one per physical core) is kept; torchrun, however, exports
    OMP_NUM_THREADS=1 to its workers, which would turn the reference arm of an N > 1 run into a single-threaded
    measurement: in that case the physical cores this process may run on are used, as in the N = 1 run."""
    if torch.get_num_threads() == 1 and "OMP_NUM_THREADS" in os.environ:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        try:
            import psutil
            n = min(n, psutil.cpu_count(logical=False) or n)
        except ImportError:
            pass
        torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


def cpu_baseline(budget_s=20.0, batch=4):
    """The same CPU reference step on a bounded sample (rank 0, N = 1 only): best of up to 3 steps after a warm-up."""
    ref = _CpuReference(min(batch, BATCH_PER_GPU))
    threads = host_threads()
    times = []
    t_begin = time.perf_counter()
    while len(times) < 4 and (time.perf_counter() - t_begin) < budget_s:
        t0 = time.perf_counter()
        ref.step()
        times.append(time.perf_counter() - t0)
    best = min(times[1:]) if len(times) > 1 else times[0]
    what = "unmodified reference modules (oracle/_ref)" if ref.kind == "reference" else "port oracle/sr_oracle.py"
    return {"value": round(ref.batch / best, 3), "unit": "images/s", "cores": threads, "kind": ref.kind,
            "sample": "%d x (zero_grad+fwd+loss+bwd+Adam of batch %d, %s, fp32), best step after 1 warm-up"
                      % (len(times), ref.batch, what)}


# --------------------------------------------------------------------------------------------------
def _time_replayed(launch_sets, reps=5):
    """Average duration (ms) of ONE launch: the launches of `launch_sets` (a list of callables, each bound to its own
    set of seeded, non-zero buffers; together they exceed the 126 MB L2, so no launch finds its operands cached by the
    previous one) are captured back to back in a CUDA graph - no host gaps inside the bracket - and the replay is timed
    with CUDA events on the stream it runs on."""
    for f in launch_sets:
        f()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    rounds = max(1, 24 // len(launch_sets))
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(rounds):
                for f in launch_sets:
                    f()
        g.replay()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            g.replay()
            e1.record(side)
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / (rounds * len(launch_sets))
            best = t if best is None else min(best, t)
    return best


def kernel_rooflines(B, dev, pk, timer=None):
    """Per-kernel rooflines of the C2 step's heaviest kernels, each timed ALONE (hence against the BURST tensor peak /
    the measured copy bandwidth) at the step's own shapes, on four rotating sets of seeded random operands.
    Algorithmic work per launch: convs 2 * N*H*W * Cin * Cout * 9 FLOP; HBM kernels: every distinct input read once and
    every output written once at the storage dtype."""
    from srk import ops
    from srk import _lib as L
    g = torch.Generator(device=dev).manual_seed(7)
    NSET = 4

    def act(n, h, w, c, scale=1.0):
        t = torch.zeros((n, h + 2, w + 2, c), dtype=torch.bfloat16, device=dev)
        t[:, 1:-1, 1:-1] = (torch.randn((n, h, w, c), generator=g, device=dev) * scale).bfloat16()
        return t

    H = LR_HW
    w64 = torch.randn((64, 64, 3, 3), generator=g, device=dev) / 24
    wup = torch.randn((256, 64, 3, 3), generator=g, device=dev) / 24
    bias = torch.randn((64,), generator=g, device=dev) * 0.1
    gamma, beta = torch.rand((64,), generator=g, device=dev) + 0.5, torch.randn((64,), generator=g, device=dev) * 0.1
    alpha = torch.full((1,), 0.25, device=dev)
    xs = [act(B, H, H, 64) for _ in range(NSET)]
    ds = [act(B, H, H, 64, 1e-3) for _ in range(NSET)]
    ys, stats = [], []
    for x in xs:
        y, _, sm = ops.conv_fprop_stats(x, w64, bias)
        _, st = ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, None, sums=sm)
        ys.append(y)
        stats.append(st)
    # The BatchNorm sums travel from the conv epilogues to the BatchNorm passes in integer accumulators (include/srk.h
    # "exact sums"): producer and consumer alternate in the step.  Timed alone, a producer's accumulator is simply left
    # unconsumed (it keeps accumulating; integer wrap-around is harmless) and a consumer reads a consumed one (zeros):
    # neither changes the work of the kernel under test, and no zero-fill launch enters the timed region.
    acc = ops.acc_acquire(dev)

    def produced(a):
        if isinstance(a, ops.Acc):
            a.dirty = False

    def consuming(f):
        def run():
            acc.dirty = True
            f(acc)
        return run
    P = B * H * H
    conv_flop = 2.0 * P * 64 * 64 * 9
    act_bytes = B * (H + 2) * (H + 2) * 64 * 2
    out = []

    def add(name, bound, work, launches, per_step, note=None):
        ms = (timer or _time_replayed)(launches)
        if bound == "tensor":
            ach, peak, unit = work / (ms * 1e-3) / 1e12, pk["tf_burst"], "TFLOP/s"
        else:
            ach, peak, unit = work / (ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
        rec = {"kernel": name, "bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit,
               "frac": round(ach / peak, 4), "avg_ms": round(ms, 4), "launches_per_step": per_step,
               ("algorithmic_flops_per_launch" if bound == "tensor" else "algorithmic_bytes_per_launch"): work,
               "traffic": ncu_traffic(name)}
        if note:
            rec["note"] = note
        out.append(rec)

    add("conv3x3_c64_fprop+bn_stats", "tensor", conv_flop,
        [lambda x=x: produced(ops.conv_fprop_stats(x, w64, bias)[2]) for x in xs], 33,
        "fold::conv3x3_fold_tc_kernel<kStats>: the variant every trunk conv of the forward pass runs (sums into an accumulator)")
    add("conv3x3_c64_dgrad+bn_bwd_reduce", "tensor", conv_flop,
        [lambda d=d, y=y, st=st: produced(ops.conv_dgrad_bnred(d, w64, y, st, gamma, beta, alpha)[1])
         for d, y, st in zip(ds, ys, stats)], 16, "conv2 dgrad of a residual block + the bn1 backward sums")
    add("conv3x3_c64_dgrad+residual+bn_bwd_reduce", "tensor", conv_flop,
        [lambda d=d, y=y, st=st, x=x: produced(ops.conv_dgrad_bnred(d, w64, y, st, gamma, beta, None, residual=x)[1])
         for d, y, st, x in zip(ds, ys, stats, xs)], 15,
        "conv1 dgrad of a residual block + skip gradient + the bn2 backward sums of the block below")
    add("conv3x3_c64_dgrad+residual", "tensor", conv_flop,
        [lambda d=d, x=x: ops.conv_dgrad(d, False, w64, x, torch.bfloat16) for d, x in zip(ds, xs)], 2)
    add("conv3x3_c64_wgrad", "tensor", conv_flop,
        [lambda x=x, d=d: ops.conv_wgrad(x, False, d, False, w64, True) for x, d in zip(xs, ds)], 33,
        "wgrad3x3_tc_kernel + wgrad_fold_kernel (two launches)")
    x128 = [act(B, 2 * H, 2 * H, 64) for _ in range(2)]
    add("conv3x3_64to256_pixelshuffle_prelu@%dx%d" % (2 * H, 2 * H), "tensor", 2.0 * B * 4 * H * H * 64 * 256 * 9,
        [lambda x=x: ops.conv_fprop(x, False, wup, None, L.ACT_PRELU, alpha, None, 2, False, torch.bfloat16) for x in x128], 1)
    del x128
    add("bn_apply_train+prelu", "hbm", 2.0 * act_bytes,
        [consuming(lambda a, y=y: ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, None, sums=a))
         for y in ys], 16)
    add("bn_apply_train+residual", "hbm", 3.0 * act_bytes,
        [consuming(lambda a, y=y, x=x: ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, None, x, sums=a))
         for y, x in zip(ys, xs)], 17)
    add("bn_bwd_apply", "hbm", 3.0 * act_bytes,
        [consuming(lambda a, d=d, y=y, st=st: ops.bn_backward(d, y, st, gamma, beta, alpha, True, pre=a))
         for d, y, st in zip(ds, ys, stats)], 31,
        "srk_bn_bwd_apply_raw: 2 reads, 1 write; the sums come out of the dgrad epilogue above it")
    add("bn_bwd_reduce+bn_bwd_apply", "hbm", 5.0 * act_bytes,
        [lambda d=d, y=y, st=st: ops.bn_backward(d, y, st, gamma, beta, None, True) for d, y, st in zip(ds, ys, stats)], 2,
        "two launches: reduce (2 reads) + apply (2 reads, 1 write): the two BatchNorms whose gradient does not come "
        "out of a 64 -> 64 dgrad (top block, bn_mid)")
    del xs, ds, ys
    # loss / metric kernels on the step's image shapes
    from src.loss import get_loss_function
    from src.metrics import psnr_ssim_sums
    HR = LR_HW * SCALE
    imgs = [(torch.rand((B, 3, HR, HR), generator=g, device=dev), torch.rand((B, 3, HR, HR), generator=g, device=dev))
            for _ in range(NSET)]
    img_bytes = B * 3 * HR * HR * 4
    crit = get_loss_function("nlpd", dev)

    def nlpd_fb(sr, hr):
        sr = sr.detach().requires_grad_(True)
        crit(sr, hr).backward()
    add("nlpd_fwd+bwd", "hbm", 3.0 * img_bytes, [lambda a=a, b=b: nlpd_fb(a, b) for a, b in imgs], 1,
        "13 launches; algorithmic bytes = sr + hr read, grad written (the pyramid intermediates are extra traffic)")
    mae = get_loss_function("mae", dev)

    def mae_fb(sr, hr):
        sr = sr.detach().requires_grad_(True)
        mae(sr, hr).backward()
    add("l1_fwd+bwd", "hbm", 5.0 * img_bytes, [lambda a=a, b=b: mae_fb(a, b) for a, b in imgs], 0,
        "forward (2 reads) + backward (2 reads, 1 write); used by config C3")
    sse = torch.empty((B,), dtype=torch.float64, device=dev)
    st = ops.stream_ptr
    add("psnr_sse", "hbm", 2.0 * img_bytes,
        [lambda a=a, b=b: L.call("srk_psnr_sse", a.data_ptr(), b.data_ptr(), B, 3 * HR * HR, 1, sse.data_ptr(), st())
         for a, b in imgs], 0, "evaluation path (config C4)")
    add("ssim", "hbm", 2.0 * img_bytes,
        [lambda a=a, b=b: L.call("srk_ssim", a.data_ptr(), b.data_ptr(), B, 3, HR, HR, 1, sse.data_ptr(), st())
         for a, b in imgs], 0, "evaluation path; fp32-FMA bound (110 FMA per pixel-channel), HBM fraction for reference")
    return out


def ncu_traffic(kernel=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` summaries
    (profiles/r2_ncu_kernels.json: {kernel name: {...}}); None when that kernel has no capture."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_kernels.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p)).get("kernels", {})
    rec = d.get(kernel) if kernel else None
    return None if rec is None else rec.get("dram_bytes")


def run_srk(args):
    import srk
    from srk import dp, ops
    from srk.trainer import GraphStep
    from src.loss import get_loss_function
    from src.models import get_model
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    srk.set_compute_dtype(args.dtype)
    torch.manual_seed(0)
    model = get_model(ARCH, SCALE, dev)
    dp.broadcast_parameters(model)
    model.train()
    crit = get_loss_function(LOSS, dev)
    averager = dp.GradAverager(model.parameters()) if world > 1 else None
    # the step object train.py uses: forward + loss + backward + all-reduce + Adam + weight re-pack, captured into ONE
    # CUDA graph after the eager warm-up steps (NCCL all-reduce included) and replayed
    trainer = GraphStep(model, crit, lr=4e-4, betas=(0.5, 0.999), averager=averager, use_graph=not args.no_graph,
                        warmup=max(args.warmup - 1, 2),
                        overlap_wgrad=os.environ.get("SRK_OVERLAP_WGRAD", "1") != "0")

    from src.dataset import synthetic_pair
    B = args.batch
    lr_h, hr_h = synthetic_pair(B, LR_HW, LR_HW, SCALE, seed=1234 + rank)
    lr_pin, hr_pin = lr_h.pin_memory(), hr_h.pin_memory()
    lr_d, hr_d = lr_pin.to(dev), hr_pin.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):      # eager steps, then the capture (inside the trainer), then one replay
        trainer(lr_d, hr_d)
    barrier()
    trainer(lr_d, hr_d)
    graphed = trainer.static_inputs(lr_d, hr_d) is not None

    # e2e: every step's batch comes from pinned host memory.  The copy of step i+1 is issued on a copy stream while
    # step i computes (staging buffers); the step then takes it with a device-to-device hand-over.
    #   --e2e-input u8 (default): what train.py ships with its GPU sample pipeline (SRK_GPU_PIPELINE=1, srk/data.py) -
    #     decoded uint8 HR crops [B, H, W, 3] + crop offsets / flip flags; srk_sr_make_batch (ToTensor + antialiased
    #     bicubic, reference dataset.py:27-41) runs inside the timed region on the device;
    #   --e2e-input fp32: the finished fp32 (lr, hr) pair, as a host-side torchvision pipeline would deliver it.
    from srk import data as srk_data
    copy_stream = torch.cuda.Stream()
    lr_n, hr_n = torch.empty_like(lr_d), torch.empty_like(hr_d)
    ev_h2d, ev_taken = torch.cuda.Event(), torch.cuda.Event()
    u8 = args.e2e_input == "u8" and SCALE in (2, 4)
    if u8:
        HRS = LR_HW * SCALE
        src_pin = torch.randint(0, 256, (B, HRS, HRS, 3), dtype=torch.uint8,
                                generator=torch.Generator().manual_seed(99 + rank)).pin_memory()
        offs_pin = torch.zeros((B, 2), dtype=torch.int32).pin_memory()
        flips_pin = (torch.arange(B) % 2).to(torch.uint8).pin_memory()
        src_n = torch.empty(src_pin.shape, dtype=torch.uint8, device=dev)
        offs_n = torch.empty(offs_pin.shape, dtype=torch.int32, device=dev)
        flips_n = torch.empty(flips_pin.shape, dtype=torch.uint8, device=dev)
        h2d_bytes = src_pin.numel() + offs_pin.numel() * 4 + flips_pin.numel()
    else:
        h2d_bytes = lr_pin.numel() * 4 + hr_pin.numel() * 4

    def prefetch():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_taken)
            if u8:
                src_n.copy_(src_pin, non_blocking=True)
                offs_n.copy_(offs_pin, non_blocking=True)
                flips_n.copy_(flips_pin, non_blocking=True)
            else:
                lr_n.copy_(lr_pin, non_blocking=True)
                hr_n.copy_(hr_pin, non_blocking=True)
            ev_h2d.record(copy_stream)

    def take_batch(lr_dst, hr_dst):
        """this step's (lr, hr) from the staging buffers the copy stream has filled, into the step's input tensors"""
        if u8:
            srk_data.make_batch_into(src_n, offs_n, flips_n, LR_HW * SCALE, SCALE, lr_dst, hr_dst)
        else:
            lr_dst.copy_(lr_n, non_blocking=True)
            hr_dst.copy_(hr_n, non_blocking=True)

    def timed(nsteps, e2e):
        barrier()
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if e2e:
            ev_taken.record(main)
            prefetch()
        for i in range(nsteps):
            if e2e:
                main.wait_event(ev_h2d)
                # the step is enqueued before the next copy is (host time between two steps is GPU idle time: the loss
                # read below drains the device every step)
                if graphed:
                    s_lr, s_hr = trainer.static_inputs(lr_d, hr_d)
                    take_batch(s_lr, s_hr)
                    ev_taken.record(main)
                    loss = trainer.replay(lr_d, hr_d)
                else:
                    take_batch(lr_d, hr_d)
                    ev_taken.record(main)
                    loss = trainer(lr_d, hr_d)
                if i + 1 < nsteps:
                    prefetch()
                loss.item()
            elif graphed:
                trainer.replay(lr_d, hr_d)
            else:
                trainer(lr_d, hr_d)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / nsteps

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step = timed(args.steps, e2e=False)
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(max(2, min(args.steps, 20)), e2e=True)
    loss_value = float((trainer.replay(lr_d, hr_d) if graphed else trainer(lr_d, hr_d)).item())
    launches = trainer.launches_per_step or 0

    if rank != 0:
        if world > 1:
            shutdown_distributed(trainer)
        return
    pk = peaks()
    value = world * B / (ms_step * 1e-3)
    roofs = None
    if world > 1:
        trainer._graphs.clear()
    if ARCH == "RESNET" and args.dtype == "bf16" and not args.no_rooflines:
        del trainer
        trainer = None
        torch.cuda.empty_cache()
        roofs = kernel_rooflines(B, dev, pk)
    roof = None
    if roofs:
        r0 = roofs[0]
        roof = {"bound": r0["bound"], "kernel": r0["kernel"], "achieved": r0["achieved"], "peak": r0["peak"],
                "unit": r0["unit"], "frac": r0["frac"], "traffic": r0["traffic"], "avg_ms": r0["avg_ms"],
                "launches_per_step": r0["launches_per_step"],
                "algorithmic_flops_per_launch": r0["algorithmic_flops_per_launch"],
                "peak_source": pk["src"] + " bf16 burst (kernel timed alone, 4 rotating operand sets > L2)"}
    step_tf = world * B * FWD_BWD_GFLOP_PER_IMG / (ms_step * 1e-3) / 1e3
    line = {"metric": "sr_train_images_per_sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": "dp%d" % world,
                       "l2": "working set (>2 GB of activations per step) exceeds the 126 MB L2; no explicit flush",
                       "cuda_graph": bool(graphed), "nccl_in_graph": bool(graphed and world > 1),
                       "final_loss": round(loss_value, 5), "step_tflops": round(step_tf, 2),
                       "step_frac_of_bf16_sustained": round(step_tf / (world * pk["tf_sustained"]), 4)},
            "e2e": {"value": round(world * B / (ms_e2e * 1e-3), 2), "unit": "images/s",
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4,
                    "input": ("uint8 HR crops + crop offsets / flips; ToTensor + antialiased bicubic on the GPU inside the "
                              "timed region (train.py's default pipeline)") if u8 else "fp32 (lr, hr) pair"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "rooflines": roofs}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_distributed(trainer)


# --------------------------------------------------------------------------------------------------
def run_eval(args):
    """BASELINE config C4 (secondary measurement, `--config C4`): ResNet-SR inference 128 -> 512 + PSNR / SSIM / NLPD
    per batch over synthetic Food101-shaped images, batches sharded over the ranks (srk/evaluate.py), one all-reduce
    of the per-batch metric sums at the end.  `value`: images/s with the batches resident in HBM; `e2e`: the same
    from pinned host memory."""
    import srk
    from srk import evaluate as ev
    from src.dataset import synthetic_pair
    from src.metrics import MetricsCalculator
    from src.models import get_model
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    srk.set_compute_dtype(args.dtype)
    torch.manual_seed(0)
    model = get_model("RESNET", 4, dev).eval()
    B = args.batch
    per_rank = args.eval_images            # images per GPU (full C4: 10 000 / 8 = 1 250)
    sizes = [B] * (per_rank // B) + ([per_rank % B] if per_rank % B else [])
    # two distinct synthetic batches are cycled (generating 1 250 x 512^2 images on the host would take minutes)
    base = [synthetic_pair(B, 128, 128, 4, seed=4321 + rank * 2 + k) for k in range(2)]
    host = [(base[i % 2][0][:n].contiguous().pin_memory(), base[i % 2][1][:n].contiguous().pin_memory())
            for i, n in enumerate(sizes)]
    resident = [(lr.to(dev), hr.to(dev)) for lr, hr in host]
    metrics_fn = MetricsCalculator(dev).compute

    def run(batches):
        # every rank evaluates ITS list (rank = 0, world = 1 inside evaluate); the cross-rank mean is one all-reduce
        return ev.evaluate(model, batches, dev, 0, 1, metrics_fn)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        run(resident[:2])
    sampler = ClockSampler(local_rank) if rank == 0 else None
    times = {}
    res = None
    for name, batches in (("resident", resident), ("e2e", host)):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = run(batches)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        times[name] = ms
    clocks = sampler.stop() if sampler else None
    if world > 1:
        buf = torch.tensor([res["psnr"], res["ssim"], float(res["batches"])], dtype=torch.float64, device=dev)
        buf[:2] *= res["batches"]
        dist.all_reduce(buf)
        res = {"psnr": float(buf[0] / buf[2]), "ssim": float(buf[1] / buf[2]), "batches": int(buf[2])}
    if rank == 0:
        total = world * per_rank
        pk = peaks()
        tf = total * 72.69 / (times["resident"] * 1e-3) / 1e3     # 72.69 GFLOP per image forward at 128 -> 512
        line = {"metric": "sr_eval_images_per_sec", "value": round(total / (times["resident"] * 1e-3), 2),
                "unit": "images/s", "n_gpus": world, "steps": len(sizes), "warmup": args.warmup,
                "ms_per_step": round(times["resident"] / len(sizes), 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": "C4: ResNet-SR 16x64ch x4 inference 128->512 + PSNR/SSIM/NLPD per batch, batch %d, "
                                       "%d images per GPU, batches sharded over ranks" % (B, per_rank),
                           "global_images": total, "parallelism": "dp%d" % world,
                           "l2": "activations of one batch (2 GB at 512^2) exceed the 126 MB L2",
                           "psnr": round(res["psnr"], 4), "ssim": round(res["ssim"], 5),
                           "step_tflops": round(tf, 2), "step_frac_of_bf16_sustained": round(tf / (world * pk["tf_sustained"]), 4)},
                "e2e": {"value": round(total / (times["e2e"] * 1e-3), 2), "unit": "images/s",
                        "h2d_bytes_per_step": int(B * 3 * (128 * 128 + 512 * 512) * 4), "d2h_bytes_per_step": 32},
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_distributed()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="srk", choices=["srk", "reference"])
    ap.add_argument("--dtype", default=os.environ.get("SRK_BENCH_DTYPE", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the config's)")
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS) + ["C4"],
                    help="BASELINE.json config (headline: C2; C4 = sharded inference + metrics)")
    ap.add_argument("--eval-images", type=int, default=1250, help="--config C4: images per GPU (10 000 / 8)")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="--impl reference: images per CPU step (default: the config's batch when it fits the time budget)")
    ap.add_argument("--e2e-input", default="u8", choices=["u8", "fp32"],
                    help="what the e2e leg ships per step: decoded uint8 crops (GPU sample pipeline) or the fp32 pair")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rooflines", action="store_true", help="skip the per-kernel roofline measurements")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.config == "C4":
        if args.batch <= 0:
            args.batch = 64
        if not torch.cuda.is_available():
            sys.exit("bench.py: no CUDA device; the SR hot path has no CPU fallback")
        _, _, world = dist_env()
        if world == 1 and args.gpus > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29511"] + sys.argv
            sys.exit(subprocess.call(cmd))
        run_eval(args)
        return
    select_config(args.config)
    if args.batch <= 0:
        args.batch = BATCH_PER_GPU
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the SR hot path has no CPU fallback (use --impl reference for the CPU port)")
    _, _, world = dist_env()
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511"] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_srk(args)


if __name__ == "__main__":
    main()
