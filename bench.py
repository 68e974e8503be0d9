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


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r1_conv3x3_ncu.json); None when the summary is missing."""
    p = os.path.join(ROOT, "profiles", "r1_conv3x3_ncu.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return {"bytes": d["dram_bytes_read"] + d["dram_bytes_write"], "algorithmic_bytes": d["algorithmic_bytes"],
            "tensor_pipe_active_pct": d["tensor_pipe_active_pct"], "source": "profiles/r1_conv3x3_ncu.json"}


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


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU port of the reference path on the host cores (rank 0 only)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import sr_oracle as O
    torch.manual_seed(0)
    threads = host_threads()
    sample_batch = args.cpu_batch
    sd = cpu_state_dict()
    lr, hr = O.synthetic_pair(sample_batch, LR_HW, LR_HW, SCALE)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running_" not in k}
    work = dict(sd)
    work.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=4e-4, betas=(0.5, 0.999))

    def step():
        opt.zero_grad()
        out = O.model_forward(ARCH, work, lr, training=True, scale_factor=SCALE)
        loss = O.loss_fn(LOSS)(out, hr)
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = sample_batch / dt
    sample = "batch %d of the C2 workload per step (fp32, torch CPU ATen ops, %d threads)" % (sample_batch, threads)
    line = {"impl": "reference", "metric": "sr_train_images_per_sec", "value": round(v, 3), "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": round(v, 3), "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_state_dict():
    """Seeded ResNet-SR weights built on the CPU through the drop-in constructors (no compute)."""
    from src.models import get_model
    torch.manual_seed(0)
    return {k: v.clone() for k, v in get_model(ARCH, SCALE, "cpu").state_dict().items()}


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


def cpu_baseline(budget_s=20.0, batch=2):
    from oracle import sr_oracle as O
    sd = cpu_state_dict()
    lr, hr = O.synthetic_pair(batch, LR_HW, LR_HW, SCALE)
    threads = host_threads()
    times = []
    t_begin = time.perf_counter()
    while len(times) < 4 and (time.perf_counter() - t_begin) < budget_s:
        t0 = time.perf_counter()
        O.train_step_grads(ARCH, sd, lr, hr, LOSS, scale_factor=SCALE)
        times.append(time.perf_counter() - t0)
    best = min(times[1:]) if len(times) > 1 else times[0]
    return {"value": round(batch / best, 3), "unit": "images/s", "cores": threads, "kind": "port",
            "sample": "%d x (fwd+loss+bwd of batch %d, C2 shapes, fp32), best step after 1 warm-up" % (len(times), batch)}


# --------------------------------------------------------------------------------------------------
def run_srk(args):
    import srk
    from srk import dp, ops
    from srk import _lib as L
    from src.loss import get_loss_function
    from src.models import get_model
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    srk.set_compute_dtype(args.dtype)
    srk.set_overlap_wgrad(os.environ.get("SRK_OVERLAP_WGRAD", "1") != "0")   # weight gradients on the side stream
    torch.manual_seed(0)
    model = get_model(ARCH, SCALE, dev)
    dp.broadcast_parameters(model)
    model.train()
    crit = get_loss_function(LOSS, dev)
    opt = srk.optim.Adam(model.parameters(), lr=4e-4, betas=(0.5, 0.999))
    averager = dp.GradAverager(model.parameters()) if world > 1 else None

    from src.dataset import synthetic_pair
    B = args.batch
    lr_h, hr_h = synthetic_pair(B, LR_HW, LR_HW, SCALE, seed=1234 + rank)
    lr_pin, hr_pin = lr_h.pin_memory(), hr_h.pin_memory()
    lr_d, hr_d = lr_pin.to(dev), hr_pin.to(dev)

    def fwd_bwd(lr, hr):
        ops.begin_step(device=dev)   # one memset serves every zero-initialised scratch tensor of the step
        opt.zero_grad()
        loss = crit(model(lr), hr)
        loss.backward()
        if averager is not None:
            averager.pack()          # gradients -> flat fp32 buckets / world
        return loss

    def finish():
        if averager is not None:
            averager.unpack()
        opt.step()
        ops.repack_all()             # all weight packs of the next step in one launch

    def step(lr, hr):
        loss = fwd_bwd(lr, hr)
        if averager is not None:
            averager.all_reduce()    # NCCL over NVLink, one collective per bucket
        finish()
        return loss

    # The step is captured once into two CUDA graphs (forward + loss + backward + weight re-pack | Adam) and
    # replayed: same kernels, same work, no per-launch host overhead.  The NCCL all-reduce between them is
    # launched eagerly.
    use_graph = not args.no_graph
    lr_s, hr_s = lr_d.clone(), hr_d.clone()
    g1 = g2 = loss_s = None

    def run_step():
        if g1 is None:
            return step(lr_s, hr_s)
        g1.replay()
        if averager is not None:
            averager.all_reduce()
        g2.replay()
        return loss_s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # e2e: every step's LR/HR batch comes from pinned host memory.  The copy of step i+1 is issued on a copy stream
    # while step i computes (staging buffers), so the step only pays a device-to-device hand-over.
    copy_stream = torch.cuda.Stream()
    lr_n, hr_n = torch.empty_like(lr_s), torch.empty_like(hr_s)
    ev_h2d, ev_taken = torch.cuda.Event(), torch.cuda.Event()

    def prefetch():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_taken)
            lr_n.copy_(lr_pin, non_blocking=True)
            hr_n.copy_(hr_pin, non_blocking=True)
            ev_h2d.record(copy_stream)

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
                lr_s.copy_(lr_n, non_blocking=True)
                hr_s.copy_(hr_n, non_blocking=True)
                ev_taken.record(main)
                if i + 1 < nsteps:
                    prefetch()
                run_step().item()
            else:
                run_step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / nsteps

    c0 = L.launch_calls
    step(lr_s, hr_s)                      # eager: also counts the libsrk launches of one step
    launches = L.launch_calls - c0
    for _ in range(max(args.warmup - 1, 2)):
        step(lr_s, hr_s)
    if use_graph:
        barrier()
        ops.repack_all()                  # packs are current; inside the graphs only finish() refreshes them
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1, stream=side, capture_error_mode="thread_local"):
                loss_s = fwd_bwd(lr_s, hr_s)
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=side, pool=g1.pool(), capture_error_mode="thread_local"):
                finish()
        torch.cuda.current_stream().wait_stream(side)
        barrier()
        run_step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step = timed(args.steps, e2e=False)
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(max(2, min(args.steps, 20)), e2e=True)
    loss_value = float(run_step().item())

    # dominant kernel (3x3 64->64 conv forward at the step's own shape): its launches inside one eager step are
    # counted, and its duration is measured with CUDA events on the launching stream around 20 back-to-back
    # launches replayed from a CUDA graph (no host launch gaps inside the bracket)
    ops.kernel_timer = ops.KernelTimer(lambda k: k[0] == "conv_fprop" and k[1:5] == (64, 64, 3, 0))
    step(lr_s, hr_s)
    torch.cuda.synchronize()
    kt = ops.kernel_timer.summary()
    ops.kernel_timer = None
    kern_ms = None
    if kt and ARCH == "RESNET":
        blk = model.res_blocks[0]
        xa = torch.zeros((B, LR_HW + 2, LR_HW + 2, 64), dtype=ops.cfg.compute_dtype, device=dev)
        one = lambda: ops.conv_fprop(xa, False, blk.conv1.weight, blk.conv1.bias, 0, None, None, 0, False, xa.dtype)
        one()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            kg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(kg, stream=side):
                for _ in range(20):
                    one()
            kg.replay()
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(side)
            kg.replay()
            k1.record(side)
            torch.cuda.synchronize()
            kern_ms = k0.elapsed_time(k1) / 20

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = world * B / (ms_step * 1e-3)
    roof = None
    if kt and kern_ms is not None:
        key, (_, count) = max(kt.items(), key=lambda kv: kv[1][0] * kv[1][1])
        avg_ms = kern_ms
        n, h, w = key[5:8]
        flops = 2.0 * n * h * w * 64 * 64 * 9
        ach = flops / (avg_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv3x3_c64_fprop(%s)" % ("tcgen05" if key[8] else "cuda-core"),
                "achieved": round(ach, 2), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tf_sustained"], 4),
                "traffic": (ncu_traffic() or {}).get("bytes"), "traffic_detail": ncu_traffic(),
                "algorithmic_flops_per_launch": flops, "avg_ms": round(avg_ms, 4),
                "launches_per_step": count, "peak_source": pk["src"] + " bf16 sustained"}
    step_tf = world * B * FWD_BWD_GFLOP_PER_IMG / (ms_step * 1e-3) / 1e3
    line = {"metric": "sr_train_images_per_sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": "dp%d" % world,
                       "l2": "working set (>2 GB of activations per step) exceeds the 126 MB L2; no explicit flush",
                       "cuda_graph": bool(g1 is not None), "final_loss": round(loss_value, 5),
                       "step_tflops": round(step_tf, 2),
                       "step_frac_of_bf16_sustained": round(step_tf / (world * pk["tf_sustained"]), 4)},
            "e2e": {"value": round(world * B / (ms_e2e * 1e-3), 2), "unit": "images/s",
                    "h2d_bytes_per_step": int(lr_pin.numel() * 4 + hr_pin.numel() * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
        dist.destroy_process_group()


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
    ap.add_argument("--cpu-batch", type=int, default=4, help="--impl reference: images per CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
