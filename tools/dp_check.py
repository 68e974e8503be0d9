"""Two-or-more-rank check of the data-parallel step on real GPUs (run under torchrun): overlapped bucketed all-reduce
inside the captured graph vs the plain pack / all-reduce / unpack path, eager vs replayed - every variant must leave
bit-identical weights on every rank, and the ranks must agree with each other.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 tools/dp_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import srk  # noqa: E402
from srk import dp  # noqa: E402
from srk.trainer import GraphStep  # noqa: E402
from src.dataset import synthetic_pair  # noqa: E402
from src.loss import get_loss_function  # noqa: E402
from src.models import ResNetSR  # noqa: E402

import faulthandler  # noqa: E402
faulthandler.dump_traceback_later(int(os.environ.get("DP_DUMP_AFTER", "45")), exit=True)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
srk.set_compute_dtype("bf16")
lr, hr = synthetic_pair(4, 32, 32, 4, seed=100 + rank)
lr, hr = lr.to(dev), hr.to(dev)
results = {}
VARIANTS = (("plain-eager", False, False), ("overlap-eager", True, False), ("overlap-graph", True, True),
            ("plain-graph", False, True))
only = os.environ.get("DP_VARIANTS")
if only:
    VARIANTS = tuple(v for v in VARIANTS if v[0] in ("plain-eager",) + tuple(only.split(",")))
for name, overlap, graph in VARIANTS:
    if rank == 0:
        print("variant", name, file=sys.stderr, flush=True)
    torch.manual_seed(0)
    model = ResNetSR(num_channels=64, num_residuals=3).to(dev).train()
    dp.broadcast_parameters(model)
    avg = dp.GradAverager(model.parameters(), bucket_bytes=256 << 10)
    step = GraphStep(model, get_loss_function("nlpd", dev), lr=4e-4, averager=avg, use_graph=graph, warmup=2, overlap_comm=overlap)
    losses = []
    for it in range(6):
        losses.append(float(step(lr, hr)))
        if rank == 0:
            print("  step", it, losses[-1], file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same_across_ranks = all(torch.equal(gathered[0], g) for g in gathered)
    results[name] = (flat.clone(), losses, same_across_ranks, len(avg.buckets))
ref = results["plain-eager"][0]
ok = True
for name, (flat, losses, same, nb) in results.items():
    eq = torch.equal(flat, ref)
    ok = ok and eq and same
    if rank == 0:
        print("%-14s buckets %d  ranks agree %s  == plain-eager %s  losses %s" % (name, nb, same, eq, ["%.5f" % l for l in losses]))
if rank == 0:
    print("DP_CHECK_OK" if ok else "DP_CHECK_FAILED", flush=True)
sys.stdout.flush()
os._exit(0 if ok else 1)      # captured graphs hold NCCL kernels: destroy_process_group() would not return
