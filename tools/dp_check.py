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

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
srk.set_compute_dtype("bf16")
lr, hr = synthetic_pair(4, 32, 32, 4, seed=100 + rank)
lr, hr = lr.to(dev), hr.to(dev)
results = {}
for name, overlap, graph in (("plain-eager", False, False), ("overlap-eager", True, False), ("overlap-graph", True, True),
                             ("plain-graph", False, True)):
    torch.manual_seed(0)
    model = ResNetSR(num_channels=64, num_residuals=3).to(dev).train()
    dp.broadcast_parameters(model)
    avg = dp.GradAverager(model.parameters(), bucket_bytes=256 << 10)
    step = GraphStep(model, get_loss_function("nlpd", dev), lr=4e-4, averager=avg, use_graph=graph, warmup=2, overlap_comm=overlap)
    losses = [float(step(lr, hr)) for _ in range(6)]
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
    print("DP_CHECK_OK" if ok else "DP_CHECK_FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
