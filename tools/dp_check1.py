"""Single-GPU isolation of the overlapped gradient averaging: a 1-rank NCCL group with the averager told world = 2
(gradients are halved, the all-reduce is the identity) - plain pack/all-reduce/unpack vs buckets launched from inside
backward, with and without side-stream weight gradients; prints every parameter whose gradient differs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import srk
from srk import dp, ops
from src.dataset import synthetic_pair
from src.loss import get_loss_function
from src.models import ResNetSR
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29611")
dev = torch.device("cuda:0")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
srk.set_compute_dtype("bf16")
lr, hr = synthetic_pair(4, 32, 32, 4, seed=100)
lr, hr = lr.to(dev), hr.to(dev)
crit = get_loss_function("nlpd", dev)
res = {}
for side in (False, True):
    for overlap in (False, True):
        srk.set_overlap_wgrad(side)
        torch.manual_seed(0)
        model = ResNetSR(num_channels=64, num_residuals=3).to(dev).train()
        avg = dp.GradAverager(model.parameters(), bucket_bytes=256 << 10)
        avg.world = 2
        loss = crit(model(lr), hr)
        if overlap:
            avg.begin_backward(); loss.backward(); avg.finish_backward()
        else:
            loss.backward(); avg.average()
        torch.cuda.synchronize()
        res[(side, overlap)] = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
ref = res[(False, False)]
for key, g in res.items():
    bad = [(k, float((g[k] - ref[k]).abs().max()), float(ref[k].abs().max())) for k in ref if not torch.equal(g[k], ref[k])]
    print("side=%s overlap=%s: %d / %d parameters differ" % (key[0], key[1], len(bad), len(ref)), bad[:8])
