"""Batch-sharded evaluation (BASELINE config C4: ResNet-SR inference + PSNR/SSIM over many images on several GPUs).

Reference semantics (train.py:148-162, 189-195): every metric is computed PER BATCH by
MetricsCalculator.compute and the per-batch values are averaged over the batches (`/ len(loader)`), a ragged
last batch counting as one batch.  To reproduce that exactly the unit of sharding is the batch: rank r evaluates
batches r, r + world, ...; each rank accumulates sums of the per-batch metrics and one all-reduce of
[sum_psnr, sum_ssim, sum_nlpd, sum_lpips, n_batches] gives the dataset means on every rank.  Batches are
independent, so there is no other collective on the path."""
import math

import torch
import torch.distributed as dist

KEYS = ("psnr", "ssim", "nlpd", "lpips")


def my_batches(n_batches, rank, world):
    return range(rank, n_batches, world)


@torch.no_grad()
def evaluate(model, batches, device, rank=0, world=1, metrics_fn=None, group=None, criterion=None):
    """batches: sequence or iterable (e.g. a DataLoader - identical on every rank) of (lr, hr) NCHW fp32 tensors (host or
    device).  Returns {"psnr","ssim","nlpd","lpips","batches"} averaged over all batches of all ranks - the same numbers
    on every rank, so that schedulers and early stopping driven by them stay in lockstep.  criterion: optional loss
    module; adds "loss" (mean of the per-batch values, reference train.py:155-162).  Infinite PSNR (identical images)
    propagates as inf."""
    if metrics_fn is None:
        from src.metrics import MetricsCalculator
        metrics_fn = MetricsCalculator(device).compute
    was_training = getattr(model, "training", False)
    if hasattr(model, "eval"):
        model.eval()
    keys = KEYS + (("loss",) if criterion is not None else ())
    sums = [0.0] * len(keys)
    count = 0
    if hasattr(batches, "__getitem__"):
        mine = (batches[i] for i in my_batches(len(batches), rank, world))
    else:
        mine = (b for i, b in enumerate(batches) if i % world == rank)
    for lr, hr in mine:
        lr, hr = lr.to(device, non_blocking=True), hr.to(device, non_blocking=True)
        sr = model(lr)
        res = dict(metrics_fn(sr, hr))
        if criterion is not None:
            res["loss"] = float(criterion(sr, hr))
        for k, key in enumerate(keys):
            v = res.get(key, float("nan"))
            sums[k] += v
        count += 1
    if was_training and hasattr(model, "train"):
        model.train()
    # inf / nan do not survive a SUM across ranks reliably: reduce the finite parts and (inf, nan) flags apart
    flags = [1000.0 if math.isnan(s) else (1.0 if math.isinf(s) else 0.0) for s in sums]
    fin = [s if math.isfinite(s) else 0.0 for s in sums]
    on_gpu = world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl"
    buf = torch.tensor(fin + flags + [float(count)], dtype=torch.float64, device=device if on_gpu else "cpu")
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    buf = buf.cpu().tolist()
    n = len(keys)
    total = buf[2 * n]
    out = {"batches": int(total)}
    for k, key in enumerate(keys):
        if buf[n + k] >= 1000.0:
            out[key] = float("nan")
        elif buf[n + k] >= 1.0:
            out[key] = float("inf")
        else:
            out[key] = buf[k] / max(total, 1.0)
    return out
