"""World-size-2 checks of the data-parallel host logic on the gloo backend (CPU): sharding, bucket
construction and gradient averaging.  The kernels are not involved."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from srk import dp
    torch.manual_seed(100 + rank)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    dp.broadcast_parameters(net)
    w0 = [p.detach().clone() for p in net.parameters()]
    torch.manual_seed(7)
    x_all, y_all = torch.randn(8, 5), torch.randn(8, 3)
    b, e = dp.shard_range(8, rank, world)
    loss = ((net(x_all[b:e]) - y_all[b:e]) ** 2).mean()
    loss.backward()
    avg = dp.GradAverager(net.parameters(), bucket_bytes=64)
    assert len(avg.buckets) >= 2
    avg.average()
    # single-process oracle over the whole batch
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    with torch.no_grad():
        for p, w in zip(ref.parameters(), w0):
            p.copy_(w)
    ((ref(x_all) - y_all) ** 2).mean().backward()
    err = max((p.grad - q.grad).abs().max().item() for p, q in zip(net.parameters(), ref.parameters()))
    out[rank] = err
    dist.destroy_process_group()


def test_grad_averaging_equals_full_batch_gradient():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29533, out), nprocs=world, join=True)
        assert len(out) == world and all(v < 1e-6 for v in out.values()), dict(out)


def test_shard_range_covers_everything_once():
    sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
    from srk import dp
    for total in (0, 1, 7, 10000, 1250):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _eval_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from srk import evaluate as ev
    torch.manual_seed(0)
    batches = [(torch.rand(4 if i < 6 else 3, 3, 8, 8), torch.rand(4 if i < 6 else 3, 3, 8, 8)) for i in range(7)]
    fake_model = lambda x: x * 2.0
    def metrics(sr, hr):  # stand-in for MetricsCalculator.compute: any per-batch function
        return {"psnr": float((sr - hr).abs().mean()), "ssim": float(sr.mean()), "nlpd": float(hr.std()), "lpips": float("nan")}
    got = ev.evaluate(fake_model, batches, "cpu", rank, world, metrics_fn=metrics)
    want = {k: sum(metrics(fake_model(a), b)[k] for a, b in batches) / len(batches) for k in ("psnr", "ssim", "nlpd")}
    out[rank] = (got["batches"], max(abs(got[k] - want[k]) for k in want), got["lpips"] != got["lpips"])
    dist.destroy_process_group()


def test_sharded_evaluation_reproduces_batch_mean_semantics():
    """7 batches (ragged last one) over 2 ranks == the single-process mean over batches (train.py:161,193-195)."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_eval_worker, args=(world, 29541, out), nprocs=world, join=True)
        assert len(out) == world
        for nb, err, lp_nan in out.values():
            assert nb == 7 and err < 1e-12 and lp_nan
