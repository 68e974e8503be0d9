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
    plain = [p.grad.detach().clone() for p in net.parameters()]
    # overlapped mode (buckets reduced from inside backward as their gradients arrive) must give the same numbers,
    # also for a parameter that receives gradient from two places of the graph (two accumulations)
    for p in net.parameters():
        p.grad = None
    avg.begin_backward()
    (((net(x_all[b:e]) - y_all[b:e]) ** 2).mean()).backward()
    avg.finish_backward()
    assert all(torch.equal(p.grad, g) for p, g in zip(net.parameters(), plain))
    twice = net[0].weight
    for p in net.parameters():
        p.grad = None
    avg.begin_backward()
    (((net(x_all[b:e]) - y_all[b:e]) ** 2).mean() + (twice ** 2).sum() * 0.5).backward()
    avg.finish_backward()
    want = plain[0] + twice.detach()
    assert torch.allclose(net[0].weight.grad, want, atol=1e-6), (net[0].weight.grad - want).abs().max()
    for p, g in zip(net.parameters(), plain):
        p.grad = g
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


def _fit_worker(rank, world, port, out):
    """train.fit (the epoch loop of train.py) on two ranks whose LOCAL validation numbers differ: the loop must run
    the same number of epochs on both, stop early on the same epoch, and never leave a rank waiting in a collective."""
    sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=__import__("datetime").timedelta(seconds=60))
    import argparse
    import train as T
    from srk import dp
    from srk import evaluate as ev
    cfg = argparse.Namespace(epochs=12, patience=2)
    net = torch.nn.Linear(4, 4)
    dp.broadcast_parameters(net)
    averager = dp.GradAverager(net.parameters())
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="max", factor=0.5, patience=2)
    # identical batches on every rank (seeded loader), last batch smaller than the world size -> skipped everywhere
    g = torch.Generator().manual_seed(5)
    data = [(torch.randn(n, 4, generator=g), torch.randn(n, 4, generator=g)) for n in (6, 5, 3, 1)]
    steps, weights = [0], []

    def step_fn(x, y, weight):
        opt.zero_grad()
        loss = ((net(x) - y) ** 2).mean() * weight
        loss.backward()
        averager.average()          # the collective every rank must reach the same number of times
        opt.step()
        steps[0] += 1
        weights.append(weight)
        return loss.detach()

    epoch = [0]
    # per-rank "validation PSNR" that would make the ranks disagree if it steered control flow directly: rank 0 keeps
    # improving, rank 1 gets worse; the all-reduced mean peaks at epoch 2, so both ranks stop after epoch 4
    local_psnr = {0: [10, 12, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23], 1: [10, 12, 14, 9, 6, 3, 1, 0, 0, 0, 0, 0]}
    val_batches = [(torch.full((2, 1), float(i)), torch.zeros(2, 1)) for i in range(4)]

    def eval_fn():
        e = epoch[0]
        epoch[0] += 1
        fake = lambda sr, hr: {"psnr": float(local_psnr[rank][e]), "ssim": 0.0, "nlpd": 0.0, "lpips": 0.0}
        return ev.evaluate(lambda x: x, val_batches, "cpu", rank, world, metrics_fn=fake,
                           criterion=lambda sr, hr: (sr - hr).abs().mean())

    saved = []
    best, epochs_run = T.fit(cfg, train_loader=data, step_fn=step_fn, eval_fn=eval_fn, scheduler=sched,
                             get_lr=lambda: opt.param_groups[0]["lr"], save_best=lambda ep: saved.append(ep), rank=rank,
                             world=world, log=lambda d: None, shard=lambda t: t)
    dist.barrier()
    w = torch.cat([p.detach().flatten() for p in net.parameters()])
    ws = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    out[rank] = (epochs_run, steps[0], best, saved, opt.param_groups[0]["lr"], float((ws[0] - ws[1]).abs().max()),
                 sorted(set(round(x, 6) for x in weights)))
    dist.destroy_process_group()


def test_training_loop_early_stops_in_lockstep_on_two_ranks():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_fit_worker, args=(world, 29547, out), nprocs=world, join=True)
        assert len(out) == world
        a, b = out[0], out[1]
        assert a[:5] == b[:5], (a, b)                      # epochs, steps, best PSNR, saved epochs, learning rate
        assert a[0] == 5 and a[1] == 5 * 3 and a[3] == [0, 1, 2]   # stopped after 2 epochs without improvement; 1-sample batch skipped
        assert abs(a[2] - 14.0) < 1e-12
        assert a[5] == 0.0 and b[5] == 0.0                 # weights stayed identical on both ranks
        # uneven shards (5 -> 3 + 2, 3 -> 2 + 1) carry weights n_local * world / n_global
        assert a[6] == [1.0, 1.2, round(4 / 3, 6)] and b[6] == [round(2 / 3, 6), 0.8, 1.0]
