"""Times the BatchNorm passes of the C2 trunk that consume conv-epilogue sums: srk_bn_apply_train and
srk_bn_bwd_apply_raw, fed by float sums or by an integer accumulator (timing is data independent: a consumed
accumulator reads as zeros).  Same harness as tools/bench_conv.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import srk  # noqa: E402
from srk import ops  # noqa: E402

srk.set_compute_dtype("bf16")
dev = torch.device("cuda:0")
B, H = int(os.environ.get("B", 64)), int(os.environ.get("H", 64))
g = torch.Generator(device=dev).manual_seed(1)


def act(scale=1.0):
    t = torch.zeros((B, H + 2, H + 2, 64), dtype=torch.bfloat16, device=dev)
    t[:, 1:-1, 1:-1] = (torch.randn((B, H, H, 64), generator=g, device=dev) * scale).bfloat16()
    return t


ys, ds, xs = [act() for _ in range(4)], [act(1e-3) for _ in range(4)], [act() for _ in range(4)]
gamma, beta = torch.rand((64,), generator=g, device=dev) + 0.5, torch.zeros((64,), device=dev)
alpha = torch.full((1,), 0.25, device=dev)
sums = torch.rand((2, 64), device=dev) * 1e5
sums[1] += 1e6
red = torch.rand((129,), device=dev)
stats = torch.stack([torch.zeros(64, device=dev), torch.ones(64, device=dev)])
acc = ops.acc_acquire(dev)


def with_acc(f):
    def run():
        acc.dirty = True
        f()
    return run


cases = {
    "bn_apply_train+prelu  float": [lambda y=y: ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, None, sums=sums) for y in ys],
    "bn_apply_train+prelu  acc": [with_acc(lambda y=y: ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, None, sums=acc)) for y in ys],
    "bn_apply_train+res    float": [lambda y=y, x=x: ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, None, x, sums=sums) for y, x in zip(ys, xs)],
    "bn_apply_train+res    acc": [with_acc(lambda y=y, x=x: ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, None, x, sums=acc)) for y, x in zip(ys, xs)],
    "bn_bwd_apply_raw+prelu float": [lambda d=d, y=y: ops.bn_backward(d, y, stats, gamma, beta, alpha, True, pre=red) for d, y in zip(ds, ys)],
    "bn_bwd_apply_raw+prelu acc": [with_acc(lambda d=d, y=y: ops.bn_backward(d, y, stats, gamma, beta, alpha, True, pre=acc)) for d, y in zip(ds, ys)],
    "bn_bwd_reduce+apply": [lambda d=d, y=y: ops.bn_backward(d, y, stats, gamma, beta, None, True) for d, y in zip(ds, ys)],
}
nbytes = B * (H + 2) * (H + 2) * 64 * 2
for name, fs in cases.items():
    ms = bench._time_replayed(fs)
    print("%-30s %7.2f us   (%d-tensor pass: %.0f GB/s)" % (name, ms * 1e3, 3 if ("res" in name or "bwd" in name) else 2,
                                                           (3 if ("res" in name or "bwd" in name) else 2) * nbytes / (ms * 1e-3) / 1e9), flush=True)
