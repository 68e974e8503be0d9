"""Times the trunk-conv variants (SRK_TC_FOLD 2 = per-tap halo slab, 4 = column strips) at the C2 layer shape:
fprop, fprop + BN statistics, dgrad + residual, dgrad + BN-backward reduction.  Four rotating operand sets (> L2),
24 launches replayed from a CUDA graph, CUDA events on the replay stream."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import srk  # noqa: E402
from srk import _lib as L  # noqa: E402
from srk import ops  # noqa: E402

srk.set_compute_dtype("bf16")
dev = torch.device("cuda:0")
B, H = int(os.environ.get("B", 64)), int(os.environ.get("H", 64))
g = torch.Generator(device=dev).manual_seed(1)


def act(scale=1.0):
    t = torch.zeros((B, H + 2, H + 2, 64), dtype=torch.bfloat16, device=dev)
    t[:, 1:-1, 1:-1] = (torch.randn((B, H, H, 64), generator=g, device=dev) * scale).bfloat16()
    return t


xs, ds = [act() for _ in range(4)], [act(1e-3) for _ in range(4)]
w = torch.randn((64, 64, 3, 3), generator=g, device=dev) / 24
bias = torch.randn((64,), generator=g, device=dev) * 0.1
gamma, beta = torch.rand((64,), generator=g, device=dev) + 0.5, torch.zeros((64,), device=dev)
alpha = torch.full((1,), 0.25, device=dev)
sums = [torch.empty((2, 64), dtype=torch.float32, device=dev) for _ in range(4)]
ys, stats = [], []
for x, sm in zip(xs, sums):
    y, _ = ops.conv_fprop(x, False, w, bias, 0, None, None, 0, False, torch.bfloat16, bn_sums=sm)
    _, st = ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, None, sums=sm)
    ys.append(y)
    stats.append(st)


def unconsumed(a):
    """timing only: the accumulator keeps accumulating (integer wrap-around is harmless here); no zero-fill launches"""
    if isinstance(a, ops.Acc):
        a.dirty = False


out = (ctypes.c_float * 2)()
flop = 2.0 * B * H * H * 64 * 64 * 9
for fold in [int(v) for v in os.environ.get("FOLDS", "2,4").split(",")]:
    L.call("srk_tc_probe", 10 + fold, out, 2)
    cases = {
        "fprop": [lambda x=x: ops.conv_fprop(x, False, w, bias, 0, None, None, 0, False, torch.bfloat16) for x in xs],
        "fprop+stats": [lambda x=x, sm=sm: ops.conv_fprop(x, False, w, bias, 0, None, None, 0, False, torch.bfloat16, bn_sums=sm)
                        for x, sm in zip(xs, sums)],
        "fprop+stats(acc)": [lambda x=x: unconsumed(ops.conv_fprop_stats(x, w, bias)[2]) for x in xs],
        "dgrad+res": [lambda d=d, x=x: ops.conv_dgrad(d, False, w, x, torch.bfloat16) for d, x in zip(ds, xs)],
        "dgrad+bnred": [lambda d=d, y=y, st=st: unconsumed(ops.conv_dgrad_bnred(d, w, y, st, gamma, beta, alpha)[1])
                        for d, y, st in zip(ds, ys, stats)],
    }
    if fold == 2:
        cases["dgrad+res+bnred"] = [lambda d=d, y=y, st=st, x=x: unconsumed(ops.conv_dgrad_bnred(d, w, y, st, gamma, beta, None, residual=x)[1])
                                    for d, y, st, x in zip(ds, ys, stats, xs)]
    for name, fs in cases.items():
        ms = bench._time_replayed(fs)
        print("fold %d  %-12s %7.2f us  %7.1f TFLOP/s" % (fold, name, ms * 1e3, flop / (ms * 1e-3) / 1e12), flush=True)
L.call("srk_tc_probe", 12, out, 2)
print("err flag", out[0])
