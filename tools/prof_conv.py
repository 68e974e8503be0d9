"""A few launches of the trunk conv (fprop, fprop + stats, dgrad + residual) per variant, for ncu."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200")); sys.path.insert(0, ROOT)
import torch
import srk
from srk import _lib as L, ops
srk.set_compute_dtype("bf16")
dev = torch.device("cuda:0")
B, H = int(os.environ.get("B", 64)), int(os.environ.get("H", 64))
g = torch.Generator(device=dev).manual_seed(1)
def act():
    t = torch.zeros((B, H + 2, H + 2, 64), dtype=torch.bfloat16, device=dev)
    t[:, 1:-1, 1:-1] = torch.randn((B, H, H, 64), generator=g, device=dev).bfloat16()
    return t
xs = [act() for _ in range(4)]
w = torch.randn((64, 64, 3, 3), generator=g, device=dev) / 24
bias = torch.randn((64,), generator=g, device=dev) * 0.1
sums = torch.empty((2, 64), dtype=torch.float32, device=dev)
out = (ctypes.c_float * 2)()
for fold in [int(v) for v in os.environ.get("FOLDS", "2,4").split(",")]:
    L.call("srk_tc_probe", 10 + fold, out, 2)
    for rep in range(2):
        for x in xs:
            ops.conv_fprop(x, False, w, bias, 0, None, None, 0, False, torch.bfloat16)
        ops.conv_fprop(xs[0], False, w, bias, 0, None, None, 0, False, torch.bfloat16, bn_sums=sums)
        ops.conv_dgrad(xs[1], False, w, xs[2], torch.bfloat16)
torch.cuda.synchronize()
