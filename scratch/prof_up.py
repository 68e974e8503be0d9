"""Upsample-conv shapes (64->256 + PixelShuffle, and its dgrad 256->64) at 64^2 and 128^2, C2 batch."""
import sys, os
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
from srk import ops, _lib as L
srk.set_compute_dtype('bf16')
dev = 'cuda'
N = 64
def act(c, h, w):
    t = torch.randn(N, h + 2, w + 2, c, device=dev).bfloat16()
    t[:, 0] = 0; t[:, -1] = 0; t[:, :, 0] = 0; t[:, :, -1] = 0
    return t
wt = torch.randn(256, 64, 3, 3, device=dev) / 24
b = torch.zeros(256, device=dev); alpha = torch.tensor([0.25], device=dev)
def t(f, n=10):
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for hw in (64, 128):
    x = act(64, hw, hw); dz = act(256, hw, hw)
    f1 = lambda: ops.conv_fprop(x, False, wt, b, L.ACT_PRELU, alpha, None, 2, False, torch.bfloat16)
    f2 = lambda: ops.conv_dgrad(dz, False, wt, None, torch.bfloat16)
    f3 = lambda: ops.conv_wgrad(x, False, dz, False, wt, True)
    print("%3d^2: fprop 64->256+shuffle %.1f us | dgrad 256->64 %.1f us | wgrad %.1f us" % (hw, t(f1), t(f2), t(f3)), flush=True)
