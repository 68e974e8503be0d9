import sys, os
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
from srk import ops
srk.set_compute_dtype('bf16')
dev = 'cuda'
def act():
    t = torch.randn(64, 66, 66, 64, device=dev).bfloat16()
    t[:, 0] = 0; t[:, -1] = 0; t[:, :, 0] = 0; t[:, :, -1] = 0
    return t
dz, z = act(), act()
wt = torch.randn(64, 64, 3, 3, device=dev) / 24
gamma = torch.rand(64, device=dev) + 0.5; beta = torch.randn(64, device=dev) * 0.1
alpha = torch.tensor([0.25], device=dev)
_, stats = ops.bn_forward(z, gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, None)
def t(f, n=20):
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def unfused():
    dx = ops.conv_dgrad(dz, False, wt, None, torch.bfloat16)
    return ops.bn_backward(dx, z, stats, gamma, beta, alpha, True)
def fused():
    dx, red = ops.conv_dgrad_bnred(dz, wt, z, stats, gamma, beta, alpha)
    return ops.bn_backward(dx, z, stats, gamma, beta, alpha, True, pre=red)
print("dgrad + BN backward: unfused %.1f us | fused %.1f us | dgrad alone %.1f | dgrad+bnred alone %.1f" % (
    t(unfused), t(fused), t(lambda: ops.conv_dgrad(dz, False, wt, None, torch.bfloat16)),
    t(lambda: ops.conv_dgrad_bnred(dz, wt, z, stats, gamma, beta, alpha))))
