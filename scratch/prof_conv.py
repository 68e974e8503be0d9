"""Kernel timing at the C2 layer shape, CPU launch overhead removed by CUDA-graph replay."""
import sys, os
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
from srk import ops
srk.set_compute_dtype('bf16')
dev = 'cuda'
N = int(os.environ.get('NIMG', '64'))
x = torch.randn(N, 66, 66, 64, device=dev).bfloat16()
x[:, 0] = 0; x[:, -1] = 0; x[:, :, 0] = 0; x[:, :, -1] = 0
dz = x.flip(0).contiguous()
wt = (torch.randn(64, 64, 3, 3, device=dev) / 24)
b = torch.zeros(64, device=dev)
sums = torch.zeros(2, 64, device=dev)
fns = {
    'fprop': lambda: ops.conv_fprop(x, False, wt, b, 0, None, None, 0, False, torch.bfloat16),
    'fprop+stats': lambda: ops.conv_fprop(x, False, wt, b, 0, None, None, 0, False, torch.bfloat16, bn_sums=sums),
    'dgrad+res': lambda: ops.conv_dgrad(dz, False, wt, x, torch.bfloat16),
    'wgrad': lambda: ops.conv_wgrad(x, False, dz, False, wt, True),
    'wgrad_nobias': lambda: ops.conv_wgrad(x, False, dz, False, wt, False),
}
only = os.environ.get('ONLY')
if only:
    fns = {k: f for k, f in fns.items() if k in only.split(',')}
for f in fns.values():
    for _ in range(2): f()
torch.cuda.synchronize()
def t(f, n=20):
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        f()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): f()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(' | '.join("%s %.2f us" % (k, 1e3 * t(f)) for k, f in fns.items() if not only or k in only.split(',')), flush=True)
