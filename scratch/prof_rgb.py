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
u = act(64, 256, 256)
dy = torch.randn(N, 3, 256, 256, device=dev)
w_out = torch.randn(3, 64, 9, 9, device=dev) * 0.01
x_lr = torch.rand(N, 3, 64, 64, device=dev)
w_in = torch.randn(64, 3, 9, 9, device=dev) * 0.05
b_in = torch.zeros(64, device=dev); alpha = torch.tensor([0.25], device=dev)
dz_lr = act(64, 64, 64)
b_out = torch.zeros(3, device=dev)
fns = {
  'outconv fprop': lambda: ops.conv_fprop(u, False, w_out, b_out, 0, None, None, 0, True, torch.float32),
  'outconv bwd (dx+dw)': lambda: ops.conv_rgbout_bwd(u, dy, w_out, True, True),
  'outconv bwd (dw only)': lambda: ops.conv_rgbout_bwd(u, dy, w_out, False, True),
  'inconv fprop': lambda: ops.conv_fprop(x_lr, True, w_in, b_in, 2, alpha, None, 0, False, torch.bfloat16),
  'inconv wgrad': lambda: ops.conv_wgrad(x_lr, True, dz_lr, False, w_in, True),
}
def t(f, n=5):
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for k, f in fns.items():
    print("%-24s %8.1f us" % (k, t(f) * 1e3), flush=True)
