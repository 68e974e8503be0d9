import sys
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
from src.loss import get_loss_function
dev = 'cuda'
crit = get_loss_function('nlpd', dev)
sr = torch.rand(64, 3, 256, 256, device=dev, requires_grad=True); hr = torch.rand(64, 3, 256, 256, device=dev)
def f():
    sr.grad = None
    crit(sr, hr).backward()
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
import os
for mode in ("0", "1", "0", "1"):
    os.environ["SRK_NLPD_2X"] = mode
    print("NLPD fwd+bwd at 64x3x256x256, SRK_NLPD_2X=%s: %.1f us" % (mode, t(f)))
