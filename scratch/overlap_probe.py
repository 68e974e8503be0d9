"""How well does the side-stream wgrad overlap the BN-backward passes?  C2 layer shape, CUDA-graph replay."""
import sys, os
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
from srk import ops, _lib as L
srk.set_compute_dtype('bf16')
dev = 'cuda'
N = 64
def act(c=64, h=64, w=64):
    t = torch.randn(N, h + 2, w + 2, c, device=dev).bfloat16()
    t[:, 0] = 0; t[:, -1] = 0; t[:, :, 0] = 0; t[:, :, -1] = 0
    return t
# several distinct buffers so that consecutive layers do not hit the same L2 lines
LAYERS = 6
xs = [act() for _ in range(LAYERS)]; ys = [act() for _ in range(LAYERS)]; douts = [act() for _ in range(LAYERS)]
wt = torch.randn(64, 64, 3, 3, device=dev) / 24
gamma = torch.rand(64, device=dev) + 0.5; beta = torch.randn(64, device=dev) * 0.1
alpha = torch.tensor([0.25], device=dev)
_, stats = ops.bn_forward(ys[0], gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, None)
side = torch.cuda.Stream()

def layer(i, mode):
    dz, _, _, _ = ops.bn_backward(douts[i], ys[i], stats, gamma, beta, alpha, True)
    if 'D' in mode:
        dx = ops.conv_dgrad(dz, False, wt, None, torch.bfloat16)
    if 'W' in mode:
        if 'o' in mode:
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
            with torch.cuda.stream(side):
                ops.conv_wgrad(xs[i], False, dz, False, wt, True)
        else:
            ops.conv_wgrad(xs[i], False, dz, False, wt, True)
    return dz

def run(mode):
    keep = []
    for i in range(LAYERS):
        keep.append(layer(i, mode))
    if 'o' in mode:
        torch.cuda.current_stream().wait_stream(side)
    return keep

def t(mode):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        run(mode); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            k = run(mode)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * LAYERS) * 1e3

for mode, what in (('B', 'BN bwd (reduce+apply)'), ('BD', 'BN + dgrad'), ('BW', 'BN + wgrad sequential'), ('BWo', 'BN + wgrad overlapped (wgrad beside next BN)'),
                   ('BDW', 'BN + dgrad + wgrad sequential'), ('BDWo', 'BN + dgrad + wgrad overlapped')):
    print("%-50s %7.1f us / layer" % (what, t(mode)), flush=True)
