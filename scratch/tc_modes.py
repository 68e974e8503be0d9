"""Bring-up: tcgen05 conv vs the CPU oracle for each A-staging mode + timing on the C2 layer shape."""
import sys, ctypes, math
sys.path.insert(0, 'tests'); sys.path.insert(0, 'food101-super-resolution_b200'); sys.path.insert(0, '.')
import torch, srk
import torch.nn.functional as F
from srk import _lib as L, ops, fn
from helpers import rel_err

def probe(mode):
    out = (ctypes.c_float * 2)()
    L.call("srk_tc_probe", mode, out, 2)
    return int(out[0]), int(out[1])

srk.set_compute_dtype('bf16')
dev = 'cuda'
cases = [(64, 64, 2, 10, 20, 'none', 0, False), (64, 64, 3, 16, 16, 'prelu', 0, True), (64, 256, 2, 8, 6, 'prelu', 2, False),
         (256, 64, 2, 9, 7, 'none', 0, False), (64, 64, 1, 64, 64, 'relu', 0, False), (64, 128, 2, 12, 33, 'none', 0, True)]
for mode in (0, 1, 2):
    probe(mode)
    for cin, cout, n, h, w, act, shuffle, use_res in cases:
        g = torch.Generator().manual_seed(cin + cout + h)
        x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
        wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).bfloat16().float()
        b = torch.randn(cout, generator=g) * 0.1
        yo = F.conv2d(x, wt, b, padding=1)
        if shuffle: yo = F.pixel_shuffle(yo, 2)
        if act == 'prelu': yo = F.prelu(yo, torch.tensor([0.25]))
        if act == 'relu': yo = F.relu(yo)
        res = torch.randn(yo.shape, generator=g).bfloat16().float() if use_res else None
        if use_res: yo = yo + res
        xa = ops.image_to_act(x.to(dev), torch.bfloat16)
        ra = ops.image_to_act(res.to(dev), torch.bfloat16) if use_res else None
        code = {'none': 0, 'relu': 1, 'prelu': 2}[act]
        if use_res and act != 'none':
            continue
        al = torch.tensor([0.25], device=dev)
        y, used = ops.conv_fprop(xa, False, wt.to(dev), b.to(dev), code, al if act == 'prelu' else None, ra, shuffle, False, torch.bfloat16)
        flag, m = probe(-1)
        yi = ops.act_to_image(y).cpu()
        border = max(float(y[:, 0].abs().max()), float(y[:, -1].abs().max()), float(y[:, :, 0].abs().max()), float(y[:, :, -1].abs().max()))
        print("mode %d tc=%s case %s: rel err %.3e  border %.1e  errflag %d" % (m, used, (cin, cout, n, h, w, act, shuffle, use_res), rel_err(yi, yo), border, flag), flush=True)
        # dgrad through the same kernel
    # timing on the C2 layer: [64, 64, 64, 64]
    x = torch.randn(64, 66, 66, 64, device=dev).bfloat16()
    x[:, 0] = 0; x[:, -1] = 0; x[:, :, 0] = 0; x[:, :, -1] = 0
    wt = (torch.randn(64, 64, 3, 3, device=dev) / 24)
    b = torch.zeros(64, device=dev)
    for _ in range(3): ops.conv_fprop(x, False, wt, b, 0, None, None, 0, False, torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.conv_fprop(x, False, wt, b, 0, None, None, 0, False, torch.bfloat16)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("mode %d: conv3x3 64->64 [64,64,64] %.4f ms  %.1f TFLOP/s  errflag %d" % (mode, ms, 19.327e9 / ms / 1e9, probe(-1)[0]), flush=True)
