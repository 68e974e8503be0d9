import sys, os, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "food101-super-resolution_b200"))
import torch
from srk import _lib as L
torch.zeros(1, device="cuda")
out = (ctypes.c_float * 2)()
for nw in (4, 8):
    for batch in (1, 2, 4):
        L.call("srk_tc_probe", 2000 + nw + 100 * batch, out, 2)
        print("ldtm warps=%d batch=%d: %.1f cycles per 4KB ld per warp, %.1f B/cycle/SM" % (nw, batch, out[0], out[1]))
for n in (64, 128, 192, 256):
    L.call("srk_tc_probe", 1000 + n // 8, out, 2)
    print("mma N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f)" % (n, out[0], out[1], 128 * n / 256))
