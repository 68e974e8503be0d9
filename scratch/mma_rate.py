import sys, ctypes
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
from srk import _lib as L
torch.zeros(1, device='cuda')
out = (ctypes.c_float * 2)()
for mn in (0, 1):
    for n in (64, 128, 256):
        for off in (0, 3):
            L.call("srk_tc_probe", 1000 + n // 8 + 100 * off + 10000 * mn, out, 2)
            print("mn_major=%d N=%3d a_row_off=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f)" % (mn, n, off, out[0], out[1], 128 * n / 256))
