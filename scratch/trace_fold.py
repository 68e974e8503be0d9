import sys, os, ctypes
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
from srk import ops, _lib as L
srk.set_compute_dtype('bf16')
dev = 'cuda'
x = torch.randn(64, 66, 66, 64, device=dev).bfloat16()
x[:, 0] = 0; x[:, -1] = 0; x[:, :, 0] = 0; x[:, :, -1] = 0
wt = (torch.randn(64, 64, 3, 3, device=dev) / 24); b = torch.zeros(64, device=dev)
for _ in range(3): ops.conv_fprop(x, False, wt, b, 0, None, None, 0, False, torch.bfloat16)
torch.cuda.synchronize()
out = (ctypes.c_float * 512)()
L.call("srk_tc_probe", 100, out, 512)
ops.conv_fprop(x, False, wt, b, 0, None, None, 0, False, torch.bfloat16)
L.call("srk_tc_probe", 102, out, 512)
names = ['prod', 'mma0', 'mma1', 'e_top', 'e_tfull', 'e_ld1', 'e_ld2', 'e_shf', 'e_bar', 'e_ofree', 'e_stg', 'e_end', 'st_go', 'st_rd']
print("tile " + " ".join("%8s" % n for n in names))
for i in range(16):
    print("%4d " % i + " ".join("%8.0f" % out[r * 32 + i] for r in range(len(names))))
