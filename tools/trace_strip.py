"""Per-tile timeline of CTA 0 of a trunk conv kernel (clock64 stamps per role; srk_tc_probe 100 / 102) on a library
built with `build.py --probes --debug-knobs` (SRK_LIB=...).  FOLD=4 (default): column-strip kernel; FOLD=2: per-tap
halo-slab kernel (general instantiation, which carries the stamps)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200")); sys.path.insert(0, ROOT)
import torch
import srk
from srk import _lib as L, ops
srk.set_compute_dtype("bf16")
dev = torch.device("cuda:0")
B, H = int(os.environ.get("B", 64)), 64
g = torch.Generator(device=dev).manual_seed(1)
x = torch.zeros((B, H + 2, H + 2, 64), dtype=torch.bfloat16, device=dev)
x[:, 1:-1, 1:-1] = torch.randn((B, H, H, 64), generator=g, device=dev).bfloat16()
w = torch.randn((64, 64, 3, 3), generator=g, device=dev) / 24
out = (ctypes.c_float * 2)()
FOLD = int(os.environ.get("FOLD", 4))
L.call("srk_tc_probe", 10 + FOLD, out, 2)
for _ in range(3):
    ops.conv_fprop(x, False, w, None, 0, None, None, 0, False, torch.bfloat16)
torch.cuda.synchronize()
buf = (ctypes.c_float * (16 * 32))()
L.call("srk_tc_probe", 100, buf, 16 * 32)
ops.conv_fprop(x, False, w, None, 0, None, None, 0, False, torch.bfloat16)
L.call("srk_tc_probe", 102, buf, 16 * 32)
names = {0: "tma issue", 1: "mma: loop top", 2: "mma: slot free", 3: "mma: slab full", 4: "mma: issued", 5: "epi: wait tfull",
         6: "epi: tfull", 7: "epi: staged", 8: "store: oready", 9: "store: read out"}
if FOLD != 4:   # stamps of fold::conv3x3_fold_tc_kernel
    names = {0: "tma: slab issued", 1: "mma: slab full", 2: "mma: tile issued", 3: "epi: wait acc", 4: "epi: acc full",
             5: "epi: acc in regs", 7: "epi: values done", 8: "epi: wait ofree", 10: "epi: ofree", 9: "epi: residual in",
             11: "epi: tile staged",
             12: "store: tile ready", 13: "store: smem read"}
print("FOLD=%d  B=%d  clock cycles since the first stamp, CTA 0, first 20 tiles" % (FOLD, B))
for r, nm in names.items():
    print("%-16s" % nm, " ".join("%6d" % int(buf[r * 32 + i]) for i in range(20)))
L.call("srk_tc_probe", 12, out, 2)
