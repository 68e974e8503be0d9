import sys, math
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
import torch.nn.functional as F
from srk import ops
srk.set_compute_dtype('bf16')
g = torch.Generator().manual_seed(1)
n, c, h, w = 2, 96, 8, 8
x = torch.randn(n, c, h, w, generator=g).bfloat16().float()
dz = torch.randn(n, c, h, w, generator=g).bfloat16().float()
wt = torch.randn(c, c, 3, 3, generator=g) * 0.05
xo = x.clone(); wo = wt.clone().requires_grad_(True)
F.conv2d(xo, wo, None, padding=1).backward(dz)
ref = wo.grad
xa = ops.image_to_act(x.cuda(), torch.bfloat16); dza = ops.image_to_act(dz.cuda(), torch.bfloat16)
dw, db = ops.conv_wgrad(xa, False, dza, False, wt.cuda(), True)
dw = dw.cpu()
print("db err", (db.cpu() - dz.sum((0, 2, 3))).abs().max().item(), "db ref max", dz.sum((0,2,3)).abs().max().item())
for co0 in (0, 64):
    for ci0 in (0, 64):
        a = dw[co0:co0+64, ci0:ci0+64]; b = ref[co0:co0+64, ci0:ci0+64]
        print("block co%d ci%d: max err %.3f  ref max %.3f  got max %.3f" % (co0, ci0, (a-b).abs().max(), b.abs().max(), a.abs().max()))
