import sys, math
sys.path.insert(0, 'tests'); sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
import torch.nn.functional as F
from srk import fn, _lib as L
from helpers import rel_err
srk.set_compute_dtype('bf16')
for (n, h, w) in ((2, 8, 8), (2, 24, 24), (4, 64, 64)):
    g = torch.Generator().manual_seed(h)
    cin = cout = 96
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    alpha = torch.tensor([0.25])
    xo, wo, bo, ao = [t.clone().requires_grad_(True) for t in (x, wt, b, alpha)]
    yo = F.prelu(F.conv2d(xo, wo, bo, padding=1), ao)
    go = torch.randn(yo.shape, generator=g).bfloat16().float()
    yo.backward(go)
    conv = torch.nn.Conv2d(cin, cout, 3, padding=1).cuda()
    with torch.no_grad(): conv.weight.copy_(wt); conv.bias.copy_(b)
    al = alpha.clone().cuda().requires_grad_(True)
    xg = x.cuda().requires_grad_(True)
    y = fn.ActToImage.apply(fn.conv_act(fn.ImageToAct.apply(xg, torch.bfloat16), conv, act=L.ACT_PRELU, alpha=al))
    y.backward(go.cuda())
    print((n, h, w), "y %.2e  dW %.2e  db %.2e  dx %.2e  dalpha %.2e" % (rel_err(y.cpu(), yo), rel_err(conv.weight.grad.cpu(), wo.grad), rel_err(conv.bias.grad.cpu(), bo.grad), rel_err(xg.grad.cpu(), xo.grad), rel_err(al.grad.cpu(), ao.grad)))
