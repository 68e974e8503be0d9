import sys, os
sys.path.insert(0, 'food101-super-resolution_b200'); sys.path.insert(0, '.')
import torch, torch.nn.functional as F
import srk
from src.loss import PerceptualLoss
from oracle import sr_oracle as O
DEV = 'cuda:0'
srk.set_compute_dtype('bf16')
torch.manual_seed(5)
crit = PerceptualLoss(DEV, weights=None)
sd = {k: v.detach().cpu() for k, v in crit.state_dict().items()}
def feats_emul(sd, x):
    # oracle with bf16 storage between layers (weights rounded to bf16 as the tensor-core path does)
    r = lambda t: t.bfloat16().float()
    first = True
    for item in O.VGG19_35:
        if item == "M":
            x = F.max_pool2d(x, 2, 2); continue
        idx, _ = item
        w = sd["vgg.%d.weight" % idx]; b = sd["vgg.%d.bias" % idx]
        x = F.conv2d(x, w if first else r(w), b, padding=1)
        first = False
        if idx != 34: x = F.relu(x)
        x = r(x)
    return x
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
def rms(a, b): return ((a - b).norm() / b.norm()).item()
g = torch.Generator().manual_seed(17)
for name, mk in (("near", lambda sr: (sr + 0.1 * torch.randn(sr.shape, generator=g)).clamp(0, 1)), ("indep", lambda sr: torch.rand(sr.shape, generator=g))):
    sr = torch.rand(2, 3, 64, 64, generator=g); hr = mk(sr)
    so = sr.clone().requires_grad_(True)
    fo = O.vgg19_features35(sd, so); lo = F.mse_loss(fo, O.vgg19_features35(sd, hr)); (go,) = torch.autograd.grad(lo, so)
    se = sr.clone().requires_grad_(True)
    fe = feats_emul(sd, se); le = F.mse_loss(fe, feats_emul(sd, hr)); (ge,) = torch.autograd.grad(le, se)
    sg = sr.to(DEV).requires_grad_(True)
    fg = crit.features(sg); lg = crit.loss(fg, crit.features(hr.to(DEV))); lg.backward()
    print(name, "loss: srk %.6e oracle %.6e emul %.6e" % (lg.item(), lo.item(), le.item()))
    print("  features: srk vs oracle max-rel %.3e rms %.3e | srk vs bf16-emulated max-rel %.3e rms %.3e | emul vs oracle rms %.3e" % (
        rel(fg.detach().cpu(), fo.detach()), rms(fg.detach().cpu(), fo.detach()), rel(fg.detach().cpu(), fe.detach()), rms(fg.detach().cpu(), fe.detach()), rms(fe.detach(), fo.detach())))
    print("  grad:     srk vs oracle max-rel %.3e rms %.3e | srk vs bf16-emulated max-rel %.3e rms %.3e | emul vs oracle max-rel %.3e rms %.3e" % (
        rel(sg.grad.cpu(), go), rms(sg.grad.cpu(), go), rel(sg.grad.cpu(), ge), rms(sg.grad.cpu(), ge), rel(ge, go), rms(ge, go)))
