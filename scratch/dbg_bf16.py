import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, 'food101-super-resolution_b200'); sys.path.insert(0, '.')
import torch, srk
from helpers import *
from oracle import sr_oracle as O
from src import models as M
from src.loss import get_loss_function
torch.manual_seed(1)
base = M.ResNetSR(num_channels=64, num_residuals=2)
sd = {k: v.detach().clone() for k, v in base.state_dict().items()}
lr, hr = O.synthetic_pair(4, 24, 24, 4, seed=8)
out_ref, _, grads_ref, _ = O.train_step_grads("RESNET", sd, lr, hr, "nlpd")
for impl in ("simt", "auto"):
    srk.set_compute_dtype("bf16"); srk.set_conv_impl(impl)
    m = M.ResNetSR(num_channels=64, num_residuals=2); m.load_state_dict(sd); m = m.cuda().train()
    out = m(lr.cuda()); get_loss_function("nlpd", "cuda")(out, hr.cuda()).backward()
    print(impl, "fwd max-rel %.3e rms-rel %.3e  mean(out-ref)=%.3e mean|ref|=%.3e" % (rel_err(out.cpu(), out_ref), rms_rel_err(out.cpu(), out_ref), (out.cpu() - out_ref).mean(), out_ref.abs().mean()))
    worst = sorted(((rel_err(p.grad.cpu(), grads_ref[k], floor=1e-4), k) for k, p in m.named_parameters()), reverse=True)[:8]
    print("   worst grads:", ["%s %.2e" % (k, e) for e, k in worst])
