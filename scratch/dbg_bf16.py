import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, 'food101-super-resolution_b200'); sys.path.insert(0, '.')
import torch, srk
from helpers import *
from test_gpu_parity import _build
from src.loss import get_loss_function
srk.set_compute_dtype('bf16')
for name in ['resnet_c32_b2', 'attn_c32_b2']:
    fix = load_golden(name)
    arch, loss_name, scale = [str(x) for x in fix['meta']]
    model, _ = _build(arch, fix, int(scale))
    lr, hr = torch.from_numpy(fix['lr']).cuda(), torch.from_numpy(fix['hr']).cuda()
    model.train()
    out = model(lr); loss = get_loss_function(loss_name, 'cuda')(out, hr); loss.backward()
    print(name, 'fwd', rel_err(out.cpu(), torch.from_numpy(fix['out_train'])))
    for k, p in model.named_parameters():
        ref = torch.from_numpy(fix['grad/'+k])
        print('  %-40s err %.3e  |ref|max %.3e' % (k, rel_err(p.grad.cpu(), ref, floor=1e-4), ref.abs().max()))
