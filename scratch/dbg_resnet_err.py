import sys, os
sys.path.insert(0, 'food101-super-resolution_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from oracle import sr_oracle as O
from helpers import rel_err, rms_rel_err
import test_gpu_parity as T
from src import models as M
for rep in range(3):
    torch.manual_seed(1)
    model = M.ResNetSR(num_channels=64, num_residuals=2)
    lr, hr = O.synthetic_pair(4, 24, 24, 4, seed=8)
    (e_max, e_rms), errs = T._bf16_vs_oracle("RESNET", model, lr, hr, "nlpd")
    top = sorted(((e, k) for k, e in errs.items() if not T._zero_grad_by_construction(k) and not k.endswith("prelu.weight") and k not in ("upsample.2.weight", "upsample.5.weight")), reverse=True)[:5]
    print("fwd %.4f %.4f | " % (e_max, e_rms) + " ".join("%s=%.4f" % (k, e) for e, k in top), flush=True)
