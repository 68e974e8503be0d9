"""Generates tests/golden/*.npz from the REFERENCE modules (run in the build container, where
/root/reference exists):   python oracle/make_golden.py

Each fixture holds the seeded inputs, the reference state_dict, and what the reference computed from
them: training-mode output, loss, all parameter gradients, BN buffers after the step, eval-mode
output.  tests/test_oracle_golden.py pins oracle/sr_oracle.py against these files; the GPU tests pin
libsrk against them directly.  Test infrastructure only."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def make_case(name, ctor, arch, loss_name, n, h, w, scale, seed):
    import src.loss as rloss
    torch.manual_seed(seed)
    model = ctor()
    # perturb BN affine / PReLU / biases away from their trivial init so every gradient path is exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    gi = torch.Generator().manual_seed(seed + 2)
    hr = torch.rand((n, 3, h * scale, w * scale), generator=gi)
    lr = torch.nn.functional.interpolate(hr, size=(h, w), mode="bicubic", align_corners=False, antialias=True)
    crit = rloss.get_loss_function(loss_name, "cpu")
    model.train()
    out = model(lr)
    loss = crit(out, hr)
    loss.backward()
    fix = {"lr": _np(lr), "hr": _np(hr), "out_train": _np(out), "loss": _np(loss)}
    for k, v in sd0.items():
        fix["sd/" + k] = _np(v)
    for k, p in model.named_parameters():
        fix["grad/" + k] = _np(p.grad)
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            fix["after/" + k] = _np(v)
    model.load_state_dict(sd0)
    model.eval()
    with torch.no_grad():
        fix["out_eval"] = _np(model(lr))
    fix["meta"] = np.array([arch, loss_name, str(scale)])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **fix)
    print(name, "loss=%.6f" % loss.item(), "%.1f KB" % (os.path.getsize(path) / 1024))


def make_loss_case():
    import src.loss as rloss
    g = torch.Generator().manual_seed(77)
    fix = {}
    for tag, shape in (("even", (2, 3, 32, 48)), ("odd", (2, 3, 25, 37)), ("native", (1, 3, 200, 200))):
        sr = torch.rand(shape, generator=g).requires_grad_(True)
        hr = torch.rand(shape, generator=g)
        fix[tag + "/sr"], fix[tag + "/hr"] = _np(sr), _np(hr)
        for lname in ("mae", "mse", "nlpd"):
            sr.grad = None
            loss = rloss.get_loss_function(lname, "cpu")(sr, hr)
            loss.backward()
            fix["%s/%s/loss" % (tag, lname)] = _np(loss)
            if tag != "native":
                fix["%s/%s/grad" % (tag, lname)] = _np(sr.grad)
            else:
                fix["%s/%s/grad_sum_abs" % (tag, lname)] = _np(sr.grad.abs().sum())
    fix["kernel"] = _np(rloss.NLPDLoss().kernel)
    path = os.path.join(OUT, "losses.npz")
    np.savez_compressed(path, **fix)
    print("losses", "%.1f KB" % (os.path.getsize(path) / 1024))


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not found at %s" % REF)
    sys.path.insert(0, REF)
    import src.models as rm
    os.makedirs(OUT, exist_ok=True)
    make_case("srcnn_x2", lambda: rm.SRCNN(scale_factor=2, hidden_dim=64), "SRCNN", "nlpd", 2, 12, 12, 2, 11)
    make_case("resnet_c32_b2", lambda: rm.ResNetSR(num_channels=32, num_residuals=2), "RESNET", "nlpd", 2, 8, 8, 4, 22)
    make_case("attn_c32_b2", lambda: rm.AttentionSR(num_channels=32, num_residuals=2), "AttentionSR", "mae", 2, 8, 8, 4, 33)
    make_loss_case()
    # seeded-init parity: first/last parameters of the reference's get_model() builds under seed 0
    fix = {}
    for arch in ("SRCNN", "RESNET", "AttentionSR"):
        torch.manual_seed(0)
        m = rm.get_model(arch, scale_factor=4)
        sd = m.state_dict()
        fix[arch + "/keys"] = np.array(list(sd.keys()))
        fix[arch + "/shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])
        fix[arch + "/dtypes"] = np.array([str(v.dtype) for v in sd.values()])
        fix[arch + "/checksum"] = np.array([float(v.double().sum()) for v in sd.values()])
    np.savez_compressed(os.path.join(OUT, "state_dicts.npz"), **fix)
    print("state_dicts ok")


if __name__ == "__main__":
    main()
