"""Pins oracle/sr_oracle.py (the CPU restatement) against fixtures produced by the reference modules
themselves (oracle/make_golden.py).  CPU only."""
import math

import pytest
import torch

from helpers import golden_state_dict, load_golden, max_abs, rel_err
from oracle import sr_oracle as O

CASES = [("srcnn_x2", 1e-5), ("resnet_c32_b2", 1e-4), ("attn_c32_b2", 1e-5)]


@pytest.mark.parametrize("name,tol", CASES)
def test_model_forward_backward_matches_reference(name, tol):
    fix = load_golden(name)
    arch, loss_name, scale = [str(x) for x in fix["meta"]]
    sd = golden_state_dict(fix)
    lr, hr = torch.from_numpy(fix["lr"]), torch.from_numpy(fix["hr"])
    out, loss, grads, work = O.train_step_grads(arch, sd, lr, hr, loss_name, scale_factor=int(scale))
    assert max_abs(out, torch.from_numpy(fix["out_train"])) <= tol
    assert abs(loss.item() - float(fix["loss"])) <= 1e-5
    n_checked = 0
    for k, v in fix.items():
        if k.startswith("grad/"):
            assert rel_err(grads[k[5:]], torch.from_numpy(v)) <= 1e-3, k
            n_checked += 1
    assert n_checked == len(grads)
    for k, v in fix.items():
        if k.startswith("after/"):
            assert max_abs(work[k[6:]].detach(), torch.from_numpy(v)) <= 1e-5, k
    with torch.no_grad():
        out_eval = O.model_forward(arch, sd, lr, training=False, scale_factor=int(scale))
    assert max_abs(out_eval, torch.from_numpy(fix["out_eval"])) <= tol


@pytest.mark.parametrize("tag", ["even", "odd", "native"])
@pytest.mark.parametrize("lname", ["mae", "mse", "nlpd"])
def test_losses_match_reference(tag, lname):
    fix = load_golden("losses")
    sr = torch.from_numpy(fix[tag + "/sr"]).requires_grad_(True)
    hr = torch.from_numpy(fix[tag + "/hr"])
    loss = O.loss_fn(lname)(sr, hr)
    loss.backward()
    assert abs(loss.item() - float(fix["%s/%s/loss" % (tag, lname)])) <= 1e-6
    if tag != "native":
        assert max_abs(sr.grad, torch.from_numpy(fix["%s/%s/grad" % (tag, lname)])) <= 1e-8
    else:
        assert abs(sr.grad.abs().sum().item() - float(fix["%s/%s/grad_sum_abs" % (tag, lname)])) <= 1e-4


def test_gaussian_kernel_matches_reference_buffer():
    fix = load_golden("losses")
    assert max_abs(O.gaussian_kernel_5x5(3), torch.from_numpy(fix["kernel"])) <= 1e-8


# ---- metrics: torchmetrics restatement, analytic known answers (parity unpinned) ----------
def test_psnr_known_answers():
    x = torch.rand(2, 3, 16, 16) * 0.8
    assert abs(O.psnr(x, x + 0.1) - 20.0) < 1e-5
    assert O.psnr(x, x) == float("inf")


def test_ssim_known_answers():
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 24, 20, generator=g)
    assert abs(O.ssim(x, x) - 1.0) < 1e-12
    a, b = torch.full((1, 3, 16, 16), 0.5), torch.full((1, 3, 16, 16), 0.6)
    expect = (2 * 0.5 * 0.6 + 1e-4) / (0.25 + 0.36 + 1e-4)
    assert abs(O.ssim(a, b) - expect) < 1e-6
    assert abs(expect - 0.983609) < 1e-6


def test_ssim_equals_valid_window_mean():
    """The reflect padding is cropped away again: SSIM == mean over the (H-10)x(W-10) valid windows."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(4)
    p, t = torch.rand(1, 1, 20, 22, generator=g).double(), torch.rand(1, 1, 20, 22, generator=g).double()
    d = torch.arange(-5, 6, dtype=torch.float64)
    k = torch.exp(-((d / 1.5) ** 2) / 2)
    k = (k / k.sum())
    k2 = (k[:, None] * k[None, :]).view(1, 1, 11, 11)
    mu_p, mu_t = F.conv2d(p, k2), F.conv2d(t, k2)
    s_pp = (F.conv2d(p * p, k2) - mu_p ** 2).clamp(min=0)
    s_tt = (F.conv2d(t * t, k2) - mu_t ** 2).clamp(min=0)
    s_pt = F.conv2d(p * t, k2) - mu_p * mu_t
    m = ((2 * mu_p * mu_t + 1e-4) * (2 * s_pt + 9e-4)) / ((mu_p ** 2 + mu_t ** 2 + 1e-4) * (s_pp + s_tt + 9e-4))
    assert abs(m.mean().item() - O.ssim(p, t)) < 1e-12


@pytest.mark.parametrize("shape", [(2, 3, 24, 20), (3, 1, 37, 53), (1, 3, 64, 64), (2, 3, 11 + 1, 11 + 6)])
def test_psnr_ssim_agree_with_independent_fp64_implementation(shape):
    """oracle/metrics_fp64.py shares no code with sr_oracle.py (numpy + scipy.ndimage.correlate1d, written from the
    SSIM paper's formulas): both must give the same PSNR and per-image SSIM, on smooth and on noisy image pairs."""
    from oracle import metrics_fp64 as M64
    g = torch.Generator().manual_seed(shape[2] * 7 + shape[3])
    hr = torch.rand(shape, generator=g)
    for noise in (0.02, 0.3):
        sr = (hr + noise * torch.randn(shape, generator=g)).clamp(0, 1)
        assert abs(O.psnr(sr, hr) - M64.psnr(sr.numpy(), hr.numpy())) <= 1e-9
        a, b = O.ssim_per_image(sr, hr).numpy(), M64.ssim_per_image(sr.numpy(), hr.numpy())
        assert abs(a - b).max() <= 1e-12, (a, b)
    # smooth content (what SR images look like): low-passed noise
    lp = torch.nn.functional.avg_pool2d(hr, 3, 1, 1)
    assert abs(O.ssim(lp, hr) - M64.ssim(lp.numpy(), hr.numpy())) <= 1e-12
    assert M64.psnr(hr.numpy(), hr.numpy()) == float("inf") and abs(M64.ssim(hr.numpy(), hr.numpy()) - 1.0) <= 1e-12


def test_metrics_compute_clamps_first():
    x = torch.full((1, 3, 16, 16), 1.5)
    y = torch.full((1, 3, 16, 16), 1.0)
    m = O.metrics_compute(x, y)
    assert m["psnr"] == float("inf") and abs(m["ssim"] - 1.0) < 1e-9 and m["nlpd"] == 0.0
    assert math.isfinite(O.metrics_compute(x * 0.3, y * 0.5)["psnr"])


def test_perceptual_oracle_matches_torchvision_vgg19_slice():
    """The reference's PerceptualLoss wraps torchvision vgg19().features[:35] (loss.py:23-24); it cannot be built
    offline (the constructor downloads the ImageNet checkpoint), so the oracle's restatement of that slice is pinned
    against the torchvision module itself with seeded random weights, forward and input gradient."""
    from torchvision.models import vgg19
    torch.manual_seed(11)
    net = vgg19(weights=None).features[:35].eval()
    sd = {"vgg." + k: v for k, v in net.state_dict().items()}
    sr = torch.rand(2, 3, 40, 36, requires_grad=True)
    hr = torch.rand(2, 3, 40, 36)
    ref = torch.nn.functional.mse_loss(net(sr), net(hr))
    (g_ref,) = torch.autograd.grad(ref, sr)
    sr2 = sr.detach().clone().requires_grad_(True)
    got = O.perceptual_loss(sd, sr2, hr)
    (g_got,) = torch.autograd.grad(got, sr2)
    assert abs(got.item() - ref.item()) <= 1e-7 * max(1.0, abs(ref.item()))
    assert max_abs(g_got, g_ref) <= 1e-9
    assert len([k for k in sd if k.endswith(".weight")]) == 16
