"""PSNR / SSIM / NLPD metrics on libsrk.  Drop-in for the reference's src/metrics.py
(reference metrics.py:6-31): MetricsCalculator(device).compute(sr, hr) -> dict of python floats.

The reference delegates to torchmetrics 1.8.2 (PeakSignalNoiseRatio / StructuralSimilarityIndexMeasure,
data_range=1) and lpips 0.1.4.  PSNR and SSIM are restated as CUDA reductions with per-image partial
sums (srk_psnr_sse / srk_ssim); LPIPS needs downloaded AlexNet weights and is reported as NaN unless
an `lpips_fn` callable is supplied."""
import math
import os

import torch

from srk import _lib as L
from srk import ops
from src.loss import NLPDLoss


def psnr_ssim_sums(sr, hr, clamp=True):
    """Per-image partial results on the device: (sse[N] float64, ssim_sum[N] float64).
    sse = sum of squared error per image; ssim_sum = sum of the SSIM map over the (H-10)x(W-10) valid
    11x11 Gaussian windows and all channels."""
    ops.require_cuda(sr, "metrics")
    ops.require_cuda(hr, "metrics")
    if sr.shape != hr.shape or sr.dim() != 4:
        raise ValueError("metrics: expected two NCHW tensors of equal shape")
    sr = sr.detach().contiguous().float()
    hr = hr.detach().contiguous().float()
    n, c, h, w = sr.shape
    out = torch.empty((2, n), dtype=torch.float64, device=sr.device)
    st = ops.stream_ptr()
    L.call("srk_psnr_sse", sr.data_ptr(), hr.data_ptr(), n, c * h * w, 1 if clamp else 0, out[0].data_ptr(), st)
    L.call("srk_ssim", sr.data_ptr(), hr.data_ptr(), n, c, h, w, 1 if clamp else 0, out[1].data_ptr(), st)
    return out[0], out[1]


def psnr_from_sse(sse_total, numel, data_range=1.0):
    """torchmetrics PSNR, dim=None, base 10: 10 log10(range^2 / mse) over the whole batch tensor."""
    mse = sse_total / numel
    if mse == 0.0:
        return float("inf")
    return 10.0 * math.log10(data_range * data_range / mse)


class MetricsCalculator:
    def __init__(self, device, lpips_fn=None):
        self.device = device
        if lpips_fn is None and os.environ.get("SRK_LPIPS_WEIGHTS"):
            # lpips.LPIPS(net='alex') (metrics.py:11) from a supplied state dict (src/lpips_alex.py); the package and its
            # weights are not available offline, in which case the key stays NaN
            from src.lpips_alex import LpipsAlex
            lpips_fn = LpipsAlex.from_file(os.environ["SRK_LPIPS_WEIGHTS"], device)
        self.lpips_fn = lpips_fn
        self.nlpd = NLPDLoss(device=device, channels=3).to(device)

    @torch.no_grad()
    def compute(self, sr, hr):
        n, c, h, w = sr.shape
        sse, ssim_sum = psnr_ssim_sums(sr, hr, clamp=True)           # clamp(0,1) fused (metrics.py:16-17)
        nlpd = self.nlpd(sr.detach(), hr.detach(), clamp01=True)
        # one device->host transfer for everything (the reference syncs four times, metrics.py:26-31)
        host = torch.cat([sse, ssim_sum, nlpd.double().reshape(1)]).cpu()
        sse_total = float(host[:n].sum())
        ssim_total = float(host[n:2 * n].sum())
        score_lpips = float("nan")
        if self.lpips_fn is not None:
            srn, hrn = sr.detach().clamp(0, 1), hr.detach().clamp(0, 1)
            score_lpips = float(self.lpips_fn(srn * 2 - 1, hrn * 2 - 1).mean().item())
        return {
            "psnr": psnr_from_sse(sse_total, n * c * h * w),
            "ssim": ssim_total / (n * c * (h - 10) * (w - 10)),
            "lpips": score_lpips,
            "nlpd": float(host[2 * n]),
        }
