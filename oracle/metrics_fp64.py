"""Independent fp64 PSNR / SSIM - TEST INFRASTRUCTURE ONLY.

Second opinion for the PSNR / SSIM restatement in oracle/sr_oracle.py (psnr(), ssim_per_image()) and for libsrk's
srk_psnr_sse / srk_ssim kernels.  The reference computes both through torchmetrics 1.8.2 (reference
src/metrics.py:9-10,19-20: PeakSignalNoiseRatio(data_range=1.0), StructuralSimilarityIndexMeasure(data_range=1.0)),
which is neither vendored nor installable here, so the parity of these two numbers stays "unpinned" against
torchmetrics itself.  What this file adds: an implementation that shares no code with sr_oracle.py and was written
from the defining formulas rather than from that restatement -

  PSNR  = 10 log10(R^2 / MSE), MSE over every element of the batch tensor (ISO/IEC definition; torchmetrics'
          dim=None, reduction='elementwise_mean' computes exactly that).
  SSIM  = Wang, Bovik, Sheikh, Simoncelli, "Image quality assessment: from error visibility to structural
          similarity", IEEE TIP 13(4), 2004, eqs. (13)-(14) with the paper's choices, which are torchmetrics'
          defaults: 11x11 circular-symmetric Gaussian window, sigma = 1.5, normalised to unit sum; K1 = 0.01,
          K2 = 0.03, C1 = (K1 R)^2, C2 = (K2 R)^2; local statistics as window-weighted moments; mean SSIM over the
          windows that lie inside the image (the paper's "valid" region; torchmetrics reflect-pads by 5 and crops 5
          again, which keeps exactly those windows); per image over all channels, then the mean over the batch.

- in numpy float64 with scipy.ndimage.correlate1d (separable window) instead of torch convolutions.
tests/test_oracle_golden.py checks sr_oracle against it; tests/test_gpu_fullsize.py checks the kernels against it."""
import numpy as np
from scipy.ndimage import correlate1d


def psnr(pred, target, data_range=1.0):
    """pred / target: arrays of equal shape (any rank).  +inf for identical inputs."""
    p, t = np.asarray(pred, dtype=np.float64), np.asarray(target, dtype=np.float64)
    mse = np.mean((p - t) ** 2)
    if mse == 0.0:
        return float("inf")
    return float(10.0 * np.log10(data_range ** 2 / mse))


def _window(size=11, sigma=1.5):
    x = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return g / g.sum()


def _local_mean(a, w):
    """Window-weighted mean over the last two axes; only the windows fully inside the image are kept."""
    r = len(w) // 2
    a = correlate1d(a, w, axis=-1, mode="mirror")
    a = correlate1d(a, w, axis=-2, mode="mirror")
    return a[..., r:-r, r:-r]


def ssim_per_image(pred, target, data_range=1.0, k1=0.01, k2=0.03):
    """pred / target: [N, C, H, W] with H, W > 10.  -> float64 [N]: mean SSIM of each image."""
    p, t = np.asarray(pred, dtype=np.float64), np.asarray(target, dtype=np.float64)
    assert p.shape == t.shape and p.ndim == 4 and p.shape[2] > 10 and p.shape[3] > 10
    w = _window()
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    mu_p, mu_t = _local_mean(p, w), _local_mean(t, w)
    var_p = np.maximum(_local_mean(p * p, w) - mu_p ** 2, 0.0)   # torchmetrics clamps the variances at 0
    var_t = np.maximum(_local_mean(t * t, w) - mu_t ** 2, 0.0)
    cov = _local_mean(p * t, w) - mu_p * mu_t
    s = ((2.0 * mu_p * mu_t + c1) * (2.0 * cov + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (var_p + var_t + c2))
    return s.reshape(s.shape[0], -1).mean(axis=1)


def ssim(pred, target, data_range=1.0):
    return float(ssim_per_image(pred, target, data_range).mean())
