"""Losses on libsrk (B200 / sm_100a).  Drop-in for the reference's src/loss.py: same class names,
constructor arguments, registered buffers and get_loss_function() names (reference loss.py:6-92).

  mae / mse   fused reduce kernels, forward and gradient          (reference loss.py:84,86)
  nlpd        Laplacian-pyramid loss, fused forward and backward  (reference loss.py:31-79)
  perceptual  VGG19.features[:35] MSE on the tcgen05 convs + max-pool kernels    (reference loss.py:19-29)
"""
import torch
import torch.nn as nn

from srk import _lib as L
from srk import ops


def _as_image(t, what):
    ops.require_cuda(t, what)
    if t.dim() != 4:
        raise ValueError("%s: expected an NCHW tensor, got shape %s" % (what, tuple(t.shape)))
    return t.contiguous().float()


class _PixelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sr, hr, mode):
        ops.require_cuda(sr, "pixel loss")
        ops.require_cuda(hr, "pixel loss")
        if sr.shape != hr.shape:
            raise ValueError("pixel loss: shape mismatch %s vs %s" % (tuple(sr.shape), tuple(hr.shape)))
        sr = sr.contiguous().float()
        hr = hr.contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=sr.device)
        scratch = torch.empty((L.cdll.srk_pixel_loss_scratch_bytes() // 8,), dtype=torch.float64, device=sr.device)
        L.call("srk_pixel_loss_fwd", sr.data_ptr(), hr.data_ptr(), sr.numel(), mode, loss.data_ptr(),
               scratch.data_ptr(), ops.stream_ptr())
        ctx.save_for_backward(sr, hr)
        ctx.mode = mode
        return loss

    @staticmethod
    def backward(ctx, gout):
        sr, hr = ctx.saved_tensors
        gout = gout.contiguous().float()
        st = ops.stream_ptr()
        gsr = ghr = None
        if ctx.needs_input_grad[0]:
            gsr = torch.empty_like(sr)
            L.call("srk_pixel_loss_bwd", sr.data_ptr(), hr.data_ptr(), sr.numel(), ctx.mode, gout.data_ptr(),
                   gsr.data_ptr(), st)
        if ctx.needs_input_grad[1]:
            ghr = torch.empty_like(hr)
            L.call("srk_pixel_loss_bwd", hr.data_ptr(), sr.data_ptr(), sr.numel(), ctx.mode, gout.data_ptr(),
                   ghr.data_ptr(), st)
        return gsr, ghr, None


class L1Loss(nn.L1Loss):
    """mean |input - target| (what get_loss_function('mae') returns; reference loss.py:84)."""

    def forward(self, input, target):
        if self.reduction != "mean":
            raise NotImplementedError("L1Loss: only reduction='mean' is implemented on libsrk")
        return _PixelLoss.apply(input, target, 0)


class MSELoss(nn.MSELoss):
    """mean (input - target)^2 (what get_loss_function('mse') returns; reference loss.py:86)."""

    def forward(self, input, target):
        if self.reduction != "mean":
            raise NotImplementedError("MSELoss: only reduction='mean' is implemented on libsrk")
        return _PixelLoss.apply(input, target, 1)


class TVLoss(nn.Module):
    """Total-variation regulariser of the GAN branch (reference loss.py:6-17); plain torch, outside the
    accelerated path."""

    def __init__(self, tv_loss_weight=1):
        super().__init__()
        self.tv_loss_weight = tv_loss_weight

    def forward(self, x):
        n, _, h, w = x.shape
        dh = (x[:, :, 1:, :] - x[:, :, : h - 1, :]).pow(2).sum()
        dw = (x[:, :, :, 1:] - x[:, :, :, : w - 1]).pow(2).sum()
        return self.tv_loss_weight * 2 * (self.tv_loss_weight * dh + self.tv_loss_weight * dw) / n


class PerceptualLoss(nn.Module):
    """MSE between VGG19.features[:35] activations of input and target (reference loss.py:19-29): sixteen 3x3
    convs (ReLU after all but the last: index 34 is the bare conv5_4) and four 2x2 max-pools, weights frozen, no
    ImageNet normalisation.  The module holds the torchvision layers under the reference's attribute name
    (`vgg`, so state_dict keys match) and runs them on libsrk: tcgen05 convs with fused bias + ReLU in the compute
    dtype (the 3 -> 64 first conv on the CUDA-core kernel), max-pool kernels, and the fused MSE reduction.

    `weights` is what torchvision.models.vgg19 takes; the reference's "DEFAULT" needs the ImageNet checkpoint
    (vgg19-dcbb9e9d.pth) in the torch hub cache - offline it raises, exactly like the reference does.  Pass
    weights=None (random init, what the parity tests use) or a state-dict path to build it without a download."""

    def __init__(self, device, weights="DEFAULT"):
        super().__init__()
        from torchvision.models import vgg19
        if isinstance(weights, str) and weights != "DEFAULT" and weights.endswith((".pth", ".pt")):
            net = vgg19(weights=None)
            net.load_state_dict(torch.load(weights, map_location="cpu"))
        else:
            net = vgg19(weights=weights)
        self.vgg = net.features[:35].eval().to(device)
        for p in self.vgg.parameters():
            p.requires_grad = False
        self.loss = MSELoss()

    def features(self, img):
        """NCHW fp32 image -> NCHW fp32 conv5_4 features, through libsrk."""
        from srk import fn
        ops.require_cuda(img, "PerceptualLoss")
        layers = list(self.vgg)
        x, x_img, i = img.contiguous().float(), True, 0
        while i < len(layers):
            layer = layers[i]
            if isinstance(layer, nn.Conv2d):
                relu = i + 1 < len(layers) and isinstance(layers[i + 1], nn.ReLU)
                x = fn.conv_act(x, layer, act=L.ACT_RELU if relu else L.ACT_NONE, x_img=x_img)
                x_img = False
                i += 2 if relu else 1
            elif isinstance(layer, nn.MaxPool2d):
                x = fn.MaxPool2.apply(x)
                i += 1
            else:
                raise RuntimeError("PerceptualLoss: unexpected layer %r in VGG19.features" % (layer,))
        return fn.ActToImage.apply(x)

    def forward(self, input, target):
        return self.loss(self.features(input), self.features(target))


class _NLPD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sr, hr, kernel, levels, alpha, clamp01):
        sr = _as_image(sr, "NLPD input")
        hr = _as_image(hr, "NLPD target")
        if sr.shape != hr.shape:
            raise ValueError("NLPD: shape mismatch %s vs %s" % (tuple(sr.shape), tuple(hr.shape)))
        n, c, h, w = sr.shape
        nbytes = L.cdll.srk_nlpd_workspace_bytes(n, c, h, w, levels)
        if nbytes < 0:
            raise ValueError("NLPD: n_levels must be in [1, 6]")
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=sr.device)
        k25 = kernel[0, 0].contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=sr.device)
        L.call("srk_nlpd_fwd", sr.data_ptr(), hr.data_ptr(), n, c, h, w, levels, alpha, k25.data_ptr(),
               1 if clamp01 else 0, ws.data_ptr(), loss.data_ptr(), ops.stream_ptr())
        ctx.save_for_backward(ws, k25)
        ctx.geo = (n, c, h, w, levels, alpha)
        return loss

    @staticmethod
    def backward(ctx, gout):
        ws, k25 = ctx.saved_tensors
        n, c, h, w, levels, alpha = ctx.geo
        if ctx.needs_input_grad[1]:
            raise RuntimeError("NLPD: gradient w.r.t. the target is not implemented")
        gout = gout.contiguous().float()
        g = torch.empty((n, c, h, w), dtype=torch.float32, device=ws.device)
        # the backward consumes the pyramid held in the workspace (one backward per forward)
        L.call("srk_nlpd_bwd", n, c, h, w, levels, alpha, k25.data_ptr(), ws.data_ptr(), gout.data_ptr(),
               g.data_ptr(), ops.stream_ptr())
        return g, None, None, None, None, None


class NLPDLoss(nn.Module):
    """alpha * L1 + (1 - alpha) * sum_l mean|Lap_l(input) - Lap_l(target)| over an n_levels Laplacian
    pyramid built with a 5x5 sigma-1 Gaussian (reference loss.py:31-79)."""

    def __init__(self, device="cpu", n_levels=4, channels=3, alpha=0.7):
        super().__init__()
        self.n_levels = n_levels
        self.channels = channels
        self.alpha = alpha
        self.mae = L1Loss()
        self.register_buffer("kernel", self._get_gaussian_kernel(channels=channels))

    def _get_gaussian_kernel(self, size=5, sigma=1.0, channels=3):
        # reference loss.py:42-55: un-normalised 2-D Gaussian (the 1/(2*pi*var) factor cancels) / its sum
        ax = torch.arange(size, dtype=torch.float32) - (size - 1) / 2.0
        g = torch.exp(-(ax[None, :] ** 2 + ax[:, None] ** 2) / (2.0 * sigma ** 2.0))
        g = (1.0 / (2.0 * 3.14159 * sigma ** 2.0)) * g
        g = g / torch.sum(g)
        return g.view(1, 1, size, size).repeat(channels, 1, 1, 1)

    def forward(self, input, target, clamp01=False):
        if input.shape[1] != self.channels:
            raise ValueError("NLPDLoss: expected %d channels, got %d" % (self.channels, input.shape[1]))
        return _NLPD.apply(input, target, self.kernel, int(self.n_levels), float(self.alpha), bool(clamp01))


def get_loss_function(name, device):
    """Same names as the reference factory (reference loss.py:81-92)."""
    name = name.lower()
    if name == "mae":
        return L1Loss()
    if name == "mse":
        return MSELoss()
    if name == "perceptual":
        # the reference always loads the ImageNet checkpoint (loss.py:23); SRK_VGG_WEIGHTS=none|<state-dict path> builds the
        # same module without a download (offline boxes, tests)
        import os
        w = os.environ.get("SRK_VGG_WEIGHTS", "DEFAULT")
        return PerceptualLoss(device, weights=None if w.lower() == "none" else w)
    if name == "nlpd":
        return NLPDLoss(device=device).to(device)
    raise ValueError(f"Unknown loss function: {name}")
