"""SR generators on libsrk (B200 / sm_100a).

Drop-in for the reference's src/models.py: same class names, constructor arguments, sub-module
names (hence the same state_dict keys, dtypes and shapes) and the same get_model() factory
(reference models.py:26-227).  The standard torch.nn containers below only HOLD parameters and
buffers; their own forward() is never used.  Every forward/backward runs through the srk autograd
nodes (srk/fn.py) on zero-bordered channels-last activations, i.e. through libsrk kernels.
Inputs and outputs are NCHW fp32 CUDA tensors like the reference's; CPU tensors are rejected
(there is no fallback path)."""
import math

import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm

from srk import _lib as L
from srk import fn, ops


def _act_in(x):
    """NCHW fp32 image -> internal activation layout (differentiable)."""
    ops.require_cuda(x, "model input")
    return fn.ImageToAct.apply(x, ops.cfg.compute_dtype)


def icnr_init(layer, scale_factor=2):
    """Sub-pixel-conv initialisation used for upsample[0] / upsample[3] (reference models.py:6-23).
    Draws one kaiming-normal sub-kernel [out_c / s^2, in_c, h, w] and tiles it s^2 times along the
    output-channel axis in block order (rows c, c + out_c/s^2, ... share a kernel), then zeroes the bias.
    RNG consumption matches the reference so seeded construction gives identical weights."""
    weight = layer.weight.data
    out_c, in_c, kh, kw = weight.shape
    groups = scale_factor ** 2
    if out_c % groups != 0:
        return
    base = torch.zeros(out_c // groups, in_c, kh, kw)
    nn.init.kaiming_normal_(base)
    tiled = torch.cat([base] * groups, dim=0)  # [out_c, in_c, kh, kw]: row o holds base[o % (out_c/groups)]
    weight.copy_(tiled)
    if layer.bias is not None:
        nn.init.zeros_(layer.bias)


def _bn_args(bn):
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return (bn.running_mean, bn.running_var, bn.num_batches_tracked), bn.eps, momentum


def _can_fold_bn(module, bn):
    """Inference with running statistics and no autograd graph: BN is a per-channel affine map of the conv output."""
    return (not module.training) and (not torch.is_grad_enabled()) and bn.track_running_stats \
        and bn.running_mean is not None


def _folded_conv_bn(conv, bn):
    """eval-mode BN(conv(x)) = conv'(x):  w' = w * gamma / sqrt(var + eps),  b' = (b - mean) * gamma / sqrt(var + eps)
    + beta  (F.batch_norm with training=False, models.py:56-57,140).  The whole residual block then is two conv
    launches with fused PReLU / skip epilogues and no elementwise pass.  The folded tensors are cached on the BN
    module and rebuilt when any source tensor changes (in-place versions; libsrk's raw-pointer writers - Adam, the
    running-statistics update - are covered by the weights epoch and by dropping the cache in training mode)."""
    srcs = (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple(-1 if t is None else t._version for t in srcs) + tuple(0 if t is None else t.data_ptr() for t in srcs) \
        + (ops._weights_epoch, bn.eps)
    ent = getattr(bn, "_srk_fold", None)
    if ent is None or ent[0] != key:
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps) if bn.weight is not None \
                else torch.rsqrt(bn.running_var + bn.eps)
            w = (conv.weight * scale.view(-1, 1, 1, 1)).contiguous()
            b0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
            b = (b0 - bn.running_mean) * scale
            if bn.bias is not None:
                b = b + bn.bias
            ent = (key, w, b.contiguous())
        bn._srk_fold = ent
    return ent[1], ent[2]


def _drop_fold(*bns):
    for bn in bns:
        if getattr(bn, "_srk_fold", None) is not None:
            bn._srk_fold = None


class SEBlock(nn.Module):
    """Squeeze-excite channel gate (reference models.py:26-41)."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel, bias=False),
            nn.Sigmoid(),
        )

    def _forward_act(self, r):
        return fn.SEGate.apply(r, self.fc[0].weight, self.fc[2].weight)

    def forward(self, x):
        return fn.ActToImage.apply(self._forward_act(_act_in(x)))


class ResidualBlock(nn.Module):
    """x + BN(conv(PReLU(BN(conv(x)))))  [optionally SE-gated]  (reference models.py:43-60)."""

    def __init__(self, channels, use_se=False):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.prelu = nn.PReLU()
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)
        self.use_se = use_se
        if use_se:
            self.se = SEBlock(channels)

    def _forward_act(self, x, link_in=None, link_out=None):
        """link_in / link_out (fn.BnLink): set by the owner of a sequential chain of blocks (ResNetSR.forward)."""
        buf1, eps1, mom1 = _bn_args(self.bn1)
        buf2, eps2, mom2 = _bn_args(self.bn2)
        if self.training:
            _drop_fold(self.bn1, self.bn2)       # the running statistics are about to change under the cache
        elif not self.use_se and _can_fold_bn(self, self.bn1) and _can_fold_bn(self, self.bn2):
            w1, b1 = _folded_conv_bn(self.conv1, self.bn1)
            w2, b2 = _folded_conv_bn(self.conv2, self.bn2)
            a = fn.ConvAct.apply(x, w1, b1, self.prelu.weight, None, L.ACT_PRELU, 0, False, False, x.dtype)
            return fn.ConvAct.apply(a, w2, b2, None, x, L.ACT_NONE, 0, False, False, x.dtype)
        if not self.use_se:
            return fn.ResBlockBN.apply(
                x, self.conv1.weight, self.conv1.bias, self.bn1.weight, self.bn1.bias, self.prelu.weight,
                self.conv2.weight, self.conv2.bias, self.bn2.weight, self.bn2.bias,
                buf1, buf2, self.training, eps1, mom1, eps2, mom2, link_in, link_out)
        a = fn.ConvBN.apply(x, self.conv1.weight, self.conv1.bias, self.bn1.weight, self.bn1.bias,
                            self.prelu.weight, None, *buf1, self.training, eps1, mom1)
        r = fn.ConvBN.apply(a, self.conv2.weight, self.conv2.bias, self.bn2.weight, self.bn2.bias,
                            None, None, *buf2, self.training, eps2, mom2)
        return fn.ActAdd.apply(x, self.se._forward_act(r))

    def forward(self, x):
        return fn.ActToImage.apply(self._forward_act(_act_in(x)))


class AttentionResidualBlock(nn.Module):
    """x + 0.1 * SE(conv(PReLU(conv(x))))  (reference models.py:62-78)."""

    def __init__(self, channels):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.prelu = nn.PReLU()
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.se = SEBlock(channels)
        self.res_scale = 0.1

    def _forward_act(self, x):
        return fn.AttnBlock.apply(x, self.conv1.weight, self.conv1.bias, self.prelu.weight,
                                  self.conv2.weight, self.conv2.bias,
                                  self.se.fc[0].weight, self.se.fc[2].weight, float(self.res_scale))

    def forward(self, x):
        return fn.ActToImage.apply(self._forward_act(_act_in(x)))


class SRCNN(nn.Module):
    """bicubic upsample -> 9x9 conv -> 1x1 conv -> 5x5 conv (reference models.py:80-102).
    The reference interpolates on the CPU (models.py:98); here it is a CUDA kernel."""

    def __init__(self, num_channels=3, scale_factor=4, hidden_dim=64):
        super().__init__()
        self.scale_factor = scale_factor
        self.conv1 = nn.Conv2d(num_channels, 64, kernel_size=9, padding=4)
        self.conv2 = nn.Conv2d(64, hidden_dim, kernel_size=1, padding=0)
        self.conv3 = nn.Conv2d(hidden_dim, num_channels, kernel_size=5, padding=2)
        self.relu = nn.ReLU(inplace=True)
        self._initialize_weights()

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        ops.require_cuda(x, "SRCNN input")
        oh = int(math.floor(x.shape[2] * self.scale_factor))
        ow = int(math.floor(x.shape[3] * self.scale_factor))
        up = fn.Bicubic.apply(x, oh, ow)
        a = fn.conv_act(up, self.conv1, act=L.ACT_RELU, x_img=True)
        a = fn.conv_act(a, self.conv2, act=L.ACT_RELU)
        return fn.conv_act(a, self.conv3, out_img=True)


def _upsample_tail(model, x):
    up = model.upsample
    x = fn.conv_act(x, up[0], act=L.ACT_PRELU, alpha=up[2].weight, shuffle=2)
    oc = model.output_conv
    if torch.is_grad_enabled() and fn.UpShuffleThenRGB.supported(x, up[3].weight, oc.weight):
        # second upsample stage + output conv as one autograd node (fused backward, see fn.UpShuffleThenRGB)
        return fn.UpShuffleThenRGB.apply(x, up[3].weight, up[3].bias, up[5].weight, oc.weight, oc.bias)
    x = fn.conv_act(x, up[3], act=L.ACT_PRELU, alpha=up[5].weight, shuffle=2)
    return fn.conv_act(x, oc, out_img=True)


def _make_upsample(num_channels):
    return nn.Sequential(
        nn.Conv2d(num_channels, 256, 3, 1, 1),
        nn.PixelShuffle(2),
        nn.PReLU(),
        nn.Conv2d(64, 256, 3, 1, 1),
        nn.PixelShuffle(2),
        nn.PReLU(),
    )


def _init_sr_weights(model):
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
    icnr_init(model.upsample[0], scale_factor=2)
    icnr_init(model.upsample[3], scale_factor=2)


class ResNetSR(nn.Module):
    """SRResNet-style x4 generator with BatchNorm residual blocks (reference models.py:104-144).
    `scale_factor` is accepted and, as in the reference, not used: the tail is always 2x PixelShuffle(2)."""

    def __init__(self, scale_factor=4, num_channels=64, num_residuals=16):
        super().__init__()
        self.input_conv = nn.Conv2d(3, num_channels, kernel_size=9, padding=4)
        self.prelu = nn.PReLU()
        self.res_blocks = nn.Sequential(*[ResidualBlock(num_channels, use_se=False) for _ in range(num_residuals)])
        self.mid_conv = nn.Conv2d(num_channels, num_channels, kernel_size=3, padding=1)
        self.bn_mid = nn.BatchNorm2d(num_channels)
        self.upsample = _make_upsample(num_channels)
        self.output_conv = nn.Conv2d(64, 3, kernel_size=9, padding=4)
        self._init_weights()

    def _init_weights(self):
        _init_sr_weights(self)

    def forward(self, x):
        ops.require_cuda(x, "ResNetSR input")
        initial = fn.conv_act(x, self.input_conv, act=L.ACT_PRELU, alpha=self.prelu.weight, x_img=True)
        r = initial
        # every block output has exactly one consumer (the next block): the bn2 backward reduction of block k rides in
        # the last dgrad of block k+1 (fn.BnLink)
        link = None
        chain = self.training and torch.is_grad_enabled()
        for i, blk in enumerate(self.res_blocks):
            nxt = fn.BnLink() if (chain and i + 1 < len(self.res_blocks)) else None
            r = blk._forward_act(r, link, nxt)
            link = nxt
        buf, eps, mom = _bn_args(self.bn_mid)
        if self.training:
            _drop_fold(self.bn_mid)
        elif _can_fold_bn(self, self.bn_mid):
            wm, bm = _folded_conv_bn(self.mid_conv, self.bn_mid)
            t = fn.ConvAct.apply(r, wm, bm, None, initial, L.ACT_NONE, 0, False, False, r.dtype)
            return _upsample_tail(self, t)
        t = fn.ConvBN.apply(r, self.mid_conv.weight, self.mid_conv.bias, self.bn_mid.weight, self.bn_mid.bias,
                            None, initial, *buf, self.training, eps, mom)
        return _upsample_tail(self, t)


class AttentionSR(nn.Module):
    """x4 generator with squeeze-excite residual blocks and no BatchNorm (reference models.py:146-189)."""

    def __init__(self, scale_factor=4, num_channels=64, num_residuals=32):
        super().__init__()
        self.input_conv = nn.Conv2d(3, num_channels, kernel_size=9, padding=4)
        self.prelu = nn.PReLU()
        self.res_blocks = nn.Sequential(*[AttentionResidualBlock(num_channels) for _ in range(num_residuals)])
        self.mid_conv = nn.Conv2d(num_channels, num_channels, kernel_size=3, padding=1)
        self.upsample = _make_upsample(num_channels)
        self.output_conv = nn.Conv2d(64, 3, kernel_size=9, padding=4)
        self._init_weights()

    def _init_weights(self):
        _init_sr_weights(self)

    def forward(self, x):
        ops.require_cuda(x, "AttentionSR input")
        initial = fn.conv_act(x, self.input_conv, act=L.ACT_PRELU, alpha=self.prelu.weight, x_img=True)
        r = initial
        for blk in self.res_blocks:
            r = blk._forward_act(r)
        t = fn.conv_act(r, self.mid_conv, residual=initial)
        return _upsample_tail(self, t)


class Discriminator(nn.Module):
    """Spectral-norm patch discriminator of the GAN branch (reference models.py:191-217).  Outside the
    accelerated hot path (no sweep config uses loss_function=gan): kept as plain torch modules so that
    `from src.models import Discriminator` (train.py:12) and its state_dict keep working."""

    def __init__(self, in_nc=3, nf=64):
        super().__init__()

        def stage(cin, cout, stride, bias, bn):
            mods = [spectral_norm(nn.Conv2d(cin, cout, 3, stride, 1, bias=bias))]
            if bn:
                mods.append(nn.BatchNorm2d(cout))
            mods.append(nn.LeakyReLU(0.2, inplace=True))
            return mods

        self.net = nn.Sequential(
            *stage(in_nc, nf, 1, True, False),
            *stage(nf, nf * 2, 2, False, True),
            *stage(nf * 2, nf * 4, 2, False, True),
            *stage(nf * 4, nf * 8, 2, False, True),
        )
        self.classifier = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            nn.Flatten(),
            spectral_norm(nn.Linear(nf * 8, 100)),
            nn.LeakyReLU(0.2, inplace=True),
            spectral_norm(nn.Linear(100, 1)),
        )

    def forward(self, x):
        return self.classifier(self.net(x))


def get_model(name, scale_factor=4, device="cpu"):
    """Factory with the reference's names and sizes (reference models.py:219-227)."""
    if name == "SRCNN":
        return SRCNN(scale_factor=scale_factor, hidden_dim=64).to(device)
    if name == "RESNET":
        return ResNetSR(scale_factor=scale_factor, num_residuals=16, num_channels=64).to(device)
    if name == "AttentionSR":
        return AttentionSR(scale_factor=scale_factor, num_residuals=32, num_channels=96).to(device)
    raise ValueError(f"Unknown architecture: {name}")
