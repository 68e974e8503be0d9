"""Autograd nodes of the SR hot path.  Each node is one fused stage of the reference graph
(/root/reference/src/models.py) whose forward and backward are libsrk kernels:

  ConvAct      conv + bias + ReLU/PReLU + PixelShuffle(2) [+ residual]         models.py:84-87,107-108,116-125
  ConvBN       conv + bias + BatchNorm2d [+ PReLU] [+ residual]                 models.py:55-60,113-114,140-141
  AttnBlock    conv + PReLU + conv + squeeze-excite gate + scaled skip          models.py:62-78
  MaxPool2     2x2 / stride 2 max pooling (VGG19 features of the perceptual loss)  loss.py:23
  ImageToAct / ActToImage                                                       layout boundary

Gradients w.r.t. parameters are returned as ordinary fp32 tensors in the state_dict layout (OIHW),
so torch.optim / utils.py (train.py:55,122-133) see what they see with the reference."""
import torch

from . import _lib as L
from . import ops


def _needs(ctx, i):
    return ctx.needs_input_grad[i]


class ImageToAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, dtype):
        ops.require_cuda(img, "image_to_act")
        return ops.image_to_act(img.contiguous().float(), dtype)

    @staticmethod
    def backward(ctx, g):
        return ops.act_to_image(g.contiguous()), None


class ActToImage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        ctx.dtype = a.dtype
        return ops.act_to_image(a)

    @staticmethod
    def backward(ctx, g):
        return ops.image_to_act(g.contiguous().float(), ctx.dtype)


class ConvAct(torch.autograd.Function):
    """y = [PixelShuffle2](act(conv(x) + b)) (+ residual).  act in {none, relu, prelu(single alpha)}."""

    @staticmethod
    def forward(ctx, x, weight, bias, alpha, residual, act, shuffle, x_img, out_img, out_dtype):
        ops.require_cuda(x, "conv")
        x = x.contiguous()
        if x_img:
            x = x.float()
        assert not (shuffle and residual is not None)
        assert not (act != L.ACT_NONE and residual is not None), "residual is added to un-activated outputs only"
        assert not (out_img and act != L.ACT_NONE), "activated outputs are kept in the act layout"
        # with pixel-shuffle the activation is applied by the conv epilogue before the store remap
        # (a single-alpha PReLU / ReLU commutes with the permutation, models.py:117-119)
        # z: pre-activation copy the epilogue fills in only while the PReLU slope is <= 0 (ops.prelu_z_like)
        y, used_tc, z = ops.conv_fprop(x, x_img, weight, bias, act, alpha, residual, shuffle, out_img, out_dtype,
                                       zsave=True)
        ctx.save_for_backward(x, weight, alpha, y if act != L.ACT_NONE else None, bias, z)
        ctx.cfg = (act, shuffle, x_img, out_img, bias is not None, residual is not None, used_tc)
        return y

    @staticmethod
    def backward(ctx, dout):
        x, weight, alpha, y, bias, z = ctx.saved_tensors
        act, shuffle, x_img, out_img, has_bias, has_res, used_tc = ctx.cfg
        dout = dout.contiguous()
        dalpha = None
        perm = False  # the tcgen05 path keeps the reference channel order
        if act != L.ACT_NONE or shuffle == 2:
            dz, dalpha = ops.act_bwd(dout, y if y is not None else dout, act, alpha, shuffle, perm, zsave=z)
            dz_img = False
        else:
            dz, dz_img = dout, out_img
        dw = db = dx = None
        if ops.rgbout_bwd_supported(x, x_img, dz_img, weight):
            dx, dw, db = ops.conv_rgbout_bwd(x, dz, weight, _needs(ctx, 0), has_bias)
            return dx, dw, db, None, None, None, None, None, None, None
        # data gradient first, then the weight gradient on the side stream: it overlaps the HBM-bound backward
        # passes of the next layer instead of competing with this layer's dgrad for the SMs
        if _needs(ctx, 0):
            dx = ops.conv_dgrad(dz, dz_img, weight, None, x.dtype if not x_img else torch.float32, perm)
            if x_img:
                dx = ops.act_to_image(dx)
        if _needs(ctx, 1) or (has_bias and _needs(ctx, 2)):
            dw, db = ops.conv_wgrad(x, x_img, dz, dz_img, weight, has_bias, perm, side=True, bias=bias)
        dres = dout if (has_res and _needs(ctx, 4)) else None
        return dx, dw, db, (dalpha if _needs(ctx, 3) else None), dres, None, None, None, None, None


class UpShuffleThenRGB(torch.autograd.Function):
    """img = output_conv(PReLU(PixelShuffle2(up_conv(x))))  -  upsample[3..5] + output_conv (models.py:120-125) as ONE
    node, so that the backward of the 64 -> 3 conv, of the PReLU and of the PixelShuffle are one kernel
    (srk_conv_rgbout_bwd_unshuffle): the 64-channel gradient at the output resolution (545 MB at C2) is never written,
    and dz of the up conv arrives with sub-pixel-major channels for its dgrad / wgrad."""

    @staticmethod
    def forward(ctx, x, w_up, b_up, alpha, w_out, b_out):
        ops.require_cuda(x, "upsample tail")
        x = x.contiguous()
        y, _, z = ops.conv_fprop(x, False, w_up, b_up, L.ACT_PRELU, alpha, None, 2, False, x.dtype, zsave=True)
        img, _ = ops.conv_fprop(y, False, w_out, b_out, L.ACT_NONE, None, None, 0, True, torch.float32)
        ctx.save_for_backward(x, y, w_up, alpha, w_out, b_up, z)
        ctx.cfg = (b_up is not None, b_out is not None)
        return img

    @staticmethod
    def backward(ctx, dimg):
        x, y, w_up, alpha, w_out, b_up, z = ctx.saved_tensors
        hb_up, hb_out = ctx.cfg
        dz, dw_out, db_out, dalpha = ops.conv_rgbout_bwd_unshuffle(y, dimg.contiguous().float(), w_out, alpha, hb_out,
                                                                   zsave=z)
        dx = ops.conv_dgrad(dz, False, w_up, None, x.dtype, perm_tc=True) if _needs(ctx, 0) else None
        dw_up, db_up = ops.conv_wgrad(x, False, dz, False, w_up, hb_up, perm_tc=True, side=True, bias=b_up)
        return dx, dw_up, db_up, dalpha, dw_out, db_out

    @staticmethod
    def supported(x, w_up, w_out):
        return (x.dtype == torch.bfloat16 and tuple(w_up.shape) == (256, 64, 3, 3) and tuple(w_out.shape[:2]) == (3, 64)
                and w_out.shape[2] == w_out.shape[3] and ops._rgb_tc_ok(w_out.shape[2], w_out.shape[3])
                and ops.tc_supported(64, 256, 3, 3, x.dtype, 2) == 1 and ops.tc_supported(256, 64, 3, 3, x.dtype, 0) == 1)


def conv_act(x, conv, act=L.ACT_NONE, alpha=None, residual=None, shuffle=0, x_img=False, out_img=False,
             out_dtype=None):
    if out_dtype is None:
        out_dtype = torch.float32 if out_img else (x.dtype if not x_img else ops.cfg.compute_dtype)
    return ConvAct.apply(x, conv.weight, conv.bias, alpha, residual, act, shuffle, x_img, out_img, out_dtype)


def _conv_bn_forward(x, w, b, bn_params, bn_buffers, training, eps, momentum, alpha, residual):
    gamma, beta = bn_params
    rm, rv, nbt = bn_buffers
    sums = None
    if ops.bn_needs_batch_stats(rm, training):  # the conv epilogue accumulates the batch statistics
        y, used_tc, sums = ops.conv_fprop_stats(x, w, b)
    else:
        y, used_tc = ops.conv_fprop(x, False, w, b, L.ACT_NONE, None, None, 0, False, x.dtype)
    out, stats = ops.bn_forward(y, gamma, beta, rm, rv, nbt, training, eps, momentum, alpha, residual, sums=sums)
    return y, out, stats


class BnLink:
    """What a residual block tells the block ABOVE it about its bn2, so that the dgrad which produces this block's
    incoming gradient (dx = dgrad(dy1) + dout of the block above, models.py:55-60) can take bn2's backward reduction
    in its epilogue (ops.conv_dgrad_bnred with a residual) - one pass over (dout, y2) less per block.  Only the owner
    of the chain (ResNetSR.forward) creates links: it knows that the block's output has exactly one consumer.
    forward fills z / stats / gamma / beta; the consumer's backward fills red (+ the identity of the gradient tensor
    the sums belong to); the producer's backward takes them if it is handed that very tensor."""
    __slots__ = ("z", "stats", "gamma", "beta", "red", "grad_ptr", "grad_version")

    def __init__(self):
        self.z = self.stats = self.gamma = self.beta = self.red = None
        self.grad_ptr = self.grad_version = None

    def offer(self, red, grad):
        self.red, self.grad_ptr, self.grad_version = red, grad.data_ptr(), grad._version

    def take(self, grad):
        red, self.red = self.red, None
        self.z = self.stats = self.gamma = self.beta = None
        if red is not None and grad.data_ptr() == self.grad_ptr and grad._version == self.grad_version:
            return red
        if isinstance(red, ops.Acc):
            ops.acc_discard(red)     # sums of a gradient nobody will ask about: the accumulator goes back clean
        return None


def _conv_bn_backward(dout, x, y, stats, w, bias, gamma, beta, alpha, batch_stats, dgrad_residual, need_dx,
                      pre=None, below=None):
    """Backward of out = [PReLU](BN(conv(x))).  pre: raw BN-backward sums of dout when the dgrad that produced dout
    already reduced them (ops.conv_dgrad_bnred).  below = (z, stats, gamma, beta, alpha) of the BatchNorm layer that
    produced x: its backward reduction is then fused into this conv's dgrad (after the skip gradient dgrad_residual, if
    any, has been added) and returned as the last item."""
    dy, dgamma, dbeta, dalpha = ops.bn_backward(dout, y, stats, gamma, beta, alpha, batch_stats, pre=pre)
    dx = red = None
    if need_dx:
        fused = ops.conv_dgrad_bnred(dy, w, *below, residual=dgrad_residual) if below is not None else None
        if fused is not None:
            dx, red = fused
        else:
            dx = ops.conv_dgrad(dy, False, w, dgrad_residual, x.dtype)
    dw, db = ops.conv_wgrad(x, False, dy, False, w, bias is not None, side=True, bias=bias)   # overlaps the next BN backward
    return dx, dw, db, dgamma, dbeta, dalpha, red


class ConvBN(torch.autograd.Function):
    """out = [PReLU](BN(conv(x) + b)) [+ residual]   (mid_conv + bn_mid + skip, models.py:140-141)"""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, alpha, residual, rm, rv, nbt, training, eps, momentum):
        ops.require_cuda(x, "conv_bn")
        x = x.contiguous()
        y, out, stats = _conv_bn_forward(x, w, b, (gamma, beta), (rm, rv, nbt), training, eps, momentum,
                                         alpha, residual)
        ctx.save_for_backward(x, y, stats, w, gamma, beta, alpha, b)
        ctx.cfg = (b is not None, residual is not None, training or rm is None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, stats, w, gamma, beta, alpha, b = ctx.saved_tensors
        has_bias, has_res, batch_stats = ctx.cfg
        dout = dout.contiguous()
        dx, dw, db, dgamma, dbeta, dalpha, _ = _conv_bn_backward(
            dout, x, y, stats, w, b, gamma, beta, alpha, batch_stats, None, _needs(ctx, 0))
        return (dx, dw, db, dgamma, dbeta, dalpha, dout if has_res else None,
                None, None, None, None, None, None)


class ResBlockBN(torch.autograd.Function):
    """out = x + BN2(conv2(PReLU(BN1(conv1(x)))))   (ResidualBlock.forward, models.py:55-60, use_se=False)"""

    @staticmethod
    def forward(ctx, x, w1, b1, g1, be1, alpha, w2, b2, g2, be2, buf1, buf2, training, eps1, mom1, eps2, mom2,
                link_in=None, link_out=None):
        """link_in: BnLink of the block that produced x (its bn2 reduction rides in this block's last dgrad);
        link_out: BnLink this block fills in for the block that consumes its output."""
        ops.require_cuda(x, "residual block")
        x = x.contiguous()
        y1, a1, st1 = _conv_bn_forward(x, w1, b1, (g1, be1), buf1, training, eps1, mom1, alpha, None)
        y2, out, st2 = _conv_bn_forward(a1, w2, b2, (g2, be2), buf2, training, eps2, mom2, None, x)
        ctx.save_for_backward(x, y1, st1, a1, y2, st2, w1, g1, be1, alpha, w2, g2, be2, b1, b2)
        ctx.cfg = (b1 is not None, b2 is not None, training or buf1[0] is None)
        ctx.link_in, ctx.link_out = link_in, link_out
        if link_out is not None:
            link_out.z, link_out.stats, link_out.gamma, link_out.beta = y2, st2, g2, be2
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y1, st1, a1, y2, st2, w1, g1, be1, alpha, w2, g2, be2, b1, b2 = ctx.saved_tensors
        hb1, hb2, batch_stats = ctx.cfg
        dout = dout.contiguous()
        # bn2's reduction was taken by the dgrad of the block above if that dgrad produced this very gradient tensor
        pre2 = ctx.link_out.take(dout) if ctx.link_out is not None else None
        # conv2's dgrad produces da1, the gradient BN1 (+ PReLU) receives: BN1's backward reduction rides in its epilogue
        da1, dw2, db2, dg2, dbe2, _, red1 = _conv_bn_backward(dout, a1, y2, st2, w2, b2, g2, be2, None,
                                                              batch_stats, None, True, pre=pre2,
                                                              below=(y1, st1, g1, be1, alpha))
        # the skip connection's gradient rides in the dgrad epilogue: dx = dgrad(dy1) + dout - and so does the bn2
        # reduction of the block below, which dx is the incoming gradient of
        lk = ctx.link_in
        below = (lk.z, lk.stats, lk.gamma, lk.beta, None) if (lk is not None and lk.z is not None) else None
        dx, dw1, db1, dg1, dbe1, dalpha, red_below = _conv_bn_backward(da1, x, y1, st1, w1, b1, g1, be1, alpha,
                                                                       batch_stats, dout, True, pre=red1, below=below)
        if red_below is not None:
            lk.offer(red_below, dx)
        return (dx, dw1, db1, dg1, dbe1, dalpha, dw2, db2, dg2, dbe2,
                None, None, None, None, None, None, None, None, None)


class AttnBlock(torch.autograd.Function):
    """out = x + scale * SE(conv2(PReLU(conv1(x))))   (AttentionResidualBlock.forward, models.py:73-78)"""

    @staticmethod
    def forward(ctx, x, w1, b1, alpha, w2, b2, fc1, fc2, scale):
        ops.require_cuda(x, "attention residual block")
        x = x.contiguous()
        a, _, za = ops.conv_fprop(x, False, w1, b1, L.ACT_PRELU, alpha, None, 0, False, x.dtype, zsave=True)
        r, _ = ops.conv_fprop(a, False, w2, b2, L.ACT_NONE, None, None, 0, False, x.dtype)
        out, pool, hidden, gate = ops.se_forward(x, r, fc1, fc2, scale)
        ctx.save_for_backward(x, a, r, pool, hidden, gate, w1, alpha, w2, fc1, fc2, b1, b2, za)
        ctx.cfg = (b1 is not None, b2 is not None, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, a, r, pool, hidden, gate, w1, alpha, w2, fc1, fc2, b1, b2, za = ctx.saved_tensors
        hb1, hb2, scale = ctx.cfg
        dout = dout.contiguous()
        dr, dfc1, dfc2 = ops.se_backward(dout, r, pool, hidden, gate, fc1, fc2, scale)
        da = ops.conv_dgrad(dr, False, w2, None, a.dtype)
        dw2, db2 = ops.conv_wgrad(a, False, dr, False, w2, hb2, side=True, bias=b2)
        dz1, dalpha = ops.act_bwd(da, a, L.ACT_PRELU, alpha, 0, zsave=za)
        dx = ops.conv_dgrad(dz1, False, w1, dout, x.dtype)
        dw1, db1 = ops.conv_wgrad(x, False, dz1, False, w1, hb1, side=True, bias=b1)
        return dx, dw1, db1, dalpha, dw2, db2, dfc1, dfc2, None


class SEGate(torch.autograd.Function):
    """out = r * sigmoid(fc2(relu(fc1(mean_hw(r)))))   (SEBlock.forward, models.py:37-41)"""

    @staticmethod
    def forward(ctx, r, fc1, fc2):
        ops.require_cuda(r, "SE block")
        r = r.contiguous()
        out, pool, hidden, gate = ops.se_forward(None, r, fc1, fc2, 1.0)
        ctx.save_for_backward(r, pool, hidden, gate, fc1, fc2)
        return out

    @staticmethod
    def backward(ctx, dout):
        r, pool, hidden, gate, fc1, fc2 = ctx.saved_tensors
        dr, dfc1, dfc2 = ops.se_backward(dout.contiguous(), r, pool, hidden, gate, fc1, fc2, 1.0)
        return dr, dfc1, dfc2


class MaxPool2(torch.autograd.Function):
    """nn.MaxPool2d(kernel_size=2, stride=2) on an act tensor (VGG19.features inside PerceptualLoss, loss.py:23)."""

    @staticmethod
    def forward(ctx, x):
        ops.require_cuda(x, "max pooling")
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.maxpool2_fwd(x)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        return ops.maxpool2_bwd(x, dout.contiguous())


class ActAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.act_add(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g, g


class Bicubic(torch.autograd.Function):
    """F.interpolate(mode='bicubic', align_corners=False) on the GPU (models.py:98 does it on the CPU).
    The SRCNN input carries no gradient, so no backward is provided."""

    @staticmethod
    def forward(ctx, img, oh, ow):
        ops.require_cuda(img, "bicubic upsample")
        return ops.bicubic_upsample(img.contiguous().float(), oh, ow)

    @staticmethod
    def backward(ctx, g):
        raise RuntimeError("bicubic upsample: gradient w.r.t. the low-resolution input is not implemented")
