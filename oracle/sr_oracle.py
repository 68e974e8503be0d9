"""CPU oracle for the SR hot path — TEST INFRASTRUCTURE ONLY.

A functional restatement, in plain torch CPU ops (fp32 or fp64), of the arithmetic of
Jaskieeeer/food101-super-resolution's src/models.py, src/loss.py and src/metrics.py.  It is what
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs compare or time;
nothing under food101-super-resolution_b200/ may import it.

Every function takes a state_dict-style mapping (the reference's key names) plus inputs and cites the
reference file:line it follows (paths relative to the reference checkout).

Pinning
  * models + mae/mse/nlpd: PINNED.  tests/golden/*.npz hold inputs, state_dicts, outputs, losses and
    gradients produced by the reference modules themselves (oracle/make_golden.py, run where
    /root/reference exists); tests/test_oracle_golden.py checks this file against them.
  * PSNR / SSIM: PARITY UNPINNED against torchmetrics.  The reference calls torchmetrics==1.8.2
    (requirements.txt:4, metrics.py:2,9-10,19-20), which is neither vendored in the reference nor installable
    here, and the reference has no test or golden value for it.  psnr()/ssim() restate torchmetrics' published
    algorithm (functional/image/psnr.py, ssim.py); they are checked against analytic known answers and against
    oracle/metrics_fp64.py, an independent numpy/scipy fp64 implementation written from the SSIM paper's
    formulas (tests/test_oracle_golden.py) - a second opinion, not a torchmetrics fixture.
  * the unmodified reference modules themselves are available as oracle/_ref (oracle/build_ref.py,
    oracle/ref_modules.py) wherever __graft_entry__.build() ran with /root/reference present.
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def conv(sd, prefix, x, padding):
    """nn.Conv2d forward, stride 1 (models.py:46,49,65,67,84-86,107,113,117,120,125)."""
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=1, padding=padding)


def prelu(sd, prefix, x):
    """nn.PReLU() with one shared slope (models.py:48,66,108,119,122)."""
    return F.prelu(x, sd[prefix + ".weight"])


def batch_norm(sd, prefix, x, training, momentum=0.1, eps=1e-5, update=True):
    """nn.BatchNorm2d (models.py:47,50,114): batch statistics + running-stat update in training mode,
    running statistics in eval mode.  Running buffers in `sd` are updated in place when update=True."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        rm_u = rm if update else rm.clone()
        rv_u = rv if update else rv.clone()
        y = F.batch_norm(x, rm_u, rv_u, w, b, True, momentum, eps)
        if update and (prefix + ".num_batches_tracked") in sd:
            sd[prefix + ".num_batches_tracked"] += 1
        return y
    return F.batch_norm(x, rm, rv, w, b, False, momentum, eps)


def se_block(sd, prefix, x):
    """SEBlock.forward (models.py:37-41): global average pool, two bias-free linears, sigmoid gate."""
    b, c = x.shape[:2]
    y = x.mean(dim=(2, 3))
    y = F.relu(F.linear(y, sd[prefix + ".fc.0.weight"]))
    y = torch.sigmoid(F.linear(y, sd[prefix + ".fc.2.weight"]))
    return x * y.view(b, c, 1, 1)


def residual_block(sd, prefix, x, training, update=True):
    """ResidualBlock.forward with use_se=False (models.py:55-60)."""
    r = conv(sd, prefix + ".conv1", x, 1)
    r = prelu(sd, prefix + ".prelu", batch_norm(sd, prefix + ".bn1", r, training, update=update))
    r = batch_norm(sd, prefix + ".bn2", conv(sd, prefix + ".conv2", r, 1), training, update=update)
    return x + r


def attention_residual_block(sd, prefix, x, res_scale=0.1):
    """AttentionResidualBlock.forward (models.py:73-78)."""
    r = conv(sd, prefix + ".conv2", prelu(sd, prefix + ".prelu", conv(sd, prefix + ".conv1", x, 1)), 1)
    return x + se_block(sd, prefix + ".se", r) * res_scale


def _num_blocks(sd):
    idx = [int(k.split(".")[1]) for k in sd if k.startswith("res_blocks.")]
    return max(idx) + 1 if idx else 0


def _upsample_tail(sd, x):
    """upsample Sequential + output_conv (models.py:116-125,142-143)."""
    x = prelu(sd, "upsample.2", F.pixel_shuffle(conv(sd, "upsample.0", x, 1), 2))
    x = prelu(sd, "upsample.5", F.pixel_shuffle(conv(sd, "upsample.3", x, 1), 2))
    return conv(sd, "output_conv", x, 4)


# ------------------------------------------------------------------------------------------------
# generators
# ------------------------------------------------------------------------------------------------
def srcnn_forward(sd, x, scale_factor):
    """SRCNN.forward (models.py:97-102)."""
    x = F.interpolate(x, scale_factor=scale_factor, mode="bicubic", align_corners=False)
    x = F.relu(conv(sd, "conv1", x, 4))
    x = F.relu(conv(sd, "conv2", x, 0))
    return conv(sd, "conv3", x, 2)


def resnet_sr_forward(sd, x, training, update=True):
    """ResNetSR.forward (models.py:137-144)."""
    initial = prelu(sd, "prelu", conv(sd, "input_conv", x, 4))
    r = initial
    for i in range(_num_blocks(sd)):
        r = residual_block(sd, "res_blocks.%d" % i, r, training, update)
    r = batch_norm(sd, "bn_mid", conv(sd, "mid_conv", r, 1), training, update=update)
    return _upsample_tail(sd, initial + r)


def attention_sr_forward(sd, x):
    """AttentionSR.forward (models.py:180-189)."""
    initial = prelu(sd, "prelu", conv(sd, "input_conv", x, 4))
    r = initial
    for i in range(_num_blocks(sd)):
        r = attention_residual_block(sd, "res_blocks.%d" % i, r)
    return _upsample_tail(sd, initial + conv(sd, "mid_conv", r, 1))


def model_forward(arch, sd, x, training=True, scale_factor=4, update=True):
    """get_model dispatch (models.py:219-227)."""
    if arch == "SRCNN":
        return srcnn_forward(sd, x, scale_factor)
    if arch == "RESNET":
        return resnet_sr_forward(sd, x, training, update)
    if arch == "AttentionSR":
        return attention_sr_forward(sd, x)
    raise ValueError("Unknown architecture: %s" % arch)


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def gaussian_kernel_5x5(channels=3, dtype=torch.float32):
    """NLPDLoss._get_gaussian_kernel (loss.py:42-55), sigma=1, size=5, normalised to sum 1."""
    ax = torch.arange(5, dtype=torch.float32) - 2.0
    g = (1.0 / (2.0 * 3.14159)) * torch.exp(-(ax[None, :] ** 2 + ax[:, None] ** 2) / 2.0)
    g = g / g.sum()
    return g.view(1, 1, 5, 5).repeat(channels, 1, 1, 1).to(dtype)


def laplacian_pyramid(img, kernel, n_levels=4):
    """NLPDLoss.get_laplacian_pyramid (loss.py:57-67)."""
    pyr, cur = [], img
    for _ in range(n_levels):
        blurred = F.conv2d(cur, kernel, padding=2, groups=img.shape[1])
        down = blurred[:, :, ::2, ::2]
        up = F.interpolate(down, size=cur.shape[2:], mode="bilinear", align_corners=False)
        pyr.append(cur - up)
        cur = down
    return pyr


def nlpd_loss(sr, hr, n_levels=4, alpha=0.7):
    """NLPDLoss.forward (loss.py:69-79)."""
    k = gaussian_kernel_5x5(sr.shape[1], sr.dtype).to(sr.device)
    l_mae = (sr - hr).abs().mean()
    l_pyr = 0
    for a, b in zip(laplacian_pyramid(sr, k, n_levels), laplacian_pyramid(hr, k, n_levels)):
        l_pyr = l_pyr + (a - b).abs().mean()
    return alpha * l_mae + (1.0 - alpha) * l_pyr


# VGG19.features[:35] (torchvision cfg "E"): conv widths, 'M' = MaxPool2d(2, 2); ReLU after every conv except the
# last one (index 34 = conv5_4, the slice ends before its ReLU).  The integer keys are the Sequential indices.
VGG19_35 = [(0, 64), (2, 64), "M", (5, 128), (7, 128), "M", (10, 256), (12, 256), (14, 256), (16, 256), "M",
            (19, 512), (21, 512), (23, 512), (25, 512), "M", (28, 512), (30, 512), (32, 512), (34, 512)]


def vgg19_features35(sd, x, prefix="vgg."):
    """torchvision vgg19().features[:35] as loss.py:23-24 slices it; sd holds '<prefix><idx>.weight/.bias'."""
    for item in VGG19_35:
        if item == "M":
            x = F.max_pool2d(x, 2, 2)
            continue
        idx, _ = item
        x = F.conv2d(x, sd["%s%d.weight" % (prefix, idx)], sd["%s%d.bias" % (prefix, idx)], padding=1)
        if idx != 34:
            x = F.relu(x)
    return x


def perceptual_loss(sd, sr, hr, prefix="vgg."):
    """PerceptualLoss.forward (loss.py:27-29): MSE between the VGG19 features of input and target."""
    return F.mse_loss(vgg19_features35(sd, sr, prefix), vgg19_features35(sd, hr, prefix))


def loss_fn(name):
    """get_loss_function (loss.py:81-92) for the names on the accelerated path."""
    name = name.lower()
    if name == "mae":
        return lambda sr, hr: (sr - hr).abs().mean()
    if name == "mse":
        return lambda sr, hr: ((sr - hr) ** 2).mean()
    if name == "nlpd":
        return nlpd_loss
    raise ValueError("Unknown loss function: %s" % name)


# ------------------------------------------------------------------------------------------------
# metrics (torchmetrics 1.8.2 restated - parity unpinned, see module docstring)
# ------------------------------------------------------------------------------------------------
def psnr(sr, hr, data_range=1.0):
    """PeakSignalNoiseRatio(data_range=1.0) (metrics.py:9,19): dim=None, base 10, elementwise_mean:
    10 log10(range^2 / mean((sr-hr)^2)) over the WHOLE batch tensor."""
    mse = ((sr.double() - hr.double()) ** 2).mean().item()
    return float("inf") if mse == 0.0 else 10.0 * math.log10(data_range ** 2 / mse)


def ssim_per_image(sr, hr, data_range=1.0, sigma=1.5, k1=0.01, k2=0.03):
    """StructuralSimilarityIndexMeasure(data_range=1.0) (metrics.py:10,20): gaussian 11x11 (sigma 1.5)
    depthwise windows over the reflect-padded images, cropped back by the pad; per-image mean of the map."""
    sr, hr = sr.double(), hr.double()
    c = sr.shape[1]
    ks = int(3.5 * sigma + 0.5) * 2 + 1
    pad = (ks - 1) // 2
    d = torch.arange((1 - ks) / 2, (1 + ks) / 2, 1, dtype=torch.float64)
    g = torch.exp(-((d / sigma) ** 2) / 2)
    g = (g / g.sum()).unsqueeze(0)
    kernel = (g.t() @ g).expand(c, 1, ks, ks)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    p = F.pad(sr, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(hr, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat((p, t, p * p, t * t, p * t))
    out = F.conv2d(stack, kernel, groups=c)
    mu_p, mu_t, e_pp, e_tt, e_pt = out.split(sr.shape[0])
    s_pp = torch.clamp(e_pp - mu_p ** 2, min=0.0)
    s_tt = torch.clamp(e_tt - mu_t ** 2, min=0.0)
    s_pt = e_pt - mu_p * mu_t
    m = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (s_pp + s_tt + c2))
    m = m[..., pad:-pad, pad:-pad]
    return m.reshape(m.shape[0], -1).mean(-1)


def ssim(sr, hr):
    return ssim_per_image(sr, hr).mean().item()


def metrics_compute(sr, hr):
    """MetricsCalculator.compute without LPIPS (metrics.py:14-31)."""
    sr, hr = sr.clamp(0, 1), hr.clamp(0, 1)
    return {"psnr": psnr(sr, hr), "ssim": ssim(sr, hr), "nlpd": nlpd_loss(sr.float(), hr.float()).item()}


# ------------------------------------------------------------------------------------------------
# helpers for tests / baselines
# ------------------------------------------------------------------------------------------------
def synthetic_pair(n, h_lr, w_lr, scale, seed=1234):
    """Food101-shaped synthetic crops: HR = 8-bit uniform noise low-passed by the NLPD
    Gaussian, LR = antialiased bicubic downsample of HR (unclamped, as dataset.py:38-39)."""
    g = torch.Generator().manual_seed(seed)
    hr = torch.randint(0, 256, (n, 3, h_lr * scale, w_lr * scale), generator=g).float() / 255.0
    hr = F.conv2d(hr, gaussian_kernel_5x5(3), padding=2, groups=3)
    lr = F.interpolate(hr, size=(h_lr, w_lr), mode="bicubic", align_corners=False, antialias=True)
    return lr.contiguous(), hr.contiguous()


def train_step_grads(arch, sd, lr_img, hr_img, loss_name, scale_factor=4, dtype=torch.float32):
    """One forward + loss + backward (train.py:116-119) on copies of `sd`; returns
    (output, loss, {param name: grad}, state after the step's BN buffer updates)."""
    work = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if v.is_floating_point():
            v = v.to(dtype)
            if not (k.endswith("running_mean") or k.endswith("running_var")):
                v.requires_grad_(True)
        work[k] = v
    out = model_forward(arch, work, lr_img.to(dtype), training=True, scale_factor=scale_factor)
    loss = loss_fn(loss_name)(out, hr_img.to(dtype))
    loss.backward()
    grads = {k: v.grad for k, v in work.items() if v.requires_grad and v.grad is not None}
    return out.detach(), loss.detach(), grads, work
