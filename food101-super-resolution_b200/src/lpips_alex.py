"""LPIPS (net='alex', v0.1) for MetricsCalculator - reference metrics.py:11,22 calls lpips.LPIPS(net='alex') from the
`lpips==0.1.4` package (requirements.txt:1), which ships linear-layer weights and downloads torchvision's ImageNet
AlexNet.  Neither is available offline, so this module restates the published computation (Zhang et al., "The
Unreasonable Effectiveness of Deep Features as a Perceptual Metric", CVPR 2018; lpips/lpips.py) over a state dict in
the lpips package's own key layout and is used when weights are supplied:

    MetricsCalculator(device)                      # SRK_LPIPS_WEIGHTS=<file.pth> in the environment, else lpips = NaN
    MetricsCalculator(device, lpips_fn=LpipsAlex.from_file(path, device))

  d(x, y) = sum_l mean_hw( w_l . (f_l(x)/|f_l(x)| - f_l(y)/|f_l(y)|)^2 ),   x, y in [-1, 1], first shifted / scaled by
  the package's ScalingLayer; f_l = AlexNet relu1..relu5; w_l = the non-negative 1x1 `lin` weights.

Outside the accelerated path: AlexNet's 11x11 stride-4 conv and 3x3 stride-2 max-pools have no libsrk
kernel; the evaluation-only metric runs on torch ops.  Parity is UNPINNED (no lpips package, no weights here): the
tests check the defining properties on random weights (zero on identical inputs, symmetry, per-layer formula)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

_SHIFT = (-0.030, -0.088, -0.188)
_SCALE = (0.458, 0.448, 0.450)
_CHANNELS = (64, 192, 384, 256, 256)


class LpipsAlex(nn.Module):
    def __init__(self):
        super().__init__()
        # torchvision alexnet().features indices 0, 3, 6, 8, 10 (lpips slices them at the ReLUs)
        self.convs = nn.ModuleList([nn.Conv2d(3, 64, 11, 4, 2), nn.Conv2d(64, 192, 5, 1, 2), nn.Conv2d(192, 384, 3, 1, 1),
                                    nn.Conv2d(384, 256, 3, 1, 1), nn.Conv2d(256, 256, 3, 1, 1)])
        self.lins = nn.ParameterList([nn.Parameter(torch.rand(1, c, 1, 1)) for c in _CHANNELS])
        self.register_buffer("shift", torch.tensor(_SHIFT).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor(_SCALE).view(1, 3, 1, 1))
        for p in self.parameters():
            p.requires_grad = False

    @classmethod
    def from_state_dict(cls, sd, device="cpu"):
        """sd: lpips.LPIPS(net='alex').state_dict() layout: net.slice{1..5}.<idx>.{weight,bias}, lin{0..4}.model.1.weight."""
        m = cls()
        idx = {1: 0, 2: 3, 3: 6, 4: 8, 5: 10}
        with torch.no_grad():
            for k, conv in enumerate(m.convs):
                conv.weight.copy_(sd["net.slice%d.%d.weight" % (k + 1, idx[k + 1])])
                conv.bias.copy_(sd["net.slice%d.%d.bias" % (k + 1, idx[k + 1])])
            for k in range(5):
                m.lins[k].copy_(sd["lin%d.model.1.weight" % k].view(1, -1, 1, 1))
        return m.to(device).eval()

    @classmethod
    def from_file(cls, path, device="cpu"):
        return cls.from_state_dict(torch.load(path, map_location="cpu"), device)

    def features(self, x):
        x = (x - self.shift) / self.scale
        feats = []
        for k, conv in enumerate(self.convs):
            if k in (1, 2):
                x = F.max_pool2d(x, 3, 2)
            x = F.relu(conv(x))
            feats.append(x)
        return feats

    @torch.no_grad()
    def forward(self, x, y):
        """x, y: NCHW in [-1, 1] -> [N, 1, 1, 1] distances (what lpips.LPIPS.forward returns)."""
        total = 0
        for fx, fy, w in zip(self.features(x), self.features(y), self.lins):
            nx = fx / (fx.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10)
            ny = fy / (fy.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10)
            total = total + ((nx - ny) ** 2 * w).sum(dim=1, keepdim=True).mean(dim=(2, 3), keepdim=True)
        return total
