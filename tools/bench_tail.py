"""Times the high-resolution tail of the C2 step (models.py:116-125): the two 64 -> 256 PixelShuffle convs, the
9x9 64 -> 3 output conv forward and its fused backward, the upsample dgrads.  Same harness as tools/bench_conv.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import srk  # noqa: E402
from srk import _lib as L  # noqa: E402
from srk import ops  # noqa: E402

srk.set_compute_dtype("bf16")
dev = torch.device("cuda:0")
B, H = int(os.environ.get("B", 64)), int(os.environ.get("H", 64))
g = torch.Generator(device=dev).manual_seed(1)


def act(h, c, scale=1.0):
    t = torch.zeros((B, h + 2, h + 2, c), dtype=torch.bfloat16, device=dev)
    t[:, 1:-1, 1:-1] = (torch.randn((B, h, h, c), generator=g, device=dev) * scale).bfloat16()
    return t


wup = torch.randn((256, 64, 3, 3), generator=g, device=dev) / 24
wout = torch.randn((3, 64, 9, 9), generator=g, device=dev) / 72
alpha = torch.full((1,), 0.25, device=dev)
x64 = [act(H, 64) for _ in range(2)]
x128 = [act(2 * H, 64) for _ in range(2)]
x256 = [act(4 * H, 64) for _ in range(2)]
gimg = [torch.randn((B, 3, 4 * H, 4 * H), generator=g, device=dev) for _ in range(2)]
d256_64 = [act(H, 256, 1e-3) for _ in range(2)]
d256_128 = [act(2 * H, 256, 1e-3) for _ in range(2)]
cases = [
    ("up1 fwd 64->256 PS PReLU @%d" % H, 2.0 * B * H * H * 64 * 256 * 9,
     [lambda x=x: ops.conv_fprop(x, False, wup, None, L.ACT_PRELU, alpha, None, 2, False, torch.bfloat16) for x in x64]),
    ("up2 fwd 64->256 PS PReLU @%d" % (2 * H), 2.0 * B * 4 * H * H * 64 * 256 * 9,
     [lambda x=x: ops.conv_fprop(x, False, wup, None, L.ACT_PRELU, alpha, None, 2, False, torch.bfloat16) for x in x128]),
    ("out conv fwd 9x9 64->3 @%d" % (4 * H), 2.0 * B * 16 * H * H * 64 * 3 * 81,
     [lambda x=x: ops.conv_fprop(x, False, wout, None, L.ACT_NONE, None, None, 0, True, torch.float32) for x in x256]),
    ("out conv bwd (+PReLU, unshuffle) @%d" % (4 * H), 2 * 2.0 * B * 16 * H * H * 64 * 3 * 81,
     [lambda x=x, gi=gi: ops.conv_rgbout_bwd_unshuffle(x, gi, wout, alpha, True) for x, gi in zip(x256, gimg)]),
    ("up2 dgrad 256->64 @%d" % (2 * H), 2.0 * B * 4 * H * H * 64 * 256 * 9,
     [lambda d=d: ops.conv_dgrad(d, False, wup, None, torch.bfloat16, perm_tc=True) for d in d256_128]),
    ("up1 dgrad 256->64 @%d" % H, 2.0 * B * H * H * 64 * 256 * 9,
     [lambda d=d: ops.conv_dgrad(d, False, wup, None, torch.bfloat16) for d in d256_64]),
    ("up2 wgrad @%d" % (2 * H), 2.0 * B * 4 * H * H * 64 * 256 * 9,
     [lambda x=x, d=d: ops.conv_wgrad(x, False, d, False, wup, True, perm_tc=True) for x, d in zip(x128, d256_128)]),
]
for name, flop, fs in cases:
    ms = bench._time_replayed(fs)
    print("%-40s %8.1f us  %7.1f TFLOP/s" % (name, ms * 1e3, flop / (ms * 1e-3) / 1e12), flush=True)
