"""HBM-bound kernels at the C2 layer shape, timed by CUDA-graph replay; prints GB/s against algorithmic bytes."""
import sys, os
sys.path.insert(0, 'food101-super-resolution_b200')
import torch, srk
from srk import ops, _lib as L
srk.set_compute_dtype('bf16')
dev = 'cuda'
N = 64
def act(c, h, w):
    t = torch.randn(N, h + 2, w + 2, c, device=dev).bfloat16()
    t[:, 0] = 0; t[:, -1] = 0; t[:, :, 0] = 0; t[:, :, -1] = 0
    return t
y, res, dout = act(64, 64, 64), act(64, 64, 64), act(64, 64, 64)
gamma = torch.rand(64, device=dev) + 0.5; beta = torch.randn(64, device=dev) * 0.1
alpha = torch.tensor([0.25], device=dev)
rm, rv, nbt = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
_, stats = ops.bn_forward(y, gamma, beta, rm, rv, nbt, True, 1e-5, 0.1, alpha, None)
big_out, big_dout = act(64, 256, 256), act(64, 256, 256)
MB = y.numel() * 2 / 1e6
fns = {
    'bn_stats': (lambda: ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, None, None), None),
    'bn_apply+prelu': (lambda: ops.bn_forward(y, gamma, beta, rm, rv, nbt, False, 1e-5, 0.1, alpha, None), 2 * MB),
    'bn_apply+res': (lambda: ops.bn_forward(y, gamma, beta, rm, rv, nbt, False, 1e-5, 0.1, None, res), 3 * MB),
    'bn_bwd(reduce+apply)': (lambda: ops.bn_backward(dout, y, stats, gamma, beta, alpha, True), 5 * MB),
    'act_bwd_unshuffle(256^2)': (lambda: ops.act_bwd(big_dout, big_out, L.ACT_PRELU, alpha, 2), 3 * big_out.numel() * 2 / 1e6),
    'act_bwd': (lambda: ops.act_bwd(dout, y, L.ACT_PRELU, alpha, 0), 3 * MB),
}
def t(f, n=10):
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for k, (f, mb) in fns.items():
    ms = t(f)
    print("%-28s %8.1f us %s" % (k, ms * 1e3, ("%7.0f GB/s (%.0f MB)" % (mb / ms, mb)) if mb else ""), flush=True)
# --- separate BN backward phases
c = 64
red = torch.zeros(2 * c + 1, device=dev)
dy = torch.empty_like(y)
st = ops.stream_ptr
def reduce_only():
    L.call("srk_bn_bwd_reduce", ops.act_desc(dout), ops.act_desc(y), stats[0].data_ptr(), stats[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), alpha.data_ptr(), red[:c].data_ptr(), red[c:2*c].data_ptr(), red[2*c:].data_ptr(), st())
def reduce_noalpha():
    L.call("srk_bn_bwd_reduce", ops.act_desc(dout), ops.act_desc(y), stats[0].data_ptr(), stats[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, red[:c].data_ptr(), red[c:2*c].data_ptr(), None, st())
def apply_only():
    L.call("srk_bn_bwd_apply", ops.act_desc(dout), ops.act_desc(y), stats[0].data_ptr(), stats[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), alpha.data_ptr(), red[:c].data_ptr(), red[c:2*c].data_ptr(), 1, ops.act_desc(dy), st())
def stats_only():
    L.call("srk_bn_stats", ops.act_desc(y), red[:c].data_ptr(), red[c:2*c].data_ptr(), st())
for k, f, mb in (("bn_bwd_reduce", reduce_only, 2 * MB), ("bn_bwd_reduce(no prelu)", reduce_noalpha, 2 * MB), ("bn_bwd_apply", apply_only, 3 * MB), ("bn_stats only", stats_only, MB)):
    ms = t(f)
    print("%-28s %8.1f us %7.0f GB/s" % (k, ms * 1e3, mb / ms), flush=True)
