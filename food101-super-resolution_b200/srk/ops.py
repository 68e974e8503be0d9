"""Thin, non-differentiable wrappers over the libsrk C ABI: descriptor construction, weight-pack
cache and one Python function per kernel family.  torch is used for device memory and the
current stream only; every arithmetic operation happens inside libsrk.

Tensor conventions
  image : NCHW fp32 contiguous (what the reference modules take / return)
  act   : zero-bordered channels-last [N, H+2, W+2, C], fp32 or bf16 (internal activations)
"""
import os
import weakref

import torch

from . import _lib as L


class _Config:
    # storage / arithmetic type of internal activations: torch.float32 (CUDA-core fp32 convs,
    # reference-exact) or torch.bfloat16 (tcgen05 convs where the shape is supported)
    compute_dtype = torch.float32
    # "auto": tcgen05 when supported and activations are bf16, else CUDA cores; "simt": never tcgen05
    conv_impl = "auto"


    # Opt-in (set_overlap_wgrad / SRK_OVERLAP_WGRAD=1; bench.py and train.py switch it on): weight-gradient kernels
    # (tensor-core / shared-memory bound) run on a side stream so that they overlap the HBM-bound BatchNorm /
    # activation backward passes of the next layer.  Their results are then NOT handed to autograd (AccumulateGrad
    # would clone or add them on the main stream before the side stream has produced them): the backward nodes return
    # None for those parameters and `.grad` is assigned once the streams have joined, at the end of the backward pass.
    # Consequences: use loss.backward() (torch.autograd.grad(..., params) sees no gradient for conv weights), and
    # parameter hooks do not fire for them (parameters with hooks stay on the ordinary path).
    overlap_wgrad = os.environ.get("SRK_OVERLAP_WGRAD", "0") == "1"
    # BatchNorm-backward reductions ride in the epilogue of the dgrad that produces their input gradient
    fuse_bn_reduce = os.environ.get("SRK_FUSE_BN_REDUCE", "1") != "0"
    # BatchNorm sums of the trunk convs through exact integer accumulators (include/srk.h "exact sums") instead of the
    # ordered float fold
    use_acc = os.environ.get("SRK_ACC", "1") != "0"


cfg = _Config()


def set_compute_dtype(dtype):
    if isinstance(dtype, str):
        dtype = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16,
                 "bfloat16": torch.bfloat16}[dtype.lower()]
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("compute dtype must be float32 or bfloat16")
    cfg.compute_dtype = dtype


def set_overlap_wgrad(on):
    """See _Config.overlap_wgrad."""
    cfg.overlap_wgrad = bool(on)


def set_conv_impl(impl):
    if impl not in ("auto", "simt"):
        raise ValueError("conv impl must be 'auto' or 'simt'")
    cfg.conv_impl = impl


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            "%s: libsrk runs on CUDA (sm_100a) tensors only; got a %s tensor. There is no CPU fallback."
            % (what, t.device))


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _dt(t):
    if t.dtype == torch.bfloat16:
        return L.BF16
    if t.dtype == torch.float32:
        return L.F32
    raise TypeError("unsupported dtype %s" % t.dtype)


def act_desc(t):
    assert t.dim() == 4 and t.is_contiguous(), "act tensors are contiguous [N,H+2,W+2,C]"
    n, hp, wp, c = t.shape
    return L.SrkTensor(t.data_ptr(), L.LAYOUT_ACT, _dt(t), n, c, hp - 2, wp - 2)


def img_desc(t):
    assert t.dim() == 4 and t.is_contiguous() and t.dtype == torch.float32, "images are contiguous NCHW fp32"
    n, c, h, w = t.shape
    return L.SrkTensor(t.data_ptr(), L.LAYOUT_IMAGE, L.F32, n, c, h, w)


def desc(t, is_image):
    return img_desc(t) if is_image else act_desc(t)


def geometry(t, is_image):
    """-> (N, C, H, W) logical sizes"""
    if is_image:
        n, c, h, w = t.shape
    else:
        n, hp, wp, c = t.shape
        h, w = hp - 2, wp - 2
    return n, c, h, w


def new_act(n, c, h, w, dtype, device):
    return torch.empty((n, h + 2, w + 2, c), dtype=dtype, device=device)


def new_image(n, c, h, w, device):
    return torch.empty((n, c, h, w), dtype=torch.float32, device=device)


def _ptr(t):
    return None if t is None else t.data_ptr()


# ---- reduce workspace ----------------------------------------------------------------------------------
# Cross-block reductions are deterministic (per-block partial rows folded in block order, include/srk.h
# "deterministic reductions"): the reducing entry points take a workspace that is zero-filled once and must not be
# shared by kernels that may run concurrently - one per (device, stream).
_reduce_ws = {}


def reduce_ws(device=None):
    st = torch.cuda.current_stream(device)
    key = (st.device.index, st.cuda_stream)
    t = _reduce_ws.get(key)
    if t is None:
        t = torch.zeros((L.cdll.srk_reduce_workspace_bytes(),), dtype=torch.uint8, device=st.device)
        _reduce_ws[key] = t
    return t.data_ptr()


# ---- accumulators (include/srk.h "exact sums") ------------------------------------------------------------
# A small ring of accumulator slots per device, zero-filled once.  A producer (conv_fprop bn_sums=Acc,
# conv_dgrad_bnred) marks its slot dirty, the consumer (bn_forward / bn_backward / acc_read), which leaves the device
# buffer zero-filled again, marks it clean.  A slot that comes round still dirty (its consumer never ran: a partial
# backward, an exception) is zero-filled before reuse.  Within a training step at most two slots are outstanding.
class Acc:
    __slots__ = ("t", "dirty")

    def __init__(self, t):
        self.t, self.dirty = t, False

    def data_ptr(self):
        return self.t.data_ptr()


_acc_pool = {}
_ACC_SLOTS = 16


def acc_acquire(device):
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    pool = _acc_pool.get(idx)
    if pool is None:
        nb = int(L.cdll.srk_acc_bytes())
        buf = torch.zeros((_ACC_SLOTS, nb), dtype=torch.uint8, device=torch.device("cuda", idx))
        pool = _acc_pool[idx] = {"slots": [Acc(buf[i]) for i in range(_ACC_SLOTS)], "next": 0}
    a = pool["slots"][pool["next"]]
    pool["next"] = (pool["next"] + 1) % _ACC_SLOTS
    if a.dirty:
        a.t.zero_()
    a.dirty = True
    return a


def acc_discard(a):
    """An accumulator whose consumer will not run: reset it now."""
    if a.dirty:
        a.t.zero_()
        a.dirty = False


def acc_read(a, nv):
    """Generic consumer -> fp32 [nv]; the accumulator is reset."""
    out = torch.empty((nv,), dtype=torch.float32, device=a.t.device)
    L.call("srk_acc_read", a.data_ptr(), nv, out.data_ptr(), stream_ptr())
    a.dirty = False
    return out


def _acc_conv_ok(x, weight):
    return (cfg.use_acc and cfg.conv_impl != "simt" and x.dtype == torch.bfloat16 and tuple(weight.shape) == (64, 64, 3, 3))


# ---- zero-initialised scratch ------------------------------------------------------------------------
# No libsrk kernel needs zero-filled outputs any more (reductions write their results); begin_step() / zeros() stay
# for callers that want many small zero tensors from one memset.  Tensors obtained this way are only valid until the
# next begin_step().
class _ZeroArena:
    buf = None
    off = 0
    active = False


def begin_step(nbytes=8 << 20, device=None):
    if _ZeroArena.buf is None or _ZeroArena.buf.numel() < nbytes or (device is not None and _ZeroArena.buf.device != torch.device(device)):
        _ZeroArena.buf = torch.empty((nbytes,), dtype=torch.uint8, device=device or "cuda")
    _ZeroArena.buf.zero_()
    _ZeroArena.off = 0
    _ZeroArena.active = True


def end_arena():
    _ZeroArena.active = False


def zeros(shape, device, dtype=torch.float32):
    if isinstance(shape, int):
        shape = (shape,)
    n = 1
    for d in shape:
        n *= d
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    a = _ZeroArena
    if a.active and a.buf is not None and a.buf.device == torch.device(device) and a.off + nbytes + 256 <= a.buf.numel():
        start = (a.off + 255) // 256 * 256
        a.off = start + nbytes
        return a.buf[start:start + nbytes].view(dtype).view(shape)
    return torch.zeros(shape, dtype=dtype, device=device)


# ---- weight packing ------------------------------------------------------------------------------
_pack_cache = {}
_weights_epoch = 0


def bump_weights_epoch():
    """Call after parameters were modified behind torch's back (raw-pointer optimizer kernels)."""
    global _weights_epoch
    _weights_epoch += 1


def _drop(key):
    _pack_cache.pop(key, None)


def packed_weight(weight, kind, shuffle):
    """OIHW fp32 master weight -> kernel operand layout, cached until the weight is modified in place
    (optimizer step bumps Tensor._version) or collected."""
    key = id(weight)
    ent = _pack_cache.get(key)
    ver = (weight._version, _weights_epoch)
    if ent is None or ent["ver"] != ver or ent["ptr"] != weight.data_ptr():
        ent = {"ver": ver, "ptr": weight.data_ptr(), "packs": {}}
        if key not in _pack_cache:
            try:
                ent["ref"] = weakref.ref(weight, lambda _r, k=key: _drop(k))
            except TypeError:
                pass
        else:
            ent["ref"] = _pack_cache[key].get("ref")
        _pack_cache[key] = ent
    pk = ent["packs"].get((kind, shuffle))
    if pk is None:
        cout, cin, r, s = weight.shape
        w = weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        nbytes = L.cdll.srk_weight_pack_bytes(cout, cin, r, s, kind)
        pk = torch.empty((nbytes,), dtype=torch.uint8, device=weight.device)
        L.call("srk_weight_pack", w.data_ptr(), pk.data_ptr(), cout, cin, r, s, kind, shuffle, stream_ptr())
        ent["packs"][(kind, shuffle)] = pk
    return pk


def repack_all():
    """Refreshes every cached weight pack in one multi-tensor launch and marks it current.  Call right after an
    optimizer step (srk.optim.Adam bumps the epoch): the next forward / backward then finds all packs ready
    instead of launching ~70 small pack kernels on first use."""
    import ctypes
    srcs, dsts, couts, cins, rs, kinds, shufs, ents = [], [], [], [], [], [], [], []
    for ent in list(_pack_cache.values()):
        ref = ent.get("ref")
        w = ref() if ref is not None else None
        if w is None or w.data_ptr() != ent["ptr"]:
            continue
        cout, cin, r, s = w.shape
        if r != s:
            continue
        for (kind, shuffle), pk in ent["packs"].items():
            srcs.append(w.data_ptr()); dsts.append(pk.data_ptr())
            couts.append(cout); cins.append(cin); rs.append(r); kinds.append(kind); shufs.append(shuffle)
        ents.append((ent, w))
    n = len(srcs)
    if n == 0:
        return 0
    vp = lambda v: (ctypes.c_void_p * n)(*v)
    ip = lambda v: (ctypes.c_int32 * n)(*v)
    L.call("srk_weight_pack_multi", n, vp(srcs), vp(dsts), ip(couts), ip(cins), ip(rs), ip(kinds), ip(shufs),
           stream_ptr())
    for ent, w in ents:
        ent["ver"] = (w._version, _weights_epoch)
    return n


def tc_supported(cin, cout, r, s, dtype, shuffle):
    """0 = CUDA cores, 1 = tcgen05 ACT->ACT conv, 2 = tcgen05 RGB-output conv (ACT -> IMAGE)"""
    if cfg.conv_impl == "simt" or dtype != torch.bfloat16:
        return 0
    return int(L.cdll.srk_conv_tc_supported(cin, cout, r, s, L.BF16, shuffle))


# ---- optional per-kernel timing (bench.py roofline) -------------------------------------------------
class KernelTimer:
    """Brackets selected libsrk launches with CUDA events on the launching (current) stream."""

    def __init__(self, select):
        self.select, self.events = select, []

    def wrap(self, key, launch):
        if not self.select(key):
            return launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = launch()
        e1.record()
        self.events.append((key, e0, e1))
        return out

    def summary(self):
        """-> {key: (average ms, launches)}; synchronises."""
        torch.cuda.synchronize()
        acc = {}
        for key, e0, e1 in self.events:
            t, n = acc.get(key, (0.0, 0))
            acc[key] = (t + e0.elapsed_time(e1), n + 1)
        return {k: (t / n, n) for k, (t, n) in acc.items()}


kernel_timer = None


def _timed(key, launch):
    if kernel_timer is None:
        return launch()
    return kernel_timer.wrap(key, launch)


# ---- side stream ------------------------------------------------------------------------------------
class _Side:
    streams = {}      # device index -> torch.cuda.Stream
    pending = []      # tensors the side stream still uses (kept alive until the join is enqueued)
    deferred = []     # (parameter, gradient computed on the side stream): assigned to .grad after the join
    listener = None   # callable(parameter, gradient): told about every side-stream gradient as it is enqueued


def set_side_grad_listener(fn):
    """srk.dp.GradAverager's overlapped mode: learn about side-stream gradients when they are ENQUEUED (they reach
    .grad only at the join), so that a gradient bucket can be all-reduced while backward is still running."""
    _Side.listener = fn


def _side_stream(device):
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    st = _Side.streams.get(idx)
    if st is None:
        st = _Side.streams[idx] = torch.cuda.Stream(device=idx)
    return st


@torch.no_grad()
def join_side_stream():
    """Make the current stream wait for everything launched through run_on_side_stream, then hand the deferred
    parameter gradients over: .grad = g, or .grad += g when a gradient is already there (accumulation)."""
    if not _Side.pending and not _Side.deferred:
        return
    main = torch.cuda.current_stream()
    main.wait_stream(_side_stream(main.device))
    for p, g in _Side.deferred:
        if p.grad is None:
            p.grad = g
        else:
            p.grad.add_(g)
    _Side.deferred.clear()
    _Side.pending.clear()   # frees happen after the join is enqueued: allocator reuse stays ordered


def run_on_side_stream(launch, keep, grads=()):
    """Run `launch()` (libsrk launches that read the current stream through stream_ptr) on the side stream,
    ordered after everything already enqueued on the current stream.  The current stream joins again at the end
    of the autograd backward pass (engine callback); outside a backward pass the join happens immediately.
    `keep` holds every tensor the launch touches and `grads` the (parameter, gradient) pairs it produces: all of
    them were allocated on the current stream and stay referenced until the join is enqueued, so the caching
    allocator cannot hand their memory to later kernels of the current stream, and the gradients reach `.grad`
    only after the join.  Works under CUDA-graph capture (event record / wait become graph edges; the join closes
    the fork before the capture ends)."""
    main = torch.cuda.current_stream()
    side = _side_stream(main.device)
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        launch()
    _Side.pending.append(keep)
    _Side.deferred.extend(grads)
    if _Side.listener is not None:
        for prm, g in grads:
            _Side.listener(prm, g)
    try:   # one callback per launch: after the first one has joined the others find nothing pending
        torch.autograd.Variable._execution_engine.queue_callback(join_side_stream)
    except RuntimeError:
        join_side_stream()


# ---- convolution -----------------------------------------------------------------------------------
def _rgb_tc_ok(r, s):
    return cfg.conv_impl != "simt" and cfg.compute_dtype == torch.bfloat16 and r == s and r in (5, 9)


def _rgb_workspace(k, device):
    return torch.empty((L.cdll.srk_conv_rgb_workspace_bytes(k),), dtype=torch.uint8, device=device)


def conv_rgbout_bwd(x, dout, weight, need_dx, need_bias):
    """Backward of a 64 -> 3 conv whose output gradient is an NCHW fp32 image (output_conv, SRCNN conv3):
    one tcgen05 kernel produces dx (bf16 act), dW and db.  -> (dx or None, dw, db or None)"""
    cout, cin, r, s = weight.shape
    n, _, h, w = geometry(x, False)
    dw = torch.empty_like(weight, memory_format=torch.contiguous_format)    # written, not accumulated
    db = torch.empty((cout,), dtype=torch.float32, device=weight.device)
    dx = new_act(n, cin, h, w, torch.bfloat16, x.device) if need_dx else None
    pk = packed_weight(weight, L.PACK_RGBOUT_DGRAD_TC, 0) if need_dx else None
    ws = _rgb_workspace(r, x.device)
    _timed(("conv_rgbout_bwd", cin, cout, r, 0, n, h, w, True),
           lambda: L.call("srk_conv_rgb_bwd", img_desc(dout), act_desc(x), _ptr(pk),
                          act_desc(dx) if need_dx else None, dw.data_ptr(), db.data_ptr(), r, 1, ws.data_ptr(),
                          stream_ptr()))
    return dx, dw, (db if need_bias else None)


def conv_rgbout_bwd_unshuffle(t64, dout_img, weight, alpha, need_bias, zsave=None):
    """Backward of the 64 -> 3 output conv fused with the PReLU + PixelShuffle(2) backward of the upsample stage that
    produced its input t64 (srk_conv_rgbout_bwd_unshuffle); zsave: that stage's prelu_z copy.
    -> (dz [N, H/2+2, W/2+2, 256] bf16 with sub-pixel-major channels, dw, db or None, dalpha[1])"""
    cout, cin, r, s = weight.shape
    n, _, h, w = geometry(t64, False)
    dw = torch.empty_like(weight, memory_format=torch.contiguous_format)
    db = torch.empty((cout,), dtype=torch.float32, device=weight.device)
    dalpha = torch.empty((1,), dtype=torch.float32, device=weight.device)
    dz = new_act(n, 4 * cin, h // 2, w // 2, torch.bfloat16, t64.device)
    pk = packed_weight(weight, L.PACK_RGBOUT_DGRAD_TC, 0)
    ws = _rgb_workspace(r, t64.device)
    _timed(("conv_rgbout_bwd_unshuffle", cin, cout, r, 2, n, h, w, True),
           lambda: L.call("srk_conv_rgbout_bwd_unshuffle", img_desc(dout_img), act_desc(t64),
                          act_desc(zsave) if zsave is not None else None, pk.data_ptr(),
                          act_desc(dz), dw.data_ptr(), db.data_ptr(), alpha.data_ptr(), dalpha.data_ptr(), r,
                          ws.data_ptr(), stream_ptr()))
    return dz, dw, (db if need_bias else None), dalpha


def rgbout_bwd_supported(x, x_img, dz_img, weight):
    cout, cin, r, s = weight.shape
    return (not x_img) and dz_img and cin == 64 and cout == 3 and x.dtype == torch.bfloat16 and _rgb_tc_ok(r, s)


def _fprop_workspace(xd, kind, device):
    nbytes = L.cdll.srk_conv_fprop_workspace_bytes(xd, kind)
    return torch.empty((nbytes,), dtype=torch.uint8, device=device) if nbytes > 0 else None


def prelu_z_like(y):
    """Buffer for the pre-activation copy a PReLU conv epilogue writes when its slope is <= 0 (include/srk.h, prelu_z).
    Allocation only: while the slope is > 0 - the usual case - nothing ever touches it."""
    return torch.empty_like(y)


def conv_fprop_stats(x, weight, bias):
    """y = conv(x, weight) + bias with the per-channel (sum, sum of squares) of y taken in the conv epilogue.
    -> (y, used_tc, sums): sums is an Acc (single-pass 64 -> 64 tensor-core conv) or an fp32 [2, Cout] tensor; both
    are what bn_forward(sums=...) takes."""
    if not x.is_contiguous():
        x = x.contiguous()
    if _acc_conv_ok(x, weight):
        n, cin, h, w = geometry(x, False)
        pk = packed_weight(weight, L.PACK_FPROP_TC, 0)
        y = new_act(n, 64, h, w, x.dtype, x.device)
        a = acc_acquire(x.device)
        rc = _timed(("conv_fprop", cin, 64, 3, 0, n, h, w, True),
                    lambda: L.cdll.srk_conv_fprop(act_desc(x), act_desc(y), pk.data_ptr(), L.PACK_FPROP_TC, 64, 3, 3, _ptr(bias),
                                                  L.ACT_NONE, None, None, 0, L.IMPL_AUTO, None, None, a.data_ptr(), None,
                                                  None, stream_ptr()))
        if rc == 0:
            L.launch_calls += 1
            return y, True, a
        a.dirty = False          # nothing was launched (rc 2: another kernel variant is selected) or the call failed
        if rc != 2:
            raise RuntimeError("srk_conv_fprop failed: %s" % L.last_error())
    sums = torch.empty((2, weight.shape[0]), dtype=torch.float32, device=x.device)
    y, used_tc = conv_fprop(x, False, weight, bias, L.ACT_NONE, None, None, 0, False, x.dtype, bn_sums=sums)
    return y, used_tc, sums


def conv_fprop(x, x_img, weight, bias, act, alpha, residual, shuffle, out_img, out_dtype, bn_sums=None, zsave=False):
    """y = [shuffle](act(conv(x, weight) + bias)) [+ residual]; stride 1, pad R//2.
    bn_sums: optional fp32 [2, Cout] that receives the per-channel sum / sum of squares of y.
    zsave=True (PReLU, ACT output): also returns the prelu_z buffer -> (y, used_tc, z)."""
    n, cin, h, w = geometry(x, x_img)
    cout, wcin, r, s = weight.shape
    assert wcin == cin, "conv: input has %d channels, weight expects %d" % (cin, wcin)
    if (x_img and cin == 3 and cout in (64, 96) and not out_img and out_dtype == torch.bfloat16 and residual is None
            and shuffle == 0 and _rgb_tc_ok(r, s)):
        pk = packed_weight(weight, L.PACK_RGBIN_TC, 0)
        y = new_act(n, cout, h, w, out_dtype, x.device)
        z = prelu_z_like(y) if (zsave and act == L.ACT_PRELU) else None
        _timed(("conv_rgbin_fprop", cin, cout, r, 0, n, h, w, True),
               lambda: L.call("srk_conv_rgb_fprop", img_desc(x), act_desc(y), pk.data_ptr(), r, _ptr(bias), act,
                              _ptr(alpha), act_desc(z) if z is not None else None, stream_ptr()))
        return (y, True, z) if zsave else (y, True)
    tc = 0 if x_img else tc_supported(cin, cout, r, s, x.dtype, shuffle)
    rgb_tc = tc == 2 and out_img and act == L.ACT_NONE and residual is None
    use_tc = rgb_tc or (tc == 1 and (not out_img) and out_dtype == torch.bfloat16)
    kind = L.PACK_FPROP_TC_N8 if rgb_tc else (L.PACK_FPROP_TC if use_tc else L.PACK_FPROP_SIMT)
    # tensor-core PixelShuffle convs take their weight rows sub-pixel-major (row = sub * Cout/4 + c): a pass of 64
    # rows then produces all channels of ONE output pixel per input pixel and stores whole 64-byte runs
    pk = packed_weight(weight, kind, 2 if (shuffle == 2 and kind == L.PACK_FPROP_TC) else 0)
    if shuffle == 2:
        oc, oh, ow = cout // 4, 2 * h, 2 * w
    else:
        oc, oh, ow = cout, h, w
    y = new_image(n, oc, oh, ow, x.device) if out_img else new_act(n, oc, oh, ow, out_dtype, x.device)
    xd, yd = desc(x, x_img), desc(y, out_img)
    rd = desc(residual, out_img) if residual is not None else None
    ws = _fprop_workspace(xd, kind, x.device)
    z = prelu_z_like(y) if (zsave and act == L.ACT_PRELU and not out_img) else None
    _timed(("conv_fprop", cin, cout, r, shuffle, n, h, w, use_tc),
           lambda: L.call("srk_conv_fprop", xd, yd, pk.data_ptr(), kind, cout, r, s, _ptr(bias), act,
                          _ptr(alpha), rd, shuffle, L.IMPL_AUTO, _ptr(bn_sums),
                          reduce_ws(x.device) if bn_sums is not None else None, None,
                          act_desc(z) if z is not None else None, _ptr(ws), stream_ptr()))
    return (y, use_tc, z) if zsave else (y, use_tc)


def conv_dgrad(dz, dz_img, weight, residual, out_dtype, perm_tc=False):
    """dx = conv_transpose(dz, weight) [+ residual]  (fprop over dz with rotated, transposed taps)."""
    n, c, h, w = geometry(dz, dz_img)
    cout, cin, r, s = weight.shape
    assert c == cout
    use_tc = (not dz_img) and tc_supported(cout, cin, r, s, dz.dtype, 0) == 1 and out_dtype == torch.bfloat16
    if perm_tc and not use_tc:
        raise RuntimeError("conv_dgrad: permuted dz requires the tcgen05 path")
    kind = L.PACK_DGRAD_TC if use_tc else L.PACK_DGRAD_SIMT
    pk = packed_weight(weight, kind, 2 if perm_tc else 0)   # perm: dz channels are sub-pixel-major
    dx = new_act(n, cin, h, w, out_dtype, dz.device)
    rd = act_desc(residual) if residual is not None else None
    dzd = desc(dz, dz_img)
    ws = _fprop_workspace(dzd, kind, dz.device)
    _timed(("conv_dgrad", cout, cin, r, 0, n, h, w, use_tc),
           lambda: L.call("srk_conv_fprop", dzd, act_desc(dx), pk.data_ptr(), kind, cin, r, s,
                          None, L.ACT_NONE, None, rd, 0, L.IMPL_AUTO, None, None, None, None, _ptr(ws), stream_ptr()))
    return dx


def conv_dgrad_bnred(dz, weight, z, stats, gamma, beta, alpha, residual=None):
    """dx = dgrad(dz) with the backward reduction of the BatchNorm layer below fused into the epilogue
    (srk_conv_dgrad_bnred).  z / stats / gamma / beta: that BN's saved input, (mean, invstd) and affine parameters;
    alpha: slope of the PReLU between the BN and this conv, or None.  residual: added to the dgrad before the
    reduction (dx = dgrad(dz) + residual: the whole gradient of a residual block's input, reduced against the bn2 of
    the block below).
    -> (dx, red) with red = [sum g | sum g*z | dalpha] (an Acc, or fp32 [2C+1] with SRK_ACC=0), or None when the
    fused kernel does not cover the shape (the caller then runs conv_dgrad and the stand-alone reduction)."""
    cout, cin, r, s = weight.shape
    if not (cfg.fuse_bn_reduce and cfg.conv_impl != "simt" and r == 3 and s == 3 and cin == 64 and cout == 64
            and dz.dtype == torch.bfloat16 and z.dtype == torch.bfloat16 and dz.shape == z.shape
            and (residual is None or (residual.dtype == torch.bfloat16 and residual.shape == dz.shape))):
        return None
    n, c, h, w = geometry(dz, False)
    pk = packed_weight(weight, L.PACK_DGRAD_TC, 0)
    dx = new_act(n, cin, h, w, torch.bfloat16, dz.device)
    rd = act_desc(residual) if residual is not None else None
    rc = 2
    if cfg.use_acc:
        red = acc_acquire(dz.device)
        rc = L.cdll.srk_conv_dgrad_bnred(act_desc(dz), act_desc(dx), pk.data_ptr(), act_desc(z), stats[0].data_ptr(),
                                         stats[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(alpha),
                                         None, None, None, rd, None, red.data_ptr(), stream_ptr())
        if rc != 0:
            red.dirty = False
    if rc == 2:
        red = torch.empty((2 * cin + 1,), dtype=torch.float32, device=dz.device)
        rc = L.cdll.srk_conv_dgrad_bnred(act_desc(dz), act_desc(dx), pk.data_ptr(), act_desc(z), stats[0].data_ptr(),
                                         stats[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(alpha),
                                         red[:cin].data_ptr(), red[cin:2 * cin].data_ptr(),
                                         red[2 * cin:].data_ptr() if alpha is not None else None, rd,
                                         reduce_ws(dz.device), None, stream_ptr())
    if rc == 2:
        return None
    if rc != 0:
        raise RuntimeError("srk_conv_dgrad_bnred failed: %s" % L.last_error())
    L.launch_calls += 1
    return dx, red


def conv_wgrad(x, x_img, dz, dz_img, weight, need_bias, perm_tc=False, side=False, bias=None):
    """-> (dW fp32 OIHW, db fp32 [Cout] or None).
    side=True asks for the side-stream path (cfg.overlap_wgrad): the kernels then run beside the following layers'
    backward passes, the gradients are assigned to weight.grad / bias.grad when the streams join at the end of the
    backward pass, and (None, None) is returned - the caller hands exactly that to autograd.  `bias` is the bias
    PARAMETER (needed to assign its gradient); parameters with hooks stay on the ordinary path."""
    # frozen layer (requires_grad = False on the weight and on the bias): no gradient is wanted - and none must appear
    # in .grad through the side-stream path, where a later optimizer over model.parameters() would pick it up
    if isinstance(weight, torch.nn.Parameter) and not weight.requires_grad and \
            (bias is None or not (torch.is_tensor(bias) and bias.requires_grad)):
        return None, None

    def _hooked(t):
        # user hooks would not fire for a gradient that bypasses autograd; the one hook srk.dp.GradAverager registers is
        # served through set_side_grad_listener instead and does not count
        if t is None:
            return False
        post = getattr(t, "_post_accumulate_grad_hooks", None)
        n_post = len(post) if post else 0
        return bool(t._backward_hooks) or n_post > (1 if getattr(t, "_srk_dp_hooked", False) else 0)
    side = (side and cfg.overlap_wgrad and torch.is_tensor(weight) and weight.is_leaf and not _hooked(weight)
            and (not need_bias or (bias is not None and bias.is_leaf and not _hooked(bias))))
    cout, cin, r, s = weight.shape

    def finish(launch, keep, dw, db):
        if not side:
            launch()
            return dw, db
        run_on_side_stream(launch, keep, [(weight, dw)] + ([(bias, db)] if db is not None else []))
        return None, None

    if not (x_img and (not dz_img) and cin == 3 and cout in (64, 96) and dz.dtype == torch.bfloat16 and _rgb_tc_ok(r, s)):
        # general path: the kernels write (accumulate = 0), so no zero-fill launches are needed
        dw = torch.empty_like(weight, memory_format=torch.contiguous_format)
        db = torch.empty((cout,), dtype=torch.float32, device=weight.device) if need_bias else None
        xd, dd = desc(x, x_img), desc(dz, dz_img)
        impl = L.IMPL_SIMT if cfg.conv_impl == "simt" else L.IMPL_AUTO
        nbytes = L.cdll.srk_conv_wgrad_workspace_bytes(xd, dd, r, s, impl)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=weight.device) if nbytes > 0 else None
        n, _, h, w = geometry(x, x_img)
        on_tc = impl != L.IMPL_SIMT and not x_img and not dz_img and x.dtype == torch.bfloat16 and r == 3
        launch = lambda: _timed(("conv_wgrad", cin, cout, r, 0, n, h, w, on_tc),
                                lambda: L.call("srk_conv_wgrad", xd, dd, dw.data_ptr(), _ptr(db), r, s, impl, 0,
                                               1 if perm_tc else 0, _ptr(ws), stream_ptr()))
        return finish(launch, (x, dz, ws), dw, db)
    dw = torch.empty_like(weight, memory_format=torch.contiguous_format)     # written, not accumulated
    db = torch.empty((cout,), dtype=torch.float32, device=weight.device) if need_bias else None
    n, _, h, w = geometry(x, True)
    ws = _rgb_workspace(r, x.device)
    launch = lambda: _timed(("conv_rgbin_wgrad", cin, cout, r, 0, n, h, w, True),
                            lambda: L.call("srk_conv_rgb_bwd", img_desc(x), act_desc(dz), None, None,
                                           dw.data_ptr(), _ptr(db), r, 0, ws.data_ptr(), stream_ptr()))
    return finish(launch, (x, dz, ws), dw, db)


def act_bwd(dout, out, act, alpha, unshuffle, perm_tc=False, zsave=None):
    """gradient of the pre-activation (conv-output geometry) from the saved post-activation tensor (and, for a PReLU
    whose slope is <= 0, from the pre-activation copy `zsave` the forward epilogue wrote)"""
    n, c, h, w = geometry(out, False)
    if unshuffle == 2:
        dz = new_act(n, 4 * c, h // 2, w // 2, out.dtype, out.device)
    else:
        dz = torch.empty_like(out)
    dalpha = torch.empty((1,), dtype=torch.float32, device=out.device) if act == L.ACT_PRELU else None
    L.call("srk_act_bwd", act_desc(dout), act_desc(out), act_desc(zsave) if zsave is not None else None, act_desc(dz),
           act, _ptr(alpha), _ptr(dalpha), unshuffle, 1 if perm_tc else 0,
           reduce_ws(out.device) if dalpha is not None else None, stream_ptr())
    return dz, dalpha


# ---- batch norm ------------------------------------------------------------------------------------
def bn_needs_batch_stats(running_mean, training):
    return training or running_mean is None


def bn_forward(y, gamma, beta, running_mean, running_var, nbt, training, eps, momentum, alpha, residual, sums=None):
    """out = [PReLU](BN(y)) [+ residual]; returns (out, stats[2, C] = mean, invstd).
    sums: per-channel (sum, sum of squares) of y when the producing conv already accumulated them."""
    n, c, h, w = geometry(y, False)
    dev = y.device
    stats = torch.empty((2, c), dtype=torch.float32, device=dev)
    mean, invstd = stats[0], stats[1]
    st = stream_ptr()
    if training or running_mean is None:
        if sums is None:
            sums = torch.empty((2, c), dtype=torch.float32, device=dev)
            L.call("srk_bn_stats", act_desc(y), sums.data_ptr(), reduce_ws(dev), st)
        upd = training and running_mean is not None
        # statistics -> mean / invstd (+ running-stat update) happen inside the apply kernel: one launch per layer
        out = torch.empty_like(y)
        is_acc = isinstance(sums, Acc)
        L.call("srk_bn_apply_train", act_desc(y), None if is_acc else sums[0].data_ptr(),
               None if is_acc else sums[1].data_ptr(), n * h * w, eps, momentum,
               _ptr(running_mean) if upd else None, _ptr(running_var) if upd else None,
               _ptr(nbt) if upd else None, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
               _ptr(alpha), act_desc(residual) if residual is not None else None, act_desc(out),
               sums.data_ptr() if is_acc else None, st)
        if is_acc:
            sums.dirty = False
        return out, stats
    else:
        if isinstance(sums, Acc):
            acc_discard(sums)
        L.call("srk_bn_eval_params", running_mean.data_ptr(), running_var.data_ptr(), c, eps,
               mean.data_ptr(), invstd.data_ptr(), st)
    out = torch.empty_like(y)
    L.call("srk_bn_apply", act_desc(y), mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
           _ptr(alpha), act_desc(residual) if residual is not None else None, act_desc(out), st)
    return out, stats


def bn_backward(dout, y, stats, gamma, beta, alpha, batch_stats, pre=None):
    """-> (dy, dgamma, dbeta, dalpha or None).  pre: the raw sums of conv_dgrad_bnred when the reduction already
    happened in the epilogue of the dgrad that produced dout."""
    c = y.shape[3]
    dev = y.device
    if isinstance(pre, Acc):
        out = torch.empty((2 * c + 1,), dtype=torch.float32, device=dev)
        dgamma, dbeta, dalpha = out[:c], out[c:2 * c], out[2 * c:]
        dy = torch.empty_like(y)
        L.call("srk_bn_bwd_apply_raw", act_desc(dout), act_desc(y), stats[0].data_ptr(), stats[1].data_ptr(),
               gamma.data_ptr(), beta.data_ptr(), _ptr(alpha), None, None, 1 if batch_stats else 0, dgamma.data_ptr(),
               act_desc(dy), pre.data_ptr(), dbeta.data_ptr(), dalpha.data_ptr() if alpha is not None else None,
               stream_ptr())
        pre.dirty = False
        return dy, dgamma, dbeta, (dalpha if alpha is not None else None)
    if pre is not None:
        sum_g, sum_gz, dalpha = pre[:c], pre[c:2 * c], pre[2 * c:]
        dgamma = torch.empty((c,), dtype=torch.float32, device=dev)
        dy = torch.empty_like(y)
        L.call("srk_bn_bwd_apply_raw", act_desc(dout), act_desc(y), stats[0].data_ptr(), stats[1].data_ptr(),
               gamma.data_ptr(), beta.data_ptr(), _ptr(alpha), sum_g.data_ptr(), sum_gz.data_ptr(),
               1 if batch_stats else 0, dgamma.data_ptr(), act_desc(dy), None, None, None, stream_ptr())
        return dy, dgamma, sum_g, (dalpha if alpha is not None else None)
    red = torch.empty((2 * c + 1,), dtype=torch.float32, device=dev)
    dgamma, dbeta, dalpha = red[:c], red[c:2 * c], red[2 * c:]
    mean, invstd = stats[0], stats[1]
    st = stream_ptr()
    L.call("srk_bn_bwd_reduce", act_desc(dout), act_desc(y), mean.data_ptr(), invstd.data_ptr(),
           gamma.data_ptr(), beta.data_ptr(), _ptr(alpha), dgamma.data_ptr(), dbeta.data_ptr(),
           dalpha.data_ptr() if alpha is not None else None, reduce_ws(dev), st)
    dy = torch.empty_like(y)
    L.call("srk_bn_bwd_apply", act_desc(dout), act_desc(y), mean.data_ptr(), invstd.data_ptr(),
           gamma.data_ptr(), beta.data_ptr(), _ptr(alpha), dgamma.data_ptr(), dbeta.data_ptr(),
           1 if batch_stats else 0, act_desc(dy), st)
    return dy, dgamma, dbeta, (dalpha if alpha is not None else None)


# ---- squeeze-excite --------------------------------------------------------------------------------
def se_forward(x, r, w1, w2, scale):
    """out = x + scale * r * sigmoid(relu(mean_hw(r) @ w1^T) @ w2^T); returns (out, pool, hidden, gate)."""
    n, c, h, w = geometry(r, False)
    cr = w1.shape[0]
    dev = r.device
    pool = torch.empty((n, c), dtype=torch.float32, device=dev)
    hidden = torch.empty((n, cr), dtype=torch.float32, device=dev)
    gate = torch.empty((n, c), dtype=torch.float32, device=dev)
    st = stream_ptr()
    L.call("srk_se_pool", act_desc(r), pool.data_ptr(), reduce_ws(dev), st)
    L.call("srk_se_fc", pool.data_ptr(), w1.data_ptr(), w2.data_ptr(), n, c, cr, hidden.data_ptr(),
           gate.data_ptr(), st)
    out = torch.empty_like(r)
    L.call("srk_se_apply", act_desc(x) if x is not None else None, act_desc(r), gate.data_ptr(), scale,
           act_desc(out), st)
    return out, pool, hidden, gate


def se_backward(dout, r, pool, hidden, gate, w1, w2, scale):
    """-> (dr, dw1, dw2)   (the skip path's gradient is dout itself)"""
    n, c, h, w = geometry(r, False)
    cr = w1.shape[0]
    dev = r.device
    st = stream_ptr()
    dgate_raw = torch.empty((n, c), dtype=torch.float32, device=dev)
    L.call("srk_se_bwd_reduce", act_desc(dout), act_desc(r), dgate_raw.data_ptr(), reduce_ws(dev), st)
    dw1 = torch.empty_like(w1, memory_format=torch.contiguous_format)     # written by the kernel
    dw2 = torch.empty_like(w2, memory_format=torch.contiguous_format)
    # dpool [n][c] followed by the kernel's scratch (dz2 [n][c], dh [n][cr])
    dpool = torch.empty((2 * n * c + n * cr,), dtype=torch.float32, device=dev)
    L.call("srk_se_fc_bwd", dgate_raw.data_ptr(), gate.data_ptr(), hidden.data_ptr(), pool.data_ptr(),
           w1.data_ptr(), w2.data_ptr(), n, c, cr, scale, dw1.data_ptr(), dw2.data_ptr(), dpool.data_ptr(), st)
    dr = torch.empty_like(r)
    L.call("srk_se_bwd_apply", act_desc(dout), gate.data_ptr(), dpool.data_ptr(), scale, act_desc(dr), st)
    return dr, dw1, dw2


# ---- layout / misc ---------------------------------------------------------------------------------
def image_to_act(img, dtype):
    n, c, h, w = img.shape
    a = new_act(n, c, h, w, dtype, img.device)
    L.call("srk_image_to_act", img_desc(img), act_desc(a), stream_ptr())
    return a


def act_to_image(a):
    n, c, h, w = geometry(a, False)
    img = new_image(n, c, h, w, a.device)
    L.call("srk_act_to_image", act_desc(a), img_desc(img), stream_ptr())
    return img


def act_add(a, b):
    out = torch.empty_like(a)
    L.call("srk_act_add", act_desc(a), act_desc(b), act_desc(out), stream_ptr())
    return out


def maxpool2_fwd(x):
    """2x2 / stride 2 max pooling (floor mode) of an act tensor."""
    n, c, h, w = geometry(x, False)
    if h < 2 or w < 2:
        raise ValueError("maxpool2: input %dx%d is smaller than the window" % (h, w))
    out = new_act(n, c, h // 2, w // 2, x.dtype, x.device)
    L.call("srk_maxpool2_fwd", act_desc(x), act_desc(out), stream_ptr())
    return out


def maxpool2_bwd(x, dout):
    dx = torch.empty_like(x)
    L.call("srk_maxpool2_bwd", act_desc(x), act_desc(dout), act_desc(dx), stream_ptr())
    return dx


def bicubic_upsample(img, oh, ow):
    n, c, h, w = img.shape
    out = new_image(n, c, oh, ow, img.device)
    L.call("srk_bicubic_upsample", img_desc(img), img_desc(out), stream_ptr())
    return out
