"""A training step on libsrk is bit-reproducible: no kernel accumulates floating-point values with atomics
(include/srk.h "deterministic reductions"): BatchNorm statistics and backward reductions, PReLU-slope and
squeeze-excite gradients, weight / bias gradients and loss values are per-block partials folded in block order.
Round 1 measured 3-4e-2 run-to-run differences on a BatchNorm gradient from float atomics amplified by bf16
rounding; here two runs must agree in every bit - outputs, loss, every gradient, the BatchNorm buffers - in both
arithmetic modes, with the weight gradients on the main or on the side stream, eagerly and from a replayed CUDA graph.

Also here: whole networks whose PReLU slopes are <= 0 against the CPU oracle (see test_gpu_parity.py for the
single-layer cases), and guard-band (canary) checks around kernel outputs."""
import pytest
import torch

from helpers import max_abs, rel_err
from oracle import sr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _modes():
    import srk
    srk.set_compute_dtype("fp32")
    srk.set_conv_impl("auto")
    srk.set_overlap_wgrad(False)
    yield
    srk.set_compute_dtype("fp32")
    srk.set_overlap_wgrad(False)


def _make(arch):
    from src import models as M
    return {"RESNET": lambda: M.ResNetSR(num_channels=64, num_residuals=3),
            "AttentionSR": lambda: M.AttentionSR(num_channels=96, num_residuals=2),
            "SRCNN": lambda: M.SRCNN(scale_factor=4)}[arch]()


def _step(arch, sd, lr, hr, loss_name):
    from src.loss import get_loss_function
    model = _make(arch)
    model.load_state_dict(sd)
    model = model.to(DEV).train()
    out = model(lr)
    loss = get_loss_function(loss_name, DEV)(out, hr)
    loss.backward()
    torch.cuda.synchronize()
    res = {"out": out.detach().clone(), "loss": loss.detach().clone()}
    for k, p in model.named_parameters():
        res["grad/" + k] = p.grad.detach().clone()
    for k, v in model.state_dict().items():
        if "running_" in k:
            res["buf/" + k] = v.detach().clone()
    return res


@pytest.mark.parametrize("overlap", [False, True])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("arch,loss_name", [("RESNET", "nlpd"), ("AttentionSR", "mae"), ("SRCNN", "mse")])
def test_training_step_is_bit_reproducible(arch, loss_name, dtype, overlap):
    import srk
    srk.set_compute_dtype(dtype)
    srk.set_overlap_wgrad(overlap)
    torch.manual_seed(13)
    sd = {k: v.clone() for k, v in _make(arch).state_dict().items()}
    lr, hr = O.synthetic_pair(5, 40, 36, 4, seed=31)      # several tiles per CTA, ragged last tile
    lr, hr = lr.to(DEV), hr.to(DEV)
    a = _step(arch, sd, lr, hr, loss_name)
    # something else on the GPU between the runs, so that block scheduling differs
    torch.randn(1 << 22, device=DEV).sum().item()
    b = _step(arch, sd, lr, hr, loss_name)
    c = _step(arch, sd, lr, hr, loss_name)
    for k in a:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), (k, max_abs(a[k].float(), b[k].float()))


def test_graph_replay_is_bit_reproducible_and_matches_eager():
    """The captured step (srk.trainer.GraphStep - what bench.py and train.py run) against the same step launched
    eagerly, and two replays against each other."""
    import srk
    from srk.trainer import GraphStep
    from src.loss import get_loss_function
    srk.set_compute_dtype("bf16")
    srk.set_overlap_wgrad(True)
    torch.manual_seed(5)
    sd = {k: v.clone() for k, v in _make("RESNET").state_dict().items()}
    lr, hr = O.synthetic_pair(4, 32, 32, 4, seed=8)
    lr, hr = lr.to(DEV), hr.to(DEV)
    runs = []
    for use_graph in (False, True, True):
        model = _make("RESNET")
        model.load_state_dict(sd)
        model = model.to(DEV).train()
        step = GraphStep(model, get_loss_function("nlpd", DEV), lr=4e-4, use_graph=use_graph, warmup=2)
        losses = [float(step(lr, hr)) for _ in range(4)]
        torch.cuda.synchronize()
        runs.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    for losses, state in runs[1:]:
        assert losses == runs[0][0], (losses, runs[0][0])
        for k, v in state.items():
            assert torch.equal(v, runs[0][1][k]), k


def _slope_net_errors(arch, loss_name, slopes):
    """Output / gradient errors of a 2-block network whose PReLU slopes cycle through `slopes`, against the fp32 CPU
    oracle -> (forward error (max-abs in fp32 mode, relative in bf16 mode), {param: relative gradient error})."""
    import srk
    from src import models as M
    from src.loss import get_loss_function
    torch.manual_seed(17)
    model = M.ResNetSR(num_channels=64, num_residuals=2) if arch == "RESNET" else M.AttentionSR(num_channels=64, num_residuals=2)
    with torch.no_grad():
        i = 0
        for k, p in model.named_parameters():
            if p.numel() == 1:
                p.fill_(slopes[i % len(slopes)])
                i += 1
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    lr, hr = O.synthetic_pair(3, 20, 24, 4, seed=41)
    out_ref, _, grads_ref, _ = O.train_step_grads(arch, sd, lr, hr, loss_name)
    model = model.to(DEV).train()
    out = model(lr.to(DEV))
    get_loss_function(loss_name, DEV)(out, hr.to(DEV)).backward()
    fwd = max_abs(out.cpu(), out_ref) if srk.cfg.compute_dtype == torch.float32 else rel_err(out.cpu(), out_ref)
    errs = {}
    for k, p in model.named_parameters():
        if k.endswith(".bias") and (".conv" in k or k.startswith("mid_conv")) and arch == "RESNET":
            continue    # analytically zero under a training-mode BatchNorm
        errs[k] = rel_err(p.grad.cpu(), grads_ref[k], floor=1e-6)
    return fwd, errs


@pytest.mark.parametrize("dtype,ftol,gtol", [("fp32", 1e-4, 2e-3), ("bf16", 1e-2, 5e-2)])
@pytest.mark.parametrize("arch,loss_name", [("RESNET", "mae"), ("AttentionSR", "mae")])
def test_networks_with_nonpositive_prelu_slopes_vs_oracle(arch, loss_name, dtype, ftol, gtol):
    """Every PReLU of the network with a slope <= 0 (alternating -0.2 and exactly 0): input conv (RGB-input kernel),
    trunk PReLUs (conv epilogue or BatchNorm-fused), both PixelShuffle stages and - in bf16 - the fused
    upsample-tail backward (srk_conv_rgbout_bwd_unshuffle) must follow the pre-activation, not the output.
    Yardstick: the same network, data and arithmetic with the usual positive slopes (0.25) - a wrong branch anywhere
    shows up as an error of order one, not as a factor on rounding noise."""
    import srk
    srk.set_compute_dtype(dtype)
    fwd_p, errs_p = _slope_net_errors(arch, loss_name, [0.25])
    fwd_n, errs_n = _slope_net_errors(arch, loss_name, [-0.2, 0.0])
    # tensor-valued gradients and the single-number slope gradients apart: a slope gradient is one heavily cancelling
    # sum (each layer's is held to 1e-2 on its own in test_prelu_epilogues_with_nonpositive_slope); through a network
    # in bf16 it is the noisiest number of the step for either sign of the slope
    is_slope = lambda k: k.endswith("prelu.weight") or k in ("upsample.2.weight", "upsample.5.weight")
    t_p = max(v for k, v in errs_p.items() if not is_slope(k))
    t_n = max(v for k, v in errs_n.items() if not is_slope(k))
    s_p = max(v for k, v in errs_p.items() if is_slope(k))
    s_n = max(v for k, v in errs_n.items() if is_slope(k))
    msg = {"fwd_pos": fwd_p, "fwd_nonpos": fwd_n, "grad_pos": t_p, "grad_nonpos": t_n, "slope_grad_pos": s_p,
           "slope_grad_nonpos": s_n, "worst": sorted(errs_n.items(), key=lambda kv: -kv[1])[:4]}
    print("\n[prelu slopes %s %s] %s" % (arch, dtype, msg))
    assert fwd_n <= max(ftol, 2.0 * fwd_p), msg
    assert t_n <= max(gtol, 2.0 * t_p), msg
    assert s_n <= max(10 * gtol, 3.0 * s_p), msg


def test_kernels_do_not_write_outside_their_outputs():
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are looked for with guard bands: outputs,
    gradients and workspaces of the tensor-core kernels are carved out of one poisoned arena with 64 KB of canary on
    either side of every buffer; after a forward / backward through each kernel family the canaries must be intact."""
    import srk
    from srk import ops
    from srk import _lib as L
    srk.set_compute_dtype("bf16")
    CANARY = 0x5A
    arena = torch.full((512 << 20,), CANARY, dtype=torch.uint8, device=DEV)
    guards, cursor = [], [0]

    def carve(shape, dtype):
        nbytes = int(torch.empty((), dtype=dtype).element_size())
        for d in shape:
            nbytes *= d
        g = 64 << 10
        start = (cursor[0] + g + 1023) // 1024 * 1024
        guards.append((cursor[0], start))
        cursor[0] = start + nbytes
        assert cursor[0] + g <= arena.numel()
        return arena[start:start + nbytes].view(dtype).view(shape)

    real_empty, real_empty_like = torch.empty, torch.empty_like

    def empty(*size, dtype=None, device=None, **kw):
        if device is None or torch.device(device).type != "cuda":
            return real_empty(*size, dtype=dtype, device=device, **kw)
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        return carve(shape, dtype or torch.float32)

    def empty_like(t, **kw):
        if not t.is_cuda:
            return real_empty_like(t, **kw)
        return carve(tuple(t.shape), kw.get("dtype", t.dtype))

    g_ = torch.Generator(device=DEV).manual_seed(3)

    def act(n, h, w, c):
        t = torch.zeros(n, h + 2, w + 2, c, device=DEV, dtype=torch.bfloat16)
        t[:, 1:-1, 1:-1] = torch.randn(n, h, w, c, generator=g_, device=DEV).bfloat16()
        return t

    torch.empty, torch.empty_like = empty, empty_like
    try:
        alpha = torch.tensor([0.25], device=DEV)
        for (n, h, w) in [(2, 13, 17), (1, 31, 9), (3, 8, 8)]:
            x64, g64 = act(n, h, w, 64), act(n, h, w, 64)
            x96, g96 = act(n, h, w, 96), act(n, h, w, 96)
            w64 = torch.randn(64, 64, 3, 3, generator=g_, device=DEV) / 24
            w96 = torch.randn(96, 96, 3, 3, generator=g_, device=DEV) / 30
            wup = torch.randn(256, 64, 3, 3, generator=g_, device=DEV) / 24
            wout = torch.randn(3, 64, 9, 9, generator=g_, device=DEV) / 72
            win = torch.randn(64, 3, 9, 9, generator=g_, device=DEV) / 16
            img = torch.rand(n, 3, h, w, generator=g_, device=DEV)
            sums = torch.empty((2, 64), dtype=torch.float32, device=DEV)
            y, _ = ops.conv_fprop(x64, False, w64, None, L.ACT_NONE, None, None, 0, False, torch.bfloat16, bn_sums=sums)
            ops.conv_fprop(x64, False, w64, None, L.ACT_PRELU, alpha, None, 0, False, torch.bfloat16)
            ops.conv_fprop(x96, False, w96, None, L.ACT_PRELU, alpha, None, 0, False, torch.bfloat16)
            up, _ = ops.conv_fprop(x64, False, wup, None, L.ACT_PRELU, alpha, None, 2, False, torch.bfloat16)
            ops.conv_fprop(up, False, wout, None, L.ACT_NONE, None, None, 0, True, torch.float32)
            yin, _ = ops.conv_fprop(img, True, win, None, L.ACT_PRELU, alpha, None, 0, False, torch.bfloat16)
            ops.conv_dgrad(g64, False, w64, x64, torch.bfloat16)
            ops.conv_dgrad(g96, False, w96, None, torch.bfloat16)
            ops.conv_wgrad(x64, False, g64, False, w64, True)
            ops.conv_wgrad(x96, False, g96, False, w96, True)
            ops.conv_wgrad(img, True, g64, False, win, True)
            gimg = torch.randn(n, 3, 2 * h, 2 * w, generator=g_, device=DEV)
            ops.conv_rgbout_bwd_unshuffle(up, gimg, wout, alpha, True)
            ops.conv_rgbout_bwd(up, gimg, wout, True, True)
            gamma, beta = torch.ones(64, device=DEV), torch.zeros(64, device=DEV)
            _, stats = ops.bn_forward(y, gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, x64, sums=sums)
            fused = ops.conv_dgrad_bnred(g64, w64, y, stats, gamma, beta, alpha)
            ops.bn_backward(fused[0], y, stats, gamma, beta, alpha, True, pre=fused[1])
            fused = ops.conv_dgrad_bnred(g64, w64, y, stats, gamma, beta, None, residual=x64)
            ops.bn_backward(fused[0], y, stats, gamma, beta, None, True, pre=fused[1])
            ya, _, acc = ops.conv_fprop_stats(x64, w64, None)       # statistics through the integer accumulator
            ops.bn_forward(ya, gamma, beta, None, None, None, True, 1e-5, 0.1, alpha, x64, sums=acc)
            ops.bn_backward(g64, y, stats, gamma, beta, alpha, True)
            ops.act_bwd(g64, yin, L.ACT_PRELU, alpha, 0)
        torch.cuda.synchronize()
    finally:
        torch.empty, torch.empty_like = real_empty, real_empty_like
    guards.append((cursor[0], cursor[0] + (64 << 10)))
    assert len(guards) > 60
    for a, b in guards:
        assert bool((arena[a:b] == CANARY).all()), "guard band [%d, %d) was overwritten" % (a, b)
