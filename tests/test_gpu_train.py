"""The reference's entry point on libsrk: train.train(config) (reference train.py:21-197) end to end on synthetic
Food101-shaped data - the captured step (srk.trainer.GraphStep), validation through the sharded evaluator, the LR
scheduler, the best-PSNR checkpoint and the final test metrics."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _config(**kw):
    cfg = dict(architecture="RESNET", batch_size=8, lr=4e-4, epochs=2, loss_function="nlpd", subset=1.0,
               pretrained_weights="", patience=5, save_name="t")
    cfg.update(kw)
    return cfg


@pytest.fixture
def synthetic_env(tmp_path, monkeypatch):
    monkeypatch.setenv("SR_SYNTHETIC_DATA", "44")
    monkeypatch.setenv("SR_CROP", "64")
    monkeypatch.setenv("WANDB_MODE", "disabled")
    monkeypatch.setenv("SRK_VGG_WEIGHTS", "none")
    monkeypatch.chdir(tmp_path)
    import srk
    yield tmp_path
    srk.set_compute_dtype("fp32")
    srk.set_overlap_wgrad(False)


@pytest.mark.parametrize("arch,loss", [("RESNET", "nlpd"), ("SRCNN", "mae")])
def test_train_entry_point_runs_two_epochs(synthetic_env, arch, loss):
    import train
    from srk import _lib as L
    c0 = L.launch_calls
    metrics = train.train(_config(architecture=arch, loss_function=loss))
    assert set(metrics) == {"psnr", "ssim", "lpips", "nlpd"}
    assert math.isfinite(metrics["psnr"]) and math.isfinite(metrics["ssim"]) and math.isfinite(metrics["nlpd"])
    assert L.launch_calls > c0, "no libsrk kernel was launched"
    ck = synthetic_env / "weights" / "t_best.pth"
    assert ck.exists()
    sd = torch.load(ck, map_location="cpu")
    # the checkpoint has the reference's state_dict keys: it loads into the drop-in and (where available) the reference
    from src.models import get_model
    get_model(arch, 4, "cpu").load_state_dict(sd, strict=True)
    from oracle import ref_modules
    if ref_modules.available():
        ref_modules.load().models.get_model(arch, 4, "cpu").load_state_dict(sd, strict=True)


def test_train_loss_decreases_and_graph_matches_eager(synthetic_env, monkeypatch):
    """Twelve steps of the captured trainer on one fixed batch: the loss must go down, and SRK_GRAPH=0 (every kernel
    launched from Python) must give the same losses bit for bit."""
    import srk
    from oracle import sr_oracle as O
    from srk.trainer import GraphStep
    from src.loss import get_loss_function
    from src.models import get_model
    srk.set_compute_dtype("bf16")
    lr, hr = O.synthetic_pair(8, 16, 16, 4, seed=3)
    lr, hr = lr.to("cuda:0"), hr.to("cuda:0")
    curves = []
    for use_graph in (True, False):
        torch.manual_seed(0)
        model = get_model("RESNET", 4, "cuda:0").train()
        step = GraphStep(model, get_loss_function("nlpd", "cuda:0"), lr=4e-4, use_graph=use_graph, warmup=3)
        curves.append([float(step(lr, hr)) for _ in range(12)])
    assert curves[0] == curves[1], curves
    assert curves[0][-1] < 0.7 * curves[0][0], curves[0]


def test_gan_branch_runs(synthetic_env):
    """loss_function=gan (reference train.py:58-65,86-114): generator + MAE / perceptual / TV on libsrk, discriminator
    on torch; one epoch must run and produce finite metrics."""
    import train
    metrics = train.train(_config(architecture="SRCNN", loss_function="gan", epochs=1, batch_size=4))
    assert math.isfinite(metrics["psnr"]) and math.isfinite(metrics["nlpd"])


@pytest.mark.parametrize("scale,crop", [(4, 200), (4, 64), (2, 48), (3, 45)])
def test_gpu_sample_pipeline_matches_torchvision(scale, crop):
    """srk.data (crop + flip + ToTensor + antialiased bicubic on the GPU) against the reference's own transforms
    (reference dataset.py:17-39: RandomCrop, RandomHorizontalFlip, ToTensor, Resize(BICUBIC) on the tensor) applied on
    the CPU to the same uint8 images with the same torch seed - identical crops / flips, HR bit-identical, LR within
    1e-5 - and the CenterCrop (test split) variant."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from srk import data as D
    from src.dataset import synthetic_image_u8
    imgs = [synthetic_image_u8(crop + 13 + 7 * k, crop + 40 - 9 * k, seed=100 + k) for k in range(5)]
    imgs.append(synthetic_image_u8(crop, crop, seed=9))            # exactly the crop size: RandomCrop draws nothing
    src, sizes = D.collate_raw(imgs)
    lr_size = crop // scale
    down = transforms.Resize((lr_size, lr_size), interpolation=transforms.InterpolationMode.BICUBIC)
    for train in (True, False):
        tf = transforms.Compose(([transforms.RandomCrop(crop), transforms.RandomHorizontalFlip()] if train
                                 else [transforms.CenterCrop(crop)]) + [transforms.ToTensor()])
        torch.manual_seed(77)
        want_hr = torch.stack([tf(Image.fromarray(im.numpy())) for im in imgs])
        want_lr = torch.stack([down(h) for h in want_hr])
        torch.manual_seed(77)
        offs, flips = D.draw_crop_params(sizes, crop, train)
        lr, hr = D.make_batch(src.to("cuda:0"), offs, flips, crop, scale)
        assert hr.shape == want_hr.shape and lr.shape == want_lr.shape
        assert torch.equal(hr.cpu(), want_hr), "HR crop / flip / ToTensor"
        assert float((lr.cpu() - want_lr).abs().max()) <= 1e-5, float((lr.cpu() - want_lr).abs().max())
        if train:
            assert int(flips.sum()) not in (0, len(imgs))            # both orientations exercised
    # channels-first uint8 input gives the same result
    lr2, hr2 = D.make_batch(src.permute(0, 3, 1, 2).contiguous().to("cuda:0"), offs, flips, crop, scale)
    assert torch.equal(hr2, hr) and torch.equal(lr2, lr)
    # the in-place form (bench.py's e2e leg writes the batch straight into the captured step's input tensors)
    lr3, hr3 = torch.full_like(lr, -1.0), torch.full_like(hr, -1.0)
    D.make_batch_into(src.to("cuda:0"), offs.to("cuda:0"), flips.to("cuda:0"), crop, scale, lr3, hr3)
    assert torch.equal(hr3, hr) and torch.equal(lr3, lr)


def test_visualize_tool_runs_on_reference_checkpoints(synthetic_env, monkeypatch):
    """visualize.py (reference visualize.py:24-61,63-122): uint8 PSNR helper, get_prediction timing with a checkpoint
    in the reference's state_dict layout, PNG outputs."""
    import importlib
    monkeypatch.setenv("NUM_EXAMPLES", "2")
    monkeypatch.setenv("SR_SYNTHETIC_DATA", "3")
    monkeypatch.setenv("OUTPUT_DIR", str(synthetic_env / "report"))
    import visualize
    importlib.reload(visualize)
    from src.models import get_model
    os.makedirs(synthetic_env / "weights", exist_ok=True)
    torch.manual_seed(0)
    torch.save(get_model("SRCNN", 4, "cpu").state_dict(), synthetic_env / "weights" / "srcnn_nlpd_best.pth")
    times = visualize.run_comparison()
    assert set(times) == {"SRCNN"} and times["SRCNN"] > 0           # the other checkpoints are missing: skipped, as upstream
    files = sorted(os.listdir(next((synthetic_env / "report").iterdir())))
    assert files == ["bicubic.png", "ground_truth.png", "input_lr_resized.png", "srcnn.png"]
    a = torch.randint(0, 256, (8, 8, 3), dtype=torch.uint8).numpy()
    assert visualize.calculate_psnr(a, a) == 100
    assert abs(visualize.calculate_psnr(a.astype("float32") + 1, a) - 20 * math.log10(255.0)) < 1e-4


def test_lpips_alex_module_properties():
    """src/lpips_alex.py restates lpips.LPIPS(net='alex') (parity unpinned: neither the package nor its weights exist
    offline): on random weights in the package's state_dict layout it must be 0 on identical inputs, symmetric,
    positive otherwise, and MetricsCalculator must report it."""
    from src.lpips_alex import LpipsAlex
    from src.metrics import MetricsCalculator
    torch.manual_seed(3)
    ref = LpipsAlex()
    idx = {1: 0, 2: 3, 3: 6, 4: 8, 5: 10}
    sd = {}
    for k, conv in enumerate(ref.convs):
        sd["net.slice%d.%d.weight" % (k + 1, idx[k + 1])] = conv.weight.detach().clone()
        sd["net.slice%d.%d.bias" % (k + 1, idx[k + 1])] = conv.bias.detach().clone()
    for k in range(5):
        sd["lin%d.model.1.weight" % k] = ref.lins[k].detach().clone()
    m = LpipsAlex.from_state_dict(sd, "cuda:0")
    x = torch.rand(2, 3, 64, 64, device="cuda:0") * 2 - 1
    y = torch.rand(2, 3, 64, 64, device="cuda:0") * 2 - 1
    assert m(x, y).shape == (2, 1, 1, 1)
    assert float(m(x, x).abs().max()) == 0.0
    assert torch.allclose(m(x, y), m(y, x)) and float(m(x, y).min()) > 0
    got = MetricsCalculator("cuda:0", lpips_fn=m).compute((x + 1) / 2, (y + 1) / 2)
    assert abs(got["lpips"] - float(m(x, y).mean())) < 1e-6
