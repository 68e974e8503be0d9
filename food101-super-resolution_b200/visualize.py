"""Qualitative comparison + inference timing with the reference's entry point (reference visualize.py:1-126): for a
number of test images it writes ground truth / nearest / bicubic / per-model super-resolved PNGs, prints the uint8
PSNR of each (peak 255, 100 dB on identical images - a different PSNR than MetricsCalculator's, reference
visualize.py:24-29) and a per-model inference-time summary.  The models are the libsrk drop-ins (src.models.get_model);
checkpoints are the reference's own state dicts (same keys).

Offline additions (environment): SR_SYNTHETIC_DATA=<n> uses <n> synthetic images instead of Food101; SRK_DTYPE picks the
arithmetic (default bf16); NUM_EXAMPLES / OUTPUT_DIR override the constants."""
import math
import os
import random
import time
from collections import defaultdict

import numpy as np
import torch

import srk
from src.models import get_model

SCALE_FACTOR = 4
NUM_EXAMPLES = int(os.environ.get("NUM_EXAMPLES", "1000"))
OUTPUT_DIR = os.environ.get("OUTPUT_DIR", "report/images")

WEIGHTS = {
    "SRCNN":           "weights/srcnn_nlpd_best.pth",
    "RESNET":          "weights/resnet_run_best.pth",
    "AttentionSR":     "weights/attentionsr_run_best.pth",
    "AttentionSR_GAN": "weights/attentionsr_gan_best.pth",
}


def calculate_psnr(img1, img2):
    """uint8 PSNR, peak 255, 100 on identical images (reference visualize.py:24-29)."""
    a = np.array(img1).astype(np.float32)
    b = np.array(img2).astype(np.float32)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return 100
    return 20 * math.log10(255.0 / math.sqrt(mse))


def to_pil(t):
    from PIL import Image
    arr = (torch.clamp(t, 0, 1) * 255.0).to(torch.uint8).permute(1, 2, 0).cpu().numpy()   # ToPILImage: mul(255).byte()
    return Image.fromarray(arr)


def get_prediction(model_name, weight_path, lr_tensor, device):
    """-> (PIL image, seconds) or (None, None) when the checkpoint is missing (reference visualize.py:31-61)."""
    arch = "AttentionSR" if "AttentionSR" in model_name else model_name
    model = get_model(arch, scale_factor=SCALE_FACTOR, device=device)
    try:
        model.load_state_dict(torch.load(weight_path, map_location=device))
    except FileNotFoundError:
        print(f"Warning: Could not find weights for {model_name} at {weight_path}")
        return None, None
    except Exception as e:  # noqa: BLE001  (the reference reports and skips)
        print(f"Error loading {model_name}: {e}")
        return None, None
    model.eval()
    torch.cuda.synchronize()
    start = time.perf_counter()
    with torch.no_grad():
        sr = model(lr_tensor)
    torch.cuda.synchronize()
    dt = time.perf_counter() - start
    return to_pil(sr.squeeze(0)), dt


def _test_images():
    n_syn = int(os.environ.get("SR_SYNTHETIC_DATA", "0"))
    if n_syn > 0:
        from src.dataset import synthetic_image_u8
        return [synthetic_image_u8(256 + 8 * (i % 5), 320 - 12 * (i % 3), seed=900 + i).permute(2, 0, 1).float() / 255.0
                for i in range(n_syn)]
    from torchvision import datasets, transforms
    ds = datasets.Food101(root="./data", split="test", download=True, transform=transforms.ToTensor())
    return ds


def run_comparison():
    if not torch.cuda.is_available():
        raise RuntimeError("visualize.py: the SR models run on CUDA (sm_100a) only")
    device = torch.device("cuda")
    srk.set_compute_dtype(os.environ.get("SRK_DTYPE", "bf16"))
    print(f"Processing images on {device}...")
    data = _test_images()
    indices = random.sample(range(len(data)), min(NUM_EXAMPLES, len(data)))
    os.makedirs(OUTPUT_DIR, exist_ok=True)
    times = defaultdict(list)
    for i, idx in enumerate(indices):
        print(f"\n--- Processing Image {i + 1}/{len(indices)} (Index: {idx}) ---")
        save = os.path.join(OUTPUT_DIR, f"image_{idx}")
        os.makedirs(save, exist_ok=True)
        item = data[idx]
        hr = item[0] if isinstance(item, (tuple, list)) else item
        _, h, w = hr.shape
        h, w = (h // SCALE_FACTOR) * SCALE_FACTOR, (w // SCALE_FACTOR) * SCALE_FACTOR
        hr_img = to_pil(hr[:, :h, :w])
        lr_img = hr_img.resize((w // SCALE_FACTOR, h // SCALE_FACTOR), resample=3)
        lr = torch.from_numpy(np.asarray(lr_img).copy()).permute(2, 0, 1).float().div(255.0).unsqueeze(0).to(device)
        hr_img.save(os.path.join(save, "ground_truth.png"))
        lr_img.resize(hr_img.size, resample=0).save(os.path.join(save, "input_lr_resized.png"))
        bic = lr_img.resize((w, h), resample=3)
        bic.save(os.path.join(save, "bicubic.png"))
        print(f"Saved Baseline | Bicubic PSNR: {calculate_psnr(bic, hr_img):.2f} dB")
        for name, path in WEIGHTS.items():
            sr_img, dt = get_prediction(name, path, lr, device)
            if sr_img is None:
                print(f"Skipped {name} (Model failed to load)")
                continue
            times[name].append(dt)
            sr_img.save(os.path.join(save, f"{name.lower()}.png"))
            print(f"Saved {name} | PSNR: {calculate_psnr(sr_img, hr_img):.2f} dB | Inference: {dt * 1000:.2f} ms")
    print(f"\n{'=' * 50}\nINFERENCE TIME SUMMARY\n{'=' * 50}")
    for name, ts in times.items():
        ms = np.array(ts) * 1000
        print(f"{name:15} | Avg: {ms.mean():7.2f} ms | Std: {ms.std():6.2f} ms | Min: {ms.min():7.2f} ms | Max: {ms.max():7.2f} ms")
    print(f"\nDone! Check the '{OUTPUT_DIR}' folder.")
    return {k: float(np.mean(v)) for k, v in times.items()}


if __name__ == "__main__":
    run_comparison()
