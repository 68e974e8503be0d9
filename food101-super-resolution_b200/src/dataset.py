"""Data for the SR trainer.  The reference's FoodSRDataset (reference dataset.py:6-44) downloads
Food101 through torchvision and produces (lr [3, crop/s, crop/s], hr [3, crop, crop]) float32 pairs in
[0, 1] with a bicubic antialiased downsample.  Offline (no network) the same interface is served from
seeded synthetic Food101-shaped crops: 8-bit uniform noise low-passed by the 5x5 sigma-1 Gaussian,
LR = antialiased bicubic downsample of HR, unclamped as in reference dataset.py:38-39."""
import os

import torch
import torch.nn.functional as F


def _lowpass_kernel():
    ax = torch.arange(5, dtype=torch.float32) - 2.0
    g = torch.exp(-(ax[None, :] ** 2 + ax[:, None] ** 2) / 2.0)
    return (g / g.sum()).view(1, 1, 5, 5).repeat(3, 1, 1, 1)


def synthetic_pair(n, h_lr, w_lr, scale, seed=1234):
    """-> (lr [n,3,h_lr,w_lr], hr [n,3,h_lr*scale,w_lr*scale]) float32 host tensors."""
    g = torch.Generator().manual_seed(seed)
    hr = torch.randint(0, 256, (n, 3, h_lr * scale, w_lr * scale), generator=g).float() / 255.0
    hr = F.conv2d(hr, _lowpass_kernel(), padding=2, groups=3)
    lr = F.interpolate(hr, size=(h_lr, w_lr), mode="bicubic", align_corners=False, antialias=True)
    return lr.contiguous(), hr.contiguous()


def synthetic_image_u8(h, w, seed):
    """A decoded-photo stand-in: uint8 [h, w, 3] (what np.asarray(PIL image) gives), low-passed 8-bit noise."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (1, 3, h, w), generator=g).float()
    img = F.conv2d(img, _lowpass_kernel(), padding=2, groups=3)
    return img[0].round().clamp(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous()


class SyntheticSRDataset(torch.utils.data.Dataset):
    """raw=False: (lr, hr) float pairs made on the host; raw=True: uint8 [h, w, 3] source images a little larger than
    the crop - crop / flip / downsample then happen on the GPU (srk.data)."""

    def __init__(self, length=1024, crop_size=200, scale_factor=4, seed=1234, raw=False):
        self.length, self.crop, self.scale, self.seed, self.raw = length, crop_size, scale_factor, seed, raw

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        if self.raw:
            return synthetic_image_u8(self.crop + 24 + (idx % 3) * 8, self.crop + 40 - (idx % 2) * 16, self.seed + idx)
        lr, hr = synthetic_pair(1, self.crop // self.scale, self.crop // self.scale, self.scale, self.seed + idx)
        return lr[0], hr[0]


class FoodSRDataset(torch.utils.data.Dataset):
    """Same constructor as the reference (split, crop_size, scale_factor).  Uses torchvision's Food101
    when the archive is already under ./data (download needs a network); with SR_SYNTHETIC_DATA=<n> it
    serves <n> synthetic crops instead."""

    def __init__(self, split="train", crop_size=200, scale_factor=4, raw=False):
        """raw=True (an addition; the reference has no such flag): __getitem__ returns the decoded image as a uint8
        [h, w, 3] tensor (up-scaled first when smaller than the crop, reference dataset.py:31-32) and the crop / flip /
        ToTensor / bicubic down-sampling run on the GPU (srk.data.GpuBatches)."""
        self.crop_size, self.scale_factor, self.raw, self.split = crop_size, scale_factor, raw, split
        n_syn = int(os.environ.get("SR_SYNTHETIC_DATA", "0"))
        if n_syn > 0:
            self.inner = SyntheticSRDataset(n_syn, crop_size, scale_factor, seed=1234 if split == "train" else 4321, raw=raw)
            self.food = None
            return
        from torchvision import transforms
        from torchvision.datasets import Food101
        assert crop_size % scale_factor == 0, "crop size must be divisible by the scale factor"
        self.food = Food101(root="./data", split=split, download=True)
        self.inner = None
        crop = [transforms.RandomCrop(crop_size), transforms.RandomHorizontalFlip()] if split == "train" \
            else [transforms.CenterCrop(crop_size)]
        self.to_hr = transforms.Compose(crop + [transforms.ToTensor()])
        self.grow = transforms.Resize(crop_size, interpolation=transforms.InterpolationMode.BICUBIC)
        lr_size = crop_size // scale_factor
        self.down = transforms.Resize((lr_size, lr_size), interpolation=transforms.InterpolationMode.BICUBIC)

    def __len__(self):
        return len(self.inner) if self.inner is not None else len(self.food)

    def __getitem__(self, idx):
        if self.inner is not None:
            return self.inner[idx]
        img, _ = self.food[idx]
        if min(img.size) < self.crop_size:
            img = self.grow(img)
        if self.raw:
            import numpy as np
            return torch.from_numpy(np.asarray(img.convert("RGB")).copy())
        hr = self.to_hr(img)
        return self.down(hr), hr
