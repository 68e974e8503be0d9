"""The sample pipeline of the reference's FoodSRDataset (reference src/dataset.py:14-41) with the pixel work on the
GPU: the host ships decoded uint8 images and draws the random crop offsets / flip flags (with torch's generator, in
the order torchvision's RandomCrop and RandomHorizontalFlip draw them); one libsrk kernel (srk_sr_make_batch) crops,
flips, converts to float / 255 (ToTensor) and produces the antialiased-bicubic LR image
(transforms.Resize((lr, lr), BICUBIC) on a tensor = F.interpolate(mode="bicubic", antialias=True))."""
import torch

from . import _lib as L
from . import ops


def draw_crop_params(sizes, crop, train, generator=None):
    """-> (offsets int32 [N, 2] (top, left), flips uint8 [N]) for images of the given (h, w) sizes.
    train: RandomCrop.get_params (two torch.randint draws: top, then left) followed by RandomHorizontalFlip
    (torch.rand(1) < 0.5), per sample, in dataset order (reference dataset.py:17-21); else CenterCrop
    (torchvision: int(round((h - crop) / 2.0)), no flip; reference dataset.py:23-26)."""
    n = len(sizes)
    offs = torch.empty((n, 2), dtype=torch.int32)
    flips = torch.zeros((n,), dtype=torch.uint8)
    for k, (h, w) in enumerate(sizes):
        h, w = int(h), int(w)
        if h < crop or w < crop:
            raise ValueError("image %dx%d is smaller than the crop (%d): resize it first (dataset.py:31-32)" % (h, w, crop))
        if train:
            if h == crop and w == crop:
                top, left = 0, 0     # RandomCrop.get_params returns (0, 0) without drawing
            else:
                top = int(torch.randint(0, h - crop + 1, size=(1,), generator=generator).item())
                left = int(torch.randint(0, w - crop + 1, size=(1,), generator=generator).item())
            flips[k] = 1 if float(torch.rand(1, generator=generator)) < 0.5 else 0
        else:
            top, left = int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))
        offs[k, 0], offs[k, 1] = top, left
    return offs, flips


def make_batch(src_u8, offsets, flips, crop, scale):
    """src_u8: CUDA uint8 [N, Hs, Ws, 3] (what np.asarray(PIL image) gives, padded to a common size) or [N, 3, Hs, Ws];
    offsets / flips: from draw_crop_params (host or device).  -> (lr [N, 3, crop / scale, crop / scale],
    hr [N, 3, crop, crop]) float32 on the device."""
    ops.require_cuda(src_u8, "make_batch")
    if src_u8.dtype != torch.uint8 or src_u8.dim() != 4:
        raise ValueError("make_batch: expected a uint8 [N, H, W, 3] or [N, 3, H, W] tensor")
    src_u8 = src_u8.contiguous()
    hwc = src_u8.shape[3] == 3
    if not hwc and src_u8.shape[1] != 3:
        raise ValueError("make_batch: expected 3 colour channels")
    n = src_u8.shape[0]
    hs, ws = (src_u8.shape[1], src_u8.shape[2]) if hwc else (src_u8.shape[2], src_u8.shape[3])
    dev = src_u8.device
    offs = offsets.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
    flp = flips.to(device=dev, dtype=torch.uint8, non_blocking=True).contiguous()
    hr = torch.empty((n, 3, crop, crop), dtype=torch.float32, device=dev)
    lr = torch.empty((n, 3, crop // scale, crop // scale), dtype=torch.float32, device=dev)
    L.call("srk_sr_make_batch", src_u8.data_ptr(), 1 if hwc else 0, n, hs, ws, offs.data_ptr(), flp.data_ptr(), crop,
           scale, hr.data_ptr(), lr.data_ptr(), ops.stream_ptr())
    return lr, hr


def make_batch_into(src_u8, offsets, flips, crop, scale, lr_out, hr_out):
    """make_batch writing into caller-owned float32 tensors (e.g. the static inputs of a captured training step):
    no allocation, no extra device-to-device copy.  offsets int32 [N, 2] / flips uint8 [N] must be on the device."""
    ops.require_cuda(src_u8, "make_batch")
    assert src_u8.dtype == torch.uint8 and src_u8.dim() == 4 and src_u8.is_contiguous()
    hwc = src_u8.shape[3] == 3
    n = src_u8.shape[0]
    hs, ws = (src_u8.shape[1], src_u8.shape[2]) if hwc else (src_u8.shape[2], src_u8.shape[3])
    assert tuple(hr_out.shape) == (n, 3, crop, crop) and tuple(lr_out.shape) == (n, 3, crop // scale, crop // scale)
    assert hr_out.dtype == torch.float32 and lr_out.dtype == torch.float32 and hr_out.is_contiguous() and lr_out.is_contiguous()
    assert offsets.is_cuda and flips.is_cuda and offsets.dtype == torch.int32 and flips.dtype == torch.uint8
    L.call("srk_sr_make_batch", src_u8.data_ptr(), 1 if hwc else 0, n, hs, ws, offsets.data_ptr(), flips.data_ptr(), crop,
           scale, hr_out.data_ptr(), lr_out.data_ptr(), ops.stream_ptr())
    return lr_out, hr_out


def collate_raw(samples):
    """DataLoader collate_fn for raw samples (uint8 [h, w, 3] tensors): pads to the largest image of the batch
    -> (uint8 [N, Hmax, Wmax, 3], sizes [(h, w), ...])."""
    sizes = [(int(s.shape[0]), int(s.shape[1])) for s in samples]
    hm, wm = max(h for h, _ in sizes), max(w for _, w in sizes)
    out = torch.zeros((len(samples), hm, wm, 3), dtype=torch.uint8)
    for k, s in enumerate(samples):
        out[k, : s.shape[0], : s.shape[1]] = s
    return out, sizes


class GpuBatches:
    """Iterates a DataLoader of raw samples (collate_raw) and yields (lr, hr) device tensors: pinned H2D copy of the
    uint8 batch on a copy stream, one batch ahead of the consumer, then the crop / flip / downsample kernel."""

    def __init__(self, loader, crop, scale, train, device, generator=None):
        self.loader, self.crop, self.scale, self.train, self.device, self.generator = loader, crop, scale, train, device, generator
        self.copy_stream = torch.cuda.Stream(device=device)

    def __len__(self):
        return len(self.loader)

    def _stage(self, item):
        src, sizes = item
        offs, flips = draw_crop_params(sizes, self.crop, self.train, self.generator)
        with torch.cuda.stream(self.copy_stream):
            dsrc = src.pin_memory().to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return dsrc, offs, flips, ev

    def __iter__(self):
        it = iter(self.loader)
        nxt = None
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dsrc, offs, flips, ev = nxt
            try:
                nxt = self._stage(next(it))      # the next batch's H2D overlaps this batch's step
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            dsrc.record_stream(torch.cuda.current_stream(self.device))
            yield make_batch(dsrc, offs, flips, self.crop, self.scale)
