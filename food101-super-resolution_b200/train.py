"""SR trainer with the reference's entry point: `python train.py --architecture ... ` and `train(config)`
(reference train.py:21-211).  Same flags, same loop structure (Adam(betas=(0.5, 0.999)), ReduceLROnPlateau on
validation PSNR, best-PSNR checkpoint, early stopping, final test metrics averaged per batch); the numerics run
on libsrk through src/models.py, src/loss.py and src/metrics.py, and the step itself is srk.trainer.GraphStep -
the captured forward + loss + backward + all-reduce + Adam + weight re-pack that bench.py measures.

Additions that do not change the reference surface (all via environment variables):
  SR_SYNTHETIC_DATA=<n>   serve <n> synthetic Food101-shaped crops instead of downloading Food101
  SR_CROP=<pixels>        HR crop size (default 200, as reference train.py:27)
  SRK_DTYPE=bf16|fp32     arithmetic of the conv stacks (default bf16 on the tensor cores)
  SRK_GRAPH=0             launch every kernel from Python instead of replaying the captured step
  SRK_GPU_PIPELINE=0      make the (lr, hr) pairs on the host like the reference; default 1: the loaders ship decoded
                          uint8 images and srk.data crops / flips / converts / down-samples on the GPU
  WANDB_MODE=disabled     wandb is optional; without the package a no-op logger is used

Data parallelism (torchrun, WORLD_SIZE > 1; the reference has none): every rank draws the SAME seeded split and the
same shuffled batches and trains on its shard of each batch; gradients are averaged with NCCL (srk/dp.py) with each
rank's loss weighted by its share of the batch, so uneven shards still give the full-batch gradient.  Everything that
steers control flow - validation PSNR / loss, hence the LR schedule, the best checkpoint and the early-stop decision -
is computed from all-reduced sums (srk/evaluate.py) and is therefore identical on every rank: no rank can leave the
loop while another waits in a collective.  A batch with fewer samples than ranks is skipped on ALL ranks.

loss_function=gan (reference train.py:58-65,86-114): the generator, the content / perceptual / TV terms and the
optimizer steps run on libsrk; the spectral-norm discriminator (models.py Discriminator) runs on torch's own kernels
(its stride-2 convs are outside the accelerated path)."""
import argparse
import os

import torch
import torch.distributed as dist
import torch.optim as optim
from torch.utils.data import DataLoader, random_split

import srk
from srk import dp
from srk import evaluate as ev
from srk.trainer import GraphStep
from src.dataset import FoodSRDataset
from src.loss import TVLoss, get_loss_function
from src.metrics import MetricsCalculator
from src.models import Discriminator, get_model
from src.utils import get_gradient_norm, get_layer_grad_ratio, get_update_ratio, save_checkpoint

try:
    import wandb
except ImportError:  # pragma: no cover
    wandb = None

SPLIT_SEED = 20240229     # shared by all ranks: the splits and the shuffle order must agree


class _Run:
    def __init__(self, config):
        self.config = argparse.Namespace(**dict(config))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _init_run(config, rank):
    if rank == 0 and wandb is not None and os.environ.get("WANDB_MODE", "") != "disabled":
        return wandb.init(config=config)
    return _Run(config)


def _log(data):
    if wandb is not None and wandb.run is not None:
        wandb.log(data)


def add_noise(img, sigma=0.15):
    if sigma <= 0:
        return img
    return img + torch.randn_like(img) * sigma


def fit(cfg, *, train_loader, step_fn, eval_fn, scheduler, get_lr, save_best, rank=0, world=1, log=_log,
        shard=None, on_log_step=None):
    """The epoch loop of reference train.py:68-180 with the compute behind callables, so that its control flow can be
    exercised without a GPU (tests/test_dp_gloo.py runs it on two gloo ranks).

      step_fn(lr, hr, weight) -> loss (tensor or float)   one optimizer step on this rank's shard; `weight` scales the
                                                          loss so that averaging over ranks gives the full-batch mean
      eval_fn() -> {"psnr": .., "loss": ..}               validation pass; MUST return the same numbers on every rank
    Returns (best_psnr, epochs_run)."""
    best_psnr, patience_counter, epochs_run = 0.0, 0, 0
    for epoch in range(cfg.epochs):
        epochs_run += 1
        for batch_idx, (lr_imgs, hr_imgs) in enumerate(train_loader):
            n_global = lr_imgs.shape[0]
            if n_global < world:
                continue      # same decision on every rank (all ranks see the same batch): nobody waits in a collective
            if shard is not None:
                b, e = dp.shard_range(n_global, rank, world)
                lr_imgs, hr_imgs = shard(lr_imgs[b:e]), shard(hr_imgs[b:e])
                weight = (e - b) * world / float(n_global)
            else:
                weight = 1.0
            loss = step_fn(lr_imgs, hr_imgs, weight)
            if batch_idx % 100 == 0 and on_log_step is not None:
                on_log_step(loss)
        res = eval_fn()
        avg_psnr, avg_val_loss = res["psnr"], res["loss"]
        scheduler.step(avg_psnr)
        if rank == 0:
            print(f"   -> Val PSNR: {avg_psnr:.2f} | Val Loss: {avg_val_loss:.4f} | LR: {get_lr()}")
            log({"epoch": epoch, "val_psnr": avg_psnr, "val_loss": avg_val_loss, "lr": get_lr()})
        if avg_psnr > best_psnr:
            best_psnr, patience_counter = avg_psnr, 0
            save_best(epoch)
        else:
            patience_counter += 1
        if patience_counter >= cfg.patience:
            if rank == 0:
                print("Early stopping triggered")
            break
    return best_psnr, epochs_run


def train(config=None):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("train.py: the SR hot path runs on CUDA (sm_100a) only; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    srk.set_compute_dtype(os.environ.get("SRK_DTYPE", "bf16"))
    crop = int(os.environ.get("SR_CROP", "200"))
    with _init_run(config, rank) as run:
        cfg = run.config
        print(f"Running on {device} | Arch: {cfg.architecture} | ranks: {world}")
        split_gen = torch.Generator().manual_seed(SPLIT_SEED)
        gpu_pipe = os.environ.get("SRK_GPU_PIPELINE", "1") != "0"
        full_train_ds = FoodSRDataset(split="train", crop_size=crop, scale_factor=4, raw=gpu_pipe)
        if cfg.subset < 1.0:
            total = len(full_train_ds)
            keep = int(total * cfg.subset)
            full_train_ds, _ = random_split(full_train_ds, [keep, total - keep], generator=split_gen)
        train_len = int(0.9 * len(full_train_ds))
        train_ds, val_ds = random_split(full_train_ds, [train_len, len(full_train_ds) - train_len], generator=split_gen)
        test_ds = FoodSRDataset(split="test", crop_size=crop, scale_factor=4, raw=gpu_pipe)
        if cfg.subset < 1.0:
            keep = int(len(test_ds) * cfg.subset)
            test_ds, _ = random_split(test_ds, [keep, len(test_ds) - keep], generator=split_gen)
        if rank == 0:
            print(f"Dataset: Train={len(train_ds)} | Val={len(val_ds)} | Test={len(test_ds)}")
        mk = lambda ds, sh: DataLoader(ds, batch_size=cfg.batch_size, shuffle=sh, num_workers=0, pin_memory=not gpu_pipe,
                                       collate_fn=srk.data.collate_raw if gpu_pipe else None,
                                       generator=torch.Generator().manual_seed(SPLIT_SEED + 1) if sh else None)
        train_loader, val_loader, test_loader = mk(train_ds, True), mk(val_ds, False), mk(test_ds, False)
        if gpu_pipe:
            # crop offsets / flips come from one seeded generator shared by all ranks (every rank sees the same batch);
            # NOTE the validation split of the reference is cut from the TRAIN dataset, so it keeps random crops
            aug_gen = torch.Generator().manual_seed(SPLIT_SEED + 2)
            train_loader = srk.data.GpuBatches(train_loader, crop, 4, True, device, aug_gen)
            val_loader = srk.data.GpuBatches(val_loader, crop, 4, True, device, aug_gen)
            test_loader = srk.data.GpuBatches(test_loader, crop, 4, False, device)

        model = get_model(cfg.architecture, scale_factor=4, device=device)
        if cfg.pretrained_weights:
            model.load_state_dict(torch.load(cfg.pretrained_weights, map_location=device), strict=False)
        dp.broadcast_parameters(model)
        averager = dp.GradAverager(model.parameters()) if world > 1 else None
        is_gan = cfg.loss_function == "gan"
        use_graph = os.environ.get("SRK_GRAPH", "1") != "0"
        metrics_calc = MetricsCalculator(device)

        if is_gan:
            criterion = get_loss_function("mae", device)             # content term / validation loss (train.py:62,156)
            gan = _GanStep(model, device, cfg.lr, averager)
            optimizer = gan.optimizer
            step_fn = gan.step
        else:
            criterion = get_loss_function(cfg.loss_function, device)
            # one rank: the criterion itself (no extra arithmetic in the step); several ranks: scaled by the shard's weight
            weighted = _Weighted(criterion) if world > 1 else None
            trainer = GraphStep(model, weighted or criterion, lr=cfg.lr, betas=(0.5, 0.999), averager=averager,
                                use_graph=use_graph)
            optimizer = trainer.optimizer

            def step_fn(lr_imgs, hr_imgs, weight):
                if weighted is not None:
                    weighted.set(weight)
                return trainer(lr_imgs, hr_imgs)
        scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="max", factor=0.5, patience=2)
        get_lr = lambda: optimizer.param_groups[0]["lr"]
        best_path = f"weights/{cfg.save_name}_best.pth"

        def eval_fn():
            dp.broadcast_buffers(model)        # BatchNorm running statistics are per rank during training (DDP semantics)
            return ev.evaluate(model, val_loader, device, rank, world, metrics_calc.compute, criterion=criterion)

        def save_best(epoch):
            if rank == 0:
                save_checkpoint(model, epoch, best_path)

        def on_log_step(loss):
            if rank != 0:
                return
            data = {"train_loss": float(loss), "dynamics/grad_norm": get_gradient_norm(model),
                    "dynamics/layer_ratio": get_layer_grad_ratio(model),
                    "dynamics/update_ratio": get_update_ratio(model, get_lr())}
            if is_gan:
                data.update({"train_loss_D": gan.loss_d, "gan_dynamics/prob_real": gan.prob_real,
                             "gan_dynamics/prob_fake": gan.prob_fake})
            _log(data)

        to_dev = lambda t: t.to(device, non_blocking=True)
        fit(cfg, train_loader=train_loader, step_fn=step_fn, eval_fn=eval_fn, scheduler=scheduler, get_lr=get_lr,
            save_best=save_best, rank=rank, world=world, shard=to_dev, on_log_step=on_log_step)

        if world > 1:
            dist.barrier()
            if not is_gan:
                trainer._graphs.clear()     # captured NCCL kernels would keep the communicator busy at teardown
        if os.path.exists(best_path):
            model.load_state_dict(torch.load(best_path, map_location=device))
        test_metrics = ev.evaluate(model, test_loader, device, rank, world, metrics_calc.compute)
        test_metrics = {k: test_metrics[k] for k in ("psnr", "ssim", "lpips", "nlpd")}
        if rank == 0:
            print(f"Final Test Results: {test_metrics}")
            _log({"test_" + k: v for k, v in test_metrics.items()})
        return test_metrics


class _Weighted(torch.nn.Module):
    """criterion scaled by a device-resident factor (a rank's share of the global batch times the world size): the
    factor can change from batch to batch while the step replays from its CUDA graph."""

    def __init__(self, criterion):
        super().__init__()
        self.criterion = criterion
        self.weight = None
        self._host = None

    def set(self, w):
        if self.weight is not None and w != self._host:
            self.weight.fill_(float(w))
        self._host = w

    def forward(self, sr, hr):
        loss = self.criterion(sr, hr)
        if self.weight is None:
            self.weight = torch.full((), float(self._host if self._host is not None else 1.0), device=sr.device)
        return loss * self.weight


class _GanStep:
    """Generator / discriminator updates of reference train.py:86-114.  Generator forward / backward, MAE, VGG
    perceptual and the Adam steps are libsrk; the discriminator and its BCE terms are torch modules."""

    def __init__(self, model, device, lr, averager):
        self.model, self.device, self.averager = model, device, averager
        self.discriminator = Discriminator().to(device)
        dp.broadcast_parameters(self.discriminator)
        self.d_averager = dp.GradAverager(self.discriminator.parameters()) if averager is not None else None
        self.optimizer = srk.optim.Adam(model.parameters(), lr=lr, betas=(0.5, 0.999))
        self.optimizer_d = optim.Adam(self.discriminator.parameters(), lr=lr * 0.1, betas=(0.5, 0.999))
        self.bce = torch.nn.BCEWithLogitsLoss()
        self.content = get_loss_function("mae", device)
        self.percep = get_loss_function("perceptual", device)
        self.tv = TVLoss(tv_loss_weight=1).to(device)
        self.batch_idx = 0
        self.loss_d, self.prob_real, self.prob_fake = 0.0, 0.5, 0.5
        self.clip = torch.zeros((1,), dtype=torch.float32, device=device)
        self.optimizer.grad_scale_dev = self.clip

    def step(self, lr_imgs, hr_imgs, weight):
        model, disc = self.model, self.discriminator
        model.train()
        disc.train()
        if self.batch_idx % 5 == 0:
            self.optimizer_d.zero_grad()
            with torch.no_grad():
                fake = model(lr_imgs)
            real_logits = disc(add_noise(hr_imgs, 0.2))
            fake_logits = disc(add_noise(fake, 0.2))
            self.prob_real = torch.sigmoid(real_logits).mean().item()
            self.prob_fake = torch.sigmoid(fake_logits).mean().item()
            d_real = self.bce(real_logits - fake_logits.mean(), torch.full_like(real_logits, 0.9))
            d_fake = self.bce(fake_logits - real_logits.mean(), torch.full_like(fake_logits, 0.1))
            loss_d = (d_real + d_fake) / 2 * weight
            loss_d.backward()
            if self.d_averager is not None:
                self.d_averager.average()
            self.optimizer_d.step()
            self.loss_d = loss_d.item()
        self.batch_idx += 1
        self.optimizer.zero_grad()
        fake = model(lr_imgs)
        fake_logits = disc(fake)
        real_logits = disc(hr_imgs).detach()
        loss_adv = self.bce(fake_logits - real_logits.mean(), torch.ones_like(fake_logits))
        loss = (1e-2 * self.content(fake, hr_imgs)) + (1.0 * self.percep(fake, hr_imgs)) + (1e-5 * loss_adv) \
            + (2e-5 * self.tv(fake))
        (loss * weight).backward()
        if self.averager is not None:
            self.averager.average()
        # clip_grad_norm_(max_norm=1.0) (train.py:113) without a host round trip: the coefficient stays on the device and
        # rides into the Adam kernel as its gradient scale
        grads = [p.grad for p in model.parameters() if p.grad is not None]
        total = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads)))
        self.clip.copy_(torch.clamp(1.0 / (total + 1e-6), max=1.0).reshape(1))
        self.optimizer.step()
        srk.ops.repack_all()
        return loss.detach()


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--architecture", type=str, default="SRCNN")
    parser.add_argument("--batch_size", type=int, default=16)
    parser.add_argument("--lr", type=float, default=0.0004)
    parser.add_argument("--epochs", type=int, default=10)
    parser.add_argument("--loss_function", type=str, default="nlpd")
    parser.add_argument("--subset", type=float, default=1.0)
    parser.add_argument("--pretrained_weights", type=str, default="")
    parser.add_argument("--patience", type=int, default=5)
    parser.add_argument("--save_name", type=str, default="model_best")
    train(config=vars(parser.parse_args()))
