"""SR trainer with the reference's entry point: `python train.py --architecture ... ` and `train(config)`
(reference train.py:21-211).  Same flags, same loop structure (Adam(betas=(0.5, 0.999)), ReduceLROnPlateau on
validation PSNR, best-PSNR checkpoint, early stopping, final test metrics averaged per batch); the numerics run
on libsrk through src/models.py, src/loss.py and src/metrics.py.

Additions that do not change the reference surface (all via environment variables):
  SR_SYNTHETIC_DATA=<n>   serve <n> synthetic Food101-shaped crops instead of downloading Food101
  SRK_DTYPE=bf16|fp32     arithmetic of the conv stacks (default bf16 on the tensor cores)
  WANDB_MODE=disabled     wandb is optional; without the package a no-op logger is used
Under torchrun (WORLD_SIZE > 1) every rank trains on its shard of each batch and gradients are averaged with NCCL
(srk/dp.py).  The GAN branch (loss_function=gan) is outside the accelerated path and not provided here."""
import argparse
import os

import torch
import torch.distributed as dist
import torch.optim as optim
from torch.utils.data import DataLoader, random_split

import srk
from srk import dp
from src.dataset import FoodSRDataset
from src.loss import get_loss_function
from src.metrics import MetricsCalculator
from src.models import get_model
from src.utils import get_gradient_norm, get_layer_grad_ratio, get_update_ratio, save_checkpoint

try:
    import wandb
except ImportError:  # pragma: no cover
    wandb = None


class _Run:
    def __init__(self, config):
        self.config = argparse.Namespace(**dict(config))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _init_run(config):
    if wandb is not None and os.environ.get("WANDB_MODE", "") != "disabled":
        return wandb.init(config=config)
    return _Run(config)


def _log(data):
    if wandb is not None and wandb.run is not None:
        wandb.log(data)


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
    # this loop only ever calls loss.backward(): weight gradients may run on the side stream (srk/ops.py)
    srk.set_overlap_wgrad(os.environ.get("SRK_OVERLAP_WGRAD", "1") != "0")
    with _init_run(config) as run:
        cfg = run.config
        if cfg.loss_function == "gan":
            raise NotImplementedError("loss_function=gan is outside the accelerated path (SURVEY 8f-3)")
        print(f"Running on {device} | Arch: {cfg.architecture} | ranks: {world}")
        full_train_ds = FoodSRDataset(split="train", crop_size=200, scale_factor=4)
        if cfg.subset < 1.0:
            total = len(full_train_ds)
            keep = int(total * cfg.subset)
            full_train_ds, _ = random_split(full_train_ds, [keep, total - keep])
        train_len = int(0.9 * len(full_train_ds))
        train_ds, val_ds = random_split(full_train_ds, [train_len, len(full_train_ds) - train_len])
        test_ds = FoodSRDataset(split="test", crop_size=200, scale_factor=4)
        if cfg.subset < 1.0:
            keep = int(len(test_ds) * cfg.subset)
            test_ds, _ = random_split(test_ds, [keep, len(test_ds) - keep])
        mk = lambda ds, sh: DataLoader(ds, batch_size=cfg.batch_size, shuffle=sh, num_workers=0, pin_memory=True)
        train_loader, val_loader, test_loader = mk(train_ds, True), mk(val_ds, False), mk(test_ds, False)

        model = get_model(cfg.architecture, scale_factor=4, device=device)
        if cfg.pretrained_weights:
            model.load_state_dict(torch.load(cfg.pretrained_weights, map_location=device), strict=False)
        dp.broadcast_parameters(model)
        averager = dp.GradAverager(model.parameters()) if world > 1 else None
        optimizer = optim.Adam(model.parameters(), lr=cfg.lr, betas=(0.5, 0.999))
        scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="max", factor=0.5, patience=2)
        criterion = get_loss_function(cfg.loss_function, device)
        metrics_calc = MetricsCalculator(device)
        best_psnr, patience_counter = 0.0, 0

        def shard(t):
            b, e = dp.shard_range(t.shape[0], rank, world)
            return t[b:e].to(device, non_blocking=True)

        for epoch in range(cfg.epochs):
            model.train()
            for batch_idx, (lr_imgs, hr_imgs) in enumerate(train_loader):
                lr_imgs, hr_imgs = shard(lr_imgs), shard(hr_imgs)
                if lr_imgs.shape[0] == 0:
                    continue
                optimizer.zero_grad()
                loss = criterion(model(lr_imgs), hr_imgs)
                loss.backward()
                if averager is not None:
                    averager.average()
                optimizer.step()
                if batch_idx % 100 == 0 and rank == 0:
                    cur_lr = optimizer.param_groups[0]["lr"]
                    _log({"train_loss": loss.item(), "dynamics/grad_norm": get_gradient_norm(model),
                          "dynamics/layer_ratio": get_layer_grad_ratio(model),
                          "dynamics/update_ratio": get_update_ratio(model, cur_lr)})
            model.eval()
            avg_psnr, avg_val_loss = 0.0, 0.0
            with torch.no_grad():
                for lr_b, hr_b in val_loader:
                    lr_b, hr_b = lr_b.to(device), hr_b.to(device)
                    sr = model(lr_b)
                    avg_psnr += metrics_calc.compute(sr, hr_b)["psnr"]
                    avg_val_loss += criterion(sr, hr_b).item()
            avg_psnr /= max(len(val_loader), 1)
            avg_val_loss /= max(len(val_loader), 1)
            scheduler.step(avg_psnr)
            if rank == 0:
                print(f"   -> Val PSNR: {avg_psnr:.2f} | Val Loss: {avg_val_loss:.4f} | LR: {optimizer.param_groups[0]['lr']}")
                _log({"epoch": epoch, "val_psnr": avg_psnr, "val_loss": avg_val_loss, "lr": optimizer.param_groups[0]["lr"]})
            if avg_psnr > best_psnr:
                best_psnr, patience_counter = avg_psnr, 0
                if rank == 0:
                    save_checkpoint(model, epoch, f"weights/{cfg.save_name}_best.pth")
            else:
                patience_counter += 1
            if patience_counter >= cfg.patience:
                print("Early stopping triggered")
                break

        if world > 1:
            dist.barrier()
        best = f"weights/{cfg.save_name}_best.pth"
        if os.path.exists(best):
            model.load_state_dict(torch.load(best, map_location=device))
        model.eval()
        test_metrics = {"psnr": 0.0, "ssim": 0.0, "lpips": 0.0, "nlpd": 0.0}
        with torch.no_grad():
            for lr_b, hr_b in test_loader:
                res = metrics_calc.compute(model(lr_b.to(device)), hr_b.to(device))
                for k in test_metrics:
                    test_metrics[k] += res[k]
        for k in test_metrics:
            test_metrics[k] /= max(len(test_loader), 1)
        if rank == 0:
            print(f"Final Test Results: {test_metrics}")
            _log({"test_" + k: v for k, v in test_metrics.items()})
        return test_metrics


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
