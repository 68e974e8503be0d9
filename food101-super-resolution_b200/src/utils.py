"""Training-dynamics helpers and checkpointing with the reference's names (reference utils.py:5-46).  They only
read `.parameters()` / `.grad`, which the libsrk modules expose as ordinary fp32 tensors in state_dict layout."""
import os

import torch


def _sq_norm_sum(tensors):
    return sum(float(t.detach().norm(2)) ** 2 for t in tensors)


def get_gradient_norm(model):
    return _sq_norm_sum(p.grad for p in model.parameters() if p.grad is not None) ** 0.5


def get_weight_norm(model):
    return _sq_norm_sum(p for p in model.parameters()) ** 0.5


def get_layer_grad_ratio(model):
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads:
        return 0.0
    return float(grads[0].norm(2)) / (float(grads[-1].norm(2)) + 1e-8)


def get_update_ratio(model, lr):
    ps = [p for p in model.parameters() if p.grad is not None]
    w = _sq_norm_sum(ps)
    if w == 0:
        return 0.0
    u = sum((float(p.grad.detach().norm(2)) * lr) ** 2 for p in ps)
    return (u ** 0.5) / (w ** 0.5)


def save_checkpoint(model, epoch, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    target = model.module if isinstance(model, torch.nn.DataParallel) else model
    torch.save(target.state_dict(), path)
    try:
        import wandb
        if wandb.run is not None:
            wandb.save(path)
    except ImportError:
        pass
