"""Adam on libsrk (reference train.py:55 uses optim.Adam(lr, betas=(0.5, 0.999)); torch.optim.Adam keeps working
on the drop-in modules - this class is the kernel-backed equivalent the trainer uses).

torch.optim.Adam semantics (no weight decay, no amsgrad): exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2*v +
(1-b2) g^2; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).  One multi-tensor launch per 48 parameters.  The step
counter AND the learning rate live on the device, so a step captured in a CUDA graph follows a scheduler
(ReduceLROnPlateau writes param_groups[0]["lr"], reference train.py:56,164) without re-capture.

A torch.optim.Optimizer subclass: lr schedulers, param_groups, state_dict() / load_state_dict() (state layout of
torch.optim.Adam: "step", "exp_avg", "exp_avg_sq" per parameter) work as with the stock optimizer."""
import ctypes

import torch

from . import _lib as L
from . import ops


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps)))
        if len(self.param_groups) != 1:
            raise ValueError("srk.optim.Adam takes a single parameter group")
        group = self.param_groups[0]
        group["params"] = [p for p in group["params"] if p.requires_grad]
        if not group["params"]:
            raise ValueError("optimizer got an empty parameter list")
        p0 = group["params"][0]
        ops.require_cuda(p0, "srk.optim.Adam")
        for p in group["params"]:
            self.state[p] = {"exp_avg": torch.zeros_like(p, memory_format=torch.contiguous_format),
                             "exp_avg_sq": torch.zeros_like(p, memory_format=torch.contiguous_format)}
        self.step_count = torch.zeros((1,), dtype=torch.int64, device=p0.device)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=p0.device)
        self._lr_host = float(lr)
        self.grad_scale_dev = None      # optional device float[1] multiplied onto every gradient (gradient clipping)

    # kept for callers of the round-1 interface
    @property
    def params(self):
        return self.param_groups[0]["params"]

    def sync_lr(self):
        """Host-side learning rate (what a scheduler edits) -> device copy.  Called by step(); a trainer that replays a
        captured step calls it before the replay (it is a no-op while the rate is unchanged)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        if closure is not None:
            raise NotImplementedError("srk.optim.Adam: closures are not supported")
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        group = self.param_groups[0]
        self.step_count += 1
        live = []
        for p in group["params"]:
            if p.grad is None:
                continue
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            st = self.state[p]
            live.append((p, g, st["exp_avg"], st["exp_avg_sq"]))
        if not live:
            return
        n = len(live)
        arr = lambda vals: (ctypes.c_void_p * n)(*vals)
        self._keep = [g for _, g, _, _ in live]  # contiguous copies must outlive the launch
        L.call("srk_adam_multi", n, arr([p.data_ptr() for p, _, _, _ in live]),
               arr([g.data_ptr() for _, g, _, _ in live]), arr([m.data_ptr() for _, _, m, _ in live]),
               arr([v.data_ptr() for _, _, _, v in live]), (ctypes.c_int64 * n)(*[p.numel() for p, _, _, _ in live]),
               float(group["lr"]), group["betas"][0], group["betas"][1], group["eps"], self.step_count.data_ptr(),
               float(grad_scale), self.lr_dev.data_ptr(),
               self.grad_scale_dev.data_ptr() if self.grad_scale_dev is not None else None, ops.stream_ptr())
        # the kernels wrote through raw pointers: invalidate the packed-weight cache
        ops.bump_weights_epoch()

    def state_dict(self):
        sd = super().state_dict()
        for st in sd["state"].values():
            st["step"] = self.step_count.detach().clone().float().reshape(())
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [st.pop("step") for st in self.state.values() if "step" in st]
        if steps:
            self.step_count.fill_(int(float(steps[0])))
        self._lr_host = None
        self.sync_lr()
