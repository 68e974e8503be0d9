"""Adam on libsrk (train.py:55 uses optim.Adam(lr, betas=(0.5, 0.999)); torch.optim.Adam keeps working
on the drop-in modules - this class is the kernel-backed equivalent used by bench.py and the DP trainer).

torch.optim.Adam semantics (no weight decay, no amsgrad): exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2*v +
(1-b2) g^2; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).  The step counter lives on the device so the
step is CUDA-graph capturable."""
import torch

from . import _lib as L
from . import ops


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        dev = self.params[0].device
        ops.require_cuda(self.params[0], "srk.optim.Adam")
        self.state = [(torch.zeros_like(p, memory_format=torch.contiguous_format),
                       torch.zeros_like(p, memory_format=torch.contiguous_format)) for p in self.params]
        self.step_count = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.param_groups = [{"lr": self.lr, "params": self.params}]

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        import ctypes
        self.step_count += 1
        lr = float(self.param_groups[0]["lr"])
        live = [(p, p.grad if p.grad.is_contiguous() else p.grad.contiguous(), m, v)
                for p, (m, v) in zip(self.params, self.state) if p.grad is not None]
        if not live:
            return
        n = len(live)
        arr = lambda vals: (ctypes.c_void_p * n)(*vals)
        self._keep = [g for _, g, _, _ in live]  # contiguous copies must outlive the launch
        L.call("srk_adam_multi", n, arr([p.data_ptr() for p, _, _, _ in live]),
               arr([g.data_ptr() for _, g, _, _ in live]), arr([m.data_ptr() for _, _, m, _ in live]),
               arr([v.data_ptr() for _, _, _, v in live]), (ctypes.c_int64 * n)(*[p.numel() for p, _, _, _ in live]),
               lr, self.betas[0], self.betas[1], self.eps, self.step_count.data_ptr(), float(grad_scale),
               ops.stream_ptr())
        # the kernels wrote through raw pointers: invalidate the packed-weight cache
        ops.bump_weights_epoch()
