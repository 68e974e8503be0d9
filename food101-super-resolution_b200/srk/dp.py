"""Data-parallel glue: one process per GPU, weights replicated, the global batch sharded over ranks,
gradients averaged with one all-reduce per flat bucket (NCCL over NVLink on GPUs; any
torch.distributed backend works - the gloo backend is what the CPU tests use).

The reference has no distributed path at all; semantics follow SURVEY 8e: losses are
batch means, so equal shards + gradient averaging reproduce the single-process gradient for
AttentionSR / SRCNN exactly, and ResNet-SR BatchNorm keeps per-rank batch statistics (DDP semantics)."""
import os

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [begin, end) slice of `total` items for `rank` (ragged tails go to low ranks)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def make_buckets(params, bucket_bytes=2 << 20):
    """Groups parameters in reverse registration order (the order backward produces gradients) into
    buckets of about `bucket_bytes`."""
    buckets, cur, size = [], [], 0
    for p in reversed([p for p in params if p.requires_grad]):
        cur.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    return buckets


class GradAverager:
    """Averages .grad of `params` across the process group through flat fp32 buckets."""

    def __init__(self, params, bucket_bytes=2 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = make_buckets(list(params), bucket_bytes)
        self.flat = [None] * len(self.buckets)

    def _flat(self, i, like):
        n = sum(p.numel() for p in self.buckets[i])
        if self.flat[i] is None or self.flat[i].device != like.device:
            self.flat[i] = torch.empty((n,), dtype=torch.float32, device=like.device)
        return self.flat[i]

    # The three phases are separate so that a trainer can capture pack() and unpack() in CUDA graphs (with the
    # backward and the optimizer step respectively) and launch only the collective itself eagerly.
    @torch.no_grad()
    def pack(self):
        """gradients -> flat buckets, pre-divided by the world size"""
        if self.world == 1:
            return
        self._views = []
        for i, bucket in enumerate(self.buckets):
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = self._flat(i, grads[0])
            views = list(flat.split([g.numel() for g in grads]))
            torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
            flat.div_(self.world)
            self._views.append(views)

    @torch.no_grad()
    def all_reduce(self):
        if self.world == 1:
            return
        works = [dist.all_reduce(f, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for f in self.flat
                 if f is not None]
        for w in works:
            w.wait()

    @torch.no_grad()
    def unpack(self):
        """flat buckets -> .grad"""
        if self.world == 1:
            return
        # .grad becomes a view of the averaged flat bucket: no copy kernels (one per parameter before), and the
        # optimizer reads the same memory the collective wrote.  The views stay valid until the next pack().
        for bucket, views in zip(self.buckets, self._views):
            for p, v in zip(bucket, views):
                p.grad = v.view_as(p)

    def average(self):
        self.pack()
        self.all_reduce()
        self.unpack()

    # ---- overlapped mode: buckets are reduced while backward is still running ------------------------------------
    # begin_backward() before loss.backward(), finish_backward() after it.  A bucket (reverse registration order =
    # the order backward produces gradients) is launched as soon as every one of its parameters has its gradient
    # ENQUEUED: on the main stream (autograd's AccumulateGrad, seen through a post-accumulate hook) or on the
    # weight-gradient side stream (ops.run_on_side_stream hands its (parameter, gradient) pairs over through
    # ops.set_side_grad_listener).  The pack + NCCL all-reduce of the bucket then run on a communication stream
    # ordered after both; finish_backward() joins that stream and points .grad at the averaged flat buckets.  All of
    # it is stream-ordered device work, so it is captured into the trainer's CUDA graph like everything else.
    def begin_backward(self):
        from . import ops
        if self.world == 1:
            return
        dev = self.buckets[0][0].device
        if getattr(self, "_owner", None) is None:
            self._comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None   # CPU (gloo tests): no streams
            self._owner = {}
            for i, bucket in enumerate(self.buckets):
                for p in bucket:
                    self._owner[p] = i
                    p.register_post_accumulate_grad_hook(self._on_autograd_grad)
                    p._srk_dp_hooked = True
        self._counted = set()
        self._side = {}
        self._launched = [False] * len(self.buckets)
        self._views = [None] * len(self.buckets)
        self._active = True
        ops.set_side_grad_listener(self._on_side_grad)

    def _on_autograd_grad(self, p):
        # the hook also fires when the backward node handed autograd an undefined gradient (what the side-stream path
        # returns for its parameters): those are announced by the listener, not here
        if getattr(self, "_active", False) and p in self._owner and p.grad is not None and p not in self._side:
            self._count(p)

    def _on_side_grad(self, p, g):
        if getattr(self, "_active", False) and p in self._owner:
            self._side[p] = g
            self._count(p)

    def _count(self, p):
        i = self._owner[p]
        if self._launched[i]:
            # a gradient contribution that arrives after its bucket went out would be lost silently
            raise RuntimeError("srk.dp: a parameter %s received gradient after its bucket was all-reduced (parameter "
                               "used in two places of the graph?); use GraphStep(overlap_comm=False)"
                               % (tuple(p.shape),))
        self._counted.add(p)
        # launch on availability, not on a count: every parameter of the bucket has its gradient enqueued
        if all((q in self._side) or (q.grad is not None) for q in self.buckets[i]):
            self._launch(i)

    @torch.no_grad()
    def _launch(self, i):
        from . import ops
        self._launched[i] = True
        bucket = self.buckets[i]
        grads = [self._side[p] if p in self._side else (p.grad if p.grad is not None else torch.zeros_like(p)) for p in bucket]
        flat = self._flat(i, grads[0])

        def reduce_bucket():
            views = list(flat.split([g.numel() for g in grads]))
            torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
            flat.div_(self.world)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            return views

        if self._comm is None:
            self._views[i] = reduce_bucket()
            return
        main = torch.cuda.current_stream()
        self._comm.wait_stream(main)
        self._comm.wait_stream(ops._side_stream(main.device))
        with torch.cuda.stream(self._comm):
            self._views[i] = reduce_bucket()
        for g in grads:
            g.record_stream(self._comm)

    @torch.no_grad()
    def finish_backward(self):
        """After loss.backward() (the side stream has been joined): launch what is left, join the communication
        stream, point .grad at the averaged buckets."""
        from . import ops
        if self.world == 1:
            return
        ops.set_side_grad_listener(None)
        self._active = False
        for i in range(len(self.buckets)):
            if not self._launched[i]:
                self._launch(i)
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)
        for bucket, views in zip(self.buckets, self._views):
            for p, v in zip(bucket, views):
                p.grad = v.view_as(p)


def broadcast_parameters(module, src=0, group=None):
    """Makes every rank start from rank `src`'s weights and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def broadcast_buffers(module, src=0, group=None):
    """BatchNorm running statistics are per rank during training (DDP semantics): before an evaluation whose
    result steers control flow, and before a checkpoint, every rank takes rank `src`'s buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in module.buffers():
        dist.broadcast(t.data, src=src, group=group)
