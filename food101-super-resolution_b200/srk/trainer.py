"""The training step of reference train.py:116-120 -

    optimizer.zero_grad(); loss = criterion(model(lr_imgs), hr_imgs); loss.backward(); optimizer.step()

- as ONE object that train.py, bench.py and the tests share, so that the entry point runs at the measured speed:

  * srk.optim.Adam (multi-tensor kernel, device-resident step counter and learning rate) followed by one
    multi-tensor re-pack of every conv weight into its tensor-core operand layouts (ops.repack_all);
  * weight gradients on the side stream (ops.run_on_side_stream, joined at the end of backward());
  * under data parallelism the gradient all-reduce (srk.dp.GradAverager: flat fp32 buckets, NCCL) between backward
    and the optimizer;
  * after `warmup` eager steps on a given input shape the whole step - forward, loss, backward, all-reduce, Adam,
    re-pack - is captured into one CUDA graph and replayed: same kernels, same arithmetic (bit-identical results,
    tests/test_gpu_determinism.py), no per-launch host work.  Other shapes (the ragged last batch of an epoch) run
    eagerly.  The learning rate is read from device memory, so ReduceLROnPlateau keeps working on the replayed graph.

Not captured, by construction: anything the caller does between steps (logging, .item()).  The returned loss is a
device scalar; under replay it is a static buffer that the next step overwrites."""
import torch

from . import dp, ops
from . import optim as srk_optim


class GraphStep:
    def __init__(self, model, criterion, lr=4e-4, betas=(0.5, 0.999), eps=1e-8, optimizer=None, averager=None,
                 use_graph=True, warmup=3, overlap_wgrad=True, extra_backward=None, overlap_comm=True):
        self.model, self.criterion = model, criterion
        self.optimizer = optimizer if optimizer is not None else srk_optim.Adam(model.parameters(), lr=lr, betas=betas,
                                                                              eps=eps)
        self.averager = averager
        self.overlap_comm = bool(overlap_comm)
        self.use_graph = bool(use_graph)
        self.warmup = max(int(warmup), 1)
        self.extra_backward = extra_backward       # optional callable(out, hr) -> extra loss term (GAN generator step)
        ops.set_overlap_wgrad(overlap_wgrad)
        self._seen = {}        # input-shape key -> eager steps run so far
        self._graphs = {}      # input-shape key -> (graph, static lr, static hr, static loss)
        self._stream = None
        self.launches_per_step = None   # libsrk launches of one eager step (bench.py's gpu_launches)

    # ---- the step itself (eager or under capture) -------------------------------------------------------------
    def _run(self, lr_imgs, hr_imgs):
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(lr_imgs)
        loss = self.criterion(out, hr_imgs)
        if self.extra_backward is not None:
            loss = loss + self.extra_backward(out, hr_imgs)
        if self.averager is not None and self.overlap_comm:
            self.averager.begin_backward()      # buckets are all-reduced on a communication stream DURING backward
            loss.backward()
            self.averager.finish_backward()
        else:
            loss.backward()
            if self.averager is not None:
                self.averager.average()
        self.optimizer.step()
        if isinstance(self.optimizer, srk_optim.Adam):
            ops.repack_all()   # the weights changed through raw pointers: refresh every cached operand pack in one launch
        return loss.detach()

    def _capture(self, key, lr_imgs, hr_imgs):
        from . import _lib as L
        dev = lr_imgs.device
        s_lr, s_hr = lr_imgs.clone(), hr_imgs.clone()
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        torch.cuda.synchronize(dev)
        ops.repack_all()       # packs are current; inside the graph only the post-Adam re-pack refreshes them
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._stream):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream, capture_error_mode="thread_local"):
                s_loss = self._run(s_lr, s_hr)
        torch.cuda.current_stream(dev).wait_stream(self._stream)
        torch.cuda.synchronize(dev)
        # capture executed nothing: the optimizer's host-side bookkeeping is unaffected, device state untouched
        self._graphs[key] = (g, s_lr, s_hr, s_loss)

    def __call__(self, lr_imgs, hr_imgs):
        from . import _lib as L
        key = (tuple(lr_imgs.shape), tuple(hr_imgs.shape), lr_imgs.dtype, ops.cfg.compute_dtype)
        ent = self._graphs.get(key)
        if ent is not None:
            g, s_lr, s_hr, s_loss = ent
            if isinstance(self.optimizer, srk_optim.Adam):
                self.optimizer.sync_lr()
            s_lr.copy_(lr_imgs, non_blocking=True)
            s_hr.copy_(hr_imgs, non_blocking=True)
            g.replay()
            return s_loss
        n = self._seen.get(key, 0)
        if self.use_graph and n >= self.warmup and isinstance(self.optimizer, srk_optim.Adam):
            self._capture(key, lr_imgs, hr_imgs)
            return self(lr_imgs, hr_imgs)
        c0 = L.launch_calls
        loss = self._run(lr_imgs, hr_imgs)
        self.launches_per_step = L.launch_calls - c0
        self._seen[key] = n + 1
        return loss

    def static_inputs(self, lr_imgs, hr_imgs):
        """The graph's own input buffers for this shape (None before capture): a data pipeline may write the next batch
        straight into them and call replay() instead of __call__ (bench.py's end-to-end leg)."""
        ent = self._graphs.get((tuple(lr_imgs.shape), tuple(hr_imgs.shape), lr_imgs.dtype, ops.cfg.compute_dtype))
        return None if ent is None else (ent[1], ent[2])

    def replay(self, lr_imgs, hr_imgs):
        """Replays the captured step of this shape on whatever its static inputs hold now -> static loss."""
        g, _, _, s_loss = self._graphs[(tuple(lr_imgs.shape), tuple(hr_imgs.shape), lr_imgs.dtype, ops.cfg.compute_dtype)]
        if isinstance(self.optimizer, srk_optim.Adam):
            self.optimizer.sync_lr()
        g.replay()
        return s_loss


def make_trainer(model, criterion, lr, world=1, use_graph=True, warmup=3, group=None):
    """Trainer for `world` data-parallel ranks: broadcasts rank 0's weights and buffers, builds the gradient averager."""
    averager = None
    if world > 1:
        dp.broadcast_parameters(model, group=group)
        averager = dp.GradAverager(model.parameters(), group=group)
    return GraphStep(model, criterion, lr=lr, betas=(0.5, 0.999), averager=averager, use_graph=use_graph, warmup=warmup)
