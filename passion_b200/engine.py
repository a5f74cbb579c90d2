"""Training-step engine: what reference train.py:198-289 does per iteration, as one object.

    trainer = Trainer(model, lr=2e-4, weight_decay=1e-4, temp=4.0, mask_type='idt', ...)
    loss, parts = trainer.step(x, target, mask)

One process per GPU.  With torch.distributed initialised (backend nccl) the step adds exactly two
exchanges: a 4-float all-reduce of rp_iter before the loss mix (rp_mask is a global-batch
statistic, train.py:265-268) and a SUM all-reduce of the gradients (the reference loss is a sum over
samples, train.py:229,260-263 — averaging would change the effective LR), bucketed and launched from
autograd hooks on a side stream so that it overlaps the rest of backward (passion_b200/ddp.py).
"""
import torch

from . import ddp as _ddp
from .train_step import loss_mix, loss_mix_baseline


class Trainer:
    def __init__(self, model, lr=2e-4, weight_decay=1e-4, temp=4.0, mask_type="idt", use_passion=True,
                 modal_weight=None, imb_beta=None, warmup=False, distributed=None, bucket_mb=4.0, use_graph=False):
        self.model = model
        dev = next(model.parameters()).device
        self.dev = dev
        model.is_training, model.use_passion, model.mask_type = True, use_passion, mask_type
        self.temp, self.mask_type, self.warmup = temp, mask_type, warmup
        self.modal_weight = (modal_weight if modal_weight is not None else torch.ones(4)).to(dev).float()
        self.imb_beta = (imb_beta if imb_beta is not None else torch.ones(4)).to(dev).float()
        # train.py:94-96
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1
        # The whole step (forward, loss mix, backward incl. the NCCL bucket all-reduces, AdamW) is sync-free and
        # shape-static, so it is replayed as ONE CUDA graph; the optimizer keeps its step counter on the device.
        self.use_graph = bool(use_graph)
        # fused=True: one multi-tensor kernel per parameter group instead of ~a dozen foreach passes (same update rule)
        self.optimizer = torch.optim.AdamW([{"params": model.parameters(), "lr": lr, "weight_decay": weight_decay}],
                                           betas=(0.9, 0.999), eps=1e-08, amsgrad=True, capturable=self.use_graph,
                                           fused=dev.type == "cuda")
        self._graph = None
        self._captured_warmup = None
        self._sig = None
        if self.use_graph:                       # the LR lives in a device tensor so that a replayed graph sees updates
            self._lr_t = torch.tensor(float(lr), device=dev)
            self.optimizer.param_groups[0]["lr"] = self._lr_t
        self.reducer = _ddp.GradReducer(model, bucket_mb=bucket_mb) if distributed else None
        if self.reducer is not None:
            self.reducer.broadcast_parameters()

    def set_lr(self, lr):
        """lr_scheduler.py:42-43 (_adjust_learning_rate)."""
        if self.use_graph:
            self._lr_t.fill_(float(lr))
        else:
            self.optimizer.param_groups[0]["lr"] = lr

    def forward_loss(self, x, target, mask):
        outs = self.model(x, mask, target=target, temp=self.temp)
        if len(outs) == 3:                   # use_passion = False (train.py:374-573)
            return loss_mix_baseline(outs, target, mask, mask_type=self.mask_type, warmup=self.warmup)
        rp_allreduce = self.reducer.allreduce_small if self.reducer is not None else None
        return loss_mix(outs, target, mask, self.imb_beta, self.modal_weight, mask_type=self.mask_type,
                        warmup=self.warmup, rp_allreduce=rp_allreduce)

    def _eager_step(self, x, target, mask):
        loss, parts = self.forward_loss(x, target, mask)
        self.optimizer.zero_grad(set_to_none=True)
        if self.reducer is not None:
            self.reducer.prepare()
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        return loss.detach(), parts

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    @staticmethod
    def _signature(*tensors):
        return tuple((tuple(t.shape), t.dtype) for t in tensors)

    def _capture(self, x, target, mask):
        self._graph = None                               # drop a previous capture (other input format) first
        self._static = tuple(t.clone() for t in (x, target, mask))
        self._sig = self._signature(x, target, mask)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                    # warm-up on a side stream (allocator, lazy inits, autotune-free)
            for _ in range(2):
                self._eager_step(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import ops
        gen = ops.scratch_generation(self.dev)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            loss, parts = self._eager_step(*self._static)
        if ops.scratch_generation(self.dev) != gen:
            # the zero-scratch arena was re-allocated inside the capture: the replayed memset would not cover what the step uses
            raise RuntimeError("passion_b200: the scratch arena grew while the step was being captured; run one more eager step first")
        self._out = (loss, {k: v.detach() for k, v in parts.items()})
        self._captured_warmup = self.warmup

    def step(self, x, target, mask):
        if not self.use_graph:
            return self._eager_step(x, target, mask)
        # the warm-up branch and the input format (one-hot float64 target or uint8 label map, batch / crop size) are baked
        # into a capture: re-capture when either changes
        if self._graph is None or self._captured_warmup != self.warmup or self._sig != self._signature(x, target, mask):
            self._capture(x, target, mask)
        for dst, src in zip(self._static, (x, target, mask)):
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._out


class DevicePrefetcher:
    """Iterates over (x, target, mask) batches that live in pinned host memory and yields them on the device, copying
    batch i+1 on a side stream while step i computes (what the reference gets from its DataLoader workers + `.cuda()`,
    train.py:200-203, minus the serialisation).  Two device buffers per tensor are reused in turn; a batch handed out is
    valid until the next-but-one `next()`."""

    def __init__(self, batches, device, like=None):
        """`like`: an example host batch; the two device buffer sets are then allocated right away (one-time setup)
        instead of on first use.  Copies start with the first `next()`."""
        self.it = iter(batches)
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [None, None]
        if like is not None:
            self.slots = [tuple(torch.empty(t.shape, dtype=t.dtype, device=device) for t in like) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [None, None]
        self.k = 0
        self.pending = None
        self.started = False

    def _issue(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.pending = None
            return
        slot = self.k % 2
        with torch.cuda.stream(self.stream):
            if self.consumed[slot] is not None:
                self.stream.wait_event(self.consumed[slot])          # the step that read this slot has finished
            if self.slots[slot] is None or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(self.slots[slot], host)):
                self.slots[slot] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host)
            for dst, src in zip(self.slots[slot], host):
                dst.copy_(src, non_blocking=True)
            self.ready[slot].record(self.stream)
        self.pending = slot
        self.k += 1

    def __iter__(self):
        return self

    def __next__(self):
        if not self.started:
            self.started = True
            self._issue()
        if self.pending is None:
            raise StopIteration
        slot = self.pending
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        batch = self.slots[slot]
        self._issue()                                               # start copying the following batch right away
        return batch

    def release(self, batch):
        """Call after the step that consumed `batch` has been enqueued: its buffers may then be overwritten."""
        for slot in (0, 1):
            if self.slots[slot] is batch:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self.consumed[slot] = ev
