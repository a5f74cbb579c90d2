"""One-process-per-GPU data parallelism for the PASSION step (replaces torch.nn.DataParallel, reference
train.py:90): NCCL over NVLink 5 / NVSwitch through torch.distributed.

Semantics follow the reference:
  * weights are broadcast ONCE at start (DataParallel re-broadcasts them every step);
  * gradients are SUM-reduced, not averaged: the reference loss is a sum over the samples of the global
    batch (train.py:229, 260-263);
  * gradients live in a few flat fp32 buckets (param.grad are views into them); a bucket's all-reduce is
    launched from a post-accumulate hook as soon as its last gradient has been written, so the
    collective runs on NCCL's stream while the rest of backward is still executing;
  * `allreduce_small` carries the 4-float rp_iter exchange that makes rp_mask a global-batch statistic.
The same class works on CPU tensors with the gloo backend (tests/test_ddp_cpu.py).
"""
import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, model, bucket_mb=4.0, group=None, overlap=None):
        """overlap: launch a bucket's all-reduce from the autograd hook of its last gradient (during backward).  Off by default
        when the conv weight gradients are scattered into the buckets by ONE launch at the END of backward (ops.BATCH_WEIGHTS):
        a bucket is then only complete after that launch, and hook counts say nothing about it (round-2 finding on 2 GPUs: the
        early all-reduce ran before the scatter and the conv gradients stayed rank-local).  All buckets are then reduced in
        finish(), right behind the scatter: 9.5 MB over NVLink is tens of microseconds."""
        self.group = group
        if overlap is None:
            from . import ops
            overlap = not ops.BATCH_WEIGHTS
        self.overlap = bool(overlap)
        self.params = [p for p in model.parameters() if p.requires_grad]
        cap = int(bucket_mb * 2 ** 20 / 4)
        # backward produces gradients roughly in reverse registration order: bucket in that order
        self.buckets = []           # each: dict(params=[...], flat=tensor, pending=int, work=None)
        cur, cur_n = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_n += p.numel()
            if cur_n >= cap:
                self.buckets.append(self._make_bucket(cur))
                cur, cur_n = [], 0
        if cur:
            self.buckets.append(self._make_bucket(cur))
        self._slot = {}                # id(param) -> (bucket, index in bucket)
        for b in self.buckets:
            for i, p in enumerate(b["params"]):
                self._slot[id(p)] = (b, i)
                p.register_post_accumulate_grad_hook(self._hook)
        self.enabled = True

    @staticmethod
    def _make_bucket(params):
        n = sum(p.numel() for p in params)
        flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
        views, off = [], 0
        for p in params:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        return {"params": list(params), "flat": flat, "views": views, "pending": len(params), "work": None}

    def broadcast_parameters(self, src=0):
        for p in self.params:
            dist.broadcast(p.data, src=src, group=self.group)

    def prepare(self):
        """Call after optimizer.zero_grad(set_to_none=True), before backward."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["work"] = None
            for p, v in zip(b["params"], b["views"]):
                p.grad = v

    def _hook(self, p):
        if not self.enabled:
            return
        b, i = self._slot[id(p)]
        v = b["views"][i]
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
            # autograd replaced the view instead of accumulating into it: copy back into the bucket
            v.copy_(p.grad)
            p.grad = v
        b["pending"] -= 1
        if b["pending"] == 0 and self.overlap:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for all bucket reductions (launching any whose hooks did not all fire)."""
        for b in self.buckets:
            if b["work"] is None:
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        for b in self.buckets:
            b["work"].wait()

    def allreduce_small(self, t):
        """SUM all-reduce of a tiny tensor that stays on the autograd tape as a constant-gradient op
        (rp_iter only feeds a comparison, so no gradient flows through it)."""
        t = t.detach().clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t
