"""Train-step tail of the reader: clip_grad_norm_(1.0) + HF AdamW(correct_bias=False) + the
WarmupLinearScheduler, fused over the flat parameter buffer, with the data-parallel gradient
all-reduce the north star adds.

Reference: train_reader.py:73-79 (backward, clip, optimizer.step, scheduler.step, zero_grad),
src/util.py:149-168 (WarmupLinearScheduler), :185-245 (AdamW(correct_bias=False), wd on both groups).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _native as nat


class WarmupLinearSchedule:
    """lr multiplier of src/util.py:149-168 (min_ratio=0, fixed_lr=False as set at util.py:243)."""

    def __init__(self, warmup_steps, scheduler_steps, min_ratio=0.0, fixed_lr=False):
        self.warmup_steps, self.scheduler_steps = warmup_steps, scheduler_steps
        self.min_ratio, self.fixed_lr = min_ratio, fixed_lr

    def __call__(self, step):
        if step < self.warmup_steps:
            return (1 - self.min_ratio) * step / float(max(1, self.warmup_steps)) + self.min_ratio
        if self.fixed_lr:
            return 1.0
        return max(0.0, 1.0 + (self.min_ratio - 1) * (step - self.warmup_steps)
                   / float(max(1.0, self.scheduler_steps - self.warmup_steps)))


class FusedAdamW:
    """One pass over (p, m, v, g): g *= min(1, clip/(||g||+1e-6)); m,v update without bias correction;
    p -= lr*m/(sqrt(v)+eps); p -= lr*wd*p; bf16 copy of p refreshed in the same kernel."""

    def __init__(self, model, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-6, clip=1.0, schedule=None,
                 process_group=None, allreduce_chunk_elems=64 * 1024 * 1024):
        self.model = model
        self.lr, self.wd, self.betas, self.eps, self.clip = lr, weight_decay, betas, eps, clip
        self.schedule = schedule
        self.step_count = 0
        self.m = torch.zeros_like(model._flat)
        self.v = torch.zeros_like(model._flat)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=model._flat.device)
        self.group = process_group
        self.chunk = allreduce_chunk_elems
        self.last_lr = lr
        self._comm_stream = None
        self._skip_allreduce = os.environ.get("LAKO_SKIP_ALLREDUCE") == "1"   # diagnostics only: wrong gradients
        if self.world() > 1 and model._flat.is_cuda and os.environ.get("LAKO_DP_OVERLAP", "1") != "0":
            # overlap: backward records an event when a range of the flat gradient is final; the all-reduce of that
            # range then runs on a communication stream while the rest of backward is still executing
            model.enable_grad_stage_events()
            self._comm_stream = torch.cuda.Stream()

    def world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def allreduce_grads(self):
        """SUM all-reduce of the flat fp32 gradient in large chunks (NCCL over NVLink when on GPUs).
        With the loss normalised by the GLOBAL target count on every rank (FiDT5.forward's
        global_target_count), the sum equals the single-process gradient of the global batch."""
        g = self.model._flat_grad
        if self._comm_stream is not None and self.model._stage_events is not None:
            evs, ranges = self.model._stage_events, self.model._stage_ranges
            main = torch.cuda.current_stream()
            with torch.cuda.stream(self._comm_stream):
                for stage, rs in ((0, ranges[0:1]), (1, ranges[1:2]), (2, ranges[2:3]), (3, ranges[3:4]), (4, ranges[4:5]),
                                  (5, ranges[5:7])):
                    self._comm_stream.wait_event(evs[stage])
                    for lo, hi in rs:
                        for off in range(lo, hi, self.chunk):
                            dist.all_reduce(g[off:min(hi, off + self.chunk)], op=dist.ReduceOp.SUM, group=self.group)
            main.wait_stream(self._comm_stream)
            return
        works = []
        for off in range(0, g.numel(), self.chunk):
            works.append(dist.all_reduce(g[off:off + self.chunk], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()

    @torch.no_grad()
    def step(self):
        model = self.model
        g = model._flat_grad
        if g is None:
            raise nat.LakoError("FusedAdamW.step() before any backward")
        if self.world() > 1 and not self._skip_allreduce:
            self.allreduce_grads()
        lr = self.lr * (self.schedule(self.step_count) if self.schedule is not None else 1.0)
        # the scheduler is stepped AFTER optimizer.step() in the reference (train_reader.py:77-78), so
        # update k uses lambda(k) with k counted from 0.
        self.last_lr = lr
        L = nat.lib()
        n = ctypes.c_size_t(g.numel())
        self.sumsq.zero_()
        nat.check(L.lako_grad_sumsq(nat.ptr(g), n, nat.ptr(self.sumsq), nat.stream_ptr()), "lako_grad_sumsq")
        if model._bf16 is None:
            model._refresh_bf16()
        nat.check(L.lako_adamw_step(nat.ptr(model._flat), nat.ptr(self.m), nat.ptr(self.v), nat.ptr(g),
                                    nat.ptr(model._bf16), n, nat.ptr(self.sumsq), ctypes.c_float(1.0),
                                    ctypes.c_float(self.clip if self.clip else 0.0), ctypes.c_float(lr),
                                    ctypes.c_float(self.betas[0]), ctypes.c_float(self.betas[1]),
                                    ctypes.c_float(self.eps), ctypes.c_float(self.wd), nat.stream_ptr()),
                  "lako_adamw_step")
        model.mark_bf16_fresh()
        self.step_count += 1

    def grad_norm(self):
        return torch.sqrt(self.sumsq)

    def zero_grad(self):
        self.model.zero_grad()


def global_target_count(labels, group=None):
    """Number of non-ignored targets over all ranks (device scalar): the CE mean's denominator, so
    that question-sharded data parallelism reproduces the single-process loss."""
    cnt = (labels != -100).sum().to(torch.float32).reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return cnt


def train_step(model, optimizer, input_ids, attention_mask, labels, group=None):
    """train_reader.py:67-79 for one batch: forward, backward, (all-reduce), clip, AdamW, zero_grad.
    Returns the (global-mean) loss as a device scalar; no host synchronisation."""
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    cnt = global_target_count(labels, group) if distributed else None
    loss = model(input_ids=input_ids, attention_mask=attention_mask, labels=labels, global_target_count=cnt)[0]
    loss.backward()
    optimizer.step()
    model.zero_grad()
    loss = loss.detach()
    if distributed:
        # each rank holds (its token-loss sum) / (global target count); the sum over ranks is the global mean
        loss = loss.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return loss
