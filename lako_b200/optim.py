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


def shard_plan(ranges, world, rank, align=8):
    """Split every gradient-stage range [lo, hi) of the flat buffer into `world` equal pieces and return
    [(lo, hi, piece_lo, piece_hi)] for `rank`, or None when some range does not divide into `align`-element
    aligned pieces (the caller then keeps the replicated all-reduce path)."""
    plan = []
    for lo, hi in ranges:
        n = hi - lo
        if n == 0:
            continue
        if n % (world * align) != 0:
            return None
        c = n // world
        plan.append((lo, hi, lo + rank * c, lo + (rank + 1) * c))
    return plan


class FusedAdamW:
    """One pass over (p, m, v, g): g *= min(1, clip/(||g||+1e-6)); m,v update without bias correction;
    p -= lr*m/(sqrt(v)+eps); p -= lr*wd*p; bf16 copy of p refreshed in the same kernel.

    Data parallel (world > 1).  Default = REPLICATED: the flat gradient is all-reduced (overlapped with backward) and
    every rank runs the full update, so every rank's fp32 masters stay current and `model.state_dict()` /
    `save_pretrained` on the main rank only (as the reference does, train_reader.py:107 `if opt.is_main`) just work.
    `shard=True` (or LAKO_DP_SHARD=1) opts into SHARDED mode: every gradient stage is REDUCE-SCATTERED while backward is
    still running, each rank clips + updates only its 1/world slice of every stage (fp32 master, moments) and the
    refreshed bf16 compute weights are ALL-GATHERED -- 6 B/parameter on the wire instead of the all-reduce's 8, and
    1/world of the optimizer's HBM traffic (the parameters the kernels read in fp32 -- shared embedding, norm gains, bias
    tables -- are gathered in fp32, _index_fp32_reads).  The price: fp32 masters of the other foreign slices go stale, and
    `sync_master_weights()` -- a COLLECTIVE, also triggered by FiDT5.state_dict / save_pretrained and by anything that
    re-casts the masters -- must then be reached by EVERY rank (call it, or state_dict(), on all ranks before a
    main-rank-only save)."""

    def __init__(self, model, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-6, clip=1.0, schedule=None,
                 process_group=None, allreduce_chunk_elems=64 * 1024 * 1024, shard=None):
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
        self._plan = None          # sharded mode: [(lo, hi, piece_lo, piece_hi)] per gradient stage range
        self._masters_stale = False
        self._skip_allreduce = os.environ.get("LAKO_SKIP_ALLREDUCE") == "1"   # diagnostics only: wrong gradients
        if self.world() > 1 and model._flat.is_cuda and os.environ.get("LAKO_DP_OVERLAP", "1") != "0":
            # overlap: backward records an event when a range of the flat gradient is final; the collective of that
            # range then runs on a communication stream while the rest of backward is still executing
            _, ranges = model.enable_grad_stage_events()
            self._comm_stream = torch.cuda.Stream()
            if shard is None:
                shard = os.environ.get("LAKO_DP_SHARD", "0") == "1"
            if shard:
                self._plan = shard_plan(ranges, self.world(), dist.get_rank(self.group))
                if self._plan is not None:
                    model._master_sync = self.sync_master_weights
                    self._index_fp32_reads()
                    # opt-in: measured at cfg4 on 2 and 8 GPUs it changes nothing (91.63 vs 91.66 ms/step at 8) -- the
                    # all-gather is not what separates the 8-GPU step from the 1-GPU step
                    if os.environ.get("LAKO_DP_GATHER_OVERLAP", "0") == "1":
                        self._init_gather_overlap(ranges)

    def _init_gather_overlap(self, ranges):
        """The all-gather of the updated weights runs on the communication stream and the next training forward starts
        without waiting for it: one event per gradient stage (recorded when that stage's weights are complete) is handed
        to lako_fidt5_forward (lako_batch.weight_events), which waits right before the first kernel that reads the
        range.  Ranges are gathered in the order the forward uses them: embedding and small vectors, encoder groups from
        the bottom up, decoder last."""
        stage_of = {lo: min(i, 5) for i, (lo, hi) in enumerate(ranges)}
        use_order = {5: 0, 4: 1, 3: 2, 2: 3, 1: 4, 0: 5}
        self._gather_order = sorted(self._plan, key=lambda e: (use_order[stage_of[e[0]]], e[0]))
        self._gather_stage = [stage_of[e[0]] for e in self._gather_order]
        self._wevents = [None] * 6
        for st in set(self._gather_stage):
            e = torch.cuda.Event()
            e.record()             # forces creation of the underlying cudaEvent_t
            self._wevents[st] = e
        self._wevent_handles = (ctypes.c_void_p * 6)(*[(e.cuda_event if e is not None else 0) for e in self._wevents])
        self._upd_event = torch.cuda.Event()

    def _index_fp32_reads(self):
        """The kernels read three kinds of parameters from the fp32 MASTERS, not from the bf16 copy: the shared embedding
        (embed_gather), every T5LayerNorm gain and the two relative-bias tables.  In sharded mode those must therefore be
        current on every rank after each step, not only the bf16 weights: the embedding range is all-gathered in fp32
        (and re-cast locally), the small vectors travel in one packed all-reduce (each rank contributes the elements it
        owns, zeros elsewhere)."""
        model = self.model
        off, shape = model._specs["shared.weight"]
        n = 1
        for d in shape:
            n *= d
        self._emb_range = next(((lo, hi) for lo, hi, _, _ in self._plan if lo == off and hi == off + n), None)
        if self._emb_range is None:
            raise nat.LakoError("sharded optimizer: the shared embedding is not a gradient stage range of its own")
        idx = []
        for name, (o, shp) in model._specs.items():
            if name.endswith("layer_norm.weight") or name.endswith("relative_attention_bias.weight"):
                cnt = 1
                for d in shp:
                    cnt *= d
                idx.append(torch.arange(o, o + cnt, dtype=torch.long))
        idx = torch.cat(idx)
        own = torch.zeros(idx.numel(), dtype=torch.bool)
        for _, _, plo, phi in self._plan:
            own |= (idx >= plo) & (idx < phi)
        dev = model._flat.device
        self._small_idx = idx.to(dev)
        self._small_own = own.to(dev, torch.float32)

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

    # ---------------------------------------------------------------- sharded data parallelism
    def _reduce_scatter(self, g, lo, hi, plo, phi):
        """SUM over ranks of g[lo:hi]; this rank needs the result only in its piece g[plo:phi] (in place)."""
        if dist.get_backend(self.group) == "nccl":
            dist.reduce_scatter_tensor(g[plo:phi], g[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        else:  # gloo (CPU tests) has no reduce-scatter: the all-reduce is a superset
            dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

    def _all_gather(self, t, lo, hi, plo, phi):
        """Every rank's piece t[plo:phi] -> the whole range t[lo:hi] on every rank (in place)."""
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(t[lo:hi], t[plo:phi], group=self.group)
        else:
            parts = [torch.empty_like(t[plo:phi]) for _ in range(self.world())]
            dist.all_gather(parts, t[plo:phi].clone(), group=self.group)
            t[lo:hi].copy_(torch.cat(parts))

    def reduce_scatter_grads(self):
        g = self.model._flat_grad
        evs = self.model._stage_events
        if self._comm_stream is None or evs is None:
            for lo, hi, plo, phi in self._plan:
                self._reduce_scatter(g, lo, hi, plo, phi)
            return
        ranges = self.model._stage_ranges
        stage_of = {}  # range start -> stage event index (ranges 5 and 6 both belong to the last event)
        for i, (lo, hi) in enumerate(ranges):
            stage_of[lo] = min(i, 5)
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._comm_stream):
            for lo, hi, plo, phi in self._plan:
                self._comm_stream.wait_event(evs[stage_of[lo]])
                self._reduce_scatter(g, lo, hi, plo, phi)
        main.wait_stream(self._comm_stream)

    def _update_piece(self, off, n, lr):
        """clip + AdamW + bf16 re-cast of flat elements [off, off+n) (sumsq already holds the global ||g||^2)."""
        model, L = self.model, nat.lib()
        nat.check(L.lako_adamw_step(nat.ptr(model._flat[off:]), nat.ptr(self.m[off:]), nat.ptr(self.v[off:]),
                                    nat.ptr(model._flat_grad[off:]), nat.ptr(model._bf16[off:]), ctypes.c_size_t(n),
                                    nat.ptr(self.sumsq), ctypes.c_float(1.0),
                                    ctypes.c_float(self.clip if self.clip else 0.0), ctypes.c_float(lr),
                                    ctypes.c_float(self.betas[0]), ctypes.c_float(self.betas[1]),
                                    ctypes.c_float(self.eps), ctypes.c_float(self.wd), nat.stream_ptr()),
                  "lako_adamw_step")

    def _sumsq_piece(self, off, n):
        nat.check(nat.lib().lako_grad_sumsq(nat.ptr(self.model._flat_grad[off:]), ctypes.c_size_t(n), nat.ptr(self.sumsq),
                                            nat.stream_ptr()), "lako_grad_sumsq")

    def _step_sharded(self, lr):
        model = self.model
        self.reduce_scatter_grads()
        self.sumsq.zero_()
        for lo, hi, plo, phi in self._plan:
            self._sumsq_piece(plo, phi - plo)
        dist.all_reduce(self.sumsq, op=dist.ReduceOp.SUM, group=self.group)   # pieces are disjoint: global ||g||^2
        for lo, hi, plo, phi in self._plan:
            self._update_piece(plo, phi - plo, lr)
        if getattr(self, "_wevents", None) is not None:
            main = torch.cuda.current_stream()
            self._upd_event.record(main)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(self._upd_event)
                self._gather_weights(self._gather_order, self._gather_stage)
            model._pending_wev = (self._wevents, self._wevent_handles)
        else:
            self._gather_weights(self._plan, None)
        self._masters_stale = True

    def _gather_weights(self, entries, stages):
        """All-gather the refreshed weights range by range (on the current stream); with `stages`, record that stage's
        event once its last range is complete."""
        model = self.model
        emb = getattr(self, "_emb_range", None)
        small = getattr(self, "_small_idx", None)
        for i, (lo, hi, plo, phi) in enumerate(entries):
            if emb == (lo, hi):   # read in fp32 by embed_gather: gather the masters, re-cast the whole range locally
                self._all_gather(model._flat, lo, hi, plo, phi)
                nat.check(nat.lib().lako_cast_f32_to_bf16(nat.ptr(model._flat[lo:]), nat.ptr(model._bf16[lo:]),
                                                          ctypes.c_size_t(hi - lo), nat.stream_ptr()), "lako_cast_f32_to_bf16")
                if small is not None:     # norm gains and bias tables, also read in fp32 (see _index_fp32_reads)
                    buf = model._flat[small] * self._small_own
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                    model._flat.index_copy_(0, small, buf)
            else:
                self._all_gather(model._bf16, lo, hi, plo, phi)
            if stages is not None and (i + 1 == len(entries) or stages[i + 1] != stages[i]):
                self._wevents[stages[i]].record()

    @torch.no_grad()
    def sync_master_weights(self):
        """All-gather the fp32 master weights (each rank only keeps its own slices current in sharded mode)."""
        if self._plan is None or not self._masters_stale:
            return
        if hasattr(self.model, "_wait_weights"):
            self.model._wait_weights()
        for lo, hi, plo, phi in self._plan:
            self._all_gather(self.model._flat, lo, hi, plo, phi)
        self._masters_stale = False
        self.model.mark_bf16_fresh()

    @torch.no_grad()
    def step(self):
        model = self.model
        g = model._flat_grad
        if g is None:
            raise nat.LakoError("FusedAdamW.step() before any backward")
        if self._plan is not None and not self._skip_allreduce:
            lr = self.lr * (self.schedule(self.step_count) if self.schedule is not None else 1.0)
            self.last_lr = lr
            model._wait_weights()   # (no forward since the last step: its all-gather may still be running)
            if model._bf16 is None:
                model._refresh_bf16()
            self._step_sharded(lr)
            model.mark_bf16_fresh()
            self.step_count += 1
            return
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

    # ---------------------------------------------------------------- checkpoint (src/util.py:105-146)
    @torch.no_grad()
    def state_dict(self):
        """What src/util.py:save stores as checkpoint["optimizer"] (and, for the built-in schedule, what it stores as
        checkpoint["scheduler"]: the step count).  In sharded data-parallel mode the Adam moments are all-gathered
        first -- a collective, call it on every rank."""
        if self._plan is not None and self.step_count > 0:
            for lo, hi, plo, phi in self._plan:
                self._all_gather(self.m, lo, hi, plo, phi)
                self._all_gather(self.v, lo, hi, plo, phi)
        return {"step": self.step_count, "exp_avg": self.m.detach().clone(), "exp_avg_sq": self.v.detach().clone(),
                "lr": self.lr, "weight_decay": self.wd, "betas": tuple(self.betas), "eps": self.eps, "clip": self.clip}

    @torch.no_grad()
    def load_state_dict(self, sd):
        if sd["exp_avg"].numel() != self.m.numel():
            raise nat.LakoError("optimizer state does not match the model's flat parameter buffer")
        self.m.copy_(sd["exp_avg"].to(self.m.device))
        self.v.copy_(sd["exp_avg_sq"].to(self.v.device))
        self.step_count = int(sd["step"])
        self.lr, self.wd, self.eps, self.clip = sd["lr"], sd["weight_decay"], sd["eps"], sd["clip"]
        self.betas = tuple(sd["betas"])

    def zero_grad(self):
        self.model.zero_grad()


def global_target_count(labels, group=None):
    """Number of non-ignored targets over all ranks (device scalar): the CE mean's denominator, so
    that question-sharded data parallelism reproduces the single-process loss."""
    cnt = (labels != -100).sum().to(torch.float32).reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return cnt


def train_step(model, optimizer, input_ids, attention_mask, labels, group=None, passage_lengths=None):
    """train_reader.py:67-79 for one batch: forward, backward, (all-reduce), clip, AdamW, zero_grad.
    Returns the (global-mean) loss as a device scalar.  With `passage_lengths` (host [B][N] valid-prefix lengths, see
    FiDT5._packing) the step contains no host synchronisation at all."""
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    cnt = global_target_count(labels, group) if distributed else None
    loss = model(input_ids=input_ids, attention_mask=attention_mask, labels=labels, global_target_count=cnt,
                 passage_lengths=passage_lengths)[0]
    loss.backward()
    optimizer.step()
    model.zero_grad()
    loss = loss.detach()
    if distributed:
        # each rank holds (its token-loss sum) / (global target count); the sum over ranks is the global mean
        loss = loss.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return loss


def gather_answers_and_scores(answers, scores, max_length, group=None):
    """Inference side of the question-sharded layout: every rank decodes its own questions and only the
    results travel -- `answers` int64 (B, <= max_length) as returned by FiDT5.generate (right-padded with pad id 0 to
    `max_length` so that every rank contributes the same shape) and `scores` (B, n_context) as returned by
    get_crossattention_scores.  Returns (answers (world*B, max_length), scores (world*B, n_context)) in rank order on
    every rank; with one process the inputs are returned (answers padded).  The reference has no counterpart: it never
    gathers predictions across ranks (test_reader.py writes per-rank files, src/util.py:248-273 reduces scalars only)."""
    B, cols = answers.shape
    if cols < max_length:
        answers = torch.nn.functional.pad(answers, (0, max_length - cols))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return answers, scores
    world = dist.get_world_size(group)
    dev = answers.device
    scores = scores.to(dev)
    out_a = torch.empty(world * B, max_length, dtype=answers.dtype, device=dev)
    out_s = torch.empty(world * B, scores.shape[1], dtype=scores.dtype, device=dev)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out_a, answers.contiguous(), group=group)
        dist.all_gather_into_tensor(out_s, scores.contiguous(), group=group)
    else:
        pa = [torch.empty_like(answers) for _ in range(world)]
        ps = [torch.empty_like(scores) for _ in range(world)]
        dist.all_gather(pa, answers.contiguous(), group=group)
        dist.all_gather(ps, scores.contiguous(), group=group)
        out_a, out_s = torch.cat(pa), torch.cat(ps)
    return out_a, out_s
