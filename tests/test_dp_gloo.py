"""Question-sharded data parallelism on CPU with gloo, world_size 2: the host-side scheme used by
lako_b200.optim (global target count + SUM all-reduce of flat gradients) reproduces the single-process
gradient of the global batch.  The oracle stands in for the gradient engine (no GPU here)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fidt5_oracle as O


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lako_b200.optim import FusedAdamW, global_target_count
    torch.set_num_threads(1)
    cfg = O.T5_TINY
    sd = O.init_state_dict(cfg, seed=1, emb_std=0.3)
    params = O.unique_params(sd)
    for p in params.values():
        p.requires_grad_(True)
    ids, mask, labels = O.synth_batch(4, 2, 24, 5, cfg.vocab, seed=9)
    sl = slice(rank * 2, rank * 2 + 2)  # this rank's questions
    cnt = global_target_count(labels[sl])
    assert float(cnt) == float((labels != -100).sum())
    _, logits = O.fidt5_forward(sd, cfg, ids[sl], mask[sl], labels[sl])
    loss_sum = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)), labels[sl].reshape(-1),
                                                 ignore_index=-100, reduction="sum")
    (loss_sum / cnt).backward()
    flat = torch.cat([p.grad.reshape(-1) for p in params.values()])
    shim = type("M", (), {})()
    shim._flat_grad = flat
    opt = FusedAdamW.__new__(FusedAdamW)
    opt.model, opt.group, opt.chunk, opt._comm_stream = shim, None, 100000, None
    opt.allreduce_grads()
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_dp2_equals_single_process(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, 29731, out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = O.T5_TINY
    sd = O.init_state_dict(cfg, seed=1, emb_std=0.3)
    params = O.unique_params(sd)
    for p in params.values():
        p.requires_grad_(True)
    ids, mask, labels = O.synth_batch(4, 2, 24, 5, cfg.vocab, seed=9)
    loss, _ = O.fidt5_forward(sd, cfg, ids, mask, labels)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in params.values()])
    assert torch.allclose(got, ref, atol=1e-6, rtol=1e-4)


def _sharded_worker(rank, world, port, out):
    """Host logic of the sharded optimizer step (reduce-scatter -> per-piece clip/update -> all-gather of the bf16
    weights -> on-demand all-gather of the fp32 masters) with torch stand-ins for the two CUDA kernels."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lako_b200.optim import FusedAdamW, shard_plan
    n = 4 * 16 * world
    ranges = [(2 * 16 * world, n), (0, 16 * world), (16 * world, 16 * world), (16 * world, 2 * 16 * world)]  # one empty
    g0 = torch.Generator().manual_seed(5)
    w = torch.randn(n, generator=g0)
    grads = [torch.randn(n, generator=g0) for _ in range(world)]
    shim = type("M", (), {})()
    shim._flat, shim._flat_grad, shim._bf16 = w.clone(), grads[rank].clone(), w.to(torch.bfloat16)
    shim._stage_events, shim._stage_ranges = None, ranges
    shim.mark_bf16_fresh = lambda: None
    opt = FusedAdamW.__new__(FusedAdamW)
    opt.model, opt.group, opt._comm_stream, opt.clip = shim, None, None, 1.0
    opt.sumsq = torch.zeros(1)
    opt._plan = shard_plan(ranges, world, rank)
    assert opt._plan is not None and len(opt._plan) == 3
    assert shard_plan([(0, 8 * world + 8)], world, rank) is None or world == 1

    def sumsq_piece(off, cnt):
        opt.sumsq += (shim._flat_grad[off:off + cnt] ** 2).sum()

    def update_piece(off, cnt, lr):  # clipped SGD stands in for the AdamW kernel
        coef = min(1.0, 1.0 / (float(opt.sumsq.sqrt()) + 1e-6))
        shim._flat[off:off + cnt] -= lr * coef * shim._flat_grad[off:off + cnt]
        shim._bf16[off:off + cnt] = shim._flat[off:off + cnt].to(torch.bfloat16)

    opt._sumsq_piece, opt._update_piece = sumsq_piece, update_piece
    opt._step_sharded(0.1)
    bf16_after = shim._bf16.clone()
    stale = shim._flat.clone()
    opt._masters_stale = True
    opt.sync_master_weights()
    if rank == 0:
        torch.save({"bf16": bf16_after, "w": shim._flat, "stale": stale, "norm2": opt.sumsq.clone(), "w0": w,
                    "gsum": sum(grads)}, out)
    dist.destroy_process_group()


def test_dp2_sharded_step_host_logic(tmp_path):
    out = str(tmp_path / "s.pt")
    mp.spawn(_sharded_worker, args=(2, 29733, out), nprocs=2, join=True)
    r = torch.load(out)
    gsum, w0 = r["gsum"], r["w0"]
    assert torch.allclose(r["norm2"], (gsum ** 2).sum().reshape(1), rtol=1e-5)
    coef = min(1.0, 1.0 / (float((gsum ** 2).sum().sqrt()) + 1e-6))
    ref = w0 - 0.1 * coef * gsum
    assert torch.allclose(r["w"], ref, atol=1e-6)                       # masters complete after the sync
    assert torch.equal(r["bf16"], ref.to(torch.bfloat16))               # compute weights identical on every rank
    assert not torch.allclose(r["stale"], ref, atol=1e-6)               # ... while foreign master slices were stale


def _gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lako_b200.optim import gather_answers_and_scores
    answers = torch.full((2, 3 + rank), rank + 1, dtype=torch.long)          # ragged widths: early stop differs per rank
    scores = torch.arange(2 * 4, dtype=torch.float64).reshape(2, 4) + 100 * rank
    a, s = gather_answers_and_scores(answers, scores, max_length=6)
    if rank == 1:
        torch.save((a, s), out)
    dist.destroy_process_group()


def test_inference_gather_of_answers_and_scores(tmp_path):
    """SURVEY 8(e) inference side: ranks decode their own questions; answers (right-padded to max_length with pad id 0)
    and per-fact scores are all-gathered in rank order."""
    out = str(tmp_path / "gather.pt")
    mp.spawn(_gather_worker, args=(2, 29735, out), nprocs=2, join=True)
    a, s = torch.load(out)
    assert a.shape == (4, 6) and s.shape == (4, 4) and s.dtype == torch.float64
    assert a[:2, :3].eq(1).all() and a[:2, 3:].eq(0).all() and a[2:, :4].eq(2).all() and a[2:, 4:].eq(0).all()
    assert torch.equal(s[2:], torch.arange(8, dtype=torch.float64).reshape(2, 4) + 100)
    from lako_b200.optim import gather_answers_and_scores
    a1, s1 = gather_answers_and_scores(torch.ones(2, 3, dtype=torch.long), torch.zeros(2, 4), max_length=5)
    assert a1.shape == (2, 5) and s1.shape == (2, 4)      # single process: padded pass-through
