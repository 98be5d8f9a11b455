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
