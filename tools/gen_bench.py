"""Greedy generate() + score capture throughput at the evaluation shapes (test_reader.py:36-84): t5-base, B=8, N=10,
L=200, max_length=50.  Prints ms per call, encoder / decode split and tokens/s."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lako_b200.model import FiDT5, FiDT5Config
size = sys.argv[1] if len(sys.argv) > 1 else "base"
B, N, L, ML = 8, 10, 200, 50
model = FiDT5(FiDT5Config.t5(size), emb_std=0.25, seed=0).cuda().eval()
g = torch.Generator().manual_seed(0)
ids = torch.randint(11, 32000, (B, N, L), generator=g)
lens = torch.randint(int(0.6 * L), L + 1, (B, N), generator=g)
mask = torch.arange(L)[None, None, :] < lens[:, :, None]
ids = (ids * mask).cuda(); mask = mask.cuda()
model.overwrite_forward_crossattention(); model.reset_score_storage()
for ml in (2, ML):
    for _ in range(2):
        out = model.generate(input_ids=ids, attention_mask=mask, max_length=ml)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        out = model.generate(input_ids=ids, attention_mask=mask, max_length=ml)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"t5-{size} generate max_length={ml}: {dt * 1e3:.2f} ms/call, produced {out.shape[1]} columns, "
          f"{B * (out.shape[1] - 1) / dt:.0f} tokens/s, {B / dt:.1f} questions/s", flush=True)
