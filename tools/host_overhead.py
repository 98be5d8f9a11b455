"""Host-side enqueue time of one training step (no device synchronisation inside the loop)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lako_b200.model import FiDT5, FiDT5Config
from lako_b200.optim import FusedAdamW, train_step
size, B, N, L, T = (sys.argv[1] if len(sys.argv) > 1 else "large"), 8, 20, 250, 10
if size == "base": N, L = 10, 200
model = FiDT5(FiDT5Config.t5(size), emb_std=0.25, seed=0).cuda().train()
opt = FusedAdamW(model, lr=1e-4, weight_decay=1e-4)
g = torch.Generator().manual_seed(0)
ids = torch.randint(11, 32000, (B, N, L), generator=g)
lens = torch.randint(int(0.6 * L), L + 1, (B, N), generator=g)
mask = torch.arange(L)[None, None, :] < lens[:, :, None]
ids = (ids * mask).cuda(); mask = mask.cuda()
labels = torch.randint(11, 32000, (B, T), generator=g).cuda()
for _ in range(3): train_step(model, opt, ids, mask, labels)
torch.cuda.synchronize()
for tag in ("no-sync", "sync-each-step"):
    t0 = time.perf_counter(); host = 0.0
    for _ in range(5):
        h0 = time.perf_counter()
        l = train_step(model, opt, ids, mask, labels)
        host += time.perf_counter() - h0
        if tag != "no-sync": l.item()
    torch.cuda.synchronize()
    print(f"{size} {tag}: wall {1e3*(time.perf_counter()-t0)/5:.2f} ms/step, host enqueue {1e3*host/5:.2f} ms/step", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): train_step(model, opt, ids, mask, labels)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
