"""Where does the e2e leg of bench.py spend its time?  Per-step wall clock of the resident and the host-fed loops."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lako_b200.model import FiDT5, FiDT5Config
from lako_b200.optim import FusedAdamW, train_step
import bench

size, B, N, L, T, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
model = FiDT5(FiDT5Config.t5(size), emb_std=0.25, seed=0).cuda().train()
opt = FusedAdamW(model, lr=1e-4, weight_decay=1e-4, clip=1.0)
host = bench.synth_host_batches(B, N, L, T, 0, 4)
dev = torch.device("cuda")
res = [(a.to(dev), b.to(dev), c.to(dev)) for a, b, c in host]
lens = [b.sum(-1, dtype=torch.int32) for _, b, _ in host]
for i in range(3):
    train_step(model, opt, *res[i % 4], passage_lengths=lens[i % 4])
torch.cuda.synchronize()
def run(tag, fn, n=8):
    ts = []
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn(i)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(tag, " ".join(f"{t:.1f}" for t in ts))
run("resident+sync      ", lambda i: train_step(model, opt, *res[i % 4], passage_lengths=lens[i % 4]))
def e2e(i, item=True, inloop_lens=True):
    a, b, c = host[i % 4]
    lh = b.sum(-1, dtype=torch.int32) if inloop_lens else lens[i % 4]
    l_ = train_step(model, opt, a.to(dev, non_blocking=True), b.to(dev, non_blocking=True), c.to(dev, non_blocking=True),
                    passage_lengths=lh)
    if item:
        l_.item()
run("host-fed, item, lens", e2e)
run("host-fed, item      ", lambda i: e2e(i, True, False))
run("host-fed            ", lambda i: e2e(i, False, False))
run("resident+sync again ", lambda i: train_step(model, opt, *res[i % 4], passage_lengths=lens[i % 4]))
t0 = time.perf_counter()
for i in range(8):
    e2e(i)
torch.cuda.synchronize()
print("8 host-fed steps back to back:", (time.perf_counter() - t0) * 1e3 / 8, "ms/step")
t0 = time.perf_counter()
for i in range(8):
    train_step(model, opt, *res[i % 4], passage_lengths=lens[i % 4])
torch.cuda.synchronize()
print("8 resident steps back to back:", (time.perf_counter() - t0) * 1e3 / 8, "ms/step")
