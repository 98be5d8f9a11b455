"""Per-kernel timeline of the real training step (PDL on, no events between kernels) through
torch.profiler/CUPTI.  Prints, per kernel name: launches/step, total us/step, avg us; then the GPU
busy time, the idle gaps, and the per-phase split (encoder fwd / decoder fwd / decoder bwd / encoder
bwd / optimizer) found from marker kernels.  Usage (GPU box):
    python tools/timeline.py [cfg2|cfg4] [out.json]
"""
import json
import os
import re
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lako_b200.model import FiDT5, FiDT5Config  # noqa: E402
from lako_b200.optim import FusedAdamW, train_step  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
out = sys.argv[2] if len(sys.argv) > 2 else None
size, B, N, L, T = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
model = FiDT5(FiDT5Config.t5(size), emb_std=bench.EMB_STD, seed=0).cuda()
model.train()
opt = FusedAdamW(model, lr=1e-4, weight_decay=1e-4, clip=1.0)
g = torch.Generator().manual_seed(1234)
ids = torch.randint(11, 32000, (B, N, L), generator=g)
lens = torch.randint(int(0.6 * L), L + 1, (B, N), generator=g)
mask = torch.arange(L)[None, None, :] < lens[:, :, None]
ids = ids * mask
labels = torch.randint(11, 32000, (B, T), generator=g)
ids, mask, labels = ids.to(dev), mask.to(dev), labels.to(dev)
for _ in range(4):
    train_step(model, opt, ids, mask, labels)
torch.cuda.synchronize()

nsteps = 3
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(nsteps):
        train_step(model, opt, ids, mask, labels)
    torch.cuda.synchronize()

trace_path = "/tmp/lako_trace.json"
prof.export_chrome_trace(trace_path)
evs = []
for e in json.load(open(trace_path))["traceEvents"]:
    if e.get("cat") == "kernel":
        grid = e.get("args", {}).get("grid", [0])
        evs.append((float(e["ts"]), float(e["ts"]) + float(e["dur"]), f'{e["name"]}@@g{grid[0]}x{grid[1] if len(grid) > 1 else 1}'))
evs.sort()


def short(n):
    n, _, grid = n.partition("@@")
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    n = n.replace("lako::", "")
    if "gemm" not in n:
        grid = ""
    return (n[:58] + " " + grid).strip()


agg = defaultdict(lambda: [0, 0.0])
for s, e, n in evs:
    a = agg[short(n)]
    a[0] += 1
    a[1] += e - s
t0, t1 = evs[0][0], max(e for _, e, _ in evs)
span = (t1 - t0) / nsteps
# union of busy intervals
busy, cur_s, cur_e = 0.0, None, None
gaps = []
for s, e, n in evs:
    if cur_e is None:
        cur_s, cur_e = s, e
    elif s <= cur_e:
        cur_e = max(cur_e, e)
    else:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, n))
        cur_s, cur_e = s, e
busy += cur_e - cur_s
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
print(f"workload {wl}: span/step {span:.1f} us, busy/step {busy / nsteps:.1f} us, "
      f"sum-of-kernels/step {sum(v[1] for v in agg.values()) / nsteps:.1f} us, kernels/step {len(evs) / nsteps:.0f}")
print(f"{'kernel':72s} {'n/step':>7s} {'us/step':>9s} {'avg us':>8s} {'share':>6s}")
for n, (c, t) in rows[:45]:
    print(f"{n:72s} {c / nsteps:7.1f} {t / nsteps:9.1f} {t / c:8.2f} {t / nsteps / span:6.3f}")
gaps.sort(reverse=True)
print("largest idle gaps (us, next kernel):")
for gdur, n in gaps[:12]:
    print(f"  {gdur:8.1f}  {short(n)}")
print(f"total idle/step {sum(x for x, _ in gaps) / nsteps:.1f} us in {len(gaps) / nsteps:.0f} gaps")
if out:
    json.dump({"workload": wl, "span_us_per_step": span, "busy_us_per_step": busy / nsteps,
               "kernels": {n: {"per_step": c / nsteps, "us_per_step": t / nsteps} for n, (c, t) in rows}},
              open(out, "w"), indent=1)
# sequential dump of the middle step (TSV: start_us, dur_us, gap_before_us, kernel) for offline reading
if out:
    n_per = len(evs) // nsteps
    seq = evs[n_per:2 * n_per]
    with open(out.replace(".json", ".tsv"), "w") as f:
        prev_end = seq[0][0]
        for s, e, n in seq:
            f.write(f"{s - seq[0][0]:.1f}\t{e - s:.1f}\t{s - prev_end:.1f}\t{short(n)}\n")
            prev_end = max(prev_end, e)
