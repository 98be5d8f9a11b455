"""Sharded (reduce-scatter / sharded AdamW / all-gather) vs replicated (all-reduce) data-parallel step on real GPUs:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
Both modes start from the same weights and see the same batches; prints the max weight difference after 3 steps."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, lr_ = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
from lako_b200.model import FiDT5, FiDT5Config  # noqa: E402
from lako_b200.optim import FusedAdamW, train_step  # noqa: E402

B, N, L, T = 4, 4, 96, 8
g = torch.Generator().manual_seed(100 + rank)
ids = torch.randint(11, 32000, (B, N, L), generator=g)
lens = torch.randint(L // 2, L + 1, (B, N), generator=g)
mask = torch.arange(L)[None, None, :] < lens[:, :, None]
ids = (ids * mask).cuda()
mask = mask.cuda()
labels = torch.randint(11, 32000, (B, T), generator=g).cuda()
res = {}
for mode in ("1", "0"):
    os.environ["LAKO_DP_SHARD"] = mode
    model = FiDT5(FiDT5Config.t5("base"), emb_std=0.25, seed=0).cuda().train()
    opt = FusedAdamW(model, lr=1e-3, weight_decay=1e-4, clip=1.0)
    assert (opt._plan is not None) == (mode == "1")
    losses, norms = [], []
    for _ in range(3):
        losses.append(float(train_step(model, opt, ids, mask, labels)))
        norms.append(float(opt.grad_norm()))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}   # triggers the master all-gather
    res[mode] = (losses, model._bf16.clone(), sd, norms)
    del model, opt
la, ba, sa, na = res["1"]
lb, bb, sb, nb = res["0"]
# AdamW without bias correction moves every touched weight by ~3*lr*sign(g) in the first steps, so elements whose
# gradient is at the fp32-atomics noise level may differ by a whole update; compare the bulk, not the maximum
diff = torch.cat([(sa[k].float() - sb[k].float()).abs().reshape(-1) for k in sa])
dw = float(diff.mean())
frac = float((diff > 1e-4).float().mean())
db = float((ba.float() - bb.float()).abs().mean())
# every rank must hold the same weights
chk = ba.float().sum().reshape(1).clone()
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"losses sharded {la} replicated {lb}")
    print(f"grad norms sharded {na} replicated {nb}")
    print(f"mean |dw| fp32 masters {dw:.3e} (fraction > 1e-4: {frac:.2e}), bf16 weights {db:.3e}; "
          f"cross-rank checksum spread {float(hi - lo):.3e}")
    ok = (frac < 5e-2 and abs(na[0] - nb[0]) < 3e-4 * nb[0] and float(hi - lo) == 0.0
          and abs(la[0] - lb[0]) < 1e-5 and all(abs(x - y) < 5e-3 for x, y in zip(la, lb)))  # later steps: AdamW chaos
    print("DP_CHECK", "OK" if ok else "FAIL")
dist.destroy_process_group()
