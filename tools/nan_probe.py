"""Small-shape training probe (t5-base, B=4, N=4, L=96, T=8; SEED selects the batch): prints three step losses.
Used to reproduce the stale-TMEM NaN (profiles/README.md item 13): small workloads leave some SMs' TMEM untouched by
any GEMM before the first attention kernel runs.  Env: SEED, UNPAD=0 (padded layout), EVAL=1 (forward only)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lako_b200.model import FiDT5, FiDT5Config
from lako_b200.optim import FusedAdamW, train_step
B, N, L, T = 4, 4, 96, 8
seed = int(os.environ.get("SEED", "100"))
g = torch.Generator().manual_seed(seed)
ids = torch.randint(11, 32000, (B, N, L), generator=g)
lens = torch.randint(L // 2, L + 1, (B, N), generator=g)
mask = torch.arange(L)[None, None, :] < lens[:, :, None]
ids = (ids * mask).cuda(); mask = mask.cuda()
labels = torch.randint(11, 32000, (B, T), generator=g).cuda()
model = FiDT5(FiDT5Config.t5("base"), emb_std=0.25, seed=0).cuda()
model.unpad = os.environ.get("UNPAD", "1") == "1"
if os.environ.get("EVAL") == "1":
    model.eval()
    with torch.no_grad():
        loss, logits = model(input_ids=ids, attention_mask=mask, labels=labels)
    print("eval loss", float(loss), "logits finite", bool(torch.isfinite(logits).all()))
else:
    model.train()
    opt = FusedAdamW(model, lr=1e-3, weight_decay=1e-4, clip=1.0)
    out = []
    for _ in range(3):
        out.append(float(train_step(model, opt, ids, mask, labels)))
    print("train losses", out, "grad norm", float(opt.grad_norm()), "lens", lens.min().item(), lens.max().item())
