"""Diagnostic: one training step on a NaN-poisoned activation workspace vs a zeroed one; lists the parameters whose
gradients differ (a kernel reading workspace bytes that nobody wrote)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_model import _load, GOLDEN
from oracle import fidt5_oracle as O
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
c, cfg, sd, model = _load(GOLDEN[0])
ids, mask, labels = O.synth_batch(2, 2, L, 6, cfg.vocab, seed=L)
model.train()
def run(poison):
    os.environ["LAKO_POISON_WORKSPACE"] = "1" if poison else "0"
    model._ws.clear()
    model.zero_grad()
    loss, logits = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), labels=labels.cuda())
    if not poison:
        model._ws["fwd"][1].zero_()
        loss, logits = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), labels=labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), logits.detach().clone(), model._flat_grad.clone()
l0, lg0, g0 = run(False)
l1, lg1, g1 = run(True)
print("loss", l0, l1, "logits finite", bool(torch.isfinite(lg1).all()), "max dlogit", (lg0 - lg1).abs().max().item())
print("grad finite", bool(torch.isfinite(g1).all()), "rel diff", ((g1 - g0).norm() / g0.norm()).item())
for name, (off, shape) in model._specs.items():
    n = 1
    for s in shape:
        n *= s
    a, b = g0[off:off + n], g1[off:off + n]
    bad = (~torch.isfinite(b)).sum().item()
    e = (a - b).norm().item() if bad == 0 else float("nan")
    if bad or e > 1e-6:
        print(f"  {name}: non-finite {bad}/{n}, |diff| {e:.3e} of {a.norm().item():.3e}")
