"""Diagnostic: repeat the same training step; report which distinct gradient results appear, with PDL on and off."""
import os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_model import _load, GOLDEN
from oracle import fidt5_oracle as O
from lako_b200 import _native as nat
L = int(sys.argv[1]) if len(sys.argv) > 1 else 300
c, cfg, sd, model = _load(GOLDEN[0])
ids, mask, labels = O.synth_batch(2, 2, L, 6, cfg.vocab, seed=L)
ids, mask, labels = ids.cuda(), mask.cuda(), labels.cuda()
model.train()
def step(junk):
    model.zero_grad()
    if junk:   # perturb launch timing / allocator state
        x = torch.randn(junk, device="cuda"); x = x * 2
    loss, logits = model(input_ids=ids, attention_mask=mask, labels=labels)
    loss.backward()
    torch.cuda.synchronize()
    return logits.detach().clone(), model._flat_grad.clone()
for pdl in (1, 0):
    nat.lib().lako_set_pdl(pdl)
    outs = []
    for i in range(int(os.environ.get("ITERS", "12"))):
        lg, g = step((i % 3) * 1000003)
        outs.append((lg, g))
    ref_lg, ref_g = outs[0]
    print("pdl", pdl, "grad rel diffs vs first:", ["%.1e" % (((g - ref_g).norm() / ref_g.norm()).item()) for _, g in outs],
          "logit max diffs:", ["%.1e" % ((lg - ref_lg).abs().max().item()) for lg, _ in outs])
