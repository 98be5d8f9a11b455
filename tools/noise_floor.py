"""bf16 noise floor of the ORACLE itself: the fp32 oracle vs the same oracle under torch CPU autocast(bfloat16)
on a golden case.  The CUDA path's deviation from the fp32 reference is reported next to these numbers
(tests/test_gpu_model.py uses them as the yardstick for gradient parity)."""
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fidt5_oracle as O  # noqa: E402

for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.pt"))):
    c = torch.load(path, weights_only=False)
    cfg = O.T5Dims(**c["cfg"])

    def run(autocast):
        sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=c["seed"], emb_std=c["emb_std"]), seed=c["seed"] + 1)
        if c.get("eos_alias") is not None:
            sd["shared.weight"][1] = sd["shared.weight"][c["eos_alias"]] * 1.05
        for p in O.unique_params(sd).values():
            p.requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            loss, logits = O.fidt5_forward(sd, cfg, c["ids"], c["mask"], c["labels"])
        loss.backward()
        return sd, logits.float().detach()

    sd32, lg32 = run(False)
    sd16, lg16 = run(True)
    print(os.path.basename(path), "logits max |d| autocast-bf16 vs fp32:", round((lg16 - lg32).abs().max().item(), 5))
    for k in c["grads"]:
        g32, g16 = sd32[k].grad, sd16[k].grad.float()
        print(f"   {k}: rel L2 {((g16 - g32).norm() / g32.norm()).item():.4f}")
