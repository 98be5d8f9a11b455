"""Workload for ncu captures of the decode kernels: t5-large, cfg3 shape, a few greedy steps with score capture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lako_b200.model import FiDT5, FiDT5Config
import bench

B, N, L = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (8, 20, 250)))
ml = int(sys.argv[4]) if len(sys.argv) >= 5 else 4
model = FiDT5(FiDT5Config.t5("large"), emb_std=0.25, seed=0).cuda().eval()
model.overwrite_forward_crossattention()
host = bench.synth_host_batches(B, N, L, 4, 0, n_batches=1, fact_format=True)
ids, mask = host[0][0].cuda(), host[0][1].cuda()
for _ in range(2):
    model.reset_score_storage()
    out = model.generate(ids, mask, max_length=ml, min_length=ml)
torch.cuda.synchronize()
print(out.shape)
