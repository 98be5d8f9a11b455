"""How much of the attention / norm / GEMM kernel time is dropout (Philox) work?  Per-category kernel times of one cfg4
training step with dropout 0.1 and with dropout 0."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lako_b200 import _native as nat
from lako_b200.model import FiDT5, FiDT5Config
from lako_b200.optim import FusedAdamW, train_step
import bench

size, B, N, L, T, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
names = ["gemm_fwd_large", "enc_attn_fwd", "enc_attn_bwd", "rmsnorm", "cross_attn", "adamw_gradnorm", "gemm_dgrad_large",
         "gemm_wgrad_large", "gemm_decoder_small"]
host = bench.synth_host_batches(B, N, L, T, 0, 2)
res = [(a.cuda(), b.cuda(), c.cuda()) for a, b, c in host]
lens = [b.sum(-1, dtype=torch.int32) for _, b, _ in host]
lib = nat.lib()
for p in (0.1, 0.0):
    conf = FiDT5Config.t5(size)
    conf.dropout_rate = p
    model = FiDT5(conf, emb_std=0.25, seed=0).cuda().train()
    opt = FusedAdamW(model, lr=1e-4, weight_decay=1e-4, clip=1.0)
    for i in range(3):
        train_step(model, opt, *res[i % 2], passage_lengths=lens[i % 2])
    torch.cuda.synchronize()
    lib.lako_prof_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2):
        train_step(model, opt, *res[i % 2], passage_lengths=lens[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms, wk, cn = (ctypes.c_double * 9)(), (ctypes.c_double * 9)(), (ctypes.c_longlong * 9)()
    lib.lako_prof_collect(ms, wk, cn)
    lib.lako_prof_enable(0)
    print(f"dropout {p}: step {e0.elapsed_time(e1) / 2:.2f} ms; " + ", ".join(f"{n} {ms[i] / 2:.2f}" for i, n in enumerate(names)))
    del model, opt
    torch.cuda.empty_cache()
