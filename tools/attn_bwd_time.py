"""Times the encoder self-attention forward and backward at one shape (default: cfg4's per-layer shape P=160 L=250 H=16),
no dropout, explicit-mask-free (packed-row semantics): python tools/attn_bwd_time.py [P L H]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lako_b200 import _native as nat
from lako_b200.model import BUCKET_CENTER as CENTER
from lako_b200.model import bucket_table
lib = nat.lib()
P, L, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (160, 250, 16)))
inner = H * 64; Mt = P * L
torch.manual_seed(0)
qkv = (torch.randn(Mt, 3 * inner, device="cuda") * 0.5).to(torch.bfloat16)
table = torch.randn(32, H, device="cuda"); buckets = bucket_table(True, 32, 128).cuda()
lut = torch.empty(H, 2 * L - 1, device="cuda")
nat.check(lib.lako_build_bias_lut(nat.ptr(table), nat.ptr(buckets), nat.ptr(lut), H, L, CENTER, nat.stream_ptr()))
out = torch.zeros(Mt, inner, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(P, H, L, device="cuda")
dout = (torch.randn(Mt, inner, device="cuda") * 0.1).to(torch.bfloat16)
dqkv = torch.zeros_like(qkv); dlut = torch.zeros(H, 2 * L - 1, device="cuda")
def fwd(n):
    for _ in range(n):
        nat.check(lib.lako_enc_attn_fwd(nat.ptr(qkv), nat.ptr(None), nat.ptr(lut), nat.ptr(out), nat.ptr(lse), P, L, H, nat.stream_ptr()))
def bwd(n):
    for _ in range(n):
        nat.check(lib.lako_enc_attn_bwd(nat.ptr(qkv), nat.ptr(out), nat.ptr(dout), nat.ptr(lse), nat.ptr(None), nat.ptr(lut),
                                        nat.ptr(dqkv), nat.ptr(dlut), P, L, H, nat.stream_ptr()))
for name, f, flop in (("fwd", fwd, 4.0), ("bwd", bwd, 10.0)):
    f(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(10); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"{name}: {us:.1f} us per launch, P={P} L={L} H={H}: {flop * P * H * L * L * 64 / us / 1e6:.0f} TF/s dense")
