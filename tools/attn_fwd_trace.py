"""Phase stamps (SM clocks, CTA 0) of the persistent encoder-attention forward (v2); needs the -DLAKO_TRACE build:
    tools/build_trace.sh; LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_trace.so python tools/attn_fwd_trace.py [P L H]
"""
import ctypes, os, sys
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
mask = None if os.environ.get("NOMASK", "1") == "1" else torch.ones(P, L, dtype=torch.bool, device="cuda")  # the model path (packed rows) passes no mask
table = torch.randn(32, H, device="cuda"); buckets = bucket_table(True, 32, 128).cuda()
lut = torch.empty(H, 2 * L - 1, device="cuda")
nat.check(lib.lako_build_bias_lut(nat.ptr(table), nat.ptr(buckets), nat.ptr(lut), H, L, CENTER, nat.stream_ptr()))
out = torch.zeros(Mt, inner, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(P, H, L, device="cuda")
def run(n):
    for _ in range(n):
        nat.check(lib.lako_enc_attn_fwd(nat.ptr(qkv), nat.ptr(mask), nat.ptr(lut), nat.ptr(out), nat.ptr(lse), P, L, H, nat.stream_ptr()))
for impl in (1, 2):
    lib.lako_set_attn_impl(impl)
    run(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(10); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"impl {impl}: {us:.1f} us per launch (no dropout), P={P} L={L} H={H}: {4.0 * P * H * L * L * 64 / us / 1e6:.0f} TF/s dense")
buf = (ctypes.c_longlong * 256)()
assert lib.lako_debug_a2_trace(buf) == 0
t = list(buf)
t0 = min(x for x in t if x > 0)
rel = lambda x: (x - t0) if x > 0 else -1
print("producer (unit: reach, stage free, staged):")
for ui in range(6):
    print("  unit", ui, [rel(t[ui * 3 + j]) for j in range(3)])
print("MMA (tile n: reach QK, slot free, -, PV issue):")
for n in range(8):
    print("  tile", n, [rel(t[64 + n * 4 + j]) for j in range(4)])
for g in (0, 1):
    print(f"softmax group {g} (k-th tile: reach, S ready, pass1 done + turn, P ready, O ready, stored):")
    for k in range(6):
        print("  ", k, [rel(t[(2 + g) * 64 + k * 6 + j]) for j in range(6)])
