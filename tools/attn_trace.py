"""Phase stamps (SM clocks) of one enc_attn_bwd CTA; needs the -DLAKO_TRACE build:
    tools/build_trace.sh; LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_trace.so python tools/attn_trace.py [P L H]
"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lako_b200 import _native as nat
from lako_b200.model import BUCKET_CENTER as CENTER
from lako_b200.model import bucket_table
lib = nat.lib()
P, L, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (80, 200, 12)))
inner = H * 64; Mt = P * L
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).to(torch.bfloat16)
qkv = bf(Mt, 3 * inner); mask = torch.ones(P, L, dtype=torch.bool, device="cuda")
table = torch.randn(32, H, device="cuda"); buckets = bucket_table(True, 32, 128).cuda()
lut = torch.empty(H, 2 * L - 1, device="cuda")
nat.check(lib.lako_build_bias_lut(nat.ptr(table), nat.ptr(buckets), nat.ptr(lut), H, L, CENTER, nat.stream_ptr()))
out = torch.zeros(Mt, inner, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(P, H, L, device="cuda")
dout = bf(Mt, inner); dqkv = torch.zeros_like(qkv); dlut = torch.zeros(H, 2 * L - 1, device="cuda")
for _ in range(3):
    nat.check(lib.lako_enc_attn_fwd(nat.ptr(qkv), nat.ptr(mask), nat.ptr(lut), nat.ptr(out), nat.ptr(lse), P, L, H, nat.stream_ptr()))
    nat.check(lib.lako_enc_attn_bwd(nat.ptr(qkv), nat.ptr(out), nat.ptr(dout), nat.ptr(lse), nat.ptr(mask), nat.ptr(lut),
                                    nat.ptr(dqkv), nat.ptr(dlut), P, L, H, nat.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    nat.check(lib.lako_enc_attn_bwd(nat.ptr(qkv), nat.ptr(out), nat.ptr(dout), nat.ptr(lse), nat.ptr(mask), nat.ptr(lut),
                                    nat.ptr(dqkv), nat.ptr(dlut), P, L, H, nat.stream_ptr()))
e1.record(); torch.cuda.synchronize()
print(f"bwd {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch, P={P} L={L} H={H}, CTAs={P * H}")
buf = (ctypes.c_longlong * 64)()
assert lib.lako_debug_eb_trace(buf) == 0
t = list(buf); t0 = t[0]
names = {0: "entry", 1: "after pdl_wait/alloc", 2: "prologue sync done", 3: "operands landed", 4: "blocks done", 5: "dQ epilogue done", 6: "exit"}
for i in (0, 1, 2, 3):
    print(f"{names[i]:28s} {t[i] - t0:8d}")
nt = (L + 127) // 128
sub = ["S/dP ready", "softmax math done", "P/dS stored", "syncthreads", "MMAs issued", "block end (dK/dV epi)"]
for blk in range(nt * nt):
    for j in range(6):
        print(f"  blk{blk} {sub[j]:24s} {t[8 + blk * 6 + j] - t0:8d}")
for i in (4, 5, 6):
    print(f"{names[i]:28s} {t[i] - t0:8d}")
