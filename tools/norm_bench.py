"""RMSNorm fwd/bwd bandwidth at encoder shapes (CUDA events, rotating buffers larger than L2)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lako_b200 import _native as nat
lib = nat.lib()

def run(rows, d, nbuf=4, reps=20):
    xs = [torch.randn(rows, d, device="cuda") for _ in range(nbuf)]
    x2 = [torch.empty(rows, d, device="cuda") for _ in range(nbuf)]
    de = [torch.randn(rows, d, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    n = [torch.empty(rows, d, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    dres = [torch.randn(rows, d, device="cuda") for _ in range(nbuf)]
    dresb = [torch.empty(rows, d, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    w = torch.ones(d, device="cuda"); rstd = torch.empty(rows, device="cuda"); dw = torch.zeros(d, device="cuda")
    def fwd(i):
        nat.check(lib.lako_rmsnorm_fwd(nat.ptr(xs[i]), nat.ptr(de[i]), nat.ptr(x2[i]), nat.ptr(w), nat.ptr(n[i]), nat.ptr(rstd),
                                       rows, d, ctypes.c_float(1e-6), ctypes.c_float(1.0), nat.stream_ptr()))
    def bwd(i):
        nat.check(lib.lako_rmsnorm_bwd(nat.ptr(de[i]), nat.ptr(x2[i]), nat.ptr(rstd), nat.ptr(w), nat.ptr(dres[i]),
                                       nat.ptr(dresb[i]), nat.ptr(dw), rows, d, 1, ctypes.c_float(1.0), nat.stream_ptr()))
    for name, fn, bpe in (("fwd", fwd, 12), ("bwd", bwd, 16)):
        for i in range(nbuf): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps): fn(r % nbuf)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"rmsnorm_{name} rows {rows} d {d}: {ms*1e3:7.1f} us  {rows*d*bpe/ms/1e6:7.0f} GB/s", flush=True)

run(16000, 768); run(40000, 1024)
