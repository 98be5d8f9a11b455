"""Decoder-row Linears (M = B*T rows, e.g. 80): the 128-row tcgen05 GEMM against the small-M mma.sync kernel that streams
the weight through every SM.  Weights rotate through a > L2-sized pool so every call reads them from HBM as in the step."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lako_b200 import _native as nat
lib = nat.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 80
for name, N, K in (("qkv", 3072, 1024), ("o/xq/xo", 1024, 1024), ("wi", 4096, 1024), ("wo", 1024, 4096), ("base qkv", 2304, 768), ("base wo", 768, 3072)):
    nW = max(2, int(300e6 // (N * K * 2)))
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16) for _ in range(nW)]
    x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    def run_gemm(i):
        nat.gemm(x, Ws[i % nW], y, M, N, K, K, K, N)
    def run_small(i):
        nat.check(lib.lako_smallm_linear(nat.ptr(None), nat.ptr(x), K, nat.ptr(None), ctypes.c_float(0.0), nat.ptr(Ws[i % nW]), nat.ptr(y),
                                         N, M, N, K, 0, ctypes.c_float(1.0), nat.stream_ptr()))
    res = {}
    for nm, f in (("tcgen05 128-row", run_gemm), ("small-M", run_small)):
        for i in range(5): f(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50): f(i)
        e1.record(); torch.cuda.synchronize()
        res[nm] = e0.elapsed_time(e1) / 50 * 1e3
    ref = x.float() @ Ws[49 % nW].float().T
    err = (y.float() - ref).abs().max().item()
    print(f"M={M} {name:9s} N={N} K={K}: " + ", ".join(f"{k} {v:.1f} us" for k, v in res.items()) + f"  (err {err:.3f}; weight {N*K*2/1e6:.1f} MB = {N*K*2/6.5e6:.1f} us at HBM peak)")
