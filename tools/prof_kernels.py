"""Runs each hot kernel once or twice at the cfg2 (t5-base, B=8, N=10, L=200) shapes so that
`ncu --set full` can capture them individually.  Usage (on the GPU box):
    ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/prof_kernels.py
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lako_b200 import _native as nat  # noqa: E402
from lako_b200.model import BUCKET_CENTER as CENTER, bucket_table  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
lib = nat.lib()
P, L, H, d, ff = (int(os.environ.get(k, v)) for k, v in (('P', 80), ('L', 200), ('H', 12), ('D', 768), ('FF', 3072)))
inner = H * 64
Mt = P * L
torch.manual_seed(0)


def bf(*shape, s=0.5):
    return (torch.randn(*shape, device="cuda") * s).to(torch.bfloat16)


if which in ("all", "attn"):
    qkv = bf(Mt, 3 * inner)
    mask = torch.ones(P, L, dtype=torch.bool, device="cuda")
    table = torch.randn(32, H, device="cuda")
    buckets = bucket_table(True, 32, 128).cuda()
    lut = torch.empty(H, 2 * L - 1, device="cuda")
    nat.check(lib.lako_build_bias_lut(nat.ptr(table), nat.ptr(buckets), nat.ptr(lut), H, L, CENTER, nat.stream_ptr()))
    out = torch.zeros(Mt, inner, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(P, H, L, device="cuda")
    dout = bf(Mt, inner)
    dqkv = torch.zeros_like(qkv)
    dlut = torch.zeros(H, 2 * L - 1, device="cuda")
    for _ in range(2):
        nat.check(lib.lako_enc_attn_fwd(nat.ptr(qkv), nat.ptr(mask), nat.ptr(lut), nat.ptr(out), nat.ptr(lse), P, L, H,
                                        nat.stream_ptr()))
        nat.check(lib.lako_enc_attn_bwd(nat.ptr(qkv), nat.ptr(out), nat.ptr(dout), nat.ptr(lse), nat.ptr(mask),
                                        nat.ptr(lut), nat.ptr(dqkv), nat.ptr(dlut), P, L, H, nat.stream_ptr()))
    torch.cuda.synchronize()

if which in ("all", "norm"):
    x = torch.randn(Mt, d, device="cuda")
    x2 = torch.empty_like(x)
    delta = bf(Mt, d)
    w = torch.ones(d, device="cuda")
    n = torch.empty(Mt, d, device="cuda", dtype=torch.bfloat16)
    rstd = torch.empty(Mt, device="cuda")
    dres = torch.randn(Mt, d, device="cuda")
    dresb = torch.empty(Mt, d, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(d, device="cuda")
    for _ in range(2):
        nat.check(lib.lako_rmsnorm_fwd(nat.ptr(x), nat.ptr(delta), nat.ptr(x2), nat.ptr(w), nat.ptr(n), nat.ptr(rstd),
                                       Mt, d, ctypes.c_float(1e-6), ctypes.c_float(1.0), nat.stream_ptr()))
        nat.check(lib.lako_rmsnorm_bwd(nat.ptr(delta), nat.ptr(x2), nat.ptr(rstd), nat.ptr(w), nat.ptr(dres),
                                       nat.ptr(dresb), nat.ptr(dw), Mt, d, 1, ctypes.c_float(1.0), nat.stream_ptr()))
    torch.cuda.synchronize()

if which in ("all", "gemm"):
    a = bf(Mt, d)
    wq = bf(3 * inner, d)
    c = torch.empty(Mt, 3 * inner, device="cuda", dtype=torch.bfloat16)
    h = bf(Mt, ff)
    wo = bf(d, ff)
    dw = torch.zeros(d, ff, device="cuda")
    dy = bf(Mt, d)
    for _ in range(2):
        nat.gemm(a, wq, c, Mt, 3 * inner, d, d, d, 3 * inner)                                   # fwd qkv
        nat.gemm(h, wo, dy, Mt, d, ff, ff, ff, d)                                               # fwd ffn out
        nat.gemm(dy, wo, h, Mt, ff, d, d, ff, ff, False, True, nat.EPI_BF16_MASK, 1.0, h, ff)   # dgrad (+relu mask)
        nat.gemm(dy, h, dw, d, ff, Mt, d, ff, ff, True, True, nat.EPI_F32_ATOMIC)               # wgrad
    torch.cuda.synchronize()

if which in ("all", "xattn"):
    lib.lako_xattn_workspace_bytes.restype = ctypes.c_size_t
    B, T, M, nl = 8, 10, 2000, 12
    ldkv = nl * 2 * inner
    q = bf(B * T, inner)
    kv = bf(B * M, ldkv)
    mask = torch.ones(B, M, dtype=torch.bool, device="cuda")
    out = torch.zeros(B * T, inner, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device="cuda")
    ws = torch.empty(lib.lako_xattn_workspace_bytes(B, H, T, M), dtype=torch.uint8, device="cuda")
    dkv = torch.zeros_like(kv)
    dq = torch.zeros(B * T, inner, device="cuda")
    dout = bf(B * T, inner)
    for _ in range(2):
        nat.check(lib.lako_xattn_fwd(nat.ptr(q), nat.ptr(kv), nat.ptr(mask), nat.ptr(out), nat.ptr(lse), None,
                                     nat.ptr(ws), B, T, M, H, inner, ldkv, 0, inner, inner, nat.stream_ptr()))
        nat.check(lib.lako_xattn_bwd(nat.ptr(q), nat.ptr(kv), nat.ptr(mask), nat.ptr(out), nat.ptr(dout), nat.ptr(lse),
                                     nat.ptr(dkv), nat.ptr(dq), B, T, M, H, inner, ldkv, 0, inner, inner,
                                     nat.stream_ptr()))
    torch.cuda.synchronize()
print("prof_kernels done")
