"""GEMM shape sweep at the cfg2 (t5-base) / cfg4 (t5-large) shapes: 1-CTA vs CTA-pair kernel vs cuBLAS."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lako_b200 import _native as nat  # noqa: E402


def bench(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bf(*shape):
    return (torch.randn(*shape, device="cuda") * 0.3).to(torch.bfloat16)


def run(tag, M, N, K, a_mn, b_mn, epi):
    A = bf(K, M) if a_mn else bf(M, K)
    B = bf(K, N) if b_mn else bf(N, K)
    out_f32 = epi in (nat.EPI_F32, nat.EPI_F32_ATOMIC)
    C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    aux = bf(M, N) if epi == nat.EPI_BF16_MASK else None
    res = []
    for bn in (256, 512):
        ms = bench(lambda: nat.gemm(A, B, C, M, N, K, A.stride(0), B.stride(0), N, a_mn, b_mn, epi, 1.0, aux,
                                    N if aux is not None else 0, 0, bn))
        res.append(2.0 * M * N * K / ms / 1e9)
    Am = A.T if a_mn else A
    Bm = B if b_mn else B.T
    ms = bench(lambda: torch.matmul(Am, Bm))
    res.append(2.0 * M * N * K / ms / 1e9)
    print(f"{tag:28s} M{M:6d} N{N:6d} K{K:6d}  1cta {res[0]:7.0f}  pair {res[1]:7.0f}  cublas {res[2]:7.0f} TF/s", flush=True)


# GB_MT_BASE / GB_MT_LARGE override the token counts (e.g. the packed row counts of the bench batches); GB_ONLY=base|large
CASES = {"base": (int(os.environ.get("GB_MT_BASE", 16000)), 768, 768, 3072, 12),
         "large": (int(os.environ.get("GB_MT_LARGE", 40000)), 1024, 1024, 4096, 24)}
for name, (Mt, d, inner, ff, nl) in CASES.items():
    if os.environ.get("GB_ONLY", name) != name:
        continue
    print("==", name, "tokens", Mt)
    run("fwd qkv", Mt, 3 * inner, d, False, False, nat.EPI_BF16)
    run("fwd o", Mt, d, inner, False, False, nat.EPI_BF16)
    run("fwd wi relu", Mt, ff, d, False, False, nat.EPI_BF16_RELU)
    run("fwd wo", Mt, d, ff, False, False, nat.EPI_BF16)
    run("fwd xkv all", Mt, nl * 2 * inner, d, False, False, nat.EPI_BF16)
    run("dgrad qkv", Mt, d, 3 * inner, False, True, nat.EPI_BF16)
    run("dgrad wo (mask)", Mt, ff, d, False, True, nat.EPI_BF16_MASK)
    run("dgrad wi", Mt, d, ff, False, True, nat.EPI_BF16)
    run("dgrad xkv all", Mt, d, nl * 2 * inner, False, True, nat.EPI_BF16)
    run("wgrad qkv", 3 * inner, d, Mt, True, True, nat.EPI_F32_ATOMIC)
    run("wgrad o", d, inner, Mt, True, True, nat.EPI_F32_ATOMIC)
    run("wgrad wi", ff, d, Mt, True, True, nat.EPI_F32_ATOMIC)
    run("wgrad wo", d, ff, Mt, True, True, nat.EPI_F32_ATOMIC)
    run("wgrad xkv all", nl * 2 * inner, d, Mt, True, True, nat.EPI_F32_ATOMIC)
