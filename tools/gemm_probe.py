"""GPU bring-up probe for the tcgen05 GEMM: runs each operand-major combination in a subprocess
(a trapped kernel poisons the CUDA context) and dumps inputs/outputs to gpurun_out/ for offline
analysis of descriptor / layout bugs."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def one(a_mn, b_mn, bn, M=256, N=512, K=256):
    import torch
    from lako_b200 import _native as nat
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    A_st = A.T.contiguous() if a_mn else A
    B_st = B.T.contiguous() if b_mn else B
    C = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    nat.gemm(A_st, B_st, C, M, N, K, A_st.stride(0), B_st.stride(0), N, a_mn, b_mn, nat.EPI_F32, 1.0, None, 0, 0, bn)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = (C - ref).abs().max().item()
    tag = f"a{int(a_mn)}b{int(b_mn)}bn{bn}"
    print(f"PROBE {tag}: max_err {err:.5f} ref_absmax {ref.abs().max().item():.3f}", flush=True)
    if err > 0.05:
        torch.save({"A": A.cpu(), "B": B.cpu(), "C": C.cpu()}, os.path.join(OUT, f"gemm_{tag}.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:
        a, b, bn = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
        one(bool(a), bool(b), bn)
    else:
        for a in (0, 1):
            for b in (0, 1):
                for bn in (128, 256):
                    try:
                        r = subprocess.run([sys.executable, __file__, str(a), str(b), str(bn)], timeout=180,
                                           capture_output=True, text=True)
                        print(r.stdout.strip()[-2000:], flush=True)
                        if r.returncode != 0:
                            print(f"PROBE a{a}b{b}bn{bn}: exit {r.returncode}\n{r.stderr.strip()[-1500:]}", flush=True)
                    except subprocess.TimeoutExpired:
                        print(f"PROBE a{a}b{b}bn{bn}: TIMEOUT", flush=True)
