"""tcgen05 GEMM parity on a B200 (bf16 operands, fp32 accumulate) against torch fp32 matmul of the
same bf16-rounded inputs, for every operand-major combination, epilogue, edge size and split-K."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * 0.5).to(torch.bfloat16)


def _run(M, N, K, a_mn, b_mn, epi, bn=0, splitk=0, alpha=1.0):
    from lako_b200 import _native as nat
    A = _mk((M, K), 1)
    B = _mk((N, K), 2)
    ref = alpha * (A.float() @ B.float().T)
    A_st = A.T.contiguous() if a_mn else A
    B_st = B.T.contiguous() if b_mn else B
    aux = None
    if epi in (nat.EPI_F32, nat.EPI_F32_ATOMIC):
        C = torch.full((M, N), 1.0 if epi == nat.EPI_F32_ATOMIC else 0.0, device="cuda", dtype=torch.float32)
        if epi == nat.EPI_F32_ATOMIC:
            ref = ref + 1.0
    else:
        C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        if epi == nat.EPI_BF16_RELU:
            ref = ref.relu()
        if epi == nat.EPI_BF16_MASK:
            aux = _mk((M, N), 3)
            ref = torch.where(aux.float() > 0, ref, torch.zeros_like(ref))
    nat.gemm(A_st, B_st, C, M, N, K, A_st.stride(0), B_st.stride(0), C.stride(0), a_mn, b_mn, epi, alpha, aux,
             aux.stride(0) if aux is not None else 0, splitk, bn)
    torch.cuda.synchronize()
    tol = 2e-2 * (K ** 0.5) * 0.25 * abs(alpha) + 1e-3
    err = (C.float() - ref).abs().max().item()
    assert err < tol, f"M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn} epi{epi} bn{bn}: max err {err} tol {tol}"
    return err


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("bn", [128, 256, 512])
def test_gemm_majors(a_mn, b_mn, bn):
    from lako_b200 import _native as nat
    _run(256, 512, 256, a_mn, b_mn, nat.EPI_F32, bn=bn)


@pytest.mark.parametrize("bn", [0, 512])
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_gemm_epilogues(epi, bn):
    _run(384, 768, 320, False, False, epi, alpha=0.5, bn=bn)


@pytest.mark.parametrize("M,N,K", [(16000, 2304, 768), (304, 520, 136), (4000, 768, 3072), (264, 264, 64),
                                    (12704, 768, 256), (2504, 2056, 320)])
def test_gemm_pair_kernel_edges(M, N, K):
    """CTA-pair (cta_group::2) kernel: ragged M/N/K, every operand-major combination.  The last two shapes leave
    a short tail round (150 and 90 tiles on 74 clusters), which the kernel cuts into 256 x 128 half tiles."""
    from lako_b200 import _native as nat
    _run(M, N, K, False, False, nat.EPI_BF16, bn=512)
    _run(M, N, K, False, True, nat.EPI_BF16_RELU, bn=512)
    _run(M, N, K, True, False, nat.EPI_F32, bn=512)
    _run(M, N, K, True, True, nat.EPI_F32_ATOMIC, bn=512, splitk=3)


@pytest.mark.parametrize("M,N,K", [(80, 768, 768), (1, 136, 64), (200, 32128 // 8, 128), (130, 264, 72),
                                    (16000, 2304, 768), (1000, 1000, 1000 // 8 * 8)])
def test_gemm_edges(M, N, K):
    from lako_b200 import _native as nat
    _run(M, N, K, False, False, nat.EPI_BF16)
    _run(M, N, K, False, True, nat.EPI_F32)


@pytest.mark.parametrize("splitk", [0, 1, 3, 16])
def test_gemm_wgrad_splitk(splitk):
    from lako_b200 import _native as nat
    # dW[N_w, K_w] = dY^T X : both operands MN-major, deep K (tokens), small output
    _run(768, 768, 8000, True, True, nat.EPI_F32_ATOMIC, splitk=splitk)
    _run(3072, 768, 4096 + 64, True, True, nat.EPI_F32_ATOMIC, splitk=splitk)


def test_gemm_linear_helper_strided():
    from lako_b200 import _native as nat
    x = _mk((512, 3 * 256), 5)
    w = _mk((256, 256), 6)
    xs = x[:, 256:512]          # strided view: lda = 768
    y = nat.linear(xs, w)
    torch.cuda.synchronize()
    ref = xs.float() @ w.float().T
    assert (y.float() - ref).abs().max().item() < 0.1


@pytest.mark.parametrize("bn", [0, 512])
def test_gemm_relu_bitmask_roundtrip(bn):
    """ReLU epilogue writes a 1-bit/element activity mask; the dgrad epilogue (epi 5) consumes it."""
    from lako_b200 import _native as nat
    M, N, K = 520, 512, 192
    A, B = _mk((M, K), 11), _mk((N, K), 12)
    h = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    bits = torch.zeros(M, N // 8, device="cuda", dtype=torch.uint8)
    nat.gemm(A, B, h, M, N, K, K, K, N, False, False, nat.EPI_BF16_RELU, 1.0, bits, N // 8, 0, bn)
    ref = (A.float() @ B.float().T).relu()
    assert (h.float() - ref).abs().max().item() < 0.1
    act = (ref > 0).view(M, N // 8, 8).to(torch.uint8)
    packed = sum(act[..., i] << i for i in range(8)).to(torch.uint8)
    agree = (packed == bits).float().mean().item()
    assert agree > 0.999, agree          # only exact-zero ties may differ
    dY = _mk((M, 256), 13)
    W = _mk((256, N), 14)               # dX[M,N] = dY[M,256] @ W[256,N], masked
    dX = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    nat.gemm(dY, W, dX, M, N, 256, 256, N, N, False, True, nat.EPI_BF16_BITMASK, 0.5, bits, N // 8, 0, bn)
    torch.cuda.synchronize()
    unpacked = torch.stack([(bits >> i) & 1 for i in range(8)], dim=-1).view(M, N).bool()
    ref2 = torch.where(unpacked, 0.5 * (dY.float() @ W.float()), torch.zeros(M, N, device="cuda"))
    assert (dX.float() - ref2).abs().max().item() < 0.1


@pytest.mark.parametrize("nb,M,N,K", [(3, 768, 3072, 80), (12, 2304, 768, 80), (26, 264, 520, 72)])
def test_gemm_wgrad_batched(nb, M, N, K):
    """nb same-shape dW_z += A_z^T B_z problems in one launch of the CTA-pair kernel (decoder weight gradients);
    26 problems exercise the chunking over the 24 descriptors one launch carries."""
    import ctypes
    from lako_b200 import _native as nat
    lib = nat.lib()
    As = [_mk((K, M), 10 + z) for z in range(nb)]   # row-major [K][M]: dy of layer z
    Bs = [_mk((K, N), 50 + z) for z in range(nb)]   # row-major [K][N]: layer input
    Cs = [torch.full((M, N), float(z), device="cuda") for z in range(nb)]
    arr = lambda ts: (ctypes.c_void_p * nb)(*[t.data_ptr() for t in ts])
    nat.check(lib.lako_gemm_bf16_wgrad_batched(arr(As), arr(Bs), arr(Cs), nb, M, N, K, M, N, N, ctypes.c_float(0.5),
                                               nat.stream_ptr()), "wgrad_batched")
    torch.cuda.synchronize()
    for z in range(nb):
        ref = 0.5 * (As[z].float().T @ Bs[z].float()) + float(z)
        err = (Cs[z] - ref).abs().max().item()
        assert err < 2e-2 * (K ** 0.5) * 0.25 + 1e-3, f"problem {z}: max err {err}"


@pytest.mark.parametrize("M,N,K", [(1000, 512, 256), (130, 256, 768), (4030, 1536, 1024)])
def test_gemm_head_major_output(M, N, K):
    """C stored [N/64][M][64] by the pair kernel's 3D TMA-store epilogue == the row-major product, re-laid out."""
    from lako_b200 import _native as nat
    torch.manual_seed(M + N)
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    cap = M + 37    # slab capacity > M: the rows beyond M must stay untouched
    C = torch.full((N // 64, cap, 64), 7.0, device="cuda", dtype=torch.bfloat16)
    nat.check(nat.lib().lako_gemm_bf16_head_major(nat.ptr(A), nat.ptr(B), nat.ptr(C), ctypes.c_int(M), ctypes.c_int(N),
                                                  ctypes.c_int(K), ctypes.c_int(K), ctypes.c_int(K), ctypes.c_int(cap),
                                                  ctypes.c_float(1.0), nat.stream_ptr()))
    torch.cuda.synchronize()
    ref = (A.float() @ B.float().T).view(M, N // 64, 64).transpose(0, 1)
    assert (C[:, :M].float() - ref).abs().max().item() < 2e-2
    assert bool((C[:, M:] == 7.0).all())
