"""ctypes binding of liblako_b200.so (include/lako_b200.h).  There is NO fallback: if the library is
missing or a call fails this module raises, it never routes work anywhere else."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LAKO_B200_LIB points at another build of the SAME library (A/B timing of kernel variants on one box)
LIB_PATH = os.environ.get("LAKO_B200_LIB") or os.path.join(_HERE, "_lib", "liblako_b200.so")

_lib = None


class LakoError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LakoError(
                f"{LIB_PATH} not found: build it with `python -m lako_b200.build` "
                "(lako_b200 has no CPU or PyTorch fallback path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.lako_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().lako_last_error()
        raise LakoError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t):
    """device pointer of a tensor (or None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


c_int = ctypes.c_int
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t

EPI_BF16, EPI_BF16_RELU, EPI_F32, EPI_F32_ATOMIC, EPI_BF16_MASK, EPI_BF16_BITMASK = range(6)


def gemm(A, B, C, M, N, K, lda, ldb, ldc, a_mn=False, b_mn=False, epi=EPI_BF16, alpha=1.0, aux=None, ld_aux=0,
         splitk=0, bn=0):
    check(lib().lako_gemm_bf16(ptr(A), ptr(B), ptr(C), c_int(M), c_int(N), c_int(K), c_int(lda), c_int(ldb),
                               c_int(ldc), c_int(int(a_mn)), c_int(int(b_mn)), c_int(epi), c_float(alpha), ptr(aux),
                               c_int(ld_aux), c_int(splitk), c_int(bn), stream_ptr()), "lako_gemm_bf16")


def linear(x, w, relu=False, out_dtype=torch.bfloat16, alpha=1.0):
    """y = alpha * x @ w.T ; x (M,K) bf16, w (N,K) bf16."""
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=x.device, dtype=out_dtype)
    epi = EPI_F32 if out_dtype == torch.float32 else (EPI_BF16_RELU if relu else EPI_BF16)
    gemm(x, w, out, M, N, K, x.stride(0), w.stride(0), out.stride(0), epi=epi, alpha=alpha)
    return out
