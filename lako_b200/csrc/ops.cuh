// Internal C++ declarations of the kernel launchers (one definition each in the .cu files).
#pragma once
#include "common.cuh"

namespace lako {

// gemm.cu
int gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
              int b_mn, int epi, float alpha, const void* aux, int ld_aux, int splitk, int bn, cudaStream_t stream);
enum : int { EPI_BF16 = 0, EPI_BF16_RELU = 1, EPI_F32 = 2, EPI_F32_ATOMIC = 3, EPI_BF16_MASK = 4 };

// elementwise.cu
int rmsnorm_fwd(float* x, const __nv_bfloat16* delta, const float* w, __nv_bfloat16* out, float* rstd, int rows, int d,
                float eps, float out_scale, cudaStream_t s);
int rmsnorm_bwd(const __nv_bfloat16* dy, const float* x, const float* rstd, const float* w, float* dres,
                __nv_bfloat16* dres_bf16, float* dw, int rows, int d, int accumulate, float dy_scale, cudaStream_t s);
int embed_gather(const long long* ids, const float* table, float* out, int rows, int d, int vocab, cudaStream_t s);
int embed_scatter_add(const long long* ids, const float* dres, float* dtable, int rows, int d, int vocab,
                      cudaStream_t s);
int shift_right(const long long* labels, long long* dec_ids, int B, int T, cudaStream_t s);
int count_valid(const long long* labels, int n, float* stats, cudaStream_t s);
int ce_fwd_bwd(const float* logits, const long long* labels, float* stats, float* row_loss, __nv_bfloat16* dlogits,
               int rows, int V, int ld, int ld_d, cudaStream_t s);
int cast_f32_to_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t s);
int transpose_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int rows, int cols, int ld_in, int ld_out,
                   cudaStream_t s);
int grad_sumsq(const float* g, size_t n, float* out, cudaStream_t s);
int adamw_step(float* p, float* m, float* v, const float* g, __nv_bfloat16* w_bf16, size_t n, const float* sumsq,
               float grad_scale, float clip, float lr, float beta1, float beta2, float eps, float wd, cudaStream_t s);

}  // namespace lako
