// Internal C++ declarations of the kernel launchers (one definition each in the .cu files).
#pragma once
#include "common.cuh"

namespace lako {

// gemm.cu
int gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
              int b_mn, int epi, float alpha, const void* aux, int ld_aux, int splitk, int bn, cudaStream_t stream,
              DropSpec drop = DropSpec{}, int c_head_major = 0);
int gemm_bf16_wgrad_batched(const void* const* A, const void* const* B, float* const* C, int nbatch, int M, int N, int K,
                            int lda, int ldb, int ldc, float alpha, cudaStream_t stream);
enum : int { EPI_BF16 = 0, EPI_BF16_RELU = 1, EPI_F32 = 2, EPI_F32_ATOMIC = 3, EPI_BF16_MASK = 4, EPI_BF16_BITMASK = 5 };

// elementwise.cu
// DropSpec{} (thresh == 0) disables a dropout site.  dd: mask on `delta` before the residual add; od: mask on the
// normalised output.  yd: mask on the incoming dy; bd: mask applied to the bf16 copy of the updated dres.
int rmsnorm_fwd(const float* x, const __nv_bfloat16* delta, float* x_out, const float* w, __nv_bfloat16* out,
                float* rstd, int rows, int d, float eps, float out_scale, cudaStream_t s, DropSpec dd = DropSpec{},
                DropSpec od = DropSpec{});
int rmsnorm_bwd(const __nv_bfloat16* dy, const float* x, const float* rstd, const float* w, float* dres,
                __nv_bfloat16* dres_bf16, float* dw, int rows, int d, int accumulate, float dy_scale, cudaStream_t s,
                DropSpec yd = DropSpec{}, DropSpec bd = DropSpec{});
int embed_gather(const long long* ids, const float* table, float* out, int rows, int d, int vocab, cudaStream_t s,
                 DropSpec ds = DropSpec{}, const int* index = nullptr);
int embed_scatter_add(const long long* ids, const float* dres, float* dtable, int rows, int d, int vocab,
                      cudaStream_t s, DropSpec ds = DropSpec{}, const int* index = nullptr);
int build_tok_index(const int* cu, int* tok_index, int P, int L, cudaStream_t s);
int shift_right(const long long* labels, long long* dec_ids, int B, int T, cudaStream_t s);
int count_valid(const long long* labels, int n, float* stats, cudaStream_t s);
int ce_fwd_bwd(const float* logits, const long long* labels, float* stats, float* row_loss, __nv_bfloat16* dlogits,
               int rows, int V, int ld, int ld_d, cudaStream_t s);
int cast_f32_to_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t s, bool rezero = false);
int dropout_mask(unsigned char* keep, size_t n, unsigned long long seed, unsigned step, unsigned site, float p,
                 cudaStream_t s);
int scale_bf16_dev(__nv_bfloat16* x, size_t n, const float* scalar, cudaStream_t s);
int transpose_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int rows, int cols, int ld_in, int ld_out,
                   cudaStream_t s);
int grad_sumsq(const float* g, size_t n, float* out, cudaStream_t s);
int adamw_step(float* p, float* m, float* v, const float* g, __nv_bfloat16* w_bf16, size_t n, const float* sumsq,
               float grad_scale, float clip, float lr, float beta1, float beta2, float eps, float wd, cudaStream_t s);


// enc_attn.cu
int enc_attn_impl();
void set_enc_attn_impl(int v);
int poison_tmem(cudaStream_t s);  // test support
int build_bias_lut(const float* table, const int* buckets, float* lut, int H, int L, int bucket_center, cudaStream_t s);
int fold_dlut(const float* dlut, const int* buckets, float* dtable, int H, int L, int bucket_center, int nb,
              cudaStream_t s);
int enc_attn_fwd(const __nv_bfloat16* qkv, const unsigned char* mask, const float* bias_lut, __nv_bfloat16* out,
                 float* lse, int P, int L, int H, cudaStream_t s, DropSpec drop = DropSpec{}, const int* cu = nullptr,
                 int packed_rows = 0, uint32_t* keepbits = nullptr);
int enc_attn_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                 const unsigned char* mask, const float* bias_lut, __nv_bfloat16* dqkv, float* dlut, int P, int L,
                 int H, cudaStream_t s, DropSpec drop = DropSpec{}, const int* cu = nullptr, int packed_rows = 0,
                 const uint32_t* keepbits = nullptr, bool d_in_out = false);
// enc_attn_long.cu: the same contract for any L <= ENC_ATTN_LONG_MAX_L (64 x 64 tiles, online softmax, fp32 SIMT); enc_attn_fwd /
// enc_attn_bwd route here when L > 256 (or when set_enc_attn_impl(3) forces it, for the tests)
constexpr int ENC_ATTN_LONG_MAX_L = 2048;
unsigned enc_attn_keep_stride(int L);  // Philox counters per query row of the probability dropout: 4 * max(8, ceil(L / 32))
int enc_attn_long_fwd(const __nv_bfloat16* qkv, const unsigned char* mask, const float* bias_lut, __nv_bfloat16* out, float* lse,
                      int P, int L, int H, cudaStream_t s, DropSpec drop, const int* cu);
int enc_attn_long_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                      const unsigned char* mask, const float* bias_lut, __nv_bfloat16* dqkv, float* dlut, int P, int L, int H,
                      cudaStream_t s, DropSpec drop, const int* cu, bool d_in_out);
// D[row][h] = sum_d dO * O written over the first H floats of O's row (H <= 16); enc_attn_bwd(..., d_in_out = true)
int attn_rowdot_inplace(__nv_bfloat16* o, const __nv_bfloat16* dout, int rows, int H, cudaStream_t s);

// dec_attn.cu
// xattn_bwd with dq_part (>= xattn_workspace_bytes) and dq_bf16: the per-chunk dq partials are summed in chunk order into
// the bf16 dq (deterministic); without them dq_f32 (pre-zeroed) is accumulated with atomics
size_t xattn_workspace_bytes(int B, int H, int T, int M);
int xattn_fwd(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, __nv_bfloat16* out, float* lse,
              float* tok_score, float* workspace, int B, int T, int M, int H, int ldq, int ldkv, int koff, int voff,
              int ldo, cudaStream_t s, DropSpec drop = DropSpec{}, const int* cu = nullptr, int npsg = 1,
              const int* tok_index = nullptr);
int xattn_bwd(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, const __nv_bfloat16* out,
              const __nv_bfloat16* dout, const float* lse, __nv_bfloat16* dkv, float* dq_f32, int B, int T, int M,
              int H, int ldq, int ldkv, int koff, int voff, int ldo, cudaStream_t s, DropSpec drop = DropSpec{},
              const int* cu = nullptr, int npsg = 1, float* dq_part = nullptr,
              __nv_bfloat16* dq_bf16 = nullptr);
// xattn_mma.cu
int xattn_fwd_mma(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, float* tok_score,
                  float* workspace, int B, int T, int M, int H, int ldq, int ldkv, int koff, int voff, cudaStream_t s,
                  DropSpec drop = DropSpec{}, const int* cu = nullptr, int npsg = 1, const int* tok_index = nullptr);
int xattn_bwd_mma(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, const __nv_bfloat16* out,
                  const __nv_bfloat16* dout, const float* lse, __nv_bfloat16* dkv, float* dq_f32, int B, int T, int M,
                  int H, int ldq, int ldkv, int koff, int voff, int ldo, cudaStream_t s, DropSpec drop = DropSpec{},
                  const int* cu = nullptr, int npsg = 1, float* dq_part = nullptr);
int dself_fwd(const __nv_bfloat16* qkv, const float* lut, __nv_bfloat16* out, float* probs, int B, int T, int H,
              cudaStream_t s, DropSpec drop = DropSpec{});
int dself_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* probs, __nv_bfloat16* dqkv, float* dlut,
              int B, int T, int H, cudaStream_t s, DropSpec drop = DropSpec{});
int dself_step(const __nv_bfloat16* qkv, __nv_bfloat16* kcache, __nv_bfloat16* vcache, const float* lut,
               __nv_bfloat16* out, int B, int pos, int Tmax, int H, cudaStream_t s, const int* pos_dev = nullptr);
int greedy_pick(const float* logits, int B, int V, int ld, long long* out_ids, int out_ld, int step, int* unfinished,
                int* n_unfinished, long long* next_ids, cudaStream_t s, int ban_eos = 0, int* step_dev = nullptr,
                int* done_dev = nullptr, int min_length = 0);

// decode.cu
int smallm_linear(const float* x_f32, const __nv_bfloat16* x_b16, int ldx, const float* ln_w, float eps,
                  const __nv_bfloat16* W, void* y, int ldy, int M, int N, int K, int epi, float alpha, cudaStream_t s,
                  const void* prefetch = nullptr, size_t prefetch_bytes = 0);
enum : int { SL_BF16 = 0, SL_BF16_RELU = 1, SL_F32 = 2, SL_F32_ADD = 3 };
size_t xattn_decode_workspace_bytes(int B, int H, int M);
int xattn_decode(const float* q, const __nv_bfloat16* kv, size_t kv_rows, const unsigned char* mask, __nv_bfloat16* out,
                 float* tok_score, const int* step, float* workspace, int B, int M, int H, int ldq, int ldkv, int koff,
                 int voff, int ldo, cudaStream_t s, const int* cu = nullptr, int npsg = 1,
                 const int* tok_index = nullptr, bool zero_q = false, bool head_major = false);
bool prof_enabled();

// fp32 evaluation mode: gemm_tf32.cu, attn_f32.cu, dec_attn.cu
int split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t s);
int gemm_tf32x3(const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo, float* C, float* C_hi,
                float* C_lo, int M, int N, int K, int lda, int ldb, int ldc, int relu, float alpha, cudaStream_t s);
int enc_attn_f32(const float* qkv, const unsigned char* mask, const float* lut, float* out, float* out_hi, float* out_lo,
                 int P, int L, int H, cudaStream_t s, const int* cu = nullptr);
size_t xattn_f32_workspace_bytes(int B, int T, int H, int M);
int xattn_f32(const float* q, const float* kv, const unsigned char* mask, float* out, float* out_hi, float* out_lo,
              float* tok_score, const int* step, float* workspace, int B, int T, int M, int H, int ldq, int ldkv, int koff,
              int voff, int ldo, cudaStream_t s, const int* cu = nullptr, int npsg = 1, const int* tok_index = nullptr);
int rmsnorm_f32(const float* x, const float* delta, float* x_out, const float* w, float* out, float* out_hi, float* out_lo,
                int rows, int d, float eps, cudaStream_t s);
int dself_fwd_f32(const float* qkv, const float* lut, float* out_hi, float* out_lo, int B, int T, int H, cudaStream_t s);
int dself_step_f32(const float* qkv, float* kcache, float* vcache, const float* lut, float* out_hi, float* out_lo, int B,
                   int pos, int Tmax, int H, cudaStream_t s, const int* pos_dev = nullptr);

}  // namespace lako
