// extern "C" boundary (include/lako_b200.h): thin, no torch types, no allocation, no sync.
#include "../../include/lako_b200.h"

#include "ops.cuh"

using namespace lako;
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define LL(p) reinterpret_cast<const long long*>(p)
#define ARCH_GUARD()         \
  do {                       \
    int rc__ = check_arch(); \
    if (rc__) return rc__;   \
  } while (0)

extern "C" {

int lako_rmsnorm_fwd(const float* x, const void* delta, float* x_out, const float* w, void* out, float* rstd, int rows,
                     int d, float eps, float out_scale, void* stream) {
  ARCH_GUARD();
  return rmsnorm_fwd(x, CBF(delta), x_out, w, BF(out), rstd, rows, d, eps, out_scale, ST(stream));
}
int lako_rmsnorm_bwd(const void* dy, const float* x, const float* rstd, const float* w, float* dres, void* dres_bf16,
                     float* dw, int rows, int d, int accumulate, float dy_scale, void* stream) {
  ARCH_GUARD();
  return rmsnorm_bwd(CBF(dy), x, rstd, w, dres, BF(dres_bf16), dw, rows, d, accumulate, dy_scale, ST(stream));
}
int lako_embed_gather(const int64_t* ids, const float* table, float* out, int rows, int d, int vocab, void* stream) {
  ARCH_GUARD();
  return embed_gather(LL(ids), table, out, rows, d, vocab, ST(stream));
}
int lako_embed_scatter_add(const int64_t* ids, const float* dres, float* dtable, int rows, int d, int vocab,
                           void* stream) {
  ARCH_GUARD();
  return embed_scatter_add(LL(ids), dres, dtable, rows, d, vocab, ST(stream));
}
int lako_shift_right(const int64_t* labels, int64_t* dec_ids, int B, int T, void* stream) {
  ARCH_GUARD();
  return shift_right(LL(labels), reinterpret_cast<long long*>(dec_ids), B, T, ST(stream));
}
int lako_count_valid(const int64_t* labels, int n, float* stats, void* stream) {
  ARCH_GUARD();
  return count_valid(LL(labels), n, stats, ST(stream));
}
int lako_ce_fwd_bwd(const float* logits, const int64_t* labels, float* stats, float* row_loss, void* dlogits, int rows,
                    int V, int ld, int ld_d, void* stream) {
  ARCH_GUARD();
  return ce_fwd_bwd(logits, LL(labels), stats, row_loss, BF(dlogits), rows, V, ld, ld_d, ST(stream));
}
int lako_cast_f32_to_bf16(const float* in, void* out, size_t n, void* stream) {
  ARCH_GUARD();
  return cast_f32_to_bf16(in, BF(out), n, ST(stream));
}
int lako_transpose_bf16(const void* in, void* out, int rows, int cols, int ld_in, int ld_out, void* stream) {
  ARCH_GUARD();
  return transpose_bf16(CBF(in), BF(out), rows, cols, ld_in, ld_out, ST(stream));
}
int lako_dropout_mask(unsigned char* keep, size_t n, unsigned long long seed, unsigned int step, unsigned int site,
                      float p, void* stream) {
  ARCH_GUARD();
  return dropout_mask(keep, n, seed, step, site, p, ST(stream));
}
int lako_grad_sumsq(const float* g, size_t n, float* out, void* stream) {
  ARCH_GUARD();
  return grad_sumsq(g, n, out, ST(stream));
}
int lako_adamw_step(float* p, float* m, float* v, const float* g, void* w_bf16, size_t n, const float* sumsq,
                    float grad_scale, float clip, float lr, float beta1, float beta2, float eps, float wd,
                    void* stream) {
  ARCH_GUARD();
  return adamw_step(p, m, v, g, BF(w_bf16), n, sumsq, grad_scale, clip, lr, beta1, beta2, eps, wd, ST(stream));
}

int lako_build_bias_lut(const float* table, const int* buckets, float* lut, int H, int L, int bucket_center,
                        void* stream) {
  ARCH_GUARD();
  return build_bias_lut(table, buckets, lut, H, L, bucket_center, ST(stream));
}
int lako_fold_dlut(const float* dlut, const int* buckets, float* dtable, int H, int L, int bucket_center, int nb,
                   void* stream) {
  ARCH_GUARD();
  return fold_dlut(dlut, buckets, dtable, H, L, bucket_center, nb, ST(stream));
}
int lako_debug_poison_tmem(void* stream) { return poison_tmem(ST(stream)); }
void lako_set_attn_impl(int v) { set_enc_attn_impl(v); }
int lako_enc_attn_fwd(const void* qkv, const unsigned char* mask, const float* bias_lut, void* out, float* lse, int P,
                      int L, int H, void* stream) {
  ARCH_GUARD();
  return enc_attn_fwd(CBF(qkv), mask, bias_lut, BF(out), lse, P, L, H, ST(stream));
}
int lako_enc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const unsigned char* mask,
                      const float* bias_lut, void* dqkv, float* dlut, int P, int L, int H, void* stream) {
  ARCH_GUARD();
  return enc_attn_bwd(CBF(qkv), CBF(out), CBF(dout), lse, mask, bias_lut, BF(dqkv), dlut, P, L, H, ST(stream));
}
int lako_enc_attn_fwd_dropout(const void* qkv, const unsigned char* mask, const float* bias_lut, void* out, float* lse,
                              unsigned int* keepbits, int P, int L, int H, unsigned long long seed, unsigned int step,
                              unsigned int site, float p, void* stream) {
  ARCH_GUARD();
  return enc_attn_fwd(CBF(qkv), mask, bias_lut, BF(out), lse, P, L, H, ST(stream), make_drop(seed, step, site, p), nullptr, 0,
                      keepbits);
}
int lako_enc_attn_bwd_dropout(const void* qkv, const void* out, const void* dout, const float* lse,
                              const unsigned char* mask, const float* bias_lut, void* dqkv, float* dlut,
                              const unsigned int* keepbits, int P, int L, int H, unsigned long long seed,
                              unsigned int step, unsigned int site, float p, void* stream) {
  ARCH_GUARD();
  return enc_attn_bwd(CBF(qkv), CBF(out), CBF(dout), lse, mask, bias_lut, BF(dqkv), dlut, P, L, H, ST(stream),
                      make_drop(seed, step, site, p), nullptr, 0, keepbits);
}
size_t lako_xattn_workspace_bytes(int B, int H, int T, int M) { return xattn_workspace_bytes(B, H, T, M); }
int lako_xattn_fwd(const void* q, const void* kv, const unsigned char* mask, void* out, float* lse, float* tok_score,
                   float* workspace, int B, int T, int M, int H, int ldq, int ldkv, int koff, int voff, int ldo,
                   void* stream) {
  ARCH_GUARD();
  return xattn_fwd(CBF(q), CBF(kv), mask, BF(out), lse, tok_score, workspace, B, T, M, H, ldq, ldkv, koff, voff, ldo,
                   ST(stream));
}
int lako_xattn_bwd(const void* q, const void* kv, const unsigned char* mask, const void* out, const void* dout,
                   const float* lse, void* dkv, float* dq_f32, int B, int T, int M, int H, int ldq, int ldkv, int koff,
                   int voff, int ldo, void* stream) {
  ARCH_GUARD();
  return xattn_bwd(CBF(q), CBF(kv), mask, CBF(out), CBF(dout), lse, BF(dkv), dq_f32, B, T, M, H, ldq, ldkv, koff, voff,
                   ldo, ST(stream));
}
int lako_dself_fwd(const void* qkv, const float* lut, void* out, float* probs, int B, int T, int H, void* stream) {
  ARCH_GUARD();
  return dself_fwd(CBF(qkv), lut, BF(out), probs, B, T, H, ST(stream));
}
int lako_dself_bwd(const void* qkv, const void* dout, const float* probs, void* dqkv, float* dlut, int B, int T, int H,
                   void* stream) {
  ARCH_GUARD();
  return dself_bwd(CBF(qkv), CBF(dout), probs, BF(dqkv), dlut, B, T, H, ST(stream));
}

int lako_smallm_linear(const float* x_f32, const void* x_bf16, int ldx, const float* ln_w, float eps, const void* W,
                       void* y, int ldy, int M, int N, int K, int epi, float alpha, void* stream) {
  ARCH_GUARD();
  return smallm_linear(x_f32, CBF(x_bf16), ldx, ln_w, eps, CBF(W), y, ldy, M, N, K, epi, alpha, ST(stream));
}
size_t lako_xattn_decode_workspace_bytes(int B, int H, int M) { return xattn_decode_workspace_bytes(B, H, M); }
int lako_xattn_decode(const float* q, const void* kv, size_t kv_rows, const unsigned char* mask, void* out,
                      float* tok_score, const int* step, float* workspace, int B, int M, int H, int ldq, int ldkv, int koff,
                      int voff, int ldo, const int* cu, int npsg, const int* tok_index, int zero_q, int head_major,
                      void* stream) {
  ARCH_GUARD();
  return xattn_decode(q, CBF(kv), kv_rows, mask, BF(out), tok_score, step, workspace, B, M, H, ldq, ldkv, koff, voff, ldo,
                      ST(stream), cu, npsg, tok_index, zero_q != 0, head_major != 0);
}
int lako_dself_step(const void* qkv, void* kcache, void* vcache, const float* lut, void* out, int B, int pos, int Tmax,
                    int H, const int* pos_dev, void* stream) {
  ARCH_GUARD();
  return dself_step(CBF(qkv), BF(kcache), BF(vcache), lut, BF(out), B, pos, Tmax, H, ST(stream), pos_dev);
}

int lako_enc_attn_f32(const float* qkv, const unsigned char* mask, const float* lut, float* out, float* out_hi,
                      float* out_lo, int P, int L, int H, const int* cu, void* stream) {
  ARCH_GUARD();
  return enc_attn_f32(qkv, mask, lut, out, out_hi, out_lo, P, L, H, ST(stream), cu);
}
size_t lako_xattn_f32_workspace_bytes(int B, int T, int H, int M) { return xattn_f32_workspace_bytes(B, T, H, M); }
int lako_xattn_f32(const float* q, const float* kv, const unsigned char* mask, float* out, float* out_hi, float* out_lo,
                   float* tok_score, const int* step, float* workspace, int B, int T, int M, int H, int ldq, int ldkv,
                   int koff, int voff, int ldo, const int* cu, int npsg, const int* tok_index, void* stream) {
  ARCH_GUARD();
  return xattn_f32(q, kv, mask, out, out_hi, out_lo, tok_score, step, workspace, B, T, M, H, ldq, ldkv, koff, voff, ldo,
                   ST(stream), cu, npsg, tok_index);
}
int lako_rmsnorm_f32(const float* x, const float* delta, float* x_out, const float* w, float* out, float* out_hi,
                     float* out_lo, int rows, int d, float eps, void* stream) {
  ARCH_GUARD();
  return rmsnorm_f32(x, delta, x_out, w, out, out_hi, out_lo, rows, d, eps, ST(stream));
}

}  // extern "C"
