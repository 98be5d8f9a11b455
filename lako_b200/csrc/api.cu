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

int lako_rmsnorm_fwd(float* x, const void* delta, const float* w, void* out, float* rstd, int rows, int d, float eps,
                     float out_scale, void* stream) {
  ARCH_GUARD();
  return rmsnorm_fwd(x, CBF(delta), w, BF(out), rstd, rows, d, eps, out_scale, ST(stream));
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

}  // extern "C"
