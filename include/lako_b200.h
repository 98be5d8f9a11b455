/* lako_b200 C ABI -- the drop-in boundary of the B200-native FiDT5 reader path.
 *
 * The reference (hackerchenzhuo/LaKo) is pure Python over PyTorch/transformers and has NO FFI of its
 * own, so there is no existing binding to mirror; every entry point below names the
 * reference computation it replaces (file:line, relative to /root/reference; "HF:" = the
 * un-vendored transformers T5 code the reference inherits, modeling_t5.py of 5.5.0).
 * INTEGRATION.md shows the ctypes stubs a maintainer adds to src/model.py to call these.
 *
 * Conventions: the caller owns every buffer (device pointers, contiguous row-major unless a leading
 * dimension is given); functions never allocate device memory, never synchronise, launch on the
 * stream passed in, and return 0 or a negative error code (lako_last_error() gives the text).
 * sm_100a only: on any other device every call returns LAKO_ERR_ARCH -- there is no fallback path.
 */
#ifndef LAKO_B200_H
#define LAKO_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LAKO_OK 0
#define LAKO_ERR_BAD_ARG (-1)
#define LAKO_ERR_UNSUPPORTED (-2)
#define LAKO_ERR_LAUNCH (-3)
#define LAKO_ERR_DRIVER (-4)
#define LAKO_ERR_ARCH (-5)
#define LAKO_ERR_WORKSPACE (-6)

const char* lako_last_error(void);
int lako_check_device(void); /* 0 iff the current device is sm_100 */

/* ---- dense contractions: every nn.Linear of the T5 blocks and its backward --------------------
 * C[M,N] = alpha * A[M,K] * B[N,K]^T with a fused epilogue, bf16 operands, fp32 accumulation in TMEM.
 * Replaces: q/k/v/o projections HF:277-299,338 ; FFN wi/wo HF:92-103 ; LM head HF:1107-1110 ;
 * cross-attention K/V projection src/model.py:311-312 ; and torch autograd's dgrad/wgrad for them.
 *  a_mn/b_mn: 0 = operand stored row-major [rows][K] (ld >= K), 1 = stored row-major [K][rows].
 *  epi: 0 bf16 store, 1 bf16 ReLU store, 2 fp32 store, 3 fp32 atomic accumulate (split-K allowed),
 *       4 bf16 store masked by aux[m,n] > 0 (ReLU backward).  bn: 0 = auto, 128 or 256. */
int lako_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                   int b_mn, int epi, float alpha, const void* aux, int ld_aux, int splitk, int bn, void* stream);

/* ---- T5LayerNorm (RMSNorm) fused with the residual add: HF:55-68 + HF:375,149 ------------------- */
int lako_rmsnorm_fwd(float* x, const void* delta_bf16, const float* w, void* out_bf16, float* rstd, int rows, int d,
                     float eps, float out_scale, void* stream);
int lako_rmsnorm_bwd(const void* dy_bf16, const float* x, const float* rstd, const float* w, float* dres,
                     void* dres_bf16, float* dw, int rows, int d, int accumulate, float dy_scale, void* stream);

/* ---- embedding HF:682 and its backward; _shift_right HF:595-614 -------------------------------- */
int lako_embed_gather(const int64_t* ids, const float* table, float* out, int rows, int d, int vocab, void* stream);
int lako_embed_scatter_add(const int64_t* ids, const float* dres, float* dtable, int rows, int d, int vocab,
                           void* stream);
int lako_shift_right(const int64_t* labels, int64_t* dec_ids, int B, int T, void* stream);

/* ---- CrossEntropyLoss(ignore_index=-100) HF:1112-1117, forward + backward fused ----------------
 * stats[0] accumulates the sum of row losses, stats[1] holds the valid-row count (lako_count_valid). */
int lako_count_valid(const int64_t* labels, int n, float* stats, void* stream);
int lako_ce_fwd_bwd(const float* logits, const int64_t* labels, float* stats, float* row_loss, void* dlogits_bf16,
                    int rows, int V, int ld, int ld_d, void* stream);

/* ---- utilities --------------------------------------------------------------------------------- */
int lako_cast_f32_to_bf16(const float* in, void* out_bf16, size_t n, void* stream);
int lako_transpose_bf16(const void* in, void* out, int rows, int cols, int ld_in, int ld_out, void* stream);

/* ---- train step tail: clip_grad_norm_ (train_reader.py:76) + HF AdamW(correct_bias=False)
 *      (src/util.py:225) + bf16 re-cast of the weights, over the flat parameter buffer ------------- */
int lako_grad_sumsq(const float* g, size_t n, float* out_scalar, void* stream);
int lako_adamw_step(float* p, float* m, float* v, const float* g, void* w_bf16, size_t n, const float* sumsq,
                    float grad_scale, float clip, float lr, float beta1, float beta2, float eps, float wd,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif
