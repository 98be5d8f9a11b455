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
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long lako_launch_count(void);
/* programmatic dependent launch of the library's kernels (default on; env LAKO_PDL=0 disables) */
void lako_set_pdl(int on);
/* optional CUDA-event timing per kernel category: 0 large forward GEMMs, 1 enc_attn_fwd, 2 enc_attn_bwd, 3 rmsnorm,
 * 4 cross-attention, 5 other (AdamW, CE, embedding, decoder self-attention), 6 large dgrad GEMMs, 7 large wgrad GEMMs,
 * 8 decoder-side (<= 128 token rows) GEMMs; work[] is FLOPs (0,1,2,4,6,7,8) or bytes (3,5). */
#define LAKO_PROF_NCAT 9
void lako_prof_enable(int on);
int lako_prof_collect(double* ms, double* work, long long* count);
int lako_check_device(void); /* 0 iff the current device is sm_100 */

/* ---- dense contractions: every nn.Linear of the T5 blocks and its backward --------------------
 * C[M,N] = alpha * A[M,K] * B[N,K]^T with a fused epilogue, bf16 operands, fp32 accumulation in TMEM.
 * Replaces: q/k/v/o projections HF:277-299,338 ; FFN wi/wo HF:92-103 ; LM head HF:1107-1110 ;
 * cross-attention K/V projection src/model.py:311-312 ; and torch autograd's dgrad/wgrad for them.
 *  a_mn/b_mn: 0 = operand stored row-major [rows][K] (ld >= K), 1 = stored row-major [K][rows].
 *  epi: 0 bf16 store, 1 bf16 ReLU store, 2 fp32 store, 3 fp32 atomic accumulate (split-K allowed),
 *       4 bf16 store masked by aux[m,n] > 0 (ReLU backward, aux bf16 [M][ld_aux]),
 *       5 bf16 store masked by a 1-bit/element mask (aux uint8 [M][ld_aux bytes], bit n&7 of byte n>>3);
 *       epi 1 with aux != NULL also WRITES that bit mask (output > 0).  bn: 0 auto, 128/256 single-CTA tile
 *       width, 512 = CTA-pair kernel (256x256 per 2-CTA cluster). */
int lako_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                   int b_mn, int epi, float alpha, const void* aux, int ld_aux, int splitk, int bn, void* stream);
/* Same product (A [M][lda], B [N][ldb], both K-major, bf16 store) with C written HEAD-MAJOR: C[n / 64][m][n % 64], one
 * contiguous slab of plane_rows >= M rows per 64-column head (N % 64 == 0; rows M..plane_rows-1 of a slab are not
 * touched), through the CTA-pair kernel's 3D TMA-store epilogue.  Used by
 * lako_fidt5_generate for the cross-attention K/V projection of all decoder layers (src/model.py:311-312), so that every
 * decode step streams contiguous K / V slabs. */
int lako_gemm_bf16_head_major(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb,
                              int plane_rows, float alpha, void* stream);
/* nbatch same-shape weight-gradient products in one launch: C_z[M][ldc] (fp32) += alpha * A_z^T B_z with
   A_z row-major [K][lda >= M], B_z row-major [K][ldb >= N] (host arrays of device pointers).  Replaces the per-layer
   dW accumulation autograd performs for every decoder Linear (HF modeling_t5.py: nn.Linear backward of
   T5Attention q/k/v/o :189-192, T5DenseReluDense wi/wo :83-84) when the per-layer problems are too small to fill
   the chip one at a time. */
int lako_gemm_bf16_wgrad_batched(const void* const* A, const void* const* B, float* const* C, int nbatch, int M, int N,
                                 int K, int lda, int ldb, int ldc, float alpha, void* stream);

/* ---- T5LayerNorm (RMSNorm) fused with the residual add: HF:55-68 + HF:375,149 ------------------- */
int lako_rmsnorm_fwd(const float* x, const void* delta_bf16, float* x_out, const float* w, void* out_bf16, float* rstd,
                     int rows, int d, float eps, float out_scale, void* stream);
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

/* ---- precision = fp32: 3xTF32-split tensor-core GEMM (the reference computes in plain fp32, train_reader.py:67) ----
 * hi[i] = x[i] with the low 13 mantissa bits cleared, lo[i] = the same truncation of x[i] - hi[i]. */
int lako_split_tf32(const float* x, float* hi, float* lo, size_t n, void* stream);
/* C[M][ldc] (and / or its hi, lo split; any of C, C_hi may be NULL) = relu?(alpha * A B^T) with A = A_hi + A_lo [M][lda],
 * B = B_hi + B_lo [N][ldb], all fp32 row-major; three tcgen05.mma.kind::tf32 products per k-step (lo*hi, hi*lo, hi*hi)
 * accumulated in fp32 TMEM. */
int lako_gemm_tf32x3(const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo, float* C, float* C_hi,
                     float* C_lo, int M, int N, int K, int lda, int ldb, int ldc, int relu, float alpha, void* stream);

/* fp32-mode attention / normalisation (SIMT fp32; outputs optionally as the (hi, lo) TF32 pair the next GEMM reads):
 * encoder self-attention HF:253-344 per passage for any L (qkv fp32 [rows][3*H*64], cu = packed row offsets or NULL),
 * cross_attention_forward src/model.py:286-349 for T query rows per question against fp32 K/V (arguments as
 * lako_xattn_decode; the capture applies to query row 0), T5LayerNorm + residual add HF:55-68. */
int lako_enc_attn_f32(const float* qkv, const unsigned char* mask, const float* lut, float* out, float* out_hi,
                      float* out_lo, int P, int L, int H, const int* cu, void* stream);
size_t lako_xattn_f32_workspace_bytes(int B, int T, int H, int M);
int lako_xattn_f32(const float* q, const float* kv, const unsigned char* mask, float* out, float* out_hi, float* out_lo,
                   float* tok_score, const int* step, float* workspace, int B, int T, int M, int H, int ldq, int ldkv,
                   int koff, int voff, int ldo, const int* cu, int npsg, const int* tok_index, void* stream);
int lako_rmsnorm_f32(const float* x, const float* delta, float* x_out, const float* w, float* out, float* out_hi,
                     float* out_lo, int rows, int d, float eps, void* stream);

/* ---- utilities --------------------------------------------------------------------------------- */
int lako_cast_f32_to_bf16(const float* in, void* out_bf16, size_t n, void* stream);
int lako_transpose_bf16(const void* in, void* out, int rows, int cols, int ld_in, int ld_out, void* stream);

/* keep[e] (0/1) of element e of dropout site `site` at (seed, step): the exact Philox4x32-7 mask the kernels
 * apply in forward and regenerate in backward (lako_b200/philox.py restates it on the host) */
int lako_dropout_mask(unsigned char* keep, size_t n, unsigned long long seed, unsigned int step, unsigned int site,
                      float p, void* stream);

/* ---- train step tail: clip_grad_norm_ (train_reader.py:76) + HF AdamW(correct_bias=False)
 *      (src/util.py:225) + bf16 re-cast of the weights, over the flat parameter buffer ------------- */
int lako_grad_sumsq(const float* g, size_t n, float* out_scalar, void* stream);
int lako_adamw_step(float* p, float* m, float* v, const float* g, void* w_bf16, size_t n, const float* sumsq,
                    float grad_scale, float clip, float lr, float beta1, float beta2, float eps, float wd,
                    void* stream);


/* ==== model-level entry points: the FiDT5 reader itself =========================================
 * One flat parameter buffer holds every weight (fp32 master `w_f32`, bf16 compute copy `w_bf16`,
 * fp32 gradients `g_f32`, all with the element offsets lako_fidt5_layout() reports); q,k,v of one
 * attention are adjacent (one fused [3*inner][d] projection) and the cross-attention k,v of ALL
 * decoder layers are adjacent (one fused [n_layers*2*inner][d] projection of the encoder memory). */
#define LAKO_MAX_LAYERS 48

typedef struct {
  int d_model, n_heads, d_ff, n_layers, vocab, n_buckets; /* d_kv is 64 (t5-small..t5-11b) */
  float eps;
} lako_dims;

typedef struct {
  size_t shared;
  size_t enc_ln0[LAKO_MAX_LAYERS], enc_qkv[LAKO_MAX_LAYERS], enc_o[LAKO_MAX_LAYERS], enc_ln1[LAKO_MAX_LAYERS],
      enc_wi[LAKO_MAX_LAYERS], enc_wo[LAKO_MAX_LAYERS];
  size_t enc_relbias, enc_final_ln;
  size_t dec_ln0[LAKO_MAX_LAYERS], dec_qkv[LAKO_MAX_LAYERS], dec_o[LAKO_MAX_LAYERS], dec_ln1[LAKO_MAX_LAYERS],
      dec_xq[LAKO_MAX_LAYERS], dec_xo[LAKO_MAX_LAYERS], dec_ln2[LAKO_MAX_LAYERS], dec_wi[LAKO_MAX_LAYERS],
      dec_wo[LAKO_MAX_LAYERS];
  size_t dec_xkv, dec_relbias, dec_final_ln;
  size_t total; /* elements */
} lako_layout;

typedef struct {
  lako_dims dims;
  const float* w_f32; /* flat fp32 master weights (norm gains, embedding gather and bias tables read these) */
  const void* w_bf16; /* flat bf16 copy (GEMM operands) */
  float* g_f32;       /* flat fp32 gradients (backward accumulates; may be NULL for inference) */
  const int* enc_buckets; /* int32 device tables: bucket(rel) at index rel + bucket_center, HF:188-234 */
  const int* dec_buckets;
  int bucket_center;
  /* precision = fp32 only (else NULL): the TF32 split of w_f32 made by lako_split_tf32 (hi = low 13 mantissa bits
   * cleared, lo = the same truncation of w - hi), same element offsets */
  const float* w_tf32_hi;
  const float* w_tf32_lo;
} lako_fidt5;

typedef struct {
  const int64_t* ids;        /* [B][N][L] passage token ids (src/data.py:148 passage_ids) */
  const unsigned char* mask; /* [B][N][L] 1 = token, 0 = pad (passage_masks, bool) */
  const int64_t* labels;     /* [B][T], -100 = ignore (target_ids); NULL for generate */
  int B, N, L, T;
  void* workspace; /* >= lako_fidt5_workspace_bytes(...) */
  size_t workspace_bytes;
  float* logits; /* out: [B*T][vocab] fp32 */
  float* stats;  /* out: stats[0] = sum of token losses, stats[1] = number of non-ignored targets */
  int count_is_set; /* 1: stats[1] already holds the (global, all-reduced) count and stats[0] is zeroed */
  /* training dropout (HF dropout_rate, 0.1 in the reference; 0 disables).  Masks come from Philox4x32-7 keyed by
   * (seed, step, call site, element) and are regenerated in backward -- pass the same values to both calls. */
  float dropout_p;
  unsigned int dropout_step;
  unsigned long long dropout_seed;
  /* Un-padded execution (optional).  When every passage's mask is a prefix (valid tokens first, pad after -- what
   * src/data.py's Collator produces), pass the int32 DEVICE array cu_seqlens[B*N+1] of exclusive prefix sums of the
   * passage lengths and packed_tokens = cu_seqlens[B*N] (host value).  All encoder-side work then runs on the
   * packed tokens only; padded positions never influence loss, logits, gradients or scores in the reference either
   * (they are masked as keys everywhere), so results are unchanged.  NULL = padded layout with `mask`. */
  const int* cu_seqlens;
  int packed_tokens;
  /* backward only, optional: LAKO_N_GRAD_STAGES cudaEvent_t recorded as gradient ranges become final
   * (lako_fidt5_grad_stages), so the caller can start the data-parallel reduce-scatter / all-reduce of a range while backward is
   * still running */
  void* stage_events;
  /* 1: `labels` holds decoder_input_ids (HF's other calling convention): no shift, no loss, logits only */
  int labels_are_decoder_ids;
  /* 0 = bf16 tensor-core operands, fp32 accumulation (default).  1 = fp32: 3xTF32-split tcgen05 GEMMs on the fp32
   * master weights + fp32 attention kernels; forward (mode 0) and generate only -- the reference is fp32-only
   * (train_reader.py:67) and north_star's 1e-4 logit tolerance is stated for this mode. */
  int precision;
  /* 1 (training, mode 1, and the matching lako_fidt5_backward): recompute the encoder layers' activations in backward
   * instead of keeping them -- the reference's CheckpointWrapper (src/model.py:236-284, --use_checkpoint).  Per layer
   * only the fp32 layer input stays resident (4 d bytes per token instead of ~29 d); backward re-runs each layer's
   * forward (minus its last GEMM) from it, with the same Philox dropout masks, so gradients are those of recompute = 0.
   * Size the workspace with lako_fidt5_workspace_bytes(..., mode = 3). */
  int recompute;
  /* training forward only, optional: LAKO_N_GRAD_STAGES cudaEvent_t (entries may be NULL), indexed like stage_events,
   * that the forward WAITS on before it first touches the weights of that range -- event 5 (shared embedding, norm
   * gains, bias tables) before the embedding gather, events 4..1 before the first layer of that encoder group, event 0
   * before the cross-K/V projection and the decoder.  Lets a sharded data-parallel optimizer all-gather the updated
   * weights on another stream while the next forward is already running (lako_b200/optim.py). */
  void* weight_events;
} lako_batch;

int lako_fidt5_layout(const lako_dims* dims, lako_layout* out);
/* mode 0 = forward only (eval loss), 1 = training (keeps activations for lako_fidt5_backward), 2 = forward only with
 * lako_batch.precision = 1 (fp32), 3 = training with lako_batch.recompute = 1 (a size query only: the calls take mode 1) */
size_t lako_fidt5_workspace_bytes(const lako_dims* dims, int B, int N, int L, int T, int mode);
/* FiDT5.forward (src/model.py:39-51) + T5 loss (HF:1105-1117) */
int lako_fidt5_forward(const lako_fidt5* m, const lako_batch* b, int mode, void* stream);
/* loss.backward() (train_reader.py:73) into g_f32; requires a mode-1 forward on the same workspace.
 * loss_grad: device scalar d(objective)/d(loss) (NULL = 1) */
int lako_fidt5_backward(const lako_fidt5* m, const lako_batch* b, const float* loss_grad, void* stream);
#define LAKO_N_GRAD_STAGES 6
/* ranges[7][2] = {begin,end} element ranges of g_f32: range i < 5 is final at stage event i (0 = decoder side,
 * 1..4 = encoder layer quarters, top first), ranges 5 and 6 at stage event 5 (end of backward) */
int lako_fidt5_grad_stages(const lako_dims* dims, size_t* ranges);
/* FiDT5.generate (src/model.py:54-60), greedy with KV cache; synchronises the stream once per token
 * to test for early stop.  tok_score_* (may be NULL) receive the step-0 cross-attention score sums
 * that src/model.py:327-329 + :152-167 derive from score_storage. */
size_t lako_fidt5_generate_workspace_bytes(const lako_dims* dims, int B, int N, int L, int max_length, int precision);
int lako_fidt5_generate(const lako_fidt5* m, const lako_batch* b, int max_length, int min_length, int64_t* out_ids,
                        int* out_len, float* tok_score_first, float* tok_score_last, void* stream);

/* ---- per-fact segmentation of the step-0 token scores: src/model.py:166-203 (+ test_reader.py:76 softmax) --------
 * tok_a (+ tok_b when not NULL) fp32 [B][N*L] as produced by lako_fidt5_generate; ids1 int64 [B][ld_ids] = token ids
 * of passage 1; scores are read from passage `score_passage` ((N+1)/2 reproduces torch.chunk(.., 2, dim=1)[1][:, 0]).
 * out float64 [B][n_context]; style 0 mean, 1 max, 2 "21mean"; missing facts -5; all / divisor; softmax != 0: softmax over
 * the n_context entries of each row. */
int lako_fact_scores(const float* tok_a, const float* tok_b, const int64_t* ids1, int ld_ids, double* out, int B, int N,
                     int L, int score_passage, int n_context, int style, double divisor, int softmax, void* stream);

/* ---- attention kernels (unit-test entry points; the model calls them internally) --------------- */
int lako_build_bias_lut(const float* table, const int* buckets, float* lut, int H, int L, int bucket_center,
                        void* stream);
int lako_fold_dlut(const float* dlut, const int* buckets, float* dtable, int H, int L, int bucket_center, int nb,
                   void* stream);
/* encoder self-attention HF:253-344 per passage (src/model.py:230-232); qkv bf16 [P*L][3*H*64] */
/* test support: overwrite every TMEM column of every SM with NaN bit patterns (tcgen05.alloc hands out columns
   uncleared; a kernel that reads columns it did not write would pick these up) */
int lako_debug_poison_tmem(void* stream);
/* encoder self-attention forward implementation: 2 (default) = persistent warp-specialised kernel with the softmax
 * probabilities kept in tensor memory; 1 = the one-shot kernel of round 1 (also env LAKO_ATTN_FWD=1); 3 = forward AND
 * backward through the tiled any-length kernels (csrc/enc_attn_long.cu) that L > 256 always takes (L <= 2048; they
 * neither write nor read the keep-word buffer of the _dropout entry points: the backward regenerates the words, with
 * 4 * max(8, ceil(L / 32)) Philox counters per query row) */
void lako_set_attn_impl(int v);
int lako_enc_attn_fwd(const void* qkv, const unsigned char* mask, const float* bias_lut, void* out, float* lse, int P,
                      int L, int H, void* stream);
int lako_enc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const unsigned char* mask,
                      const float* bias_lut, void* dqkv, float* dlut, int P, int L, int H, void* stream);
/* the same with dropout on the attention probabilities (HF modeling_t5.py:332, active in model.train()): keep decisions
 * from Philox4x32-7 keyed by (seed, step, site), one bit-sliced comparison per (query row, 32-key chunk)
 * (csrc/common.cuh drop_keep32_sliced; lako_b200/philox.py attn_keep_words restates it).  keepbits [P][H][L][8] receives
 * (forward) / supplies (backward; NULL = regenerate) the keep words. */
int lako_enc_attn_fwd_dropout(const void* qkv, const unsigned char* mask, const float* bias_lut, void* out, float* lse,
                              unsigned int* keepbits, int P, int L, int H, unsigned long long seed, unsigned int step,
                              unsigned int site, float p, void* stream);
int lako_enc_attn_bwd_dropout(const void* qkv, const void* out, const void* dout, const float* lse,
                              const unsigned char* mask, const float* bias_lut, void* dqkv, float* dlut,
                              const unsigned int* keepbits, int P, int L, int H, unsigned long long seed,
                              unsigned int step, unsigned int site, float p, void* stream);
/* cross-attention over the fused memory, src/model.py:286-349 */
size_t lako_xattn_workspace_bytes(int B, int H, int T, int M);
int lako_xattn_fwd(const void* q, const void* kv, const unsigned char* mask, void* out, float* lse, float* tok_score,
                   float* workspace, int B, int T, int M, int H, int ldq, int ldkv, int koff, int voff, int ldo,
                   void* stream);
int lako_xattn_bwd(const void* q, const void* kv, const unsigned char* mask, const void* out, const void* dout,
                   const float* lse, void* dkv, float* dq_f32, int B, int T, int M, int H, int ldq, int ldkv, int koff,
                   int voff, int ldo, void* stream);
/* ---- KV-cached single-token decode (greedy generate, src/model.py:54-60 -> HF greedy search with use_cache) ----
 * Few-row Linear: y[M][ldy] = epi(alpha * norm?(x) W[N][K]^T) for the M = batch rows of a decode step (any M; rows are
 * processed 64 at a time).  x_f32 != NULL: x is the fp32 residual stream [M][K], RMS-normalised with gain ln_w / eps on
 * the fly when ln_w != NULL (T5LayerNorm HF:55-68 fused as the prologue); else x_bf16 [M][ldx].
 * epi: 0 bf16 store, 1 bf16 ReLU store, 2 fp32 store, 3 fp32 y += (residual add, HF:375,149). */
int lako_smallm_linear(const float* x_f32, const void* x_bf16, int ldx, const float* ln_w, float eps, const void* W,
                       void* y, int ldy, int M, int N, int K, int epi, float alpha, void* stream);
/* cross_attention_forward (src/model.py:286-349) for ONE query token per question (q fp32 [B][ldq]) against the bf16
 * K/V rows of the fused memory (kv_rows rows of pitch ldkv, head h at column koff/voff + 64 h; tiles arrive by TMA).
 * Packed memory: question b owns rows [cu[b*npsg], cu[(b+1)*npsg]) and tok_index maps a packed row to its index in
 * the padded [B][M] score buffer; cu == NULL: padded rows [b*M, (b+1)*M) with the 0/1 key mask.  tok_score (may be
 * NULL) += the pre-softmax scores of valid keys (summed over heads) when *step == 0 (step == NULL: always) -- the
 * capture of src/model.py:327-329.  The flash-decoding combine of the 256-key chunks is done by the last chunk CTA
 * to arrive; `workspace` (lako_xattn_decode_workspace_bytes) must be ZERO before the first call and is left with its
 * arrival counters zeroed.  zero_q != 0: q[b][64 h .. 64 h + 64) is cleared once (b, h) is finished (the K-split
 * projection that produces q accumulates into it).  head_major != 0: kv is stored [ldkv/64][kv_rows][64] (one contiguous
 * slab per 64-column head, what lako_fidt5_generate's K/V projection writes) instead of row-major [kv_rows][ldkv].
 * The 128-row tiles read up to 127 rows past a question's last token: rows that belong to no question (the unused tail
 * of a slab) must hold finite values (zero them once), they are multiplied by zero probabilities. */
size_t lako_xattn_decode_workspace_bytes(int B, int H, int M);
int lako_xattn_decode(const float* q, const void* kv, size_t kv_rows, const unsigned char* mask, void* out,
                      float* tok_score, const int* step, float* workspace, int B, int M, int H, int ldq, int ldkv, int koff,
                      int voff, int ldo, const int* cu, int npsg, const int* tok_index, int zero_q, int head_major,
                      void* stream);
/* decoder self-attention of the token at position pos (or *pos_dev when not NULL) against the KV cache
 * [B][H][Tmax][64] (appends this step's k, v first); qkv bf16 [B][3*H*64], lut fp32 [H][2*Tmax-1] causal buckets */
int lako_dself_step(const void* qkv, void* kcache, void* vcache, const float* lut, void* out, int B, int pos, int Tmax,
                    int H, const int* pos_dev, void* stream);
/* replay decode steps from a captured CUDA graph (default on; also env LAKO_DECODE_GRAPH=0) */
void lako_set_decode_graph(int on);
/* counters since process start: graphs captured, decode steps replayed from a graph, failed captures */
void lako_decode_graph_stats(long long* captures, long long* launches, long long* failures);
/* causal decoder self-attention HF:253-344,438-445 */
int lako_dself_fwd(const void* qkv, const float* lut, void* out, float* probs, int B, int T, int H, void* stream);
int lako_dself_bwd(const void* qkv, const void* dout, const float* probs, void* dqkv, float* dlut, int B, int T, int H,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif
