// Host-side orchestration of the FiDT5 reader path: enqueues the kernels of one forward (with the
// loss), one backward and the greedy decode loop on the caller's stream.  Pure C++ so that a step is
// ~1k kernel launches issued back-to-back without an interpreter in between.
//
// Reference call stacks:
//   FiDT5.forward            /root/reference/src/model.py:39-51   -> lako_fidt5_forward
//   EncoderWrapper.forward   /root/reference/src/model.py:227-234 -> encode()
//   loss.backward()          /root/reference/train_reader.py:73   -> lako_fidt5_backward
//   FiDT5.generate           /root/reference/src/model.py:54-60   -> lako_fidt5_generate
//   score capture            /root/reference/src/model.py:327-329 -> tok_score accumulation in generate
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ops.cuh"

using namespace lako;

// ------------------------------------------------------------------------------------------------
// flat parameter layout
// ------------------------------------------------------------------------------------------------
extern "C" int lako_fidt5_layout(const lako_dims* dm, lako_layout* lo) {
  if (!dm || !lo) return LAKO_ERR_BAD_ARG;
  LAKO_REQUIRE(dm->n_layers > 0 && dm->n_layers <= LAKO_MAX_LAYERS, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE(dm->d_model % 128 == 0 && dm->d_model <= 1024, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE(dm->d_ff % 8 == 0 && dm->vocab % 8 == 0 && dm->n_heads > 0, LAKO_ERR_UNSUPPORTED);
  const size_t d = dm->d_model, inner = static_cast<size_t>(dm->n_heads) * 64, ff = dm->d_ff;
  size_t off = 0;
  auto take = [&](size_t n) {
    const size_t o = off;
    off += (n + 7) / 8 * 8;
    return o;
  };
  lo->shared = take(static_cast<size_t>(dm->vocab) * d);
  for (int l = 0; l < dm->n_layers; ++l) {
    lo->enc_ln0[l] = take(d);
    lo->enc_qkv[l] = take(3 * inner * d);
    lo->enc_o[l] = take(d * inner);
    lo->enc_ln1[l] = take(d);
    lo->enc_wi[l] = take(ff * d);
    lo->enc_wo[l] = take(d * ff);
  }
  lo->enc_relbias = take(static_cast<size_t>(dm->n_buckets) * dm->n_heads);
  lo->enc_final_ln = take(d);
  for (int l = 0; l < dm->n_layers; ++l) {
    lo->dec_ln0[l] = take(d);
    lo->dec_qkv[l] = take(3 * inner * d);
    lo->dec_o[l] = take(d * inner);
    lo->dec_ln1[l] = take(d);
    lo->dec_xq[l] = take(inner * d);
    lo->dec_xo[l] = take(d * inner);
    lo->dec_ln2[l] = take(d);
    lo->dec_wi[l] = take(ff * d);
    lo->dec_wo[l] = take(d * ff);
  }
  lo->dec_xkv = take(static_cast<size_t>(dm->n_layers) * 2 * inner * d);
  lo->dec_relbias = take(static_cast<size_t>(dm->n_buckets) * dm->n_heads);
  lo->dec_final_ln = take(d);
  lo->total = off;
  return LAKO_OK;
}

static inline int enc_group_begin(int nl, int k) {  // first layer of encoder group k (k = 0 is the TOP quarter)
  const int gsz = (nl + 3) / 4;
  const int lo = nl - (k + 1) * gsz;
  return lo < 0 ? 0 : lo;
}

namespace {

struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

typedef __nv_bfloat16 bf16;

struct EncLayerAct {
  unsigned char* hbits;  // ReLU-activity bitmask [tokens][d_ff/8]
  bf16 *n_a, *qkv, *ao, *n_f, *h;
  float *lse, *rstd_a, *rstd_f;
  uint32_t* akeep;  // attention-dropout keep masks [P][H][L][8] (forward -> backward)
};
struct DecLayerAct {
  unsigned char* hbits;
  bf16 *n_s, *qkv, *ao, *n_c, *q, *xo, *n_f, *h;
  float *probs, *lse, *rstd_s, *rstd_c, *rstd_f;
};

struct DecGrad {  // per-layer bf16 gradient operands of the decoder (kept so wgrads can run off the critical path)
  bf16 *g0, *g1, *g2, *dh, *dq, *dqkv;
};

struct Work {
  std::vector<DecGrad> dg;
  // sizes
  int B, N, L, T, P, M, Mtok, BT;
  bool recompute;  // mode 3: encoder activations are recomputed layer by layer in backward
  // packed (un-padded) execution: Mrows = number of token rows actually processed
  int Mrows;
  const int* cu;
  int* tok_index;
  // encoder
  std::vector<float*> X;  // 2*nl+1 residual-stream snapshots (train) or 2 ping-pong (eval)
  std::vector<EncLayerAct> enc;
  float *rstd_enc_final, *lut_enc, *dlut_enc;
  bf16 *mem, *delta_enc, *kv_all;
  // decoder
  std::vector<float*> Y;  // 3*nl+1
  std::vector<DecLayerAct> dec;
  float *rstd_dec_final, *lut_dec, *dlut_dec, *logits, *stats, *row_loss, *xattn_ws;
  bf16 *hn, *delta_dec, *dlogits;
  long long* dec_ids;
  // backward temporaries
  float *dres_enc, *dres_dec, *dq_f32;
  float* dsk_f32;  // [BT][d] split-K accumulation scratch of the decoder's deep-K dgrads (kept zeroed between uses)
  bf16 *dres_enc_b, *dn_enc, *dh_enc, *dao_enc, *dqkv_enc, *dkv_all, *dmem;
  bf16 *dres_dec_b, *dn_dec, *dh_dec, *dao_dec, *dqkv_dec, *dq_dec, *dhn;
  size_t bytes;
};

// mode: 0 = eval forward (no activations kept), 1 = train (keep everything backward needs)
void carve(const lako_dims& dm, int B, int N, int L, int T, int mode, void* base, Work& w) {
  const int nl = dm.n_layers;
  const size_t d = dm.d_model, inner = static_cast<size_t>(dm.n_heads) * 64, ff = dm.d_ff, H = dm.n_heads;
  w.B = B; w.N = N; w.L = L; w.T = T;
  w.P = B * N; w.M = N * L; w.Mtok = B * N * L; w.BT = B * T;
  const size_t Mt = w.Mtok, BT = w.BT;
  Bump b(base);
  // mode 3: training with the encoder layers' activations recomputed in backward (the reference's CheckpointWrapper,
  // src/model.py:236-284 under --use_checkpoint): per layer only its fp32 input X[2l] is kept; everything else a layer's
  // backward reads lives in ONE set of buffers that enc_layer_recompute() refills from X[2l] just before that backward
  const bool recompute = mode == 3;
  const bool train = mode == 1 || recompute;
  w.recompute = recompute;
  w.X.resize(2 * nl + 1);
  if (recompute) {
    float* odd = b.take<float>(Mt * d);  // X[2l+1] (residual stream between a layer's two sub-blocks) is recomputed too
    for (int i = 0; i <= 2 * nl; ++i) w.X[i] = (i & 1) ? odd : b.take<float>(Mt * d);
  } else {
    const int nX = train ? 2 * nl + 1 : 2;
    std::vector<float*> bufs(nX);
    for (int i = 0; i < nX; ++i) bufs[i] = b.take<float>(Mt * d);
    for (int i = 0; i <= 2 * nl; ++i) w.X[i] = bufs[train ? i : (i & 1)];
  }
  w.enc.resize(nl);
  EncLayerAct shared_e{};
  if (!train || recompute) {
    shared_e.n_a = b.take<bf16>(Mt * d);
    shared_e.qkv = b.take<bf16>(Mt * 3 * inner);
    shared_e.ao = b.take<bf16>(Mt * inner);
    shared_e.n_f = recompute ? b.take<bf16>(Mt * d) : shared_e.n_a;
    shared_e.h = b.take<bf16>(Mt * ff);
    shared_e.lse = nullptr;
    shared_e.akeep = nullptr;
    shared_e.rstd_a = shared_e.rstd_f = nullptr;
    if (recompute) {
      shared_e.hbits = b.take<unsigned char>(Mt * ff / 8);
      shared_e.lse = b.take<float>(static_cast<size_t>(w.P) * H * L);
      shared_e.akeep = L <= 256 ? b.take<uint32_t>(static_cast<size_t>(w.P) * H * L * 8) : nullptr;
      shared_e.rstd_a = b.take<float>(Mt);
      shared_e.rstd_f = b.take<float>(Mt);
    }
  }
  for (int l = 0; l < nl; ++l) {
    if (!train || recompute) {
      w.enc[l] = shared_e;
      continue;
    }
    EncLayerAct& a = w.enc[l];
    a.n_a = b.take<bf16>(Mt * d);
    a.qkv = b.take<bf16>(Mt * 3 * inner);
    a.ao = b.take<bf16>(Mt * inner);
    a.n_f = b.take<bf16>(Mt * d);
    a.h = b.take<bf16>(Mt * ff);
    a.hbits = b.take<unsigned char>(Mt * ff / 8);
    a.lse = b.take<float>(static_cast<size_t>(w.P) * H * L);
    a.akeep = L <= 256 ? b.take<uint32_t>(static_cast<size_t>(w.P) * H * L * 8) : nullptr;  // longer passages regenerate the words
    a.rstd_a = b.take<float>(Mt);
    a.rstd_f = b.take<float>(Mt);
  }
  w.rstd_enc_final = train ? b.take<float>(Mt) : nullptr;
  w.lut_enc = b.take<float>(H * (2 * L - 1));
  w.mem = b.take<bf16>(Mt * d);
  w.delta_enc = b.take<bf16>(Mt * d);
  w.kv_all = b.take<bf16>(Mt * nl * 2 * inner);
  // decoder
  const int nY = train ? 3 * nl + 1 : 2;
  w.Y.resize(3 * nl + 1);
  {
    std::vector<float*> bufs(nY);
    for (int i = 0; i < nY; ++i) bufs[i] = b.take<float>(BT * d);
    for (int i = 0; i <= 3 * nl; ++i) w.Y[i] = bufs[train ? i : (i & 1)];
  }
  w.dec.resize(nl);
  DecLayerAct shared_d{};
  if (!train) {
    shared_d.n_s = b.take<bf16>(BT * d);
    shared_d.qkv = b.take<bf16>(BT * 3 * inner);
    shared_d.ao = b.take<bf16>(BT * inner);
    shared_d.n_c = shared_d.n_s;
    shared_d.q = b.take<bf16>(BT * inner);
    shared_d.xo = b.take<bf16>(BT * inner);
    shared_d.n_f = shared_d.n_s;
    shared_d.h = b.take<bf16>(BT * ff);
    shared_d.probs = nullptr;
    shared_d.lse = nullptr;
  }
  for (int l = 0; l < nl; ++l) {
    if (!train) {
      w.dec[l] = shared_d;
      continue;
    }
    DecLayerAct& a = w.dec[l];
    a.n_s = b.take<bf16>(BT * d);
    a.qkv = b.take<bf16>(BT * 3 * inner);
    a.ao = b.take<bf16>(BT * inner);
    a.n_c = b.take<bf16>(BT * d);
    a.q = b.take<bf16>(BT * inner);
    a.xo = b.take<bf16>(BT * inner);
    a.n_f = b.take<bf16>(BT * d);
    a.h = b.take<bf16>(BT * ff);
    a.hbits = b.take<unsigned char>(BT * ff / 8);
    a.probs = b.take<float>(static_cast<size_t>(B) * H * T * T);
    a.lse = b.take<float>(static_cast<size_t>(B) * H * T);
    a.rstd_s = b.take<float>(BT);
    a.rstd_c = b.take<float>(BT);
    a.rstd_f = b.take<float>(BT);
  }
  w.rstd_dec_final = train ? b.take<float>(BT) : nullptr;
  w.lut_dec = b.take<float>(H * (2 * T - 1));
  w.hn = b.take<bf16>(BT * d);
  w.delta_dec = b.take<bf16>(BT * d);
  w.logits = nullptr;  // caller-provided
  w.stats = nullptr;   // caller-provided
  w.row_loss = b.take<float>(BT);
  w.xattn_ws = b.take<float>(xattn_workspace_bytes(B, dm.n_heads, T, w.M) / sizeof(float));
  w.dec_ids = b.take<long long>(BT);
  w.tok_index = b.take<int>(Mt);
  w.Mrows = w.Mtok;
  w.cu = nullptr;
  if (train) {
    const size_t vpad = (static_cast<size_t>(dm.vocab) + 7) / 8 * 8;
    w.dlogits = b.take<bf16>(BT * vpad);
    w.dlut_enc = b.take<float>(H * (2 * L - 1));
    w.dlut_dec = b.take<float>(H * (2 * T - 1));
    w.dres_enc = b.take<float>(Mt * d);
    w.dres_enc_b = b.take<bf16>(Mt * d);
    w.dn_enc = b.take<bf16>(Mt * d);
    w.dh_enc = b.take<bf16>(Mt * ff);
    w.dao_enc = b.take<bf16>(Mt * inner);
    w.dqkv_enc = b.take<bf16>(Mt * 3 * inner);
    w.dkv_all = b.take<bf16>(Mt * nl * 2 * inner);
    w.dmem = b.take<bf16>(Mt * d);
    w.dres_dec = b.take<float>(BT * d);
    w.dres_dec_b = b.take<bf16>(BT * d);
    w.dg.resize(nl);
    for (int l = 0; l < nl; ++l) {
      w.dg[l].g0 = b.take<bf16>(BT * d);
      w.dg[l].g1 = b.take<bf16>(BT * d);
      w.dg[l].g2 = b.take<bf16>(BT * d);
      w.dg[l].dh = b.take<bf16>(BT * ff);
      w.dg[l].dq = b.take<bf16>(BT * inner);
      w.dg[l].dqkv = b.take<bf16>(BT * 3 * inner);
    }
    w.dn_dec = b.take<bf16>(BT * d);
    w.dh_dec = b.take<bf16>(BT * ff);
    w.dao_dec = b.take<bf16>(BT * inner);
    w.dqkv_dec = b.take<bf16>(BT * 3 * inner);
    w.dq_dec = b.take<bf16>(BT * inner);
    w.dq_f32 = b.take<float>(BT * inner);
    w.dsk_f32 = b.take<float>(BT * d);
    w.dhn = b.take<bf16>(BT * d);
  }
  w.bytes = (b.off + 255) & ~size_t(255);
}

#define RC(x)            \
  do {                   \
    int rc__ = (x);      \
    if (rc__) return rc__; \
  } while (0)

// y[M,N] = x[M,K] W[N,K]^T
inline int lin_fwd(const bf16* x, const bf16* W, void* y, int M, int N, int K, int epi, float alpha, cudaStream_t s,
                   DropSpec drop = DropSpec{}, unsigned char* relu_bits = nullptr) {
  return gemm_bf16(x, W, y, M, N, K, K, K, N, 0, 0, epi, alpha, relu_bits, N / 8, 1, 0, s, drop);
}
// dx[M,K] = dy[M,N] W[N,K]   (optionally masked by aux > 0)
inline int lin_dgrad(const bf16* dy, const bf16* W, bf16* dx, int M, int N, int K, const unsigned char* relu_bits,
                     float alpha, cudaStream_t s) {
  return gemm_bf16(dy, W, dx, M, K, N, N, K, K, 0, 1, relu_bits ? EPI_BF16_BITMASK : EPI_BF16, alpha, relu_bits, K / 8, 1,
                   0, s);
}
// Same product for the decoder's few-row, deep-N cases (LM head, FFN-in, fused qkv): an 80-row dgrad has only a handful of
// output tiles, so a single CTA per tile would stream all of N serially.  Split N over many CTAs accumulating into the
// zeroed fp32 scratch, then round to bf16 (the cast leaves the scratch zeroed again).
inline int lin_dgrad_splitk(const bf16* dy, int ld_dy, const bf16* W, bf16* dx, float* scratch, int M, int N, int K,
                            float alpha, cudaStream_t s) {
  int rc = gemm_bf16(dy, W, scratch, M, K, N, ld_dy, K, K, 0, 1, EPI_F32_ATOMIC, alpha, nullptr, 0, 0, 0, s);
  if (rc) return rc;
  return cast_f32_to_bf16(scratch, dx, static_cast<size_t>(M) * K, s, true);
}
// dW[N,K] += alpha * dy[M,N]^T x[M,K]
inline int lin_wgrad(const bf16* dy, const bf16* x, float* dW, int M, int N, int K, float alpha, cudaStream_t s) {
  return gemm_bf16(dy, x, dW, N, K, M, N, K, K, 1, 1, EPI_F32_ATOMIC, alpha, nullptr, 0, 0, 0, s);
}

enum DropKind : unsigned {
  E_EMB = 0, E_ATTN, E_O, E_RELU, E_FF, E_FINAL, D_EMB, D_SELF, D_O, D_XP, D_XO, D_RELU, D_FF, D_FINAL
};
struct Dropper {
  unsigned long long seed;
  unsigned step;
  float p;
  DropSpec operator()(int layer, unsigned kind) const {
    return p > 0.f ? make_drop(seed, step, static_cast<unsigned>(layer) * 16u + kind, p) : DropSpec{};
  }
  DropSpec none() const { return DropSpec{}; }
};

inline int set_packing(Work& w, const lako_batch* bt, cudaStream_t s) {
  w.cu = nullptr;
  w.Mrows = w.Mtok;
  if (bt->cu_seqlens) {
    LAKO_REQUIRE(bt->packed_tokens > 0 && bt->packed_tokens <= w.Mtok, LAKO_ERR_BAD_ARG);
    w.cu = bt->cu_seqlens;
    w.Mrows = bt->packed_tokens;
  }
  (void)s;
  return LAKO_OK;
}

struct Ptrs {
  const lako_dims* dm;
  lako_layout lo;
  const float* wf;
  const bf16* wb;
  float* g;
  int d, inner, ff, H, nl, V;
  float eps;
};

int make_ptrs(const lako_fidt5* m, Ptrs& p) {
  LAKO_REQUIRE(m && m->w_f32 && m->w_bf16, LAKO_ERR_BAD_ARG);
  p.dm = &m->dims;
  RC(lako_fidt5_layout(&m->dims, &p.lo));
  p.wf = m->w_f32;
  p.wb = static_cast<const bf16*>(m->w_bf16);
  p.g = m->g_f32;
  p.d = m->dims.d_model;
  p.H = m->dims.n_heads;
  p.inner = p.H * 64;
  p.ff = m->dims.d_ff;
  p.nl = m->dims.n_layers;
  p.V = m->dims.vocab;
  p.eps = m->dims.eps;
  return LAKO_OK;
}

// One encoder layer from its normed input n_a (already in place) up to the ReLU output h: q|k|v, self-attention, o-projection,
// residual + second norm, wi + ReLU (+ dropout).  Shared by the forward pass and by the recompute of mode 3.
int enc_layer_body(const Ptrs& p, Work& w, int l, const unsigned char* amask, cudaStream_t s, const Dropper& dr) {
  const int d = p.d, inner = p.inner, ff = p.ff, Mt = w.Mrows;
  EncLayerAct& a = w.enc[l];
  RC(lin_fwd(a.n_a, p.wb + p.lo.enc_qkv[l], a.qkv, Mt, 3 * inner, d, EPI_BF16, 1.f, s));
  RC(enc_attn_fwd(a.qkv, amask, w.lut_enc, a.ao, a.lse, w.P, w.L, p.H, s, dr(l, E_ATTN), w.cu, Mt, a.akeep));
  RC(lin_fwd(a.ao, p.wb + p.lo.enc_o[l], w.delta_enc, Mt, d, inner, EPI_BF16, 1.f, s));
  RC(rmsnorm_fwd(w.X[2 * l], w.delta_enc, w.X[2 * l + 1], p.wf + p.lo.enc_ln1[l], a.n_f, a.rstd_f, Mt, d, p.eps, 1.f, s,
                 dr(l, E_O)));
  RC(lin_fwd(a.n_f, p.wb + p.lo.enc_wi[l], a.h, Mt, ff, d, EPI_BF16_RELU, 1.f, s, dr(l, E_RELU), a.hbits));
  return LAKO_OK;
}
// mode 3 (CheckpointWrapper): refill the shared activation buffers with layer l's forward from its saved input X[2l].
// The dropout masks are functions of (seed, step, site, element), so the recomputed activations are the forward's, bit
// for bit; the layer's wo GEMM is not repeated (its output only fed X[2l+2], which is saved).
int enc_layer_recompute(const Ptrs& p, Work& w, int l, const unsigned char* amask, cudaStream_t s, const Dropper& dr) {
  EncLayerAct& a = w.enc[l];
  RC(rmsnorm_fwd(w.X[2 * l], nullptr, nullptr, p.wf + p.lo.enc_ln0[l], a.n_a, a.rstd_a, w.Mrows, p.d, p.eps, 1.f, s));
  return enc_layer_body(p, w, l, amask, s, dr);
}

// ---- encoder: ids (P*L) -> mem bf16 [Mtok][d] and kv_all bf16 [Mtok][nl*2*inner]
int encode(const Ptrs& p, const lako_fidt5* m, Work& w, const long long* ids, const unsigned char* mask, bool train,
           cudaStream_t s, const Dropper& dr, bool kv_head_major = false, const cudaEvent_t* wev = nullptr) {
  const int d = p.d, inner = p.inner, ff = p.ff, nl = p.nl, Mt = w.Mrows;
  const unsigned char* amask = w.cu ? nullptr : mask;  // packed rows are all valid tokens
  // weight-ready events of an overlapped all-gather (lako_batch.weight_events): wait right before the first use
  auto wait_w = [&](int stage) {
    if (wev && wev[stage]) cudaStreamWaitEvent(s, wev[stage], 0);
  };
  if (w.cu) RC(build_tok_index(w.cu, w.tok_index, w.P, w.L, s));
  wait_w(5);
  RC(embed_gather(ids, p.wf + p.lo.shared, w.X[0], Mt, d, p.V, s, dr(0, E_EMB), w.cu ? w.tok_index : nullptr));
  RC(build_bias_lut(p.wf + p.lo.enc_relbias, m->enc_buckets, w.lut_enc, p.H, w.L, m->bucket_center, s));
  for (int l = 0; l < nl; ++l) {
    EncLayerAct& a = w.enc[l];
    for (int k = 0; k < 4; ++k)
      if (l == enc_group_begin(nl, k) && (k == 0 || enc_group_begin(nl, k - 1) > l)) wait_w(1 + k);
    if (l == 0)
      RC(rmsnorm_fwd(w.X[0], nullptr, nullptr, p.wf + p.lo.enc_ln0[l], a.n_a, a.rstd_a, Mt, d, p.eps, 1.f, s));
    else
      RC(rmsnorm_fwd(w.X[2 * l - 1], w.delta_enc, w.X[2 * l], p.wf + p.lo.enc_ln0[l], a.n_a, a.rstd_a, Mt, d, p.eps,
                     1.f, s, dr(l - 1, E_FF)));
    RC(enc_layer_body(p, w, l, amask, s, dr));
    RC(lin_fwd(a.h, p.wb + p.lo.enc_wo[l], w.delta_enc, Mt, d, ff, EPI_BF16, 1.f, s));
  }
  RC(rmsnorm_fwd(w.X[2 * nl - 1], w.delta_enc, w.X[2 * nl], p.wf + p.lo.enc_final_ln, w.mem, w.rstd_enc_final, Mt, d,
                 p.eps, 1.f, s, dr(nl - 1, E_FF), dr(0, E_FINAL)));
  wait_w(0);
  // cross-attention K/V of ALL decoder layers in one GEMM (13.8 % of the model's FLOPs)
  // (generate: written head-major, [layer][k|v][head][token][64], so that a decode step streams contiguous slabs)
  RC(gemm_bf16(w.mem, p.wb + p.lo.dec_xkv, w.kv_all, Mt, nl * 2 * inner, d, d, d, nl * 2 * inner, 0, 0, EPI_BF16, 1.f, nullptr,
               0, 1, 0, s, DropSpec{}, kv_head_major ? w.Mtok : 0));  // slab capacity = the padded token count, so the
                                                                       // layout does not depend on the batch's lengths
  (void)train;
  return LAKO_OK;
}

#include "model_f32.inc"

}  // namespace

// mode 0 eval forward, 1 training forward (+ backward), 2 eval forward with precision = fp32
extern "C" size_t lako_fidt5_workspace_bytes(const lako_dims* dm, int B, int N, int L, int T, int mode) {
  if (mode == 2) {
    WorkF32 wf;
    carve_f32(*dm, B, N, L, T, T, nullptr, wf);
    return wf.bytes + 256;
  }
  Work w;
  carve(*dm, B, N, L, T, mode, nullptr, w);
  return w.bytes + 256;
}

// FiDT5.forward(input_ids, attention_mask, labels) -> loss (stats[0]/stats[1]) and logits
extern "C" int lako_fidt5_forward(const lako_fidt5* m, const lako_batch* bt, int mode, void* stream) {
  RC(check_arch());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Ptrs p;
  RC(make_ptrs(m, p));
  LAKO_REQUIRE(bt && bt->ids && bt->mask && bt->labels && bt->workspace && bt->logits && bt->stats, LAKO_ERR_BAD_ARG);
  if (bt->precision == 1) {
    LAKO_REQUIRE(mode == 0, LAKO_ERR_UNSUPPORTED);  // evaluation only
    return forward_f32(m, bt, p, s);
  }
  // L > 256 runs the encoder attention through the tiled kernels of enc_attn_long.cu; the bucket tables cover offsets up
  // to +-bucket_center
  LAKO_REQUIRE(bt->T >= 1 && bt->T <= 64 && bt->L >= 1 && bt->L <= ENC_ATTN_LONG_MAX_L && bt->L - 1 <= m->bucket_center,
               LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE(mode == 0 || mode == 1, LAKO_ERR_BAD_ARG);
  if (mode == 1 && bt->recompute) mode = 3;  // same pass, smaller activation footprint (see carve)
  Work w;
  carve(m->dims, bt->B, bt->N, bt->L, bt->T, mode, bt->workspace, w);
  LAKO_REQUIRE(w.bytes <= bt->workspace_bytes, LAKO_ERR_WORKSPACE);
  const bool train = mode == 1 || mode == 3;
  const int d = p.d, inner = p.inner, ff = p.ff, nl = p.nl, BT = w.BT;
  Dropper dr{bt->dropout_seed, bt->dropout_step, train ? bt->dropout_p : 0.f};
  LAKO_REQUIRE(dr.p >= 0.f && dr.p < 0.9f, LAKO_ERR_BAD_ARG);
  RC(set_packing(w, bt, s));
  const unsigned char* xmask = w.cu ? nullptr : bt->mask;
  RC(encode(p, m, w, reinterpret_cast<const long long*>(bt->ids), bt->mask, train, s, dr, false,
            train ? reinterpret_cast<const cudaEvent_t*>(bt->weight_events) : nullptr));

  const long long* labels = reinterpret_cast<const long long*>(bt->labels);
  if (bt->labels_are_decoder_ids) {
    LAKO_REQUIRE(!train, LAKO_ERR_BAD_ARG);
    cudaMemcpyAsync(w.dec_ids, labels, sizeof(long long) * BT, cudaMemcpyDeviceToDevice, s);
  } else {
    RC(shift_right(labels, w.dec_ids, bt->B, bt->T, s));
  }
  RC(embed_gather(w.dec_ids, p.wf + p.lo.shared, w.Y[0], BT, d, p.V, s, dr(0, D_EMB)));
  RC(build_bias_lut(p.wf + p.lo.dec_relbias, m->dec_buckets, w.lut_dec, p.H, bt->T, m->bucket_center, s));
  for (int l = 0; l < nl; ++l) {
    DecLayerAct& a = w.dec[l];
    if (l == 0)
      RC(rmsnorm_fwd(w.Y[0], nullptr, nullptr, p.wf + p.lo.dec_ln0[l], a.n_s, a.rstd_s, BT, d, p.eps, 1.f, s));
    else
      RC(rmsnorm_fwd(w.Y[3 * l - 1], w.delta_dec, w.Y[3 * l], p.wf + p.lo.dec_ln0[l], a.n_s, a.rstd_s, BT, d, p.eps,
                     1.f, s, dr(l - 1, D_FF)));
    RC(lin_fwd(a.n_s, p.wb + p.lo.dec_qkv[l], a.qkv, BT, 3 * inner, d, EPI_BF16, 1.f, s));
    RC(dself_fwd(a.qkv, w.lut_dec, a.ao, a.probs, bt->B, bt->T, p.H, s, dr(l, D_SELF)));
    RC(lin_fwd(a.ao, p.wb + p.lo.dec_o[l], w.delta_dec, BT, d, inner, EPI_BF16, 1.f, s));
    RC(rmsnorm_fwd(w.Y[3 * l], w.delta_dec, w.Y[3 * l + 1], p.wf + p.lo.dec_ln1[l], a.n_c, a.rstd_c, BT, d, p.eps, 1.f,
                   s, dr(l, D_O)));
    RC(lin_fwd(a.n_c, p.wb + p.lo.dec_xq[l], a.q, BT, inner, d, EPI_BF16, 1.f, s));
    RC(xattn_fwd(a.q, w.kv_all, xmask, a.xo, a.lse, nullptr, w.xattn_ws, bt->B, bt->T, w.M, p.H, inner,
                 nl * 2 * inner, l * 2 * inner, l * 2 * inner + inner, inner, s, dr(l, D_XP), w.cu, bt->N, nullptr));
    RC(lin_fwd(a.xo, p.wb + p.lo.dec_xo[l], w.delta_dec, BT, d, inner, EPI_BF16, 1.f, s));
    RC(rmsnorm_fwd(w.Y[3 * l + 1], w.delta_dec, w.Y[3 * l + 2], p.wf + p.lo.dec_ln2[l], a.n_f, a.rstd_f, BT, d, p.eps,
                   1.f, s, dr(l, D_XO)));
    RC(lin_fwd(a.n_f, p.wb + p.lo.dec_wi[l], a.h, BT, ff, d, EPI_BF16_RELU, 1.f, s, dr(l, D_RELU), a.hbits));
    RC(lin_fwd(a.h, p.wb + p.lo.dec_wo[l], w.delta_dec, BT, d, ff, EPI_BF16, 1.f, s));
  }
  RC(rmsnorm_fwd(w.Y[3 * nl - 1], w.delta_dec, w.Y[3 * nl], p.wf + p.lo.dec_final_ln, w.hn, w.rstd_dec_final, BT, d,
                 p.eps, 1.f, s, dr(nl - 1, D_FF), dr(0, D_FINAL)));
  // tied LM head: logits = (h * d^-0.5) E^T  (HF:1107-1110), fp32 out
  const float alpha = 1.0f / sqrtf(static_cast<float>(d));
  RC(lin_fwd(w.hn, p.wb + p.lo.shared, bt->logits, BT, p.V, d, EPI_F32, alpha, s));
  if (bt->labels_are_decoder_ids) return LAKO_OK;  // logits only
  if (!bt->count_is_set) RC(count_valid(labels, BT, bt->stats, s));
  const size_t vpad = (static_cast<size_t>(p.V) + 7) / 8 * 8;
  RC(ce_fwd_bwd(bt->logits, labels, bt->stats, w.row_loss, train ? w.dlogits : nullptr, BT, p.V, p.V,
                static_cast<int>(vpad), s));
  return LAKO_OK;
}

// loss.backward(): accumulates into the flat gradient buffer g_f32 (caller zeroes it: model.zero_grad())
// Gradient-ready stages for overlapping the data-parallel all-reduce with the rest of backward.  When
// bt->stage_events holds LAKO_N_GRAD_STAGES cudaEvent_t: event 0 is recorded once the decoder-side gradients (flat
// range [dec_ln0[0], total)) are final, events 1..4 as each quarter of the encoder layers (top quarter first)
// becomes final, event 5 at the end (shared embedding, encoder bias table and final norm).
// lako_fidt5_grad_stages() reports the element ranges: ranges[i] for i < 5 belongs to stage i, ranges[5] and
// ranges[6] to stage 5.
extern "C" int lako_fidt5_grad_stages(const lako_dims* dm, size_t* ranges /* [7][2] begin,end */) {
  lako_layout lo;
  RC(lako_fidt5_layout(dm, &lo));
  const int nl = dm->n_layers;
  ranges[0] = lo.dec_ln0[0];
  ranges[1] = lo.total;
  for (int k = 0; k < 4; ++k) {
    const int b = enc_group_begin(nl, k);
    const int e = (k == 0) ? nl : enc_group_begin(nl, k - 1);
    ranges[2 + 2 * k] = lo.enc_ln0[b];
    ranges[3 + 2 * k] = (e >= nl) ? lo.enc_relbias : lo.enc_ln0[e];
    if (b >= e) ranges[3 + 2 * k] = ranges[2 + 2 * k];  // empty group (fewer than 4 layers)
  }
  ranges[10] = 0;
  ranges[11] = lo.enc_ln0[0];
  ranges[12] = lo.enc_relbias;
  ranges[13] = lo.dec_ln0[0];
  return LAKO_OK;
}

extern "C" int lako_fidt5_backward(const lako_fidt5* m, const lako_batch* bt, const float* loss_grad, void* stream) {
  RC(check_arch());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaEvent_t* stage_ev = reinterpret_cast<cudaEvent_t*>(bt ? bt->stage_events : nullptr);
  Ptrs p;
  RC(make_ptrs(m, p));
  LAKO_REQUIRE(p.g != nullptr && bt && bt->workspace, LAKO_ERR_BAD_ARG);
  Work w;
  carve(m->dims, bt->B, bt->N, bt->L, bt->T, bt->recompute ? 3 : 1, bt->workspace, w);
  LAKO_REQUIRE(w.bytes <= bt->workspace_bytes, LAKO_ERR_WORKSPACE);
  RC(set_packing(w, bt, s));
  const unsigned char* xmask = w.cu ? nullptr : bt->mask;
  const int d = p.d, inner = p.inner, ff = p.ff, nl = p.nl, BT = w.BT, Mt = w.Mrows, H = p.H;
  const size_t vpad = (static_cast<size_t>(p.V) + 7) / 8 * 8;
  const float alpha = 1.0f / sqrtf(static_cast<float>(d));
  float* g = p.g;
  Dropper dr{bt->dropout_seed, bt->dropout_step, bt->dropout_p};
  const float dsc = dr.p > 0.f ? dr(0, 0).scale : 1.f;  // 1/(1-p), same for every site

  cudaMemsetAsync(w.dlut_enc, 0, sizeof(float) * H * (2 * w.L - 1), s);
  cudaMemsetAsync(w.dlut_dec, 0, sizeof(float) * H * (2 * w.T - 1), s);
  cudaMemsetAsync(w.dsk_f32, 0, sizeof(float) * BT * d, s);

  // ---- LM head
  if (loss_grad) RC(scale_bf16_dev(w.dlogits, static_cast<size_t>(BT) * vpad, loss_grad, s));
  RC(lin_dgrad_splitk(w.dlogits, static_cast<int>(vpad), p.wb + p.lo.shared, w.dhn, w.dsk_f32, BT, p.V, d, alpha, s));
  RC(gemm_bf16(w.dlogits, w.hn, g + p.lo.shared, p.V, d, BT, static_cast<int>(vpad), d, d, 1, 1, EPI_F32_ATOMIC, alpha,
               nullptr, 0, 1, 0, s));
  RC(rmsnorm_bwd(w.dhn, w.Y[3 * nl], w.rstd_dec_final, p.wf + p.lo.dec_final_ln, w.dres_dec, w.dg[nl - 1].g0,
                 g + p.lo.dec_final_ln, BT, d, 0, dsc, s, dr(0, D_FINAL), dr(nl - 1, D_FF)));
  // ---- decoder layers: the dgrad chain only; the gradient operands are kept per layer and all layers' weight
  //      gradients are formed afterwards, one batched launch per weight kind
  for (int l = nl - 1; l >= 0; --l) {
    DecLayerAct& a = w.dec[l];
    DecGrad& dg = w.dg[l];
    // FFN
    RC(lin_dgrad(dg.g0, p.wb + p.lo.dec_wo[l], dg.dh, BT, d, ff, a.hbits, dsc, s));
    RC(lin_dgrad_splitk(dg.dh, ff, p.wb + p.lo.dec_wi[l], w.dn_dec, w.dsk_f32, BT, ff, d, 1.f, s));
    RC(rmsnorm_bwd(w.dn_dec, w.Y[3 * l + 2], a.rstd_f, p.wf + p.lo.dec_ln2[l], w.dres_dec, dg.g1,
                   g + p.lo.dec_ln2[l], BT, d, 1, 1.f, s, dr.none(), dr(l, D_XO)));
    // cross-attention
    RC(lin_dgrad(dg.g1, p.wb + p.lo.dec_xo[l], w.dao_dec, BT, d, inner, nullptr, 1.f, s));
    // dq: per-chunk partials in the (now idle) forward workspace, summed in chunk order -- no atomics, so the gradient
    // signal below this point does not depend on the order the chunk CTAs finish in
    RC(xattn_bwd(a.q, w.kv_all, xmask, a.xo, w.dao_dec, a.lse, w.dkv_all, nullptr, bt->B, bt->T, w.M, H, inner,
                 nl * 2 * inner, l * 2 * inner, l * 2 * inner + inner, inner, s, dr(l, D_XP), w.cu, bt->N, w.xattn_ws, dg.dq));
    RC(lin_dgrad(dg.dq, p.wb + p.lo.dec_xq[l], w.dn_dec, BT, inner, d, nullptr, 1.f, s));
    RC(rmsnorm_bwd(w.dn_dec, w.Y[3 * l + 1], a.rstd_c, p.wf + p.lo.dec_ln1[l], w.dres_dec, dg.g2,
                   g + p.lo.dec_ln1[l], BT, d, 1, 1.f, s, dr.none(), dr(l, D_O)));
    // self-attention
    RC(lin_dgrad(dg.g2, p.wb + p.lo.dec_o[l], w.dao_dec, BT, d, inner, nullptr, 1.f, s));
    RC(dself_bwd(a.qkv, w.dao_dec, a.probs, dg.dqkv, w.dlut_dec, bt->B, bt->T, H, s, dr(l, D_SELF)));
    RC(lin_dgrad_splitk(dg.dqkv, 3 * inner, p.wb + p.lo.dec_qkv[l], w.dn_dec, w.dsk_f32, BT, 3 * inner, d, 1.f, s));
    RC(rmsnorm_bwd(w.dn_dec, w.Y[3 * l], a.rstd_s, p.wf + p.lo.dec_ln0[l], w.dres_dec,
                   l > 0 ? w.dg[l - 1].g0 : w.dres_dec_b, g + p.lo.dec_ln0[l], BT, d, 1, 1.f, s, dr.none(),
                   l > 0 ? dr(l - 1, D_FF) : dr.none()));
  }
  {
    // dW[N][K] += dy[BT][N]^T x[BT][K] for every layer at once (each problem alone is a few microseconds of work
    // behind a fixed launch + pipeline-fill cost)
    const void* dy[LAKO_MAX_LAYERS];
    const void* xin[LAKO_MAX_LAYERS];
    float* dW[LAKO_MAX_LAYERS];
    auto batched = [&](auto get_dy, auto get_x, const size_t* off, int N, int K) -> int {
      for (int l = 0; l < nl; ++l) {
        dy[l] = get_dy(l);
        xin[l] = get_x(l);
        dW[l] = g + off[l];
      }
      return gemm_bf16_wgrad_batched(dy, xin, dW, nl, N, K, BT, N, K, K, 1.f, s);
    };
    RC(batched([&](int l) { return w.dg[l].g0; }, [&](int l) { return w.dec[l].h; }, p.lo.dec_wo, d, ff));
    RC(batched([&](int l) { return w.dg[l].dh; }, [&](int l) { return w.dec[l].n_f; }, p.lo.dec_wi, ff, d));
    RC(batched([&](int l) { return w.dg[l].g1; }, [&](int l) { return w.dec[l].xo; }, p.lo.dec_xo, d, inner));
    RC(batched([&](int l) { return w.dg[l].dq; }, [&](int l) { return w.dec[l].n_c; }, p.lo.dec_xq, inner, d));
    RC(batched([&](int l) { return w.dg[l].g2; }, [&](int l) { return w.dec[l].ao; }, p.lo.dec_o, d, inner));
    RC(batched([&](int l) { return w.dg[l].dqkv; }, [&](int l) { return w.dec[l].n_s; }, p.lo.dec_qkv, 3 * inner, d));
  }
  RC(embed_scatter_add(w.dec_ids, w.dres_dec, g + p.lo.shared, BT, d, p.V, s, dr(0, D_EMB)));
  RC(fold_dlut(w.dlut_dec, m->dec_buckets, g + p.lo.dec_relbias, H, w.T, m->bucket_center, m->dims.n_buckets, s));
  // ---- fused memory: all decoder layers' cross K/V projections at once
  const int nkv = nl * 2 * inner;
  RC(lin_dgrad(w.dkv_all, p.wb + p.lo.dec_xkv, w.dmem, Mt, nkv, d, nullptr, 1.f, s));
  RC(lin_wgrad(w.dkv_all, w.mem, g + p.lo.dec_xkv, Mt, nkv, d, 1.f, s));
  if (stage_ev) cudaEventRecord(stage_ev[0], s);
  RC(rmsnorm_bwd(w.dmem, w.X[2 * nl], w.rstd_enc_final, p.wf + p.lo.enc_final_ln, w.dres_enc, w.dres_enc_b,
                 g + p.lo.enc_final_ln, Mt, d, 0, dsc, s, dr(0, E_FINAL), dr(nl - 1, E_FF)));
  // ---- encoder layers
  for (int l = nl - 1; l >= 0; --l) {
    EncLayerAct& a = w.enc[l];
    if (w.recompute) RC(enc_layer_recompute(p, w, l, xmask, s, dr));
    RC(lin_dgrad(w.dres_enc_b, p.wb + p.lo.enc_wo[l], w.dh_enc, Mt, d, ff, a.hbits, dsc, s));
    RC(lin_wgrad(w.dres_enc_b, a.h, g + p.lo.enc_wo[l], Mt, d, ff, 1.f, s));
    RC(lin_dgrad(w.dh_enc, p.wb + p.lo.enc_wi[l], w.dn_enc, Mt, ff, d, nullptr, 1.f, s));
    RC(lin_wgrad(w.dh_enc, a.n_f, g + p.lo.enc_wi[l], Mt, ff, d, 1.f, s));
    RC(rmsnorm_bwd(w.dn_enc, w.X[2 * l + 1], a.rstd_f, p.wf + p.lo.enc_ln1[l], w.dres_enc, w.dres_enc_b,
                   g + p.lo.enc_ln1[l], Mt, d, 1, 1.f, s, dr.none(), dr(l, E_O)));
    RC(lin_dgrad(w.dres_enc_b, p.wb + p.lo.enc_o[l], w.dao_enc, Mt, d, inner, nullptr, 1.f, s));
    RC(lin_wgrad(w.dres_enc_b, a.ao, g + p.lo.enc_o[l], Mt, d, inner, 1.f, s));
    // D = rowsum(dO * O) once per (token, head), over the now dead attention output (its wgrad was the last reader)
    const bool d_pre = H <= 16;
    if (d_pre) RC(attn_rowdot_inplace(a.ao, w.dao_enc, Mt, H, s));
    RC(enc_attn_bwd(a.qkv, a.ao, w.dao_enc, a.lse, xmask, w.lut_enc, w.dqkv_enc, w.dlut_enc, w.P, w.L, H, s,
                    dr(l, E_ATTN), w.cu, Mt, a.akeep, d_pre));
    RC(lin_dgrad(w.dqkv_enc, p.wb + p.lo.enc_qkv[l], w.dn_enc, Mt, 3 * inner, d, nullptr, 1.f, s));
    RC(lin_wgrad(w.dqkv_enc, a.n_a, g + p.lo.enc_qkv[l], Mt, 3 * inner, d, 1.f, s));
    RC(rmsnorm_bwd(w.dn_enc, w.X[2 * l], a.rstd_a, p.wf + p.lo.enc_ln0[l], w.dres_enc, w.dres_enc_b,
                   g + p.lo.enc_ln0[l], Mt, d, 1, 1.f, s, dr.none(), l > 0 ? dr(l - 1, E_FF) : dr.none()));
    if (stage_ev)
      for (int k = 0; k < 4; ++k)
        if (l == enc_group_begin(nl, k) && (k == 0 || enc_group_begin(nl, k - 1) > l))
          cudaEventRecord(stage_ev[1 + k], s);  // encoder group k (layers >= l of that quarter) is final
  }
  RC(embed_scatter_add(reinterpret_cast<const long long*>(bt->ids), w.dres_enc, g + p.lo.shared, Mt, d, p.V, s,
                       dr(0, E_EMB), w.cu ? w.tok_index : nullptr));
  RC(fold_dlut(w.dlut_enc, m->enc_buckets, g + p.lo.enc_relbias, H, w.L, m->bucket_center, m->dims.n_buckets, s));
  if (stage_ev) cudaEventRecord(stage_ev[5], s);
  return LAKO_OK;
}

// ------------------------------------------------------------------------------------------------
// greedy generate with KV cache (+ optional step-0 cross-attention token scores)
// ------------------------------------------------------------------------------------------------
namespace {
struct GenWork {
  Work enc;  // encoder part (eval carve with T=1)
  bf16 *kc, *vc;  // [nl][B][H][Tmax][64]
  float *lut_dec, *logits, *xdec_ws, *tok_first, *tok_last, *q_f32;
  int *unfinished, *n_unf, *d_step, *d_done, *cu_copy;
  unsigned char* mask_copy;
  long long *next_ids, *out_ids;
  size_t bytes;
};
void carve_gen(const lako_dims& dm, int B, int N, int L, int Tmax, void* base, GenWork& gw) {
  carve(dm, B, N, L, 1, 0, base, gw.enc);
  Bump b(base);
  b.off = gw.enc.bytes;
  const size_t H = dm.n_heads, nl = dm.n_layers;
  gw.kc = b.take<bf16>(nl * B * H * Tmax * 64);
  gw.vc = b.take<bf16>(nl * B * H * Tmax * 64);
  gw.lut_dec = b.take<float>(H * (2 * Tmax - 1));
  gw.logits = b.take<float>(static_cast<size_t>(B) * dm.vocab);
  gw.unfinished = b.take<int>(B);
  gw.n_unf = b.take<int>(1);
  gw.d_step = b.take<int>(1);
  gw.d_done = b.take<int>(1);
  gw.next_ids = b.take<long long>(B);
  gw.out_ids = b.take<long long>(static_cast<size_t>(B) * Tmax);
  gw.cu_copy = b.take<int>(static_cast<size_t>(B) * N + 1);
  gw.mask_copy = b.take<unsigned char>(static_cast<size_t>(B) * N * L);
  gw.tok_first = b.take<float>(static_cast<size_t>(B) * N * L);
  gw.tok_last = b.take<float>(static_cast<size_t>(B) * N * L);
  gw.xdec_ws = b.take<float>(xattn_decode_workspace_bytes(B, dm.n_heads, N * L) / sizeof(float) + 1);
  gw.q_f32 = b.take<float>(static_cast<size_t>(B) * H * 64);
  gw.bytes = (b.off + 255) & ~size_t(255);
}
__global__ void gen_init_kernel(int* unfinished, int* n_unf, int* d_step, int* d_done, long long* next_ids,
                                long long* out_ids, int B, int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    unfinished[i] = 1;
    next_ids[i] = 0;  // decoder_start_token_id = pad = 0
    for (int t = 0; t < out_ld; ++t) out_ids[static_cast<size_t>(i) * out_ld + t] = 0;
  }
  if (i == 0) {
    *n_unf = B;
    *d_step = 0;
    *d_done = 0;
  }
}

struct DecodeCtx {
  const Ptrs* p;
  GenWork* gw;
  int B, N, max_length, min_length;
  bool packed, capture;
};

// One cached decoder pass over the last generated token of every question (HF T5Block with use_cache, HF:411-498):
// 10 launches per layer -- fused RMSNorm -> projection GEMVs, KV-cached self-attention, split cross-attention over the
// packed memory (+ step-0 score capture), residual-add epilogues -- then the tied LM head and the greedy pick.  Every
// step-dependent quantity (cache position, output column, EOS ban, "is this step 0") is read from the device counter
// d_step, so the enqueued work is identical for every step and can be replayed from a captured CUDA graph.
int enqueue_decode_step(const DecodeCtx& c, cudaStream_t s) {
  const Ptrs& p = *c.p;
  GenWork& gw = *c.gw;
  Work& w = gw.enc;
  const int d = p.d, inner = p.inner, ff = p.ff, nl = p.nl, B = c.B, H = p.H;
  const int first_half = (nl + 1) / 2;
  const size_t cache_l = static_cast<size_t>(B) * H * c.max_length * 64;
  const int* cu = c.packed ? gw.cu_copy : nullptr;
  const unsigned char* xmask = c.packed ? nullptr : gw.mask_copy;
  DecLayerAct& a = w.dec[0];
  float* Y = w.Y[0];
  RC(embed_gather(gw.next_ids, p.wf + p.lo.shared, Y, B, d, p.V, s));
  // Weight prefetch to L2 (prefetch.global.L2 from the first instructions of a Linear): the Linear before the
  // cross-attention asks for the three weights after it (xo, wi, wo: one contiguous range of the flat buffer) -- they
  // arrive while the cross-attention streams K / V, whose own loads are evict_first -- and the last Linear of a layer for
  // the next layer's qkv .. xq range.  A launch here otherwise lives 5-8 us, mostly the latency of its first bytes.
  auto range_bytes = [&](size_t first, size_t last, size_t last_elems) { return (last + last_elems - first) * 2; };
  for (int l = 0; l < nl; ++l) {
    RC(smallm_linear(Y, nullptr, 0, p.wf + p.lo.dec_ln0[l], p.eps, p.wb + p.lo.dec_qkv[l], a.qkv, 3 * inner, B, 3 * inner,
                     d, SL_BF16, 1.f, s));
    RC(dself_step(a.qkv, gw.kc + l * cache_l, gw.vc + l * cache_l, gw.lut_dec, a.ao, B, 0, c.max_length, H, s, gw.d_step));
    RC(smallm_linear(nullptr, a.ao, inner, nullptr, 0.f, p.wb + p.lo.dec_o[l], Y, d, B, d, inner, SL_F32_ADD, 1.f, s));
    // q accumulates (K-split, atomic) into the zeroed fp32 buffer; the cross-attention kernel clears it again after use
    RC(smallm_linear(Y, nullptr, 0, p.wf + p.lo.dec_ln1[l], p.eps, p.wb + p.lo.dec_xq[l], gw.q_f32, inner, B, inner, d,
                     SL_F32_ADD, 1.f, s, p.wb + p.lo.dec_xo[l],
                     range_bytes(p.lo.dec_xo[l], p.lo.dec_wo[l], static_cast<size_t>(d) * ff)));
    float* ts = c.capture ? ((l < first_half) ? gw.tok_first : gw.tok_last) : nullptr;
    RC(xattn_decode(gw.q_f32, w.kv_all, static_cast<size_t>(w.Mtok), xmask, a.xo, ts, gw.d_step, gw.xdec_ws, B, w.M, H,
                    inner, nl * 2 * inner, l * 2 * inner, l * 2 * inner + inner, inner, s, cu, c.N,
                    c.packed ? w.tok_index : nullptr, true, true));
    RC(smallm_linear(nullptr, a.xo, inner, nullptr, 0.f, p.wb + p.lo.dec_xo[l], Y, d, B, d, inner, SL_F32_ADD, 1.f, s));
    RC(smallm_linear(Y, nullptr, 0, p.wf + p.lo.dec_ln2[l], p.eps, p.wb + p.lo.dec_wi[l], a.h, ff, B, ff, d, SL_BF16_RELU,
                     1.f, s));
    const bool more = l + 1 < nl;
    RC(smallm_linear(nullptr, a.h, ff, nullptr, 0.f, p.wb + p.lo.dec_wo[l], Y, d, B, d, ff, SL_F32_ADD, 1.f, s,
                     more ? p.wb + p.lo.dec_qkv[l + 1] : nullptr,
                     more ? range_bytes(p.lo.dec_qkv[l + 1], p.lo.dec_xq[l + 1], static_cast<size_t>(inner) * d) : 0));
  }
  // final norm + tied LM head: logits = (norm(h) * d^-0.5) E^T  (HF:1107-1110), fp32
  RC(smallm_linear(Y, nullptr, 0, p.wf + p.lo.dec_final_ln, p.eps, p.wb + p.lo.shared, gw.logits, p.V, B, p.V, d, SL_F32,
                   1.0f / sqrtf(static_cast<float>(d)), s));
  RC(greedy_pick(gw.logits, B, p.V, p.V, gw.out_ids, c.max_length, 0, gw.unfinished, gw.n_unf, gw.next_ids, s, 0,
                 gw.d_step, gw.d_done, c.min_length));
  return LAKO_OK;
}

// One cached cudaGraphExec_t per host thread: a decode step replayed max_length - 2 times per call and unchanged
// between calls as long as the buffers it was captured on are the same (the Python side keeps one workspace per shape).
struct GraphKey {
  const void *wf, *wb, *ws;
  int d, H, ff, nl, V, B, N, L, max_length, min_length, packed, capture, precision;
  bool operator==(const GraphKey& o) const {
    return wf == o.wf && wb == o.wb && ws == o.ws && d == o.d && H == o.H && ff == o.ff && nl == o.nl && V == o.V &&
           B == o.B && N == o.N && L == o.L && max_length == o.max_length && min_length == o.min_length &&
           packed == o.packed && capture == o.capture && precision == o.precision;
  }
};
struct DecodeGraphCache {
  bool valid = false, broken = false;
  GraphKey key{};
  cudaGraphExec_t exec = nullptr;
  long long kernels = 0;  // kernel nodes of the captured step
  int* h_unf = nullptr;  // pinned: n_unfinished after every step
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t stream = nullptr;  // the decode loop's own stream: the caller's may be the legacy default stream, which
  cudaEvent_t ev_in = nullptr;    // cannot be captured into a graph
};
thread_local DecodeGraphCache g_dgraph;
long long g_graph_captures = 0, g_graph_launches = 0, g_graph_failures = 0;

struct GenWorkF32 {
  WorkF32 w;
  float *kc, *vc, *logits, *tok_first, *tok_last;
  int *unfinished, *n_unf, *d_step, *d_done;
  long long *next_ids, *out_ids;
  size_t bytes;
};
void carve_gen_f32(const lako_dims& dm, int B, int N, int L, int Tmax, void* base, GenWorkF32& gw) {
  carve_f32(dm, B, N, L, 1, Tmax, base, gw.w);
  Bump b(base);
  b.off = gw.w.bytes;
  const size_t H = dm.n_heads, nl = dm.n_layers;
  gw.kc = b.take<float>(nl * B * H * Tmax * 64);
  gw.vc = b.take<float>(nl * B * H * Tmax * 64);
  gw.logits = b.take<float>(static_cast<size_t>(B) * dm.vocab);
  gw.tok_first = b.take<float>(static_cast<size_t>(B) * N * L);
  gw.tok_last = b.take<float>(static_cast<size_t>(B) * N * L);
  gw.unfinished = b.take<int>(B);
  gw.n_unf = b.take<int>(1);
  gw.d_step = b.take<int>(1);
  gw.d_done = b.take<int>(1);
  gw.next_ids = b.take<long long>(B);
  gw.out_ids = b.take<long long>(static_cast<size_t>(B) * Tmax);
  gw.bytes = (b.off + 255) & ~size_t(255);
}

int g_decode_graph = -1;
bool decode_graph_enabled() {
  if (g_decode_graph < 0) {
    const char* e = getenv("LAKO_DECODE_GRAPH");
    g_decode_graph = (e && e[0] == '0') ? 0 : 1;
  }
  return g_decode_graph != 0;
}
}  // namespace

extern "C" void lako_set_decode_graph(int on) { g_decode_graph = on ? 1 : 0; }
extern "C" void lako_decode_graph_stats(long long* captures, long long* launches, long long* failures) {
  if (captures) *captures = g_graph_captures;
  if (launches) *launches = g_graph_launches;
  if (failures) *failures = g_graph_failures;
}

extern "C" size_t lako_fidt5_generate_workspace_bytes(const lako_dims* dm, int B, int N, int L, int max_length,
                                                      int precision) {
  if (precision == 1) {
    GenWorkF32 gf;
    carve_gen_f32(*dm, B, N, L, max_length, nullptr, gf);
    return gf.bytes + 256;
  }
  GenWork gw;
  carve_gen(*dm, B, N, L, max_length, nullptr, gw);
  return gw.bytes + 256;
}

// out_ids int64 [B][max_length] (device), *out_len (host) = number of generated columns incl. the start
// token.  tok_score_first/last fp32 [B][N*L] (device, may be null): sum over heads and over the first
// ceil(nl/2) / the remaining decoder layers of the masked step-0 pre-softmax cross-attention scores.
// min_length: HF's MinLengthLogitsProcessor (EOS cannot be produced while the decoder length is below it); the
// reference never sets it (0), bench.py uses it to time full-length decodes with random-init weights.
// Host synchronisation: none per token.  The count of unfinished rows is copied to pinned host memory after every step
// and examined two steps later (HF's early stop, pipelined), and the stream is synchronised once at the end because
// *out_len is a host value.
extern "C" int lako_fidt5_generate(const lako_fidt5* m, const lako_batch* bt, int max_length, int min_length,
                                   int64_t* out_ids, int* out_len, float* tok_score_first, float* tok_score_last,
                                   void* stream) {
  RC(check_arch());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Ptrs p;
  RC(make_ptrs(m, p));
  LAKO_REQUIRE(bt && bt->ids && bt->mask && bt->workspace && out_ids && out_len, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(max_length >= 2 && max_length <= 1024 && max_length - 1 <= m->bucket_center && bt->L >= 1 &&
                   bt->L - 1 <= m->bucket_center && (bt->L <= ENC_ATTN_LONG_MAX_L || bt->precision == 1),
               LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE((tok_score_first == nullptr) == (tok_score_last == nullptr), LAKO_ERR_BAD_ARG);
  const int B = bt->B, H = p.H;
  const size_t ntok = static_cast<size_t>(B) * bt->N * bt->L;
  const bool capture = tok_score_first != nullptr;
  const bool f32 = bt->precision == 1;
  GenWork gw;
  GenWorkF32 gf;
  PtrsF32 pw{m->w_tf32_hi, m->w_tf32_lo};
  bool packed = false;
  int *d_n_unf = nullptr;
  long long* d_out_ids = nullptr;
  float *d_tok_first = nullptr, *d_tok_last = nullptr;
  if (f32) {
    LAKO_REQUIRE(m->w_tf32_hi && m->w_tf32_lo, LAKO_ERR_BAD_ARG);
    carve_gen_f32(m->dims, B, bt->N, bt->L, max_length, bt->workspace, gf);
    LAKO_REQUIRE(gf.bytes <= bt->workspace_bytes, LAKO_ERR_WORKSPACE);
    WorkF32& w = gf.w;
    if (bt->cu_seqlens) {
      LAKO_REQUIRE(bt->packed_tokens > 0 && bt->packed_tokens <= w.Mtok, LAKO_ERR_BAD_ARG);
      w.cu = bt->cu_seqlens;
      w.Mrows = bt->packed_tokens;
    }
    packed = w.cu != nullptr;
    if (capture) {
      cudaMemsetAsync(gf.tok_first, 0, sizeof(float) * ntok, s);
      cudaMemsetAsync(gf.tok_last, 0, sizeof(float) * ntok, s);
    }
    RC(encode_f32(p, pw, m, w, reinterpret_cast<const long long*>(bt->ids), bt->mask, s));
    RC(build_bias_lut(p.wf + p.lo.dec_relbias, m->dec_buckets, w.lut_dec, H, max_length, m->bucket_center, s));
    launch_pdl(gen_init_kernel, dim3((B + 127) / 128), dim3(128), 0, s, gf.unfinished, gf.n_unf, gf.d_step, gf.d_done,
               gf.next_ids, gf.out_ids, B, max_length);
    LAKO_CHECK_LAUNCH();
    d_n_unf = gf.n_unf;
    d_out_ids = gf.out_ids;
    d_tok_first = gf.tok_first;
    d_tok_last = gf.tok_last;
  } else {
    carve_gen(m->dims, bt->B, bt->N, bt->L, max_length, bt->workspace, gw);
    LAKO_REQUIRE(gw.bytes <= bt->workspace_bytes, LAKO_ERR_WORKSPACE);
    Work& w = gw.enc;
    RC(set_packing(w, bt, s));
    packed = w.cu != nullptr;
    if (packed)
      cudaMemcpyAsync(gw.cu_copy, bt->cu_seqlens, sizeof(int) * (static_cast<size_t>(B) * bt->N + 1), cudaMemcpyDeviceToDevice, s);
    else
      cudaMemcpyAsync(gw.mask_copy, bt->mask, ntok, cudaMemcpyDeviceToDevice, s);
    if (capture) {
      cudaMemsetAsync(gw.tok_first, 0, sizeof(float) * ntok, s);
      cudaMemsetAsync(gw.tok_last, 0, sizeof(float) * ntok, s);
    }
    cudaMemsetAsync(gw.xdec_ws, 0, xattn_decode_workspace_bytes(B, H, bt->N * bt->L), s);  // arrival counters
    cudaMemsetAsync(gw.q_f32, 0, sizeof(float) * B * H * 64, s);
    RC(encode(p, m, w, reinterpret_cast<const long long*>(bt->ids), bt->mask, false, s, Dropper{0, 0, 0.f}, true));
    if (w.Mrows < w.Mtok) {
      // the decode kernel's 128-row TMA boxes reach up to 127 rows past the last token of a head's slab: those rows
      // were never written (workspace memory) and must not hold NaN bit patterns (0 * NaN in the P V product)
      const size_t tail = static_cast<size_t>(std::min(128, w.Mtok - w.Mrows));
      cudaMemset2DAsync(w.kv_all + static_cast<size_t>(w.Mrows) * 64, static_cast<size_t>(w.Mtok) * 128, 0, tail * 128,
                        static_cast<size_t>(p.nl) * 2 * H, s);
    }
    RC(build_bias_lut(p.wf + p.lo.dec_relbias, m->dec_buckets, gw.lut_dec, H, max_length, m->bucket_center, s));
    launch_pdl(gen_init_kernel, dim3((B + 127) / 128), dim3(128), 0, s, gw.unfinished, gw.n_unf, gw.d_step, gw.d_done,
               gw.next_ids, gw.out_ids, B, max_length);
    LAKO_CHECK_LAUNCH();
    d_n_unf = gw.n_unf;
    d_out_ids = gw.out_ids;
    d_tok_first = gw.tok_first;
    d_tok_last = gw.tok_last;
  }
  // one fp32-mode decode step: embedding of the last token, cached decoder pass, greedy pick (no graph: this mode is
  // for numerics, its launches go straight to the stream)
  auto enqueue_step_f32 = [&]() -> int {
    WorkF32& w = gf.w;
    RC(embed_gather(gf.next_ids, p.wf + p.lo.shared, w.Y, B, p.d, p.V, s));
    RC(decoder_f32(p, pw, w, packed ? nullptr : bt->mask, gf.logits, bt->N, gf.kc, gf.vc, max_length, gf.d_step,
                   capture ? gf.tok_first : nullptr, capture ? gf.tok_last : nullptr, s));
    RC(greedy_pick(gf.logits, B, p.V, p.V, gf.out_ids, max_length, 0, gf.unfinished, gf.n_unf, gf.next_ids, s, 0, gf.d_step,
                   gf.d_done, min_length));
    return LAKO_OK;
  };

  DecodeGraphCache& gc = g_dgraph;
  if (!gc.h_unf) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&gc.h_unf), sizeof(int) * 1024, cudaHostAllocDefault) != cudaSuccess) {
      set_last_error("cudaHostAlloc failed", __FILE__, __LINE__);
      return LAKO_ERR_DRIVER;
    }
    for (auto& e : gc.ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&gc.ev_in, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&gc.stream, cudaStreamNonBlocking) != cudaSuccess) {
      set_last_error("cudaStreamCreate failed", __FILE__, __LINE__);
      return LAKO_ERR_DRIVER;
    }
  }
  // everything enqueued so far (encoder, K/V projection, init) is ordered before the decode loop, which runs on the
  // library's own stream and is joined (synchronised) before this function returns
  cudaStream_t s_user = s;
  cudaEventRecord(gc.ev_in, s_user);
  s = gc.stream;
  cudaStreamWaitEvent(s, gc.ev_in, 0);
  DecodeCtx ctx{&p, &gw, B, bt->N, max_length, min_length, packed, capture};
  const GraphKey key{p.wf, p.wb, bt->workspace, p.d, p.H, p.ff, p.nl, p.V, B, bt->N, bt->L, max_length, min_length,
                     packed ? 1 : 0, capture ? 1 : 0, bt->precision};
  const bool want_graph = decode_graph_enabled() && !gc.broken && !prof_enabled() && !f32;
  const int nsteps = max_length - 1;
  constexpr int LAG = 2;
  int stop_at = -1;  // first step after which no row is unfinished
  for (int step = 0; step < nsteps; ++step) {
    bool launched = false;
    if (want_graph && step >= 1) {
      if (!(gc.valid && gc.key == key)) {
        // (re)capture: step 0 ran directly, so every lazily configured kernel attribute is set by now
        if (gc.exec) cudaGraphExecDestroy(gc.exec);
        gc.exec = nullptr;
        gc.valid = false;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          const unsigned long long n0 = launch_count_value();
          const int rc = enqueue_decode_step(ctx, s);
          gc.kernels = static_cast<long long>(launch_count_value() - n0);
          count_launches(-gc.kernels);  // captured, not executed
          const cudaError_t ee = cudaStreamEndCapture(s, &graph);
          if (rc == LAKO_OK && ee == cudaSuccess && graph &&
              cudaGraphInstantiate(&gc.exec, graph, 0) == cudaSuccess) {
            gc.valid = true;
            gc.key = key;
            ++g_graph_captures;
          }
          if (graph) cudaGraphDestroy(graph);
        }
        if (!gc.valid) {
          cudaGetLastError();
          gc.broken = true;  // this driver cannot capture the step: plain stream launches from now on
          ++g_graph_failures;
        }
      }
      if (gc.valid) {
        if (cudaGraphLaunch(gc.exec, s) != cudaSuccess) {
          set_last_error("cudaGraphLaunch failed", __FILE__, __LINE__);
          return LAKO_ERR_LAUNCH;
        }
        count_launches(gc.kernels);
        ++g_graph_launches;
        launched = true;
      }
    }
    if (!launched) {
      if (f32)
        RC(enqueue_step_f32());
      else
        RC(enqueue_decode_step(ctx, s));
    }
    cudaMemcpyAsync(&gc.h_unf[step], d_n_unf, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaEventRecord(gc.ev[step & 3], s);
    if (step >= LAG) {
      cudaEventSynchronize(gc.ev[(step - LAG) & 3]);
      if (gc.h_unf[step - LAG] == 0) {
        stop_at = step - LAG;
        break;
      }
    }
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
    return LAKO_ERR_LAUNCH;
  }
  if (stop_at < 0)
    for (int step = 0; step < nsteps; ++step)
      if (gc.h_unf[step] == 0) {
        stop_at = step;
        break;
      }
  // columns: start token + one per executed step up to and including the step that finished the last row; steps
  // launched beyond it only wrote pad ids (every row was finished), exactly what HF's early stop leaves out
  *out_len = (stop_at >= 0) ? stop_at + 2 : max_length;
  s = s_user;  // the decode stream is idle (synchronised above): results are copied out on the caller's stream
  cudaMemcpyAsync(out_ids, d_out_ids, sizeof(long long) * B * max_length, cudaMemcpyDeviceToDevice, s);
  if (capture) {
    cudaMemcpyAsync(tok_score_first, d_tok_first, sizeof(float) * ntok, cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(tok_score_last, d_tok_last, sizeof(float) * ntok, cudaMemcpyDeviceToDevice, s);
  }
  return LAKO_OK;
}
