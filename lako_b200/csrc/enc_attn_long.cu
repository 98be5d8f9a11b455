// T5 encoder self-attention for passages longer than the 256 tokens the tensor-memory kernels of enc_attn.cu hold in one
// softmax tile (HF:modeling_t5.py:253-344 via src/model.py:227-234; the reference takes any --text_maxlength,
// src/options.py:60).  Same contract as enc_attn_fwd / enc_attn_bwd -- bf16 q|k|v rows in, no 1/sqrt(dk), bucketed
// relative bias from the [H][2L-1] table, additive finfo.min key mask, dropout on the probabilities with the keep words
// of common.cuh drop_keep32_sliced -- for any L <= 2048, as 64 x 64 tiles with an online softmax in fp32 SIMT arithmetic.
// This is the coverage path, not the hot path: the reader's headline shapes (L = 200 / 250) never come here, and nothing
// below is tuned beyond conflict-free shared-memory tiles.  Three kernels:
//   forward   one CTA per (64-query tile, head, passage), loops over key tiles; writes O (bf16) and the row lse
//   dq        same grid; recomputes P from lse, forms dS, accumulates dQ = dS K and the per-offset bias gradient
//   dkv       one CTA per (64-key tile, head, passage), loops over query tiles; dV = Pd^T dO, dK = dS^T Q
// dq and dkv each recompute S and dO V^T (7 tile products instead of the minimal 5) so that neither needs atomics on
// the bf16 gradient rows and both are deterministic; only the bias gradient uses (fp32) atomics, as in enc_attn.cu.
#include "ops.cuh"

namespace lako {

namespace {

constexpr int LT = 64;   // tile edge (queries, keys, head dim)
constexpr int LPT = 68;  // pitch of the transposed tiles (floats): float4 reads along the 64 index, conflict-free
constexpr float LONG_MASK_MIN = -3.4028234663852886e38f;  // torch.finfo(float32).min, HF's additive mask

struct LongAttnParams {
  const __nv_bfloat16* qkv;   // [rows][3*inner]
  const unsigned char* mask;  // [P][L] (padded layout) or null
  const float* lut;           // [H][2L-1]
  __nv_bfloat16* out;         // forward: [rows][inner]
  float* lse;                 // [P][H][L]
  const __nv_bfloat16* o_in;  // backward: forward output (or D in its first H floats per row when d_in_out)
  const __nv_bfloat16* dout;  // [rows][inner]
  __nv_bfloat16* dqkv;        // [rows][3*inner]
  float* dlut;                // [H][2L-1] +=
  const int* cu;              // packed layout: passage p owns rows [cu[p], cu[p+1])
  int L, H, inner, d_in_out;
  unsigned kstride;           // Philox counters per query row (4 per 32-key chunk)
  DropSpec drop;
};

// 64 x 64 bf16 tile at `base` (row pitch ld elements) -> fp32 shared memory, transposed (dstT[c * LPT + r]) and / or
// row-major (dstR[r * 64 + c]); rows >= nvalid read as zero
__device__ __forceinline__ void load_tile(const __nv_bfloat16* base, size_t ld, int nvalid, float* dstT, float* dstR, int tid) {
  for (int idx = tid; idx < 512; idx += 256) {
    const int r = idx & 63, u = idx >> 6;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < nvalid) v = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(r) * ld + u * 8));
    const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y), bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
    if (dstT) {
#pragma unroll
      for (int e = 0; e < 8; ++e) dstT[(u * 8 + e) * LPT + r] = f[e];
    }
    if (dstR) {
      *reinterpret_cast<float4*>(dstR + r * 64 + u * 8) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(dstR + r * 64 + u * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
}

// acc[i][j] += sum_k A[k * pa + ty * 4 + i] * B[k * pb + tx * 4 + j]
__device__ __forceinline__ void mm64(const float* A, int pa, const float* B, int pb, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll 8
  for (int k = 0; k < 64; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(A + k * pa + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(B + k * pb + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

__device__ __forceinline__ void zero44(float (&a)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// keep words of a 64-query x 64-key tile: sKeep[r * 2 + w] for query q0 + r and 32-key chunk kt * 2 + w (threads 0..127)
__device__ __forceinline__ void tile_keep_words(const LongAttnParams& p, int psg, int h, int q0, int kt, int Lp, uint32_t* sKeep, int tid) {
  if (tid < 128) {
    const int r = tid >> 1, w = tid & 1;
    const int qc = min(q0 + r, Lp - 1);
    const unsigned long long row = static_cast<unsigned long long>(psg * p.H + h) * p.L + qc;
    sKeep[tid] = drop_keep32_sliced(p.drop, row * p.kstride + static_cast<unsigned>(kt * 2 + w) * 4u);
  }
}
__device__ __forceinline__ bool keep_of(const uint32_t* sKeep, int r, int kl) {
  return (sKeep[r * 2 + (kl >> 5)] >> sliced_bit(kl & 31)) & 1u;
}

// additive key term of a key tile: 0 / finfo.min by the padding mask, -inf beyond the passage
__device__ __forceinline__ void tile_key_mask(const LongAttnParams& p, int psg, int kt, int Lp, float* sKm, int tid) {
  if (tid < 64) {
    const int key = kt * LT + tid;
    float km = -INFINITY;
    if (key < Lp) km = (p.mask == nullptr || p.mask[static_cast<size_t>(psg) * p.L + key]) ? 0.f : LONG_MASK_MIN;
    sKm[tid] = km;
  }
}

// per-row lse and D = sum_d dO * O of a query tile (rows beyond the passage: lse = +inf so that P = 0, D = 0)
__device__ __forceinline__ void tile_row_stats(const LongAttnParams& p, int psg, int h, int row0, int q0, int Lp, float* sLse,
                                               float* sD, int tid) {
  if (tid < 64) {
    const int q = q0 + tid;
    sLse[tid] = (q < Lp) ? p.lse[(static_cast<size_t>(psg) * p.H + h) * p.L + q] : INFINITY;
    if (p.d_in_out)
      sD[tid] = (q < Lp) ? reinterpret_cast<const float*>(p.o_in + static_cast<size_t>(row0 + q) * p.inner)[h] : 0.f;
  }
  if (!p.d_in_out) {
    for (int idx = tid; idx < 512; idx += 256) {  // 8 consecutive lanes share a row
      const int r = idx >> 3, u = idx & 7;
      float acc = 0.f;
      if (q0 + r < Lp) {
        const size_t off = static_cast<size_t>(row0 + q0 + r) * p.inner + h * 64 + u * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.o_in + off)), b = __ldg(reinterpret_cast<const uint4*>(p.dout + off));
        acc = bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
              bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (u == 0) sD[r] = acc;
    }
  }
}

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* dst, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int LF_SMEM_FLOATS = 3 * 64 * LPT + 64 * 64 + 64 + 128;  // + (2L-1) LUT
template <bool kDrop>
__global__ void __launch_bounds__(256) enc_attn_long_fwd_kernel(const LongAttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float lsm[];
  float* sQt = lsm;              // [d][row]
  float* sKt = sQt + 64 * LPT;   // [d][key]
  float* sPt = sKt + 64 * LPT;   // [key][row]
  float* sV = sPt + 64 * LPT;    // [key][d]
  float* sKm = sV + 64 * 64;
  uint32_t* sKeep = reinterpret_cast<uint32_t*>(sKm + 64);
  float* sLut = reinterpret_cast<float*>(sKeep + 128);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int q0 = blockIdx.x * LT, h = blockIdx.y, psg = blockIdx.z;
  const int Lmax = p.L;
  int row0 = psg * Lmax, Lp = Lmax;
  if (p.cu) {
    row0 = p.cu[psg];
    Lp = p.cu[psg + 1] - row0;
  }
  if (q0 >= Lp) return;
  const size_t ld = 3 * static_cast<size_t>(p.inner);
  for (int i = tid; i < 2 * Lmax - 1; i += 256) sLut[i] = p.lut[static_cast<size_t>(h) * (2 * Lmax - 1) + i];
  load_tile(p.qkv + static_cast<size_t>(row0 + q0) * ld + h * 64, ld, Lp - q0, sQt, nullptr, tid);
  float m[4], l[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
  }
  zero44(o);
  const int ktiles = (Lp + LT - 1) / LT;
  for (int kt = 0; kt < ktiles; ++kt) {
    __syncthreads();  // previous tile consumed (and Q / LUT staged before first use)
    const __nv_bfloat16* kbase = p.qkv + static_cast<size_t>(row0 + kt * LT) * ld + h * 64;
    load_tile(kbase + p.inner, ld, Lp - kt * LT, sKt, nullptr, tid);
    load_tile(kbase + 2 * p.inner, ld, Lp - kt * LT, nullptr, sV, tid);
    tile_key_mask(p, psg, kt, Lp, sKm, tid);
    if (kDrop) tile_keep_words(p, psg, h, q0, kt, Lp, sKeep, tid);
    __syncthreads();
    float s[4][4];
    zero44(s);
    mm64(sQt, LPT, sKt, LPT, ty, tx, s);
    float mt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qpos = min(q0 + ty * 4 + i, Lp - 1);
      mt[i] = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kl = tx * 4 + j;
        const int kpos = min(kt * LT + kl, Lmax - 1);
        s[i][j] += sLut[kpos - qpos + Lmax - 1] + sKm[kl];
        mt[i] = fmaxf(mt[i], s[i][j]);
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) mt[i] = fmaxf(mt[i], __shfl_xor_sync(0xffffffffu, mt[i], off));
    }
    float pr[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float m_new = fmaxf(m[i], mt[i]);
      const float sc = (m[i] == -INFINITY) ? 0.f : expf(m[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float e = expf(s[i][j] - m_new);  // -inf (beyond the passage) -> 0
        rs += e;  // the normaliser is the sum before dropout (HF:331-332: softmax, then dropout)
        pr[i][j] = (!kDrop || keep_of(sKeep, ty * 4 + i, tx * 4 + j)) ? e : 0.f;
        o[i][j] *= sc;
      }
      l[i] = l[i] * sc + rs;
      m[i] = m_new;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(sPt + (tx * 4 + j) * LPT + ty * 4) = make_float4(pr[0][j], pr[1][j], pr[2][j], pr[3][j]);
    __syncthreads();
    mm64(sPt, LPT, sV, 64, ty, tx, o);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) l[i] += __shfl_xor_sync(0xffffffffu, l[i], off);
    const int q = q0 + ty * 4 + i;
    if (q >= Lp) continue;
    const float inv = (kDrop ? p.drop.scale : 1.f) / l[i];
    store_bf16x4(p.out + static_cast<size_t>(row0 + q) * p.inner + h * 64 + tx * 4, o[i][0] * inv, o[i][1] * inv, o[i][2] * inv,
                 o[i][3] * inv);
    if (p.lse && tx == 0) p.lse[(static_cast<size_t>(psg) * p.H + h) * Lmax + q] = m[i] + logf(l[i]);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ, dbias
constexpr int LQ_SMEM_FLOATS = 5 * 64 * LPT + 64 * 64 + 3 * 64 + 128;  // + 2 (2L-1): LUT and its gradient
template <bool kDrop>
__global__ void __launch_bounds__(256) enc_attn_long_dq_kernel(const LongAttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float lsm[];
  float* sQt = lsm;               // [d][row]
  float* sdOt = sQt + 64 * LPT;   // [d][row]
  float* sKt = sdOt + 64 * LPT;   // [d][key]
  float* sVt = sKt + 64 * LPT;    // [d][key]
  float* sdSt = sVt + 64 * LPT;   // [key][row]
  float* sK = sdSt + 64 * LPT;    // [key][d]
  float* sKm = sK + 64 * 64;
  float* sLse = sKm + 64;
  float* sD = sLse + 64;
  uint32_t* sKeep = reinterpret_cast<uint32_t*>(sD + 64);
  float* sLut = reinterpret_cast<float*>(sKeep + 128);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int q0 = blockIdx.x * LT, h = blockIdx.y, psg = blockIdx.z;
  const int Lmax = p.L, nlut = 2 * Lmax - 1;
  float* sDl = sLut + nlut;
  int row0 = psg * Lmax, Lp = Lmax;
  if (p.cu) {
    row0 = p.cu[psg];
    Lp = p.cu[psg + 1] - row0;
  }
  if (q0 >= Lp) return;
  const size_t ld = 3 * static_cast<size_t>(p.inner);
  for (int i = tid; i < nlut; i += 256) {
    sLut[i] = p.lut[static_cast<size_t>(h) * nlut + i];
    sDl[i] = 0.f;
  }
  load_tile(p.qkv + static_cast<size_t>(row0 + q0) * ld + h * 64, ld, Lp - q0, sQt, nullptr, tid);
  load_tile(p.dout + static_cast<size_t>(row0 + q0) * p.inner + h * 64, p.inner, Lp - q0, sdOt, nullptr, tid);
  tile_row_stats(p, psg, h, row0, q0, Lp, sLse, sD, tid);
  float dq[4][4];
  zero44(dq);
  const float keep_scale = kDrop ? p.drop.scale : 1.f;
  const int ktiles = (Lp + LT - 1) / LT;
  for (int kt = 0; kt < ktiles; ++kt) {
    __syncthreads();
    const __nv_bfloat16* kbase = p.qkv + static_cast<size_t>(row0 + kt * LT) * ld + h * 64;
    load_tile(kbase + p.inner, ld, Lp - kt * LT, sKt, sK, tid);
    load_tile(kbase + 2 * p.inner, ld, Lp - kt * LT, sVt, nullptr, tid);
    tile_key_mask(p, psg, kt, Lp, sKm, tid);
    if (kDrop) tile_keep_words(p, psg, h, q0, kt, Lp, sKeep, tid);
    __syncthreads();
    float s[4][4], dp[4][4];
    zero44(s);
    zero44(dp);
    mm64(sQt, LPT, sKt, LPT, ty, tx, s);
    mm64(sdOt, LPT, sVt, LPT, ty, tx, dp);
    float diag[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // this thread's dS summed along j - i
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const int qpos = min(q0 + r, Lp - 1);
      const float lse = sLse[r], D = sD[r];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kl = tx * 4 + j;
        const int kpos = min(kt * LT + kl, Lmax - 1);
        const float pv = expf(s[i][j] + sLut[kpos - qpos + Lmax - 1] + sKm[kl] - lse);
        const float kf = (!kDrop || keep_of(sKeep, r, kl)) ? keep_scale : 0.f;
        const float ds = pv * (kf * dp[i][j] - D);
        s[i][j] = ds;
        diag[j - i + 3] += ds;
      }
    }
    if (p.dlut) {
      const int base = (kt * LT + tx * 4) - (q0 + ty * 4) + Lmax - 1;
#pragma unroll
      for (int d = 0; d < 7; ++d) {
        const int idx = base + d - 3;
        if (idx >= 0 && idx < nlut && diag[d] != 0.f) atomicAdd(sDl + idx, diag[d]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(sdSt + (tx * 4 + j) * LPT + ty * 4) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();
    mm64(sdSt, LPT, sK, 64, ty, tx, dq);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q < Lp) store_bf16x4(p.dqkv + static_cast<size_t>(row0 + q) * ld + h * 64 + tx * 4, dq[i][0], dq[i][1], dq[i][2], dq[i][3]);
  }
  if (p.dlut) {
    __syncthreads();
    for (int i = tid; i < nlut; i += 256)
      if (sDl[i] != 0.f) atomicAdd(p.dlut + static_cast<size_t>(h) * nlut + i, sDl[i]);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
constexpr int LK_SMEM_FLOATS = 6 * 64 * LPT + 2 * 64 * 64 + 3 * 64 + 128;  // + (2L-1) LUT
template <bool kDrop>
__global__ void __launch_bounds__(256) enc_attn_long_dkv_kernel(const LongAttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float lsm[];
  float* sKt = lsm;               // [d][key]
  float* sVt = sKt + 64 * LPT;    // [d][key]
  float* sQt = sVt + 64 * LPT;    // [d][row]
  float* sdOt = sQt + 64 * LPT;   // [d][row]
  float* sPd = sdOt + 64 * LPT;   // [row][key]  dropped probabilities
  float* sdS = sPd + 64 * LPT;    // [row][key]
  float* sQ = sdS + 64 * LPT;     // [row][d]
  float* sdO = sQ + 64 * 64;      // [row][d]
  float* sKm = sdO + 64 * 64;
  float* sLse = sKm + 64;
  float* sD = sLse + 64;
  uint32_t* sKeep = reinterpret_cast<uint32_t*>(sD + 64);
  float* sLut = reinterpret_cast<float*>(sKeep + 128);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;  // ty: keys, tx: query rows
  const int kt = blockIdx.x, h = blockIdx.y, psg = blockIdx.z;
  const int Lmax = p.L, nlut = 2 * Lmax - 1;
  int row0 = psg * Lmax, Lp = Lmax;
  if (p.cu) {
    row0 = p.cu[psg];
    Lp = p.cu[psg + 1] - row0;
  }
  if (kt * LT >= Lp) return;
  const size_t ld = 3 * static_cast<size_t>(p.inner);
  for (int i = tid; i < nlut; i += 256) sLut[i] = p.lut[static_cast<size_t>(h) * nlut + i];
  const __nv_bfloat16* kbase = p.qkv + static_cast<size_t>(row0 + kt * LT) * ld + h * 64;
  load_tile(kbase + p.inner, ld, Lp - kt * LT, sKt, nullptr, tid);
  load_tile(kbase + 2 * p.inner, ld, Lp - kt * LT, sVt, nullptr, tid);
  tile_key_mask(p, psg, kt, Lp, sKm, tid);
  float dk[4][4], dv[4][4];
  zero44(dk);
  zero44(dv);
  const float keep_scale = kDrop ? p.drop.scale : 1.f;
  const int qtiles = (Lp + LT - 1) / LT;
  for (int qt = 0; qt < qtiles; ++qt) {
    const int q0 = qt * LT;
    __syncthreads();
    load_tile(p.qkv + static_cast<size_t>(row0 + q0) * ld + h * 64, ld, Lp - q0, sQt, sQ, tid);
    load_tile(p.dout + static_cast<size_t>(row0 + q0) * p.inner + h * 64, p.inner, Lp - q0, sdOt, sdO, tid);
    tile_row_stats(p, psg, h, row0, q0, Lp, sLse, sD, tid);
    if (kDrop) tile_keep_words(p, psg, h, q0, kt, Lp, sKeep, tid);
    __syncthreads();
    float s[4][4], dp[4][4];  // [key i][row j]
    zero44(s);
    zero44(dp);
    mm64(sKt, LPT, sQt, LPT, ty, tx, s);
    mm64(sVt, LPT, sdOt, LPT, ty, tx, dp);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = tx * 4 + j;
      const int qpos = min(q0 + r, Lp - 1);
      const float lse = sLse[r], D = sD[r];
      float pd[4], ds[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kl = ty * 4 + i;
        const int kpos = min(kt * LT + kl, Lmax - 1);
        const float pv = expf(s[i][j] + sLut[kpos - qpos + Lmax - 1] + sKm[kl] - lse);
        const float kf = (!kDrop || keep_of(sKeep, r, kl)) ? keep_scale : 0.f;
        pd[i] = pv * kf;
        ds[i] = pv * (kf * dp[i][j] - D);
      }
      *reinterpret_cast<float4*>(sPd + r * LPT + ty * 4) = make_float4(pd[0], pd[1], pd[2], pd[3]);
      *reinterpret_cast<float4*>(sdS + r * LPT + ty * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    __syncthreads();
    mm64(sPd, LPT, sdO, 64, ty, tx, dv);
    mm64(sdS, LPT, sQ, 64, ty, tx, dk);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = kt * LT + ty * 4 + i;
    if (key >= Lp) continue;
    __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(row0 + key) * ld + h * 64 + tx * 4;
    store_bf16x4(dst + p.inner, dk[i][0], dk[i][1], dk[i][2], dk[i][3]);
    store_bf16x4(dst + 2 * p.inner, dv[i][0], dv[i][1], dv[i][2], dv[i][3]);
  }
}

template <typename K>
int configure_smem(K kern, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) {
    set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
    return LAKO_ERR_LAUNCH;
  }
  return LAKO_OK;
}

}  // namespace

unsigned enc_attn_keep_stride(int L) { return 4u * static_cast<unsigned>(std::max(8, (L + 31) / 32)); }

int enc_attn_long_fwd(const __nv_bfloat16* qkv, const unsigned char* mask, const float* bias_lut, __nv_bfloat16* out, float* lse,
                      int P, int L, int H, cudaStream_t s, DropSpec drop, const int* cu) {
  int rc = check_arch();
  if (rc) return rc;
  LAKO_REQUIRE(P > 0 && H > 0 && L > 0 && qkv && bias_lut && out, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(L <= ENC_ATTN_LONG_MAX_L && P <= 65535 && H <= 65535, LAKO_ERR_UNSUPPORTED);
  LongAttnParams p{};
  p.qkv = qkv;
  p.mask = mask;
  p.lut = bias_lut;
  p.out = out;
  p.lse = lse;
  p.cu = cu;
  p.L = L;
  p.H = H;
  p.inner = H * 64;
  p.kstride = enc_attn_keep_stride(L);
  p.drop = drop;
  const size_t smem = (LF_SMEM_FLOATS + 2 * L - 1) * sizeof(float);
  auto kern = drop.thresh ? enc_attn_long_fwd_kernel<true> : enc_attn_long_fwd_kernel<false>;
  rc = configure_smem(kern, smem);
  if (rc) return rc;
  ProfScope prof(PROF_ENC_ATTN_FWD, 4.0 * P * H * static_cast<double>(L) * L * 64, s);
  launch_pdl(kern, dim3((L + LT - 1) / LT, H, P), dim3(256), smem, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

int enc_attn_long_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                      const unsigned char* mask, const float* bias_lut, __nv_bfloat16* dqkv, float* dlut, int P, int L, int H,
                      cudaStream_t s, DropSpec drop, const int* cu, bool d_in_out) {
  int rc = check_arch();
  if (rc) return rc;
  LAKO_REQUIRE(P > 0 && H > 0 && L > 0 && qkv && out && dout && lse && bias_lut && dqkv, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(L <= ENC_ATTN_LONG_MAX_L && P <= 65535 && H <= 65535, LAKO_ERR_UNSUPPORTED);
  LongAttnParams p{};
  p.qkv = qkv;
  p.mask = mask;
  p.lut = bias_lut;
  p.lse = const_cast<float*>(lse);
  p.o_in = out;
  p.dout = dout;
  p.dqkv = dqkv;
  p.dlut = dlut;
  p.cu = cu;
  p.L = L;
  p.H = H;
  p.inner = H * 64;
  p.d_in_out = d_in_out ? 1 : 0;
  p.kstride = enc_attn_keep_stride(L);
  p.drop = drop;
  const size_t smem_q = (LQ_SMEM_FLOATS + 2 * (2 * L - 1)) * sizeof(float);
  const size_t smem_k = (LK_SMEM_FLOATS + 2 * L - 1) * sizeof(float);
  auto kq = drop.thresh ? enc_attn_long_dq_kernel<true> : enc_attn_long_dq_kernel<false>;
  auto kk = drop.thresh ? enc_attn_long_dkv_kernel<true> : enc_attn_long_dkv_kernel<false>;
  rc = configure_smem(kq, smem_q);
  if (rc) return rc;
  rc = configure_smem(kk, smem_k);
  if (rc) return rc;
  ProfScope prof(PROF_ENC_ATTN_BWD, 10.0 * P * H * static_cast<double>(L) * L * 64, s);
  const dim3 grid((L + LT - 1) / LT, H, P);
  launch_pdl(kq, grid, dim3(256), smem_q, s, p);
  LAKO_CHECK_LAUNCH();
  launch_pdl(kk, grid, dim3(256), smem_k, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
