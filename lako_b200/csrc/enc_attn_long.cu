// T5 encoder self-attention for passages longer than the 256 tokens the tensor-memory kernels of enc_attn.cu hold in one
// softmax tile (HF:modeling_t5.py:253-344 via src/model.py:227-234; the reference takes any --text_maxlength,
// src/options.py:60).  Same contract as enc_attn_fwd / enc_attn_bwd -- bf16 q|k|v rows in, no 1/sqrt(dk), bucketed
// relative bias from the [H][2L-1] table, additive finfo.min key mask, dropout on the probabilities with the keep words
// of common.cuh drop_keep32_sliced -- for any L <= 2048: 64 x 64 tiles, online softmax, warp-level tensor-core MMAs
// (mma.sync m16n8k16 bf16, fp32 accumulation; mma_sync.cuh).  This is the coverage path, not the hot path: the reader's
// headline shapes (L = 200 / 250) never come here.  Three kernels, 4 warps each, a warp owns 16 rows of the CTA's tile:
//   forward   one CTA per (64-query tile, head, passage), loops over key tiles; P stays in registers (the S accumulators
//             are re-packed as the A fragments of P V); writes O (bf16) and the row lse
//   dq        same grid; recomputes P from lse, forms dS, accumulates dQ = dS K and the per-offset bias gradient
//   dkv       one CTA per (64-key tile, head, passage), loops over query tiles, works on the transposed products
//             S^T = K Q^T, dP^T = V dO^T; dV = Pd^T dO, dK = dS^T Q
// dq and dkv each recompute S and dO V^T (7 tile products instead of the minimal 5) so that neither needs atomics on
// the bf16 gradient rows and both are deterministic; only the bias gradient uses (fp32) atomics, as in enc_attn.cu.
#include "mma_sync.cuh"
#include "ops.cuh"

namespace lako {

namespace {

constexpr int LT = 64;          // tile edge (queries, keys, head dim)
constexpr int LTILE_B = 8192;   // one [64][64] bf16 tile, 128B-swizzled rows
constexpr int LDS_P = 72;       // pitch of the fp32 dS tile: float2 stores of a half-warp (rows g, column pairs t4) and the
                                // diagonal walks are both conflict-free
constexpr float LONG_MASK_MIN = -3.4028234663852886e38f;  // torch.finfo(float32).min, HF's additive mask
constexpr float LONG_LOG2E = 1.4426950408889634f;
// exp(x) as one MUFU.EX2 (the tcgen05 kernels do the same); exp(-inf) = 0
__device__ __forceinline__ float exp_fast(float x) { return fast_exp2(x * LONG_LOG2E); }

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

__device__ __forceinline__ void zero84(float (&a)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// acc[nt] (16 x 8 tiles over 64 columns) += A[arow0 .. +16][64] * B[0 .. 64][64]^T: both operands are [row][64] bf16
// swizzled tiles, the contraction runs over their 64 columns (S = Q K^T and its transposes)
__device__ __forceinline__ void mma_rowrow(float (&acc)[8][4], uint32_t aA, int arow0, uint32_t aB, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, aA + swz(arow0 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b4[4];
      ldsm_x4(b4, aB + swz(np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, ks * 2 + ((lane >> 3) & 1)));
      mma16816(acc[np * 2], a, b4[0], b4[1]);
      mma16816(acc[np * 2 + 1], a, b4[2], b4[3]);
    }
  }
}
// acc[nt] (16 x 8 tiles over the 64 columns of B) += A(16 x 64, register fragments a[kk]) * B[0 .. 64][64]: the
// contraction runs over B's rows (P V, dS K, Pd^T dO, dS^T Q)
__device__ __forceinline__ void mma_fragrow(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t aB, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b4[4];
      ldsm_x4_t(b4, aB + swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
      mma16816(acc[dp * 2], a[kk], b4[0], b4[1]);
      mma16816(acc[dp * 2 + 1], a[kk], b4[2], b4[3]);
    }
  }
}
// accumulator tile pair (nt = 2kk, 2kk+1) -> the A fragment of contraction step kk
__device__ __forceinline__ void acc_to_afrag(uint32_t (&a)[4][4], const float (&v)[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    a[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(v[nt][0], v[nt][1]);
    a[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(v[nt][2], v[nt][3]);
  }
}

// keep words of a 64-query x 64-key tile: sKeep[r * 2 + w] for query q0 + r and 32-key chunk kt * 2 + w (128 threads)
__device__ __forceinline__ void tile_keep_words(const LongAttnParams& p, int psg, int h, int q0, int kt, int Lp, uint32_t* sKeep, int tid) {
  const int r = tid >> 1, w = tid & 1;
  const int qc = min(q0 + r, Lp - 1);
  const unsigned long long row = static_cast<unsigned long long>(psg * p.H + h) * p.L + qc;
  sKeep[tid] = drop_keep32_sliced(p.drop, row * p.kstride + static_cast<unsigned>(kt * 2 + w) * 4u);
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
    for (int idx = tid; idx < 512; idx += 128) {  // 8 consecutive lanes share a row
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

// rows r0 = g, r1 = g + 8 of a warp's 16 x 64 accumulator block -> bf16 rows at dst (row pitch ld), rows >= nvalid skipped
__device__ __forceinline__ void store_acc_rows(__nv_bfloat16* dst, size_t ld, const float (&acc)[8][4], int g, int t4, int nvalid,
                                               float s0 = 1.f, float s1 = 1.f) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (g < nvalid)
      *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(g) * ld + nt * 8 + t4 * 2) = pack_bf16x2(acc[nt][0] * s0, acc[nt][1] * s0);
    if (g + 8 < nvalid)
      *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(g + 8) * ld + nt * 8 + t4 * 2) = pack_bf16x2(acc[nt][2] * s1, acc[nt][3] * s1);
  }
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int LF_SMEM_BYTES = 3 * LTILE_B + (64 + 128) * 4;  // + (2L-1) LUT floats
template <bool kDrop>
__global__ void __launch_bounds__(128) enc_attn_long_fwd_kernel(const LongAttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t lsm[];
  uint8_t* sQ = lsm;
  uint8_t* sK = sQ + LTILE_B;
  uint8_t* sV = sK + LTILE_B;
  float* sKm = reinterpret_cast<float*>(sV + LTILE_B);
  uint32_t* sKeep = reinterpret_cast<uint32_t*>(sKm + 64);
  float* sLut = reinterpret_cast<float*>(sKeep + 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * LT, h = blockIdx.y, psg = blockIdx.z;
  const int Lmax = p.L;
  int row0 = psg * Lmax, Lp = Lmax;
  if (p.cu) {
    row0 = p.cu[psg];
    Lp = p.cu[psg + 1] - row0;
  }
  if (q0 >= Lp) return;
  const size_t ld = 3 * static_cast<size_t>(p.inner);
  for (int i = tid; i < 2 * Lmax - 1; i += 128) sLut[i] = p.lut[static_cast<size_t>(h) * (2 * Lmax - 1) + i];
  stage_sw(sQ, p.qkv + static_cast<size_t>(row0 + q0) * ld + h * 64, ld, 64, Lp - q0, tid, 128);
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
  const int rl[2] = {warp * 16 + g, warp * 16 + g + 8};  // this thread's two rows within the tile
  const int qpos[2] = {min(q0 + rl[0], Lp - 1), min(q0 + rl[1], Lp - 1)};
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, o[8][4];
  zero84(o);
  const int ktiles = (Lp + LT - 1) / LT;
  for (int kt = 0; kt < ktiles; ++kt) {
    __syncthreads();  // previous tile consumed
    const __nv_bfloat16* kbase = p.qkv + static_cast<size_t>(row0 + kt * LT) * ld + h * 64;
    stage_sw(sK, kbase + p.inner, ld, 64, Lp - kt * LT, tid, 128);
    stage_sw(sV, kbase + 2 * p.inner, ld, 64, Lp - kt * LT, tid, 128);
    tile_key_mask(p, psg, kt, Lp, sKm, tid);
    if (kDrop) tile_keep_words(p, psg, h, q0, kt, Lp, sKeep, tid);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4];
    zero84(s);
    mma_rowrow(s, aQ, warp * 16, aK, lane);
    float mt[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kl = nt * 8 + t4 * 2 + (e & 1), r = e >> 1;
        const int kpos = min(kt * LT + kl, Lmax - 1);
        s[nt][e] += sLut[kpos - qpos[r] + Lmax - 1] + sKm[kl];
        mt[r] = fmaxf(mt[r], s[nt][e]);
      }
    float sc[2], m_new[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mt[r] = fmaxf(mt[r], __shfl_xor_sync(0xffffffffu, mt[r], 1));
      mt[r] = fmaxf(mt[r], __shfl_xor_sync(0xffffffffu, mt[r], 2));
      m_new[r] = fmaxf(m[r], mt[r]);
      sc[r] = (m[r] == -INFINITY) ? 0.f : exp_fast(m[r] - m_new[r]);
      m[r] = m_new[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const float ex = exp_fast(s[nt][e] - m_new[r]);  // -inf (beyond the passage) -> 0
        rs[r] += ex;  // the normaliser is the sum before dropout (HF:331-332: softmax, then dropout)
        s[nt][e] = (!kDrop || keep_of(sKeep, rl[r], nt * 8 + t4 * 2 + (e & 1))) ? ex : 0.f;
        o[nt][e] *= sc[r];
      }
    l[0] = l[0] * sc[0] + rs[0];  // per-thread partial sums; the quad is reduced once at the end
    l[1] = l[1] * sc[1] + rs[1];
    uint32_t pa[4][4];
    acc_to_afrag(pa, s);
    mma_fragrow(o, pa, aV, lane);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  const float dscale = kDrop ? p.drop.scale : 1.f;
  store_acc_rows(p.out + static_cast<size_t>(row0 + q0 + warp * 16) * p.inner + h * 64, p.inner, o, g, t4, Lp - q0 - warp * 16,
                 dscale / l[0], dscale / l[1]);
  if (p.lse && t4 == 0) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
      if (q0 + rl[r] < Lp) p.lse[(static_cast<size_t>(psg) * p.H + h) * Lmax + q0 + rl[r]] = m[r] + logf(l[r]);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ, dbias
constexpr int LQ_SMEM_BYTES = 4 * LTILE_B + (64 * LDS_P + 3 * 64 + 128) * 4;  // + 2 (2L-1) floats: LUT and its gradient
template <bool kDrop>
__global__ void __launch_bounds__(128) enc_attn_long_dq_kernel(const LongAttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t lsm[];
  uint8_t* sQ = lsm;
  uint8_t* sdO = sQ + LTILE_B;
  uint8_t* sK = sdO + LTILE_B;
  uint8_t* sV = sK + LTILE_B;
  float* sDS = reinterpret_cast<float*>(sV + LTILE_B);  // [row][key] fp32, pitch LDS_P (bias gradient only)
  float* sKm = sDS + 64 * LDS_P;
  float* sLse = sKm + 64;
  float* sD = sLse + 64;
  uint32_t* sKeep = reinterpret_cast<uint32_t*>(sD + 64);
  float* sLut = reinterpret_cast<float*>(sKeep + 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
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
  for (int i = tid; i < nlut; i += 128) {
    sLut[i] = p.lut[static_cast<size_t>(h) * nlut + i];
    sDl[i] = 0.f;
  }
  stage_sw(sQ, p.qkv + static_cast<size_t>(row0 + q0) * ld + h * 64, ld, 64, Lp - q0, tid, 128);
  stage_sw(sdO, p.dout + static_cast<size_t>(row0 + q0) * p.inner + h * 64, p.inner, 64, Lp - q0, tid, 128);
  tile_row_stats(p, psg, h, row0, q0, Lp, sLse, sD, tid);
  const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV);
  const int rl[2] = {warp * 16 + g, warp * 16 + g + 8};
  const int qpos[2] = {min(q0 + rl[0], Lp - 1), min(q0 + rl[1], Lp - 1)};
  float dq[8][4];
  zero84(dq);
  const float keep_scale = kDrop ? p.drop.scale : 1.f;
  const int ktiles = (Lp + LT - 1) / LT;
  for (int kt = 0; kt < ktiles; ++kt) {
    __syncthreads();  // previous tile (K, V, dS) consumed; first pass: the row statistics are in place
    const __nv_bfloat16* kbase = p.qkv + static_cast<size_t>(row0 + kt * LT) * ld + h * 64;
    stage_sw(sK, kbase + p.inner, ld, 64, Lp - kt * LT, tid, 128);
    stage_sw(sV, kbase + 2 * p.inner, ld, 64, Lp - kt * LT, tid, 128);
    tile_key_mask(p, psg, kt, Lp, sKm, tid);
    if (kDrop) tile_keep_words(p, psg, h, q0, kt, Lp, sKeep, tid);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4], dp[8][4];
    zero84(s);
    zero84(dp);
    mma_rowrow(s, aQ, warp * 16, aK, lane);
    mma_rowrow(dp, adO, warp * 16, aV, lane);
    const float lse[2] = {sLse[rl[0]], sLse[rl[1]]}, D[2] = {sD[rl[0]], sD[rl[1]]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kl = nt * 8 + t4 * 2 + (e & 1), r = e >> 1;
        const int kpos = min(kt * LT + kl, Lmax - 1);
        const float pv = exp_fast(s[nt][e] + sLut[kpos - qpos[r] + Lmax - 1] + sKm[kl] - lse[r]);
        const float kf = (!kDrop || keep_of(sKeep, rl[r], kl)) ? keep_scale : 0.f;
        s[nt][e] = pv * (kf * dp[nt][e] - D[r]);
      }
    if (p.dlut) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<float2*>(sDS + rl[0] * LDS_P + nt * 8 + t4 * 2) = make_float2(s[nt][0], s[nt][1]);
        *reinterpret_cast<float2*>(sDS + rl[1] * LDS_P + nt * 8 + t4 * 2) = make_float2(s[nt][2], s[nt][3]);
      }
    }
    uint32_t dsa[4][4];
    acc_to_afrag(dsa, s);
    mma_fragrow(dq, dsa, aK, lane);
    if (p.dlut) {
      __syncthreads();
      if (tid < 127) {  // one thread per diagonal key - query = dd of the tile
        const int dd = tid - 63;
        float acc = 0.f;
        for (int r = max(0, -dd); r < min(64, 64 - dd); ++r) acc += sDS[r * LDS_P + r + dd];
        const int idx = (kt * LT - q0) + dd + Lmax - 1;
        if (idx >= 0 && idx < nlut) sDl[idx] += acc;  // pairs outside the passage carry dS = 0
      }
    }
  }
  store_acc_rows(p.dqkv + static_cast<size_t>(row0 + q0 + warp * 16) * ld + h * 64, ld, dq, g, t4, Lp - q0 - warp * 16);
  if (p.dlut) {
    __syncthreads();
    for (int i = tid; i < nlut; i += 128)
      if (sDl[i] != 0.f) atomicAdd(p.dlut + static_cast<size_t>(h) * nlut + i, sDl[i]);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
constexpr int LK_SMEM_BYTES = 4 * LTILE_B + (3 * 64 + 128) * 4;  // + (2L-1) LUT floats
template <bool kDrop>
__global__ void __launch_bounds__(128, 3) enc_attn_long_dkv_kernel(const LongAttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t lsm[];
  uint8_t* sK = lsm;
  uint8_t* sV = sK + LTILE_B;
  uint8_t* sQ = sV + LTILE_B;
  uint8_t* sdO = sQ + LTILE_B;
  float* sKm = reinterpret_cast<float*>(sdO + LTILE_B);
  float* sLse = sKm + 64;
  float* sD = sLse + 64;
  uint32_t* sKeep = reinterpret_cast<uint32_t*>(sD + 64);
  float* sLut = reinterpret_cast<float*>(sKeep + 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int kt = blockIdx.x, h = blockIdx.y, psg = blockIdx.z;
  const int Lmax = p.L, nlut = 2 * Lmax - 1;
  int row0 = psg * Lmax, Lp = Lmax;
  if (p.cu) {
    row0 = p.cu[psg];
    Lp = p.cu[psg + 1] - row0;
  }
  if (kt * LT >= Lp) return;
  const size_t ld = 3 * static_cast<size_t>(p.inner);
  for (int i = tid; i < nlut; i += 128) sLut[i] = p.lut[static_cast<size_t>(h) * nlut + i];
  const __nv_bfloat16* kbase = p.qkv + static_cast<size_t>(row0 + kt * LT) * ld + h * 64;
  stage_sw(sK, kbase + p.inner, ld, 64, Lp - kt * LT, tid, 128);
  stage_sw(sV, kbase + 2 * p.inner, ld, 64, Lp - kt * LT, tid, 128);
  tile_key_mask(p, psg, kt, Lp, sKm, tid);
  const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV);
  const int kl2[2] = {warp * 16 + g, warp * 16 + g + 8};  // this thread's two keys within the tile
  const int kpos[2] = {min(kt * LT + kl2[0], Lmax - 1), min(kt * LT + kl2[1], Lmax - 1)};
  float dk[8][4], dv[8][4];
  zero84(dk);
  zero84(dv);
  const float keep_scale = kDrop ? p.drop.scale : 1.f;
  const int qtiles = (Lp + LT - 1) / LT;
  for (int qt = 0; qt < qtiles; ++qt) {
    const int q0 = qt * LT;
    __syncthreads();
    stage_sw(sQ, p.qkv + static_cast<size_t>(row0 + q0) * ld + h * 64, ld, 64, Lp - q0, tid, 128);
    stage_sw(sdO, p.dout + static_cast<size_t>(row0 + q0) * p.inner + h * 64, p.inner, 64, Lp - q0, tid, 128);
    tile_row_stats(p, psg, h, row0, q0, Lp, sLse, sD, tid);
    if (kDrop) tile_keep_words(p, psg, h, q0, kt, Lp, sKeep, tid);
    cp_async_wait_all();
    __syncthreads();
    float st[8][4], dpt[8][4];  // [key][query row]
    zero84(st);
    zero84(dpt);
    mma_rowrow(st, aK, warp * 16, aQ, lane);
    mma_rowrow(dpt, aV, warp * 16, adO, lane);
    const float km[2] = {sKm[kl2[0]], sKm[kl2[1]]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = nt * 8 + t4 * 2 + (e & 1), k = e >> 1;
        const int qp = min(q0 + r, Lp - 1);
        const float pv = exp_fast(st[nt][e] + sLut[kpos[k] - qp + Lmax - 1] + km[k] - sLse[r]);
        const float kf = (!kDrop || keep_of(sKeep, r, kl2[k])) ? keep_scale : 0.f;
        st[nt][e] = pv * kf;                           // dropped probabilities (transposed)
        dpt[nt][e] = pv * (kf * dpt[nt][e] - sD[r]);   // dS^T
      }
    uint32_t fa[4][4];
    acc_to_afrag(fa, st);
    mma_fragrow(dv, fa, adO, lane);
    acc_to_afrag(fa, dpt);
    mma_fragrow(dk, fa, aQ, lane);
  }
  __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(row0 + kt * LT + warp * 16) * ld + h * 64;
  store_acc_rows(dst + p.inner, ld, dk, g, t4, Lp - kt * LT - warp * 16);
  store_acc_rows(dst + 2 * p.inner, ld, dv, g, t4, Lp - kt * LT - warp * 16);
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
  const size_t smem = LF_SMEM_BYTES + (2 * L - 1) * sizeof(float);
  auto kern = drop.thresh ? enc_attn_long_fwd_kernel<true> : enc_attn_long_fwd_kernel<false>;
  rc = configure_smem(kern, smem);
  if (rc) return rc;
  ProfScope prof(PROF_ENC_ATTN_FWD, 4.0 * P * H * static_cast<double>(L) * L * 64, s);
  launch_pdl(kern, dim3((L + LT - 1) / LT, H, P), dim3(128), smem, s, p);
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
  const size_t smem_q = LQ_SMEM_BYTES + 2 * (2 * L - 1) * sizeof(float);
  const size_t smem_k = LK_SMEM_BYTES + (2 * L - 1) * sizeof(float);
  auto kq = drop.thresh ? enc_attn_long_dq_kernel<true> : enc_attn_long_dq_kernel<false>;
  auto kk = drop.thresh ? enc_attn_long_dkv_kernel<true> : enc_attn_long_dkv_kernel<false>;
  rc = configure_smem(kq, smem_q);
  if (rc) return rc;
  rc = configure_smem(kk, smem_k);
  if (rc) return rc;
  ProfScope prof(PROF_ENC_ATTN_BWD, 10.0 * P * H * static_cast<double>(L) * L * 64, s);
  const dim3 grid((L + LT - 1) / LT, H, P);
  launch_pdl(kq, grid, dim3(128), smem_q, s, p);
  LAKO_CHECK_LAUNCH();
  launch_pdl(kk, grid, dim3(128), smem_k, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
