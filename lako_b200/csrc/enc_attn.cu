// Passage-batched T5 encoder self-attention on tcgen05/TMEM, forward and backward.
//
// Reference: T5Attention.forward HF:modeling_t5.py:253-344 as driven per passage by
// EncoderWrapper.forward (/root/reference/src/model.py:227-234): scores = q k^T (NO 1/sqrt(dk)) +
// bucketed relative-position bias (shared by all layers, HF:755-758) + key-padding mask (finfo.min),
// fp32 softmax, probs @ v.   One "unit" = (passage, head); L = text_maxlength <= 256, d_kv = 64.
//
// Because a whole passage fits one tile (L <= 256 keys), softmax is single pass (no online
// rescaling): S = Q K^T lands in TMEM (128 lanes x LK fp32 columns), each of the 128 threads owns
// one query row, reads it with tcgen05.ld, adds the bias from a per-head LUT in shared memory
// (2L-1 entries), writes un-normalised bf16 probabilities into 128B-swizzled shared memory, and
// the second MMA O = P V (V as an MN-major operand straight from the qkv buffer) accumulates into
// the same TMEM columns.  Q/K/V tiles come in by TMA from the fused qkv activation
// [tokens][3*inner]; two CTAs are co-resident per SM (256 TMEM columns, ~97 KB smem each) so one
// CTA's softmax overlaps the other's loads/MMAs.
#include "ops.cuh"

namespace lako {

constexpr float MASK_MIN = -3.4028234663852886e38f;  // torch.finfo(float32).min, as HF's additive mask

// bias_lut[h][rel + L - 1] = table[bucket(rel)][h], rel = key_pos - query_pos
__global__ void build_bias_lut_kernel(const float* __restrict__ table, const int* __restrict__ buckets,
                                      float* __restrict__ lut, int H, int L, int bucket_center) {
  const int n = 2 * L - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * n) return;
  const int h = i / n, j = i - h * n;
  const int rel = j - (L - 1);
  lut[i] = table[buckets[rel + bucket_center] * H + h];
}

int build_bias_lut(const float* table, const int* buckets, float* lut, int H, int L, int bucket_center,
                   cudaStream_t s) {
  const int n = H * (2 * L - 1);
  build_bias_lut_kernel<<<(n + 255) / 256, 256, 0, s>>>(table, buckets, lut, H, L, bucket_center);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

struct EncAttnParams {
  const unsigned char* mask;  // [P][L] 1 = valid key
  const float* bias_lut;      // [H][2L-1]
  __nv_bfloat16* out;         // [P*L][inner]
  float* lse;                 // [P][H][L]
  int L, H, inner, LK;        // LK = L rounded up to 16
};

// smem map (bytes from the 1024-aligned base):
//   [0, 32K)   K tile rows 0..255   | after S is read: P chunks 0,1
//   [32K, 48K) Q tile               | P chunk 2
//   [48K, 64K) spare                | P chunk 3
//   [64K, 96K) V tile rows 0..255
//   then: lut (2048 B), mask (256 B), barriers
constexpr int EA_K_OFF = 0, EA_Q_OFF = 32768, EA_V_OFF = 65536, EA_LUT_OFF = 98304, EA_MASK_OFF = 98304 + 2048,
              EA_BAR_OFF = EA_MASK_OFF + 256;
constexpr int EA_SMEM_BYTES = EA_BAR_OFF + 64 + 1024;

__global__ void __launch_bounds__(128, 2)
enc_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const EncAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_lut = reinterpret_cast<float*>(smem + EA_LUT_OFF);
  unsigned char* s_mask = smem + EA_MASK_OFF;
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + EA_BAR_OFF);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_o = bar_qk + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, psg = blockIdx.z;
  const int L = p.L, LK = p.LK;
  const int row0 = psg * L;
  const int ktiles = (LK + 127) / 128;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  // stage the per-head bias LUT and the key mask
  for (int i = tid; i < 2 * L - 1; i += 128) s_lut[i] = p.bias_lut[h * (2 * L - 1) + i];
  for (int i = tid; i < 256; i += 128) s_mask[i] = (i < L) ? p.mask[psg * L + i] : 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_qk, 16384 * (1 + ktiles));
    tma_load_2d(smem + EA_Q_OFF, &tm_qkv, bar_qk, h * 64, row0 + q0);
    for (int t = 0; t < ktiles; ++t)
      tma_load_2d(smem + EA_K_OFF + t * 16384, &tm_qkv, bar_qk, p.inner + h * 64, row0 + t * 128);
    mbar_arrive_expect_tx(bar_v, 16384 * ktiles);
    for (int t = 0; t < ktiles; ++t)
      tma_load_2d(smem + EA_V_OFF + t * 16384, &tm_qkv, bar_v, 2 * p.inner + h * 64, row0 + t * 128);
    // S = Q K^T : M=128, N=LK, K=64
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, LK, false, false);
    const uint32_t qa = smem_u32(smem + EA_Q_OFF), ka = smem_u32(smem + EA_K_OFF);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc,
                k > 0);
    umma_commit(bar_s);
  }

  // ---- softmax: thread == query row
  const int r = tid;           // row within tile == TMEM lane
  const int q = q0 + r;        // query position in the passage
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  mbar_wait(bar_s, 0);
  tc_fence_after();
  const int nchunks = (LK + 31) / 32;
  const float* lut_q = s_lut + (L - 1 - q);  // lut_q[k] = bias(k - q); only dereferenced for q < L, k < L
  float m = -INFINITY;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(t_row + c * 32, v);
    tmem_ld_wait();
    if (q < L) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k = c * 32 + i;
        if (k < L) {
          const float s = s_mask[k] ? __uint_as_float(v[i]) + lut_q[k] : MASK_MIN;
          m = fmaxf(m, s);
        }
      }
    }
  }
  float lsum = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(t_row + c * 32, v);
    tmem_ld_wait();
    float pr[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = c * 32 + i;
      float e = 0.f;
      if (q < L && k < L) {
        const float s = s_mask[k] ? __uint_as_float(v[i]) + lut_q[k] : MASK_MIN;
        e = __expf(s - m);
      }
      pr[i] = e;
      lsum += e;
    }
    // P[r][k] bf16 into the K-major 128B-swizzled layout: 64-key chunks of [128 rows][128 B]
    uint8_t* chunk = smem + (c >> 1) * 16384 + r * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int unit = (c & 1) * 4 + u;  // 16-byte unit index within the 128-byte row
      uint4 o;
      o.x = pack_bf16x2(pr[u * 8 + 0], pr[u * 8 + 1]);
      o.y = pack_bf16x2(pr[u * 8 + 2], pr[u * 8 + 3]);
      o.z = pack_bf16x2(pr[u * 8 + 4], pr[u * 8 + 5]);
      o.w = pack_bf16x2(pr[u * 8 + 6], pr[u * 8 + 7]);
      *reinterpret_cast<uint4*>(chunk + ((unit ^ (r & 7)) << 4)) = o;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    // O = P V : M=128, N=64, K=LK ; A = P (K-major), B = V (MN-major: [key][64 d] rows of 128 B)
    tc_fence_after();
    mbar_wait(bar_v, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, false, true);
    const uint32_t pa = smem_u32(smem), va = smem_u32(smem + EA_V_OFF);
    const int ksteps = LK / 16;
    for (int j = 0; j < ksteps; ++j)
      umma_bf16(tmem, make_smem_desc_sw128(pa + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                make_smem_desc_sw128(va + j * 2048, 8192, 1024), idesc, j > 0);
    umma_commit(bar_o);
  }
  mbar_wait(bar_o, 0);
  tc_fence_after();
  {
    uint32_t o0[32], o1[32];
    tmem_ld_32x32(t_row, o0);
    tmem_ld_32x32(t_row + 32, o1);
    tmem_ld_wait();
    if (q < L) {
      const float inv = 1.f / lsum;
      __nv_bfloat16* orow = p.out + static_cast<size_t>(row0 + q) * p.inner + h * 64;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(o0[u * 8 + 0]) * inv, __uint_as_float(o0[u * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(o0[u * 8 + 2]) * inv, __uint_as_float(o0[u * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(o0[u * 8 + 4]) * inv, __uint_as_float(o0[u * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(o0[u * 8 + 6]) * inv, __uint_as_float(o0[u * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + u * 8) = o;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(o1[u * 8 + 0]) * inv, __uint_as_float(o1[u * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(o1[u * 8 + 2]) * inv, __uint_as_float(o1[u * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(o1[u * 8 + 4]) * inv, __uint_as_float(o1[u * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(o1[u * 8 + 6]) * inv, __uint_as_float(o1[u * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + 32 + u * 8) = o;
      }
      if (p.lse) p.lse[(static_cast<size_t>(psg) * p.H + h) * L + q] = m + logf(lsum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// qkv bf16 [P*L][3*inner]; out bf16 [P*L][inner]; lse fp32 [P][H][L] (may be null)
int enc_attn_fwd(const __nv_bfloat16* qkv, const unsigned char* mask, const float* bias_lut, __nv_bfloat16* out,
                 float* lse, int P, int L, int H, cudaStream_t s) {
  int rc = check_arch();
  if (rc) return rc;
  LAKO_REQUIRE(P > 0 && H > 0 && L > 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(L <= 256, LAKO_ERR_UNSUPPORTED);  // single-tile softmax; longer passages need the looped variant
  const int inner = H * 64;
  CUtensorMap tm;
  rc = get_tensor_map_2d(&tm, qkv, 3 * inner, static_cast<uint64_t>(P) * L, 3 * inner, 64, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(enc_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EA_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return LAKO_ERR_LAUNCH;
    }
    configured = true;
  }
  EncAttnParams p;
  p.mask = mask;
  p.bias_lut = bias_lut;
  p.out = out;
  p.lse = lse;
  p.L = L;
  p.H = H;
  p.inner = inner;
  p.LK = (L + 15) / 16 * 16;
  ProfScope prof(PROF_ENC_ATTN_FWD, 4.0 * P * H * static_cast<double>(L) * L * 64, s);
  dim3 grid((L + 127) / 128, H, P);
  enc_attn_fwd_kernel<<<grid, 128, EA_SMEM_BYTES, s>>>(tm, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako

// ================================================================================================
// backward
// ================================================================================================
namespace lako {

struct EncAttnBwdParams {
  const unsigned char* mask;     // [P][L]
  const float* bias_lut;         // [H][2L-1]
  const float* lse;              // [P][H][L]
  const __nv_bfloat16* out;      // [P*L][inner]  forward output (for D = rowsum(dO * O))
  const __nv_bfloat16* dout;     // [P*L][inner]
  __nv_bfloat16* dqkv;           // [P*L][3*inner]
  float* dlut;                   // [H][2L-1]  += sum over (q,k) diagonals of dS (bias gradient, pre-bucket)
  int L, H, inner;
};

constexpr int EB_Q_OFF = 0, EB_DO_OFF = 32768, EB_K_OFF = 65536, EB_V_OFF = 98304, EB_P_OFF = 131072,
              EB_DS_OFF = 163840, EB_LUT_OFF = 196608, EB_DLUT_OFF = EB_LUT_OFF + 2048, EB_MASK_OFF = EB_DLUT_OFF + 2048,
              EB_LSE_OFF = EB_MASK_OFF + 256, EB_D_OFF = EB_LSE_OFF + 1024, EB_BAR_OFF = EB_D_OFF + 1024;
constexpr int EB_SMEM_BYTES = EB_BAR_OFF + 64 + 1024;
constexpr int TM_S = 0, TM_DP = 128, TM_DQ = 256, TM_DK = 384, TM_DV = 448;

__global__ void __launch_bounds__(256, 1)
enc_attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                    const EncAttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_lut = reinterpret_cast<float*>(smem + EB_LUT_OFF);
  float* s_dlut = reinterpret_cast<float*>(smem + EB_DLUT_OFF);
  unsigned char* s_mask = smem + EB_MASK_OFF;
  float* s_lse = reinterpret_cast<float*>(smem + EB_LSE_OFF);
  float* s_D = reinterpret_cast<float*>(smem + EB_D_OFF);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + EB_BAR_OFF);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, chalf = warp >> 2;
  const int h = blockIdx.x, psg = blockIdx.y;
  const int L = p.L;
  const int row0 = psg * L;
  const int nt = (L + 127) / 128;  // 128-row tiles of queries == of keys
  const int nlut = 2 * L - 1;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < nlut; i += 256) {
    s_lut[i] = p.bias_lut[h * nlut + i];
    s_dlut[i] = 0.f;
  }
  {
    const int q = tid;  // 256 threads cover the (<=256) query rows: D[q] = sum_d dO*O, lse[q]
    s_mask[q] = (q < L) ? p.mask[psg * L + q] : 0;
    float dsum = 0.f, l = 0.f;
    if (q < L) {
      const uint4* o4 = reinterpret_cast<const uint4*>(p.out + static_cast<size_t>(row0 + q) * p.inner + h * 64);
      const uint4* d4 = reinterpret_cast<const uint4*>(p.dout + static_cast<size_t>(row0 + q) * p.inner + h * 64);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint4 a = o4[u], b = d4[u];
        dsum += bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) +
                bf16_hi(a.y) * bf16_hi(b.y) + bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) +
                bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
      }
      l = p.lse[(static_cast<size_t>(psg) * p.H + h) * L + q];
    }
    s_D[q] = dsum;
    s_lse[q] = l;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const uint32_t qa = smem_u32(smem + EB_Q_OFF), doa = smem_u32(smem + EB_DO_OFF), ka = smem_u32(smem + EB_K_OFF),
                 va = smem_u32(smem + EB_V_OFF), pa = smem_u32(smem + EB_P_OFF), dsa = smem_u32(smem + EB_DS_OFF);
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);    // S, dP : K-major x K-major
  constexpr uint32_t idesc_tn = make_idesc_bf16(128, 64, true, true);      // dV, dK: MN-major x MN-major
  constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64, false, true);     // dQ    : K-major x MN-major

  auto issue_s_dp = [&](int qt, int kt) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem + TM_S, make_smem_desc_sw128(qa + qt * 16384 + k * 32, 16, 1024),
                make_smem_desc_sw128(ka + kt * 16384 + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem + TM_DP, make_smem_desc_sw128(doa + qt * 16384 + k * 32, 16, 1024),
                make_smem_desc_sw128(va + kt * 16384 + k * 32, 16, 1024), idesc_s, k > 0);
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_load, 16384 * 4 * nt);
    for (int t = 0; t < nt; ++t) {
      tma_load_2d(smem + EB_Q_OFF + t * 16384, &tm_qkv, bar_load, h * 64, row0 + t * 128);
      tma_load_2d(smem + EB_K_OFF + t * 16384, &tm_qkv, bar_load, p.inner + h * 64, row0 + t * 128);
      tma_load_2d(smem + EB_V_OFF + t * 16384, &tm_qkv, bar_load, 2 * p.inner + h * 64, row0 + t * 128);
      tma_load_2d(smem + EB_DO_OFF + t * 16384, &tm_do, bar_load, h * 64, row0 + t * 128);
    }
    mbar_wait(bar_load, 0);
    tc_fence_after();
    issue_s_dp(0, 0);
    umma_commit(bar_mma);
  }

  const int r = quarter * 32 + lane;  // row within a 128-row tile == TMEM lane
  const uint32_t t_row = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
  uint32_t mma_phase = 0;

  for (int kt = 0; kt < nt; ++kt) {
    for (int qt = 0; qt < nt; ++qt) {
      const int q = qt * 128 + r;
      mbar_wait(bar_mma, mma_phase);
      mma_phase ^= 1;
      tc_fence_after();
      const float lse_q = s_lse[q < 256 ? q : 0];
      const float D_q = s_D[q < 256 ? q : 0];
      const float* lut_q = s_lut + (L - 1 - q);
      uint8_t* prow = smem + EB_P_OFF + chalf * 16384 + r * 128;
      uint8_t* dsrow = smem + EB_DS_OFF + chalf * 16384 + r * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dpv[32];
        tmem_ld_32x32(t_row + TM_S + chalf * 64 + c * 32, sv);
        tmem_ld_32x32(t_row + TM_DP + chalf * 64 + c * 32, dpv);
        tmem_ld_wait();
        float pr[32], ds[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int k = kt * 128 + chalf * 64 + c * 32 + i;
          float pv = 0.f;
          if (q < L && k < L) {
            const float s = s_mask[k] ? __uint_as_float(sv[i]) + lut_q[k] : MASK_MIN;
            pv = __expf(s - lse_q);
          }
          pr[i] = pv;
          ds[i] = pv * (__uint_as_float(dpv[i]) - D_q);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int unit = ((c * 4 + u) ^ (r & 7)) << 4;
          uint4 o;
          o.x = pack_bf16x2(pr[u * 8 + 0], pr[u * 8 + 1]);
          o.y = pack_bf16x2(pr[u * 8 + 2], pr[u * 8 + 3]);
          o.z = pack_bf16x2(pr[u * 8 + 4], pr[u * 8 + 5]);
          o.w = pack_bf16x2(pr[u * 8 + 6], pr[u * 8 + 7]);
          *reinterpret_cast<uint4*>(prow + unit) = o;
          o.x = pack_bf16x2(ds[u * 8 + 0], ds[u * 8 + 1]);
          o.y = pack_bf16x2(ds[u * 8 + 2], ds[u * 8 + 3]);
          o.z = pack_bf16x2(ds[u * 8 + 4], ds[u * 8 + 5]);
          o.w = pack_bf16x2(ds[u * 8 + 6], ds[u * 8 + 7]);
          *reinterpret_cast<uint4*>(dsrow + unit) = o;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        // dV[kt] += P^T dO[qt] ; dK[kt] += dS^T Q[qt]   (M = 128 keys, N = 64, K = 128 queries)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16(tmem + TM_DV, make_smem_desc_sw128(pa + j * 2048, 16384, 1024),
                    make_smem_desc_sw128(doa + qt * 16384 + j * 2048, 8192, 1024), idesc_tn, (qt > 0 || j > 0));
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16(tmem + TM_DK, make_smem_desc_sw128(dsa + j * 2048, 16384, 1024),
                    make_smem_desc_sw128(qa + qt * 16384 + j * 2048, 8192, 1024), idesc_tn, (qt > 0 || j > 0));
        // dQ[qt] += dS K[kt]   (M = 128 queries, N = 64, K = 128 keys)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16(tmem + TM_DQ + qt * 64, make_smem_desc_sw128(dsa + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(ka + kt * 16384 + j * 2048, 8192, 1024), idesc_dq, (kt > 0 || j > 0));
        // next block's S / dP go to different TMEM columns: queue them behind the gradient MMAs
        int nq = qt + 1, nk = kt;
        if (nq == nt) {
          nq = 0;
          nk = kt + 1;
        }
        if (nk < nt) issue_s_dp(nq, nk);
        umma_commit(bar_mma);
      }
      // bias gradient: diagonal sums of this 128x128 dS block, read back (bf16) from shared memory.
      // thread t owns the diagonal d = t - 127 (k_local - q_local); consecutive threads read
      // consecutive keys of one row, so the reads are conflict-free.
      {
        const int d = tid - 127;
        float acc = 0.f;
        if (d < 128) {
          const int qlo = d < 0 ? -d : 0, qhi = d > 0 ? 128 - d : 128;
          for (int qq = qlo; qq < qhi; ++qq) {
            const int kk = qq + d;
            const uint8_t* a = smem + EB_DS_OFF + (kk >> 6) * 16384 + qq * 128 + ((((kk & 63) >> 3) ^ (qq & 7)) << 4) +
                               ((kk & 7) << 1);
            acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(a));
          }
          const int rel = (kt - qt) * 128 + d;
          if (rel > -L && rel < L && acc != 0.f) atomicAdd(&s_dlut[rel + L - 1], acc);
        }
      }
      if (qt == nt - 1) {
        // dK[kt], dV[kt] are final once the MMAs just issued complete
        mbar_wait(bar_mma, mma_phase);
        tc_fence_after();
        const int key = kt * 128 + r;
        uint32_t a[32], b[32];
        tmem_ld_32x32(t_row + TM_DK + chalf * 32, a);
        tmem_ld_32x32(t_row + TM_DV + chalf * 32, b);
        tmem_ld_wait();
        if (key < L) {
          __nv_bfloat16* dk = p.dqkv + static_cast<size_t>(row0 + key) * 3 * p.inner + p.inner + h * 64 + chalf * 32;
          __nv_bfloat16* dv = dk + p.inner;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(a[u * 8 + 0]), __uint_as_float(a[u * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(a[u * 8 + 2]), __uint_as_float(a[u * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(a[u * 8 + 4]), __uint_as_float(a[u * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(a[u * 8 + 6]), __uint_as_float(a[u * 8 + 7]));
            *reinterpret_cast<uint4*>(dk + u * 8) = o;
            o.x = pack_bf16x2(__uint_as_float(b[u * 8 + 0]), __uint_as_float(b[u * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(b[u * 8 + 2]), __uint_as_float(b[u * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(b[u * 8 + 4]), __uint_as_float(b[u * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(b[u * 8 + 6]), __uint_as_float(b[u * 8 + 7]));
            *reinterpret_cast<uint4*>(dv + u * 8) = o;
          }
        }
        // the wait above did not consume the phase: the loop head waits on the same completion
        tc_fence_before();
      }
      __syncthreads();  // diagonal reads of dS done before the next block overwrites it
    }
  }
  // all MMAs are complete here only if the last loop iteration waited (qt == nt-1 -> yes)
  {
    tc_fence_after();
#pragma unroll 1
    for (int qt = 0; qt < nt; ++qt) {
      const int q = qt * 128 + r;
      uint32_t a[32];
      tmem_ld_32x32(t_row + TM_DQ + qt * 64 + chalf * 32, a);
      tmem_ld_wait();
      if (q < L) {
        __nv_bfloat16* dq = p.dqkv + static_cast<size_t>(row0 + q) * 3 * p.inner + h * 64 + chalf * 32;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(a[u * 8 + 0]), __uint_as_float(a[u * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(a[u * 8 + 2]), __uint_as_float(a[u * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(a[u * 8 + 4]), __uint_as_float(a[u * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(a[u * 8 + 6]), __uint_as_float(a[u * 8 + 7]));
          *reinterpret_cast<uint4*>(dq + u * 8) = o;
        }
      }
    }
  }
  if (p.dlut) {
    for (int i = tid; i < nlut; i += 256) {
      const float v = s_dlut[i];
      if (v != 0.f) atomicAdd(&p.dlut[h * nlut + i], v);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int enc_attn_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                 const unsigned char* mask, const float* bias_lut, __nv_bfloat16* dqkv, float* dlut, int P, int L,
                 int H, cudaStream_t s) {
  int rc = check_arch();
  if (rc) return rc;
  LAKO_REQUIRE(P > 0 && H > 0 && L > 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(L <= 256, LAKO_ERR_UNSUPPORTED);
  const int inner = H * 64;
  CUtensorMap tm_qkv, tm_do;
  rc = get_tensor_map_2d(&tm_qkv, qkv, 3 * inner, static_cast<uint64_t>(P) * L, 3 * inner, 64, 128);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tm_do, dout, inner, static_cast<uint64_t>(P) * L, inner, 64, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(enc_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EB_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return LAKO_ERR_LAUNCH;
    }
    configured = true;
  }
  EncAttnBwdParams p;
  p.mask = mask;
  p.bias_lut = bias_lut;
  p.lse = lse;
  p.out = out;
  p.dout = dout;
  p.dqkv = dqkv;
  p.dlut = dlut;
  p.L = L;
  p.H = H;
  p.inner = inner;
  ProfScope prof(PROF_ENC_ATTN_BWD, 10.0 * P * H * static_cast<double>(L) * L * 64, s);
  dim3 grid(H, P);
  enc_attn_bwd_kernel<<<grid, 256, EB_SMEM_BYTES, s>>>(tm_qkv, tm_do, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// dtable[bucket][h] += sum over rel of dlut[h][rel] (fold the per-offset bias gradient into the 32 buckets)
__global__ void fold_dlut_kernel(const float* __restrict__ dlut, const int* __restrict__ buckets,
                                 float* __restrict__ dtable, int H, int L, int bucket_center, int nb) {
  const int h = blockIdx.x;
  const int n = 2 * L - 1;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j)
      if (buckets[j - (L - 1) + bucket_center] == b) acc += dlut[h * n + j];
    dtable[b * H + h] += acc;
  }
}
int fold_dlut(const float* dlut, const int* buckets, float* dtable, int H, int L, int bucket_center, int nb,
              cudaStream_t s) {
  fold_dlut_kernel<<<H, 32, 0, s>>>(dlut, buckets, dtable, H, L, bucket_center, nb);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
