// Passage-batched T5 encoder self-attention on tcgen05/TMEM, forward and backward.
//
// Reference: T5Attention.forward HF:modeling_t5.py:253-344 as driven per passage by
// EncoderWrapper.forward (/root/reference/src/model.py:227-234): scores = q k^T (NO 1/sqrt(dk)) +
// bucketed relative-position bias (shared by all layers, HF:755-758) + key-padding mask (finfo.min),
// fp32 softmax, probs @ v.   One "unit" = (passage, head); L = text_maxlength <= 256, d_kv = 64.
//
// Because a whole passage fits one key tile row (L <= 256 keys) the softmax needs no online rescaling: S = Q K^T lands
// in tensor memory (128 lanes x LK fp32 columns) and two threads share a query row (one key half each).
//   * forward, default (enc_attn_fwd2_kernel): persistent warp-specialised CTA per SM and head -- TMA producer warp,
//     MMA warp, two softmax groups ping-ponging over two 256-column TMEM slots; probabilities stay in tensor memory
//     (bf16 pairs written in place over the consumed scores, O = P V with P as the TMEM A operand).
//   * forward, round 1 (enc_attn_fwd_kernel; LAKO_ATTN_FWD=1): one-shot CTA per (query tile, head, passage), P through
//     128B-swizzled shared memory, two CTAs per SM.
//   * backward (enc_attn_bwd_kernel): persistent CTA per SM and head, 16 softmax warps + an issuing warp, all 512 TMEM
//     columns (S | dP | dQ x2 | dK | dV).
// Q/K/V tiles come in by TMA from the fused qkv activation [tokens][3*inner]; V is an MN-major operand straight from
// that buffer.  DESIGN.md section 3.2 has the measurements behind each choice.
#include <algorithm>
#include <cstdlib>

#include "ops.cuh"

namespace lako {

// forward implementation: 1 = one-shot CTA per (query tile, head, passage); 2 = persistent warp-specialised kernel with
// the probabilities kept in tensor memory (default).  LAKO_ATTN_FWD=1 / lako_set_attn_impl(1) selects the former;
// 3 sends forward AND backward through the any-length tiled kernels of enc_attn_long.cu (what L > 256 always uses).
static int g_attn_impl = -1;
int enc_attn_impl() {
  if (g_attn_impl < 0) {
    const char* e = getenv("LAKO_ATTN_FWD");
    g_attn_impl = (e && e[0] == '1') ? 1 : (e && e[0] == '3') ? 3 : 2;
  }
  return g_attn_impl;
}
void set_enc_attn_impl(int v) { g_attn_impl = (v == 1 || v == 3) ? v : 2; }

constexpr float MASK_MIN = -3.4028234663852886e38f;  // torch.finfo(float32).min, as HF's additive mask

// bias_lut[h][rel + L - 1] = table[bucket(rel)][h], rel = key_pos - query_pos
__global__ void build_bias_lut_kernel(const float* __restrict__ table, const int* __restrict__ buckets,
                                      float* __restrict__ lut, int H, int L, int bucket_center) {
  pdl_launch_dependents();
  pdl_wait();
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
  launch_pdl(build_bias_lut_kernel, dim3((n + 255) / 256), dim3(256), 0, s, table, buckets, lut, H, L, bucket_center);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

struct EncAttnParams {
  const unsigned char* mask;  // [P][L] 1 = valid key
  const float* bias_lut;      // [H][2L-1]
  __nv_bfloat16* out;         // [P*L][inner]
  float* lse;                 // [P][H][L]
  int L, H, inner;            // L = text_maxlength (LUT size, lse pitch); per-passage length from cu when packed
  const int* cu;              // packed (un-padded) token layout: passage p owns rows [cu[p], cu[p+1]); null = padded
  DropSpec drop;              // dropout on the attention probabilities (HF:332); group = ((psg*H+h)*L+q)*32 + key/8
  uint32_t* keepbits;         // [P][H][L][8] keep mask of the 256 keys of every query row, saved for the backward
};

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;

// Packed fp32 pairs (sm_100: FFMA2 / FADD2 work on 64-bit register pairs): (a0, a1) = (a0, a1) * s + (b0, b1)
__device__ __forceinline__ void ffma2(uint32_t& a0, uint32_t& a1, float s, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rs, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rs, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\t"
      "fma.rn.f32x2 ra, ra, rs, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "+r"(a0), "+r"(a1)
      : "f"(s), "f"(b0), "f"(b1));
}
// (d0, d1) = (a0, a1) * (b0, b1) + (c0, c1)
__device__ __forceinline__ void ffma2v(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 ra, ra, rb, rc;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
// (d0, d1) = (a0, a1) * (b0, b1)
__device__ __forceinline__ void fmul2v(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// (a0, a1) += (b0, b1)
__device__ __forceinline__ void fadd2(uint32_t& a0, uint32_t& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+r"(a0), "+r"(a1)
      : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fadd2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(b0), "f"(b1));
}
constexpr int LUTCOPY = 520;  // floats per shifted LUT copy (max index (L-1) + 255 + 3 < 520 for L <= 256)

// Four copies of the per-head bias LUT, copy j shifted left by j floats, so that every thread can fetch
// the 4 consecutive biases of 4 consecutive keys with ONE aligned 128-bit shared load whatever its
// query position is:  copy_j[x] = lut[x + j];  thread with base b = L-1-q uses j = b & 3 at offset b - j.
__device__ __forceinline__ void stage_lut4(float* s_lut4, const float* __restrict__ lut, int nlut, int tid,
                                           int nthreads) {
  for (int i = tid; i < 4 * LUTCOPY; i += nthreads) {
    const int j = i / LUTCOPY, x = i - j * LUTCOPY;
    const int src = x + j;
    s_lut4[i] = (src < nlut) ? lut[src] : 0.f;
  }
}
// The same biases as a table of overlapping windows, win[x] = (lut[x], lut[x+1], lut[x+2], lut[x+3]) * scale: the thread
// of query q fetches the biases of keys k..k+3 (k % 4 == 0) as the 16-byte entry x = (L-1-q) + k.  Consecutive lanes
// (consecutive queries) read consecutive entries, so every 8 lanes cover 128 contiguous bytes: conflict free (the 4
// shifted copies above cost 2 wavefronts per 8 lanes).
constexpr int LUTWIN = 512;  // entries; x <= 255 + 252
__device__ __forceinline__ void stage_lut_win(float* s_win, const float* __restrict__ lut, int nlut, int tid, int nthreads,
                                              float scale) {
  for (int i = tid; i < 4 * LUTWIN; i += nthreads) {
    const int src = (i >> 2) + (i & 3);
    s_win[i] = (src < nlut) ? lut[src] * scale : 0.f;
  }
}
// additive key mask: 0 valid, finfo.min for padded keys (HF semantics; it stays finfo.min in the log2 domain of the v2
// kernel, so a fully masked row still degrades to the uniform distribution), -inf beyond the passage
__device__ __forceinline__ void stage_kmask(float* s_kmask, const unsigned char* __restrict__ mask, int L, int tid,
                                            int nthreads) {
  for (int i = tid; i < 256; i += nthreads)
    s_kmask[i] = (i < L) ? ((mask == nullptr || mask[i]) ? 0.f : MASK_MIN) : -INFINITY;
}

// smem map (bytes from the 1024-aligned base):
//   [0, 32K)   K tile rows 0..255   | after S is read: P chunks 0,1
//   [32K, 48K) Q tile               | P chunk 2
//   [48K, 64K) spare                | P chunk 3
//   [64K, 96K) V tile rows 0..255
//   then: 4 shifted LUT copies, additive key mask, cross-half reduction scratch, barriers
constexpr int EA_K_OFF = 0, EA_Q_OFF = 32768, EA_V_OFF = 65536, EA_LUT_OFF = 98304,
              EA_KMASK_OFF = EA_LUT_OFF + 4 * LUTCOPY * 4, EA_RED_OFF = EA_KMASK_OFF + 1024,
              EA_BAR_OFF = EA_RED_OFF + 2048;
constexpr int EA_SMEM_BYTES = EA_BAR_OFF + 64 + 1024;

// 256 threads: warp w owns TMEM lane quarter (w & 3) = 32 query rows, and key half (w >> 2).
__global__ void __launch_bounds__(256, 2)
enc_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const EncAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_lut4 = reinterpret_cast<float*>(smem + EA_LUT_OFF);
  float* s_kmask = reinterpret_cast<float*>(smem + EA_KMASK_OFF);
  float* s_red_m = reinterpret_cast<float*>(smem + EA_RED_OFF);  // row maxima [2][128], then row sums [2][128] (red_a + 1024)
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + EA_BAR_OFF);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_o = bar_qk + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, chalf = warp >> 2;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, psg = blockIdx.z;
  const int Lmax = p.L;  // LUT geometry / lse pitch
  pdl_launch_dependents();
  int row0 = psg * Lmax, L = Lmax;  // L = this passage's token count
  if (p.cu) {
    pdl_wait();
    row0 = p.cu[psg];
    L = p.cu[psg + 1] - row0;
    if (q0 >= L) return;  // whole query tile beyond the passage (nothing allocated yet)
  }
  // Keys are processed in 32-column chunks, so S must be DEFINED for a whole number of chunks: with N rounded to 16
  // only, the tail columns of the last chunk would be whatever the TMEM columns held before (tcgen05.alloc does not
  // clear them; NaN bits left there by another kernel or process survive "+ bias - inf" and poison the row sum).
  // Rows of the K tile past the passage are the next passage's rows or TMA zero fill -- finite, and masked below.
  const int LK = (L + 31) / 32 * 32;
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
  pdl_wait();
  stage_lut4(s_lut4, p.bias_lut + h * (2 * Lmax - 1), 2 * Lmax - 1, tid, 256);
  stage_kmask(s_kmask, p.mask ? p.mask + static_cast<size_t>(psg) * Lmax : nullptr, L, tid, 256);
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

  const int r = quarter * 32 + lane;   // query row within the tile == TMEM lane
  const int q = q0 + r;                // query position in the passage
  const int qc = min(q, L - 1);        // rows beyond the passage compute on a clamped position, never stored
  const int b = Lmax - 1 - qc;
  const uint32_t lut_a = smem_u32(s_lut4) + ((b & 3) * LUTCOPY + (b & ~3)) * 4;  // explicit shared addresses (common.cuh)
  const uint32_t kmask_a = smem_u32(s_kmask), red_a = smem_u32(s_red_m), smem_a = smem_u32(smem);
  const uint32_t t_row = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
  const int nchunks = (LK + 31) / 32;
  const int c_begin = chalf * 4, c_end = min(nchunks, chalf * 4 + 4);

  mbar_wait(bar_s, 0);
  tc_fence_after();
  // Warps whose 32 query rows all lie beyond the passage do no softmax work: their P rows only feed O rows that
  // are never stored (the MMA is row-independent), so whatever bits sit in shared memory there are harmless.
  const bool warp_live = (q0 + quarter * 32) < L;
  const int c_lim = warp_live ? c_end : c_begin;
  // ---- pass 1: s = qk + bias + mask, row max; s goes back to TMEM
  float m = -INFINITY;
  for (int c = c_begin; c < c_lim; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(t_row + c * 32, v);
    tmem_ld_wait();
    const bool need_km = (p.mask != nullptr) || (c * 32 + 32 > L);  // warp-uniform
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 lb = lds_f4(lut_a + (c * 32 + g * 4) * 4);
      float s0 = __uint_as_float(v[g * 4 + 0]) + lb.x;
      float s1 = __uint_as_float(v[g * 4 + 1]) + lb.y;
      float s2 = __uint_as_float(v[g * 4 + 2]) + lb.z;
      float s3 = __uint_as_float(v[g * 4 + 3]) + lb.w;
      if (need_km) {
        const float4 km = lds_f4(kmask_a + (c * 32 + g * 4) * 4);
        s0 += km.x;
        s1 += km.y;
        s2 += km.z;
        s3 += km.w;
      }
      m = fmaxf(m, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
      v[g * 4 + 0] = __float_as_uint(s0);
      v[g * 4 + 1] = __float_as_uint(s1);
      v[g * 4 + 2] = __float_as_uint(s2);
      v[g * 4 + 3] = __float_as_uint(s3);
    }
    tmem_st_32x32(t_row + c * 32, v);
  }
  tmem_st_wait();
  sts_f(red_a + (chalf * 128 + r) * 4, m);
  __syncthreads();
  m = fmaxf(lds_f(red_a + r * 4), lds_f(red_a + (128 + r) * 4));
  // ---- pass 2: p = exp(s - m) (un-normalised), row sum, P -> shared (bf16, K-major 128B swizzle)
  const float nm = -m * LOG2E_F;
  float lsum = 0.f;
  for (int c = c_begin; c < c_lim; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(t_row + c * 32, v);
    tmem_ld_wait();
    const uint32_t chunk_a = smem_a + (c >> 1) * 16384 + r * 128;
    uint32_t kword = 0xffffffffu;  // keep bits of the chunk's 32 keys (drop_keep32_sliced order)
    if (p.drop.thresh)
      kword = drop_keep32_sliced(p.drop, (static_cast<unsigned long long>(psg * p.H + h) * Lmax + qc) * 32 + c * 4);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        e[i] = fast_exp2(fmaf(__uint_as_float(v[u * 8 + i]), LOG2E_F, nm));
        lsum += e[i];
      }
      if (p.drop.thresh) {  // the row sum normalises the UNdropped probabilities; only the P V operand is masked
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ((kword >> sliced_bit(u * 8 + i)) & 1u) ? e[i] * p.drop.scale : 0.f;
      }
      const int unit = (c & 1) * 4 + u;  // 16-byte unit index within the 128-byte row
      uint4 o;
      o.x = pack_bf16x2(e[0], e[1]);
      o.y = pack_bf16x2(e[2], e[3]);
      o.z = pack_bf16x2(e[4], e[5]);
      o.w = pack_bf16x2(e[6], e[7]);
      sts_u4(chunk_a + ((unit ^ (r & 7)) << 4), o);
    }
    if (p.keepbits && p.drop.thresh && q < L)
      p.keepbits[((static_cast<size_t>(psg) * p.H + h) * Lmax + q) * 8 + c] = kword;
  }
  sts_f(red_a + (256 + chalf * 128 + r) * 4, lsum);
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
  lsum = lds_f(red_a + (256 + r) * 4) + lds_f(red_a + (256 + 128 + r) * 4);
  mbar_wait(bar_o, 0);
  tc_fence_after();
  {
    uint32_t o0[32];
    tmem_ld_32x32(t_row + chalf * 32, o0);  // this warp's 32 of the 64 head dims
    tmem_ld_wait();
    if (q < L) {
      const float inv = 1.f / lsum;
      __nv_bfloat16* orow = p.out + static_cast<size_t>(row0 + q) * p.inner + h * 64 + chalf * 32;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(o0[u * 8 + 0]) * inv, __uint_as_float(o0[u * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(o0[u * 8 + 2]) * inv, __uint_as_float(o0[u * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(o0[u * 8 + 4]) * inv, __uint_as_float(o0[u * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(o0[u * 8 + 6]) * inv, __uint_as_float(o0[u * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + u * 8) = o;
      }
      if (p.lse && chalf == 0) p.lse[(static_cast<size_t>(psg) * p.H + h) * Lmax + q] = m + logf(lsum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ================================================================================================
// forward, version 2: persistent, warp-specialised, probabilities kept in tensor memory
// ================================================================================================
// One CTA per SM serves ONE head (CTA c -> head c % H) and walks that head's passages, so the shifted bias LUTs are
// staged once per CTA.  Roles: warp 0 = producer (TMA loads of a passage's Q / K / V into a 2-deep ring of
// shared-memory stages), warp 1 = one thread issuing every MMA, warps 2-9 and 10-17 = two softmax groups, each bound to
// one 256-column TMEM slot and taking every other 128-query tile:
//   S = Q K^T (SS MMA, N = keys rounded to 32) -> slot columns [0, LK)
//   two threads per query row (TMEM lane), one per half of the keys: pass 1 row max of s + bias + mask (exchanged
//   through shared memory), pass 2 exp -> bf16 pairs written IN PLACE over the consumed S columns by tcgen05.st
//   (keys 0-127 -> columns [0, 64), keys 128-255 -> columns [128, 192))
//   O = P V with P as the TMEM A operand (TS MMA) -> slot columns [64, 128); read back, scaled by 1 / row sum, stored.
// While one group runs its SIMT softmax the other group's MMAs / epilogue proceed and the next passage's tiles are
// already in flight; nothing is allocated or staged per tile.  (The one-shot kernel above spends most of a CTA's life
// in exposed latencies -- TMEM allocation, LUT staging, TMA, MMA and tcgen05.ld round trips -- at 2 CTAs per SM.)
constexpr int A2_STAGE = 98304;  // Q0 16K | Q1 16K | K 32K | V 32K
constexpr int A2_K_OFF = 32768, A2_V_OFF = 65536;
constexpr int A2_LUT_OFF = 2 * A2_STAGE;                      // 4 shifted LUT copies of this CTA's head
constexpr int A2_KMASK_OFF = A2_LUT_OFF + 4 * LUTWIN * 4;    // 2 stages x 256 floats (explicit key masks only)
constexpr int A2_RED_OFF = A2_KMASK_OFF + 2 * 1024;           // row max / row sum exchange: [2 groups][2][2 halves][128]
constexpr int A2_BAR_OFF = A2_RED_OFF + 2 * 2 * 2 * 128 * 4;
constexpr int A2_SMEM_BYTES = A2_BAR_OFF + 256 + 1024;
constexpr int A2_THREADS = 64 + 512;
constexpr int A2_O_COL = 64;  // O accumulator columns inside a slot

#ifdef LAKO_TRACE
__device__ long long g_a2_trace[256];
// role: 0 producer, 1 MMA, 2 softmax group 0, 3 softmax group 1; 64 stamps each (CTA 0 only)
#define A2T(role, i, cond) do { if (blockIdx.x == 0 && (cond) && (i) < 64) g_a2_trace[(role) * 64 + (i)] = clock64(); } while (0)
#else
#define A2T(role, i, cond) do { } while (0)
#endif
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool kMask, bool kDrop>
__global__ void __launch_bounds__(A2_THREADS, 1)
enc_attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const EncAttnParams p, const int P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_lut4 = reinterpret_cast<float*>(smem + A2_LUT_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A2_BAR_OFF);
  uint64_t* stage_full = bars;        // [2] TMA bytes of a passage landed (+ its key mask staged)
  uint64_t* stage_free = bars + 2;    // [2] every MMA reading the stage has completed
  uint64_t* s_full = bars + 4;        // [2] S of a tile is in the slot
  uint64_t* p_ready = bars + 6;       // [2] P of a tile is in the slot (256 softmax threads)
  uint64_t* o_full = bars + 8;        // [2] O of a tile is in the slot
  uint64_t* slot_free = bars + 10;    // [2] the slot's O has been read out (256 softmax threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  // warp index through a shuffle: provably warp-uniform, so roles, TMEM addresses and loop state live in uniform registers
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int Lmax = p.L, H = p.H;
  const int h = blockIdx.x % H;                  // this CTA's head
  const int pstep = gridDim.x / H;               // CTAs serving this head
  const int p_first = blockIdx.x / H;            // first passage of this CTA
  pdl_launch_dependents();
  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&stage_full[i], 2);   // the TMA's expect_tx arrival + the producer warp's "mask staged" arrival
      mbar_init(&stage_free[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&slot_free[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();
  stage_lut_win(s_lut4, p.bias_lut + h * (2 * Lmax - 1), 2 * Lmax - 1, tid, A2_THREADS, LOG2E_F);  // log2 domain
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // passage geometry, recomputed identically by every role
  auto psg_geom = [&](int psg, int& row0, int& L) {
    row0 = psg * Lmax;
    L = Lmax;
    if (p.cu) {
      row0 = p.cu[psg];
      L = p.cu[psg + 1] - row0;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    int ui = 0;
    for (int psg = p_first; psg < P; psg += pstep) {
      int row0, L;
      psg_geom(psg, row0, L);
      if (L <= 0) continue;
      const int stage = ui & 1;
      const int LK = (L + 31) / 32 * 32, ktiles = (LK + 127) / 128, nt = (L + 127) / 128;
      A2T(0, ui * 3 + 0, lane == 0);
      mbar_wait(&stage_free[stage], ((ui >> 1) & 1) ^ 1);
      A2T(0, ui * 3 + 1, lane == 0);
      uint8_t* st = smem + stage * A2_STAGE;
      if (lane == 0) {
        mbar_arrive_expect_tx(&stage_full[stage], 16384 * (nt + 2 * ktiles));
        for (int t = 0; t < nt; ++t) tma_load_2d(st + t * 16384, &tm_qkv, &stage_full[stage], h * 64, row0 + t * 128);
        for (int t = 0; t < ktiles; ++t) {
          tma_load_2d(st + A2_K_OFF + t * 16384, &tm_qkv, &stage_full[stage], p.inner + h * 64, row0 + t * 128);
          tma_load_2d(st + A2_V_OFF + t * 16384, &tm_qkv, &stage_full[stage], 2 * p.inner + h * 64, row0 + t * 128);
        }
      }
      if (kMask)  // explicit key mask (padded layout): additive form for this passage
        stage_kmask(reinterpret_cast<float*>(smem + A2_KMASK_OFF) + stage * 256, p.mask + static_cast<size_t>(psg) * Lmax, L,
                    lane, 32);
      __syncwarp();
      if (lane == 0) mbar_arrive(&stage_full[stage]);  // release: the warp's mask writes precede it
      A2T(0, ui * 3 + 2, lane == 0);
      ++ui;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);
      int ui = 0, n = 0;
      bool pend = false;
      int pend_slot = 0, pend_stage = 0, pend_LK = 0, pend_k = 0;
      bool pend_last = false;
      auto do_pv = [&]() {
        mbar_wait(&p_ready[pend_slot], pend_k & 1);
        tc_fence_after();
        A2T(1, (pend_k * 2 + pend_slot) * 4 + 2, true);
        const uint32_t t_slot = tmem + pend_slot * 256;
        const uint32_t va = smem_u32(smem + pend_stage * A2_STAGE + A2_V_OFF);
        const int ksteps = pend_LK / 16;
        for (int j = 0; j < ksteps; ++j)  // P: keys 0-127 at columns [0, 64), keys 128-255 at [128, 192)
          umma_bf16_ts(t_slot + A2_O_COL, t_slot + (j < 8 ? j * 8 : 128 + (j - 8) * 8),
                       make_smem_desc_sw128(va + j * 2048, 8192, 1024), idesc_pv, j > 0);
        umma_commit(&o_full[pend_slot]);
        if (pend_last) umma_commit(&stage_free[pend_stage]);
        pend = false;
      };
      for (int psg = p_first; psg < P; psg += pstep) {
        int row0, L;
        psg_geom(psg, row0, L);
        if (L <= 0) continue;
        const int stage = ui & 1;
        const int LK = (L + 31) / 32 * 32, nt = (L + 127) / 128;
        mbar_wait(&stage_full[stage], (ui >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + stage * A2_STAGE), ka = qa + A2_K_OFF;
        const uint32_t idesc_s = make_idesc_bf16(128, LK, false, false);
        for (int t = 0; t < nt; ++t) {
          const int slot = n & 1, k = n >> 1;
          A2T(1, n * 4 + 0, true);
          mbar_wait(&slot_free[slot], (k & 1) ^ 1);
          tc_fence_after();
          A2T(1, n * 4 + 1, true);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem + slot * 256, make_smem_desc_sw128(qa + t * 16384 + kk * 32, 16, 1024),
                      make_smem_desc_sw128(ka + kk * 32, 16, 1024), idesc_s, kk > 0);
          umma_commit(&s_full[slot]);
          if (pend) do_pv();  // the previous tile's P V goes behind this tile's Q K^T: its softmax ran meanwhile
          pend = true;
          pend_slot = slot;
          pend_stage = stage;
          pend_LK = LK;
          pend_k = k;
          pend_last = (t == nt - 1);
          ++n;
        }
        ++ui;
      }
      if (pend) do_pv();
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int sw = warp - 2;
    const int group = sw >> 3;           // == TMEM slot
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int chalf = (sw >> 2) & 1;     // key half: chunks [4 chalf, 4 chalf + 4)
    const int r = quarter * 32 + lane;   // query row within the tile == TMEM lane
    const uint32_t t_row = tmem + group * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    float* s_red = reinterpret_cast<float*>(smem + A2_RED_OFF) + group * 512;  // [max | sum][2 halves][128]
    int ui = 0, n = 0;
    for (int psg = p_first; psg < P; psg += pstep) {
      int row0, L;
      psg_geom(psg, row0, L);
      if (L <= 0) continue;
      const int stage = ui & 1;
      const int LK = (L + 31) / 32 * 32, nt = (L + 127) / 128, nchunks = LK / 32;
      const float* s_kmask = reinterpret_cast<const float*>(smem + A2_KMASK_OFF) + stage * 256;
      const int c_begin = chalf * 4, c_end = min(nchunks, chalf * 4 + 4);
      for (int t = 0; t < nt; ++t, ++n) {
        if ((n & 1) != group) continue;
        const int k = n >> 1;
        const int q0 = t * 128;
        const int q = q0 + r, qc = min(q, L - 1);
        const int b = Lmax - 1 - qc;
        const uint32_t lut_a = smem_u32(s_lut4) + b * 16;  // shared address of this row's bias windows
        const bool trace = sw == (group * 8) && lane == 0;
        (void)trace;
        A2T(2 + group, k * 6 + 0, trace);
        // dropout keep words of this thread's four 32-key chunks, drawn under the wait for S
        uint32_t kw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (kDrop) {
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
            if (c_begin + ci < c_end)
              kw[ci] = drop_keep32_sliced(p.drop, (static_cast<unsigned long long>(psg * H + h) * Lmax + qc) * 32 + (c_begin + ci) * 4);
        }
        if (kMask) mbar_wait(&stage_full[stage], (ui >> 1) & 1);  // the passage's key mask is staged
        mbar_wait(&s_full[group], k & 1);
        tc_fence_after();
        A2T(2 + group, k * 6 + 1, trace);
        // warps whose 32 query rows all lie beyond the passage do no arithmetic (the MMA is row independent: whatever their
        // P rows hold only reaches O rows that are never stored) but keep the protocol (barriers, arrivals)
        const bool warp_live = (q0 + quarter * 32) < L;
        const int nlive = warp_live ? (c_end - c_begin) : 0;  // 32-key chunks of this half: 0..4, warp-uniform
        const uint32_t t_half = t_row + chalf * 128;            // this half's S columns
        const uint32_t lut_h = lut_a + chalf * 2048;            // ... and its (log2e-scaled) bias windows
        const uint32_t km_h = smem_u32(s_kmask) + chalf * 512;
        // ---- pass 1: t = s * log2(e) + bias' (+ mask), row max; t goes back to its TMEM columns so that pass 2 is
        //      exp2(t - max) only.  Everything is unrolled over the 4 chunks: TMEM / shared addresses are immediates.
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          if (ci < nlive) {
            uint32_t v[32];
            tmem_ld_32x32(t_half + ci * 32, v);
            tmem_ld_wait();
            if (!kMask && (c_begin + ci) * 32 + 32 > L) {  // packed rows: only the passage's last chunk has dead keys
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if ((c_begin + ci) * 32 + i >= L) v[i] = 0xff800000u;  // -inf
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 lb = lds_f4(lut_h + (ci * 32 + g * 4) * 16);
              ffma2(v[g * 4 + 0], v[g * 4 + 1], LOG2E_F, lb.x, lb.y);
              ffma2(v[g * 4 + 2], v[g * 4 + 3], LOG2E_F, lb.z, lb.w);
              if (kMask) {
                const float4 km = lds_f4(km_h + (ci * 32 + g * 4) * 4);
                fadd2(v[g * 4 + 0], v[g * 4 + 1], km.x, km.y);
                fadd2(v[g * 4 + 2], v[g * 4 + 3], km.z, km.w);
              }
              m0 = fmaxf(m0, fmaxf(__uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1])));
              m1 = fmaxf(m1, fmaxf(__uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3])));
            }
            tmem_st_32x32(t_half + ci * 32, v);
          }
        }
        float m = fmaxf(m0, m1);
        const uint32_t red_a = smem_u32(s_red);
        sts_f(red_a + (chalf * 128 + r) * 4, m);
        tmem_st_wait();
        named_bar_sync(1 + group, 256);
        m = fmaxf(lds_f(red_a + r * 4), lds_f(red_a + (128 + r) * 4));  // log2 domain
        const float nm = -m;
        A2T(2 + group, k * 6 + 2, trace);
        // ---- pass 2: p = exp2(t - max) (un-normalised), row sum; bf16 pairs overwrite consumed S columns in place:
        //      chunk ci of this half (S columns chalf*128 + 32 ci ..) -> P columns chalf*128 + 16 ci ..  Dropout only
        //      zeroes entries here (the 1/(1-p) factor is applied with 1 / row sum in the epilogue).
        float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          if (ci < nlive) {
            uint32_t v[32];
            tmem_ld_32x32(t_half + ci * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              fadd2(v[g * 4 + 0], v[g * 4 + 1], nm, nm);
              fadd2(v[g * 4 + 2], v[g * 4 + 3], nm, nm);
              float e0 = fast_exp2(__uint_as_float(v[g * 4 + 0])), e1 = fast_exp2(__uint_as_float(v[g * 4 + 1]));
              float e2 = fast_exp2(__uint_as_float(v[g * 4 + 2])), e3 = fast_exp2(__uint_as_float(v[g * 4 + 3]));
              fadd2(ls0, ls1, e0, e1);
              fadd2(ls2, ls3, e2, e3);
              pk[g * 2 + 0] = pack_bf16x2(e0, e1);
              pk[g * 2 + 1] = pack_bf16x2(e2, e3);
            }
            if (kDrop) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] &= sliced_pair_mask(kw[ci], i);
              if (p.keepbits && q < L) p.keepbits[((static_cast<size_t>(psg) * H + h) * Lmax + q) * 8 + c_begin + ci] = kw[ci];
            }
            tmem_st_32x16(t_row + chalf * 128 + ci * 16, pk);
          }
        }
        const float lsum_half = (ls0 + ls1) + (ls2 + ls3);
        sts_f(red_a + (256 + chalf * 128 + r) * 4, lsum_half);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[group]);
        A2T(2 + group, k * 6 + 3, trace);
        // ---- epilogue: O / row sum (each half reads 32 of the 64 head dims)
        mbar_wait(&o_full[group], k & 1);
        tc_fence_after();
        A2T(2 + group, k * 6 + 4, trace);
        named_bar_sync(1 + group, 256);  // both halves' row sums are in shared memory
        const float lsum = lds_f(red_a + (256 + r) * 4) + lds_f(red_a + (256 + 128 + r) * 4);
        {
          uint32_t o0[32];
          tmem_ld_32x32(t_row + A2_O_COL + chalf * 32, o0);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&slot_free[group]);  // the slot may take the next tile's S while the stores below drain
          if (q < L) {
            const float inv = (kDrop ? p.drop.scale : 1.f) / lsum;
            __nv_bfloat16* orow = p.out + static_cast<size_t>(row0 + q) * p.inner + h * 64 + chalf * 32;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(o0[uu * 8 + 0]) * inv, __uint_as_float(o0[uu * 8 + 1]) * inv);
              w.y = pack_bf16x2(__uint_as_float(o0[uu * 8 + 2]) * inv, __uint_as_float(o0[uu * 8 + 3]) * inv);
              w.z = pack_bf16x2(__uint_as_float(o0[uu * 8 + 4]) * inv, __uint_as_float(o0[uu * 8 + 5]) * inv);
              w.w = pack_bf16x2(__uint_as_float(o0[uu * 8 + 6]) * inv, __uint_as_float(o0[uu * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + uu * 8) = w;
            }
            if (p.lse && chalf == 0) p.lse[(static_cast<size_t>(psg) * H + h) * Lmax + q] = (m + log2f(lsum)) * LN2_F;
          }
        }
        A2T(2 + group, k * 6 + 5, trace);
      }
      ++ui;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Test support: fill all 512 TMEM columns of every SM with NaN bit patterns (what a previous tenant of the GPU may
// leave behind), so that tests can prove no kernel consumes TMEM it did not write.
__global__ void __launch_bounds__(128, 1) poison_tmem_kernel() {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0x7fc00000u;
  for (int c = 0; c < 16; ++c) tmem_st_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
  tmem_st_wait();
  // keep the SM occupied long enough that the block scheduler places one CTA on every SM
  const long long t0 = clock64();
  while (clock64() - t0 < 200000) {
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
int poison_tmem(cudaStream_t s) {
  int rc = check_arch();
  if (rc) return rc;
  cudaFuncSetAttribute(poison_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  poison_tmem_kernel<<<device_sm_count() * 2, 128, 200 * 1024, s>>>();  // 200 KB dynamic smem: one CTA per SM at a time
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// qkv bf16 [P*L][3*inner]; out bf16 [P*L][inner]; lse fp32 [P][H][L] (may be null)
int enc_attn_fwd(const __nv_bfloat16* qkv, const unsigned char* mask, const float* bias_lut, __nv_bfloat16* out,
                 float* lse, int P, int L, int H, cudaStream_t s, DropSpec drop, const int* cu, int packed_rows,
                 uint32_t* keepbits) {
  int rc = check_arch();
  if (rc) return rc;
  LAKO_REQUIRE(P > 0 && H > 0 && L > 0, LAKO_ERR_BAD_ARG);
  // the kernels below hold a passage's whole key range in one softmax tile; longer passages take the tiled path (which
  // regenerates the keep words in its backward and leaves `keepbits` untouched)
  if (L > 256 || enc_attn_impl() == 3) return enc_attn_long_fwd(qkv, mask, bias_lut, out, lse, P, L, H, s, drop, cu);
  const int inner = H * 64;
  CUtensorMap tm;
  const uint64_t total_rows = cu ? static_cast<uint64_t>(packed_rows) : static_cast<uint64_t>(P) * L;
  LAKO_REQUIRE(total_rows > 0, LAKO_ERR_BAD_ARG);
  rc = get_tensor_map_2d(&tm, qkv, 3 * inner, total_rows, 3 * inner, 64, 128);
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
  p.cu = cu;
  p.drop = drop;
  p.keepbits = keepbits;
  ProfScope prof(PROF_ENC_ATTN_FWD, 4.0 * P * H * static_cast<double>(L) * L * 64, s);
  if (enc_attn_impl() == 2 && H <= device_sm_count()) {
    using Kern = void (*)(CUtensorMap, EncAttnParams, int);
    static const Kern kerns[4] = {enc_attn_fwd2_kernel<false, false>, enc_attn_fwd2_kernel<false, true>,
                                  enc_attn_fwd2_kernel<true, false>, enc_attn_fwd2_kernel<true, true>};
    static bool configured2 = false;
    if (!configured2) {
      for (Kern kf : kerns) {
        cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM_BYTES);
        if (e != cudaSuccess) {
          set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
          return LAKO_ERR_LAUNCH;
        }
      }
      configured2 = true;
    }
    // every CTA serves one head: (SMs / H) CTAs per head (at most one per passage)
    const int per_head = std::max(1, std::min(P, device_sm_count() / H));
    launch_pdl(kerns[(mask ? 2 : 0) + (drop.thresh ? 1 : 0)], dim3(per_head * H), dim3(A2_THREADS), A2_SMEM_BYTES, s, tm, p, P);
    LAKO_CHECK_LAUNCH();
    return LAKO_OK;
  }
  dim3 grid((L + 127) / 128, H, P);
  launch_pdl(enc_attn_fwd_kernel, grid, dim3(256), EA_SMEM_BYTES, s, tm, p);
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
  const int* cu;                 // packed layout offsets (see EncAttnParams); null = padded
  DropSpec drop;
  const uint32_t* keepbits;      // keep masks saved by the forward (null: regenerate them with Philox)
  int d_in_out;                  // 1: `out` rows no longer hold O but, in their first H floats, D[q][h] = sum_d dO * O
                                 // (attn_rowdot_inplace); 0: the kernel forms D from O and dO itself
};

// bias-gradient bins in shared memory: bin EB_DLUT_PAD + (k - q + L - 1); the padding takes the (all-zero) sums of the
// diagonals of a 32x32 sub-block that fall outside the table, so no sum needs a range check
constexpr int EB_DLUT_PAD = 32, EB_DLUT_BINS = 576;
constexpr int EB_Q_OFF = 0, EB_DO_OFF = 32768, EB_K_OFF = 65536, EB_V_OFF = 98304, EB_P_OFF = 131072,
              EB_DS_OFF = 163840, EB_LUT_OFF = 196608, EB_DLUT_OFF = EB_LUT_OFF + 4 * LUTWIN * 4,
              EB_KMASK_OFF = EB_DLUT_OFF + EB_DLUT_BINS * 4, EB_LSE_OFF = EB_KMASK_OFF + 1024, EB_D_OFF = EB_LSE_OFF + 1024,
              EB_SCR_OFF = EB_D_OFF + 1024, EB_BAR_OFF = EB_SCR_OFF + 16 * 128;
constexpr int EB_SMEM_BYTES = EB_BAR_OFF + 64 + 1024;
constexpr int TM_S = 0, TM_DP = 128, TM_DQ = 256, TM_DK = 384, TM_DV = 448;

// One persistent CTA per SM, bound to a head, walks that head's passages; 16 softmax warps (the arithmetic between the
// MMAs is latency bound, so the CTA carries
// as many warps as the register file allows) + 1 warp that issues the TMA loads and every MMA: softmax warp w owns
// TMEM lane quarter (w & 3) = 32 query rows and key columns (w >> 2)*32..+32 of each 128x128 (query tile x key tile)
// block.  The issuing warp runs ahead: S/dP of block i+1 and the 24 gradient MMAs of block i overlap the softmax
// arithmetic of block i+1 (a softmax thread that also issued the MMAs held the whole CTA back ~35 % per block).
// TMEM: S 128 | dP 128 | dQ 2x64 | dK 64 | dV 64 columns.
constexpr int EB_SOFTMAX_THREADS = 512;
constexpr int EB_THREADS = EB_SOFTMAX_THREADS + 32;
#ifdef LAKO_TRACE
__device__ long long g_eb_trace[64];
#define EB_TRACE(i) do { if (tid == 0 && blockIdx.x == 0) g_eb_trace[i] = clock64(); } while (0)
#else
#define EB_TRACE(i) do { } while (0)
#endif
template <bool kDrop>
__global__ void __launch_bounds__(EB_THREADS, 1)
enc_attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                    const EncAttnBwdParams p, const int P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_lut4 = reinterpret_cast<float*>(smem + EB_LUT_OFF);
  float* s_dlut = reinterpret_cast<float*>(smem + EB_DLUT_OFF);
  float* s_kmask = reinterpret_cast<float*>(smem + EB_KMASK_OFF);
  float* s_lse = reinterpret_cast<float*>(smem + EB_LSE_OFF);
  float* s_D = reinterpret_cast<float*>(smem + EB_D_OFF);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + EB_BAR_OFF);
  uint64_t* bar_mma = bar_load + 1;
  uint64_t* bar_grad = bar_load + 2;
  uint64_t* bar_ready = bar_load + 3;  // all softmax threads: P / dS of the block stored
  uint64_t* bar_sfree = bar_load + 4;  // all softmax threads: S / dP of the block are in registers (TMEM columns free)
  uint64_t* bar_kv = bar_load + 5;     // dK / dV of a key tile are final (committed before that block's 8 dQ MMAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 6);

  // warp index through a shuffle: provably warp-uniform (the issuing warp's loops then run on uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int quarter = warp & 3, cq = warp >> 2;
  const bool issuer = tid == EB_SOFTMAX_THREADS;  // lane 0 of the 17th warp
  // Persistent: CTA c serves head c % H and walks that head's passages (c / H, + CTAs per head, ...): barriers, tensor
  // memory, the bias windows and the bias-gradient accumulator are set up once per CTA, and the next passage's operand
  // tiles stream in (TMA) underneath the current passage's dQ read-out and the next one's lse / D staging.  (One CTA per
  // (head, passage) spent ~35 % of its life in that un-overlapped prologue and epilogue.)
  const int H = p.H;
  const int h = blockIdx.x % H;
  const int p_first = blockIdx.x / H;
  // CTAs per head = the passage stride of this CTA, kept in shared memory and re-read at each passage change: as a
  // register it was spilled, and its reload (local memory = the global-store path) sat behind the dQ / dK / dV stores
  // for ~8 % of the kernel's warp samples.
  int* s_misc = reinterpret_cast<int*>(tmem_slot + 1);
  if (tid == 0) s_misc[0] = gridDim.x / H;
  auto pstep_v = [&]() -> int { return static_cast<int>(lds_u32(smem_u32(s_misc))); };
  const int Lmax = p.L;  // LUT geometry / lse pitch
  const int nlut = 2 * Lmax - 1;
  EB_TRACE(0);
  pdl_launch_dependents();
  if (issuer) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_grad, 1);
    mbar_init(bar_ready, EB_SOFTMAX_THREADS);
    mbar_init(bar_sfree, EB_SOFTMAX_THREADS);
    mbar_init(bar_kv, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  pdl_wait();
  EB_TRACE(1);
  stage_lut_win(s_lut4, p.bias_lut + h * nlut, nlut, tid, EB_THREADS, LOG2E_F);  // log2 domain
  for (int i = tid; i < EB_DLUT_BINS; i += EB_THREADS) s_dlut[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  EB_TRACE(2);
  const uint32_t tmem = *tmem_slot;
  auto psg_geom = [&](int psg, int& row0, int& L) {  // recomputed identically by every role
    row0 = psg * Lmax;
    L = Lmax;
    if (p.cu) {
      row0 = p.cu[psg];
      L = p.cu[psg + 1] - row0;
    }
    row0 = __shfl_sync(0xffffffffu, row0, 0);  // provably warp-uniform
    L = __shfl_sync(0xffffffffu, L, 0);
  };

  const uint32_t qa = smem_u32(smem + EB_Q_OFF), doa = smem_u32(smem + EB_DO_OFF), ka = smem_u32(smem + EB_K_OFF),
                 va = smem_u32(smem + EB_V_OFF), pa = smem_u32(smem + EB_P_OFF), dsa = smem_u32(smem + EB_DS_OFF);
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);    // S, dP : K-major x K-major
  constexpr uint32_t idesc_tn = make_idesc_bf16(128, 64, true, true);      // dV, dK: MN-major x MN-major
  constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64, false, true);     // dQ    : K-major x MN-major

  // descriptor start-address fields count 16-byte units: + 2 per 32-byte K step of a K-major tile, + 128 per 2048-byte
  // step of an MN-major one, + 1024 per 16 KB tile
  const uint64_t qd = make_smem_desc_sw128(qa, 16, 1024), dod = make_smem_desc_sw128(doa, 16, 1024),
                 kd = make_smem_desc_sw128(ka, 16, 1024), vd = make_smem_desc_sw128(va, 16, 1024);
  const uint64_t p_t = make_smem_desc_sw128(pa, 16384, 1024), ds_t = make_smem_desc_sw128(dsa, 16384, 1024),
                 ds_k = make_smem_desc_sw128(dsa, 16, 1024);
  const uint64_t do_mn = make_smem_desc_sw128(doa, 8192, 1024), q_mn = make_smem_desc_sw128(qa, 8192, 1024),
                 k_mn = make_smem_desc_sw128(ka, 8192, 1024);
  auto issue_s_dp = [&](int qt, int kt) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_S, qd + qt * 1024 + k * 2, kd + kt * 1024 + k * 2, idesc_s, k > 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_DP, dod + qt * 1024 + k * 2, vd + kt * 1024 + k * 2, idesc_s, k > 0);
  };

  if (warp == EB_SOFTMAX_THREADS / 32) {
    // ---------------- issuing warp: the whole warp walks the loop and waits (uniform control flow, descriptors in uniform
    // registers), one elected lane issues.  With the loop inside an `if (one thread)` region every MMA is wrapped in a
    // divergence "waterfall" (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~130 cycles per MMA): 32 MMAs per block then cost more
    // issue time than the block's softmax arithmetic.
    const bool el = elect_one();  // (MMA / TMA issue only inside real branches on it; see DESIGN.md on elected producers)
    uint32_t r_phase = 0, l_phase = 0;
    uint32_t n_commits = 0;  // gradient commits (= bar_grad phases) issued so far
    for (int psg = p_first; psg < P; psg += pstep_v()) {
      int row0, L;
      psg_geom(psg, row0, L);
      if (L <= 0) continue;
      const int nt = (L + 127) / 128;
      // the previous passage's gradient MMAs have finished reading Q / K / V / dO / P / dS: its tiles may be replaced
      // (only the LAST commit can still be pending: every earlier phase completed before the softmax threads released
      // the block after it; an mbarrier parity wait can only tell the current phase from the one before)
      if (n_commits > 0) mbar_wait_warp(bar_grad, (n_commits - 1) & 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_load, 16384 * 4 * nt);
        for (int t = 0; t < nt; ++t) {
          tma_load_2d(smem + EB_Q_OFF + t * 16384, &tm_qkv, bar_load, h * 64, row0 + t * 128);
          tma_load_2d(smem + EB_K_OFF + t * 16384, &tm_qkv, bar_load, p.inner + h * 64, row0 + t * 128);
          tma_load_2d(smem + EB_V_OFF + t * 16384, &tm_qkv, bar_load, 2 * p.inner + h * 64, row0 + t * 128);
          tma_load_2d(smem + EB_DO_OFF + t * 16384, &tm_do, bar_load, h * 64, row0 + t * 128);
        }
      }
      __syncwarp();
      {
        // The NEXT passage's tiles go to L2 now: all CTAs change passage at about the same time, and without this the
        // ~190 KB each then asks for arrive as one HBM burst (~28 MB chip-wide) that the staging has to sit out.
        int nrow0 = 0, nL = 0;
        const int npsg = psg + pstep_v();
        if (npsg < P) psg_geom(npsg, nrow0, nL);
        if (lane == 0 && nL > 0) {
          for (int t = 0; t < (nL + 127) / 128; ++t) {
            tma_prefetch_2d(&tm_qkv, h * 64, nrow0 + t * 128);
            tma_prefetch_2d(&tm_qkv, p.inner + h * 64, nrow0 + t * 128);
            tma_prefetch_2d(&tm_qkv, 2 * p.inner + h * 64, nrow0 + t * 128);
            tma_prefetch_2d(&tm_do, h * 64, nrow0 + t * 128);
          }
        }
        __syncwarp();
      }
      mbar_wait_warp(bar_load, l_phase);
      l_phase ^= 1;
      tc_fence_after();
      if (el) {
        issue_s_dp(0, 0);
        umma_commit(bar_mma);
      }
      for (int kt = 0; kt < nt; ++kt) {
        for (int qt = 0; qt < nt; ++qt) {
          // The softmax threads release the S / dP columns as soon as their values are in registers (about half way
          // through the block's arithmetic): the next block's S / dP are computed underneath the rest of it.
          mbar_wait_warp(bar_sfree, r_phase);
          tc_fence_after();
          int nq = qt + 1, nk = kt;
          if (nq == nt) {
            nq = 0;
            nk = kt + 1;
          }
          if (nk < nt && el) {
            issue_s_dp(nq, nk);
            umma_commit(bar_mma);
          }
          mbar_wait_warp(bar_ready, r_phase);  // P / dS of this block are in shared memory
          r_phase ^= 1;
          tc_fence_after();
          if (el) {
            // dV[kt] += P^T dO[qt] ; dK[kt] += dS^T Q[qt]   (M = 128 keys, N = 64, K = 128 queries)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              umma_bf16(tmem + TM_DV, p_t + j * 128, do_mn + qt * 1024 + j * 128, idesc_tn, (qt > 0 || j > 0));
#pragma unroll
            for (int j = 0; j < 8; ++j)
              umma_bf16(tmem + TM_DK, ds_t + j * 128, q_mn + qt * 1024 + j * 128, idesc_tn, (qt > 0 || j > 0));
            // the key tile's last block: its dK / dV can be read out while the 8 dQ MMAs below still run
            if (qt == nt - 1) umma_commit(bar_kv);
            // dQ[qt] += dS K[kt]   (M = 128 queries, N = 64, K = 128 keys)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              umma_bf16(tmem + TM_DQ + qt * 64, ds_k + (j >> 2) * 1024 + (j & 3) * 2, k_mn + kt * 1024 + j * 128, idesc_dq,
                        (kt > 0 || j > 0));
            umma_commit(bar_grad);
          }
          ++n_commits;
        }
      }
    }
  } else {
  // ---------------- softmax warps
  const int r = quarter * 32 + lane;  // row within a 128-row tile == TMEM lane
  const uint32_t t_row = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
  const float keep = (lane == 0) ? 0.f : 1.f;  // systolic diagonal sums: lane 0 starts a new diagonal
  // Two completion streams: bar_mma = "S and dP of the next block are in TMEM", bar_grad = "the gradient MMAs of the
  // previous block have finished reading P / dS from shared memory".  The S/dP MMAs of block i+1 are issued AHEAD of
  // the gradient MMAs of block i, so the softmax arithmetic of block i+1 (registers only) overlaps the 24 gradient
  // MMAs of block i; only the P / dS stores wait for them.
  uint32_t s_phase = 0, g_phase = 0, kv_phase = 0;
  bool g_pending = false;
  float pre_lse = INFINITY, pre_D = 0.f;  // (lse, D) of this thread's row in the NEXT passage (d_in_out mode)
  bool have_pre = false;

  for (int psg = p_first; psg < P; psg += pstep_v()) {
  int row0, L;
  psg_geom(psg, row0, L);
  if (L <= 0) continue;
  const int nt = (L + 127) / 128;  // 128-row tiles of queries == of keys
  EB_TRACE(3);
  {
    // Per-passage staging: additive key mask, lse and D[q] = sum_d dO * O (two threads per query row, 32 head dims
    // each).  No thread gets here before every thread has read the previous passage's values for the last time (the last
    // block's gradient MMAs, which everybody waits for, are only issued once all have arrived on bar_ready).
    unsigned char mk = 1;
    if (tid < L && p.mask) mk = __ldg(p.mask + static_cast<size_t>(psg) * Lmax + tid);
    if (tid < 256) s_kmask[tid] = (tid < L) ? (mk ? 0.f : MASK_MIN) : -INFINITY;
    int nrow0 = 0, nL = 0;  // the next passage of this CTA
    const int npsg = psg + pstep_v();
    if (npsg < P) psg_geom(npsg, nrow0, nL);
    if (p.d_in_out) {
      // D was formed by attn_rowdot_inplace: one (lse, D) pair per query row, and the NEXT passage's pair is fetched into
      // two registers a whole passage ahead, so this staging exposes no memory latency at all.
      if (!have_pre && tid < L) {
        pre_lse = __ldg(p.lse + (static_cast<size_t>(psg) * H + h) * Lmax + tid);
        pre_D = __ldg(reinterpret_cast<const float*>(p.out + static_cast<size_t>(row0 + tid) * p.inner) + h);
      }
      if (tid < 256) {
        s_lse[tid] = (tid < L) ? pre_lse : INFINITY;  // rows beyond the passage -> P = exp(s - inf) = 0
        s_D[tid] = (tid < L) ? pre_D : 0.f;
      }
      named_bar_sync(1, EB_SOFTMAX_THREADS);
      have_pre = true;
      if (tid < nL) {
        pre_lse = __ldg(p.lse + (static_cast<size_t>(npsg) * H + h) * Lmax + tid);
        pre_D = __ldg(reinterpret_cast<const float*>(p.out + static_cast<size_t>(nrow0 + tid) * p.inner) + h);
      }
    } else {
      const int q = tid >> 1, part = tid & 1;  // two threads per query row, 32 head dims each
      uint4 ov[4], dv[4];
      float lse_q = INFINITY;
      if (q < L) {
        const uint4* o4 = reinterpret_cast<const uint4*>(p.out + static_cast<size_t>(row0 + q) * p.inner + h * 64) + part * 4;
        const uint4* d4 = reinterpret_cast<const uint4*>(p.dout + static_cast<size_t>(row0 + q) * p.inner + h * 64) + part * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ov[u] = __ldg(o4 + u);
          dv[u] = __ldg(d4 + u);
        }
        if (part == 0) lse_q = __ldg(p.lse + (static_cast<size_t>(psg) * H + h) * Lmax + q);
      }
      float dsum = 0.f;
      if (q < L) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 a = ov[u], b = dv[u];
          dsum += bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) +
                  bf16_hi(a.y) * bf16_hi(b.y) + bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) +
                  bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
        }
      }
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      if (part == 0) {
        s_D[q] = dsum;
        s_lse[q] = lse_q;
      }
      named_bar_sync(1, EB_SOFTMAX_THREADS);
      // the next passage's O rows (one 128-byte line per query and head) and lse to L2
      if (part == 0 && q < nL) prefetch_l2_line(p.out + static_cast<size_t>(nrow0 + q) * p.inner + h * 64);
      if (tid < 8 && tid * 32 < nL) prefetch_l2_line(p.lse + (static_cast<size_t>(npsg) * H + h) * Lmax + tid * 32);
    }
  }
  EB_TRACE(7);

  // dK / dV of key tile `kt_done` out of tensor memory (its last gradient MMAs have completed)
  auto store_dk_dv = [&](int kt_done) {
    tc_fence_after();
    const int key = kt_done * 128 + r;
    uint32_t a[16], bb[16];
    tmem_ld_32x16(t_row + TM_DK + cq * 16, a);
    tmem_ld_32x16(t_row + TM_DV + cq * 16, bb);
    tmem_ld_wait();
    if (key < L) {
      __nv_bfloat16* dk = p.dqkv + static_cast<size_t>(row0 + key) * 3 * p.inner + p.inner + h * 64 + cq * 16;
      __nv_bfloat16* dv = dk + p.inner;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(a[u * 8 + 0]), __uint_as_float(a[u * 8 + 1]));
        o.y = pack_bf16x2(__uint_as_float(a[u * 8 + 2]), __uint_as_float(a[u * 8 + 3]));
        o.z = pack_bf16x2(__uint_as_float(a[u * 8 + 4]), __uint_as_float(a[u * 8 + 5]));
        o.w = pack_bf16x2(__uint_as_float(a[u * 8 + 6]), __uint_as_float(a[u * 8 + 7]));
        *reinterpret_cast<uint4*>(dk + u * 8) = o;
        o.x = pack_bf16x2(__uint_as_float(bb[u * 8 + 0]), __uint_as_float(bb[u * 8 + 1]));
        o.y = pack_bf16x2(__uint_as_float(bb[u * 8 + 2]), __uint_as_float(bb[u * 8 + 3]));
        o.z = pack_bf16x2(__uint_as_float(bb[u * 8 + 4]), __uint_as_float(bb[u * 8 + 5]));
        o.w = pack_bf16x2(__uint_as_float(bb[u * 8 + 6]), __uint_as_float(bb[u * 8 + 7]));
        *reinterpret_cast<uint4*>(dv + u * 8) = o;
      }
    }
    tc_fence_before();
  };

  for (int kt = 0; kt < nt; ++kt) {
    for (int qt = 0; qt < nt; ++qt) {
      const int q = qt * 128 + r;
      const int qc = min(q, L - 1);
      const int b = Lmax - 1 - qc;
      const uint32_t lut_a = smem_u32(s_lut4) + b * 16;  // this row's bias windows (stage_lut_win)
      const bool rows_live = (qt * 128 + quarter * 32) < L;  // warp-uniform: any of this warp's queries in the passage
      const int kbase = kt * 128 + cq * 32;                  // first key of this warp's 32-column chunk
      // the forward's keep mask of this row's 32 keys (fetched ahead of the MMA wait)
      uint32_t kword = 0xffffffffu;  // drop_keep32_sliced order
      if (kDrop && q < L && kbase < L)
        kword = p.keepbits ? __ldg(p.keepbits + ((static_cast<size_t>(psg) * p.H + h) * Lmax + q) * 8 + (kbase >> 5))
                           : drop_keep32_sliced(p.drop, (static_cast<unsigned long long>(psg * p.H + h) * Lmax + qc) * 32 +
                                                            (kbase >> 3));
      mbar_wait(bar_mma, s_phase);
      s_phase ^= 1;
      tc_fence_after();
      EB_TRACE(8 + (kt * 2 + qt) * 6 + 0);
      const float nlse_q = -lds_f(smem_u32(s_lse) + q * 4) * LOG2E_F;  // lse = +inf beyond the passage -> exp2(-inf) = 0
      const float D_q = lds_f(smem_u32(s_D) + q * 4);
      const bool live = rows_live && kbase < L;
      const bool need_km = (p.mask != nullptr) || (kbase + 32 > L);
      // !live: nothing of this 32x32 sub-block lies inside the passage: P = dS = 0 (dV / dK sum over all 128 rows of
      // the tile, so the zeros must really be written), no exp / Philox / diagonal sums.
      // bias-gradient index of the diagonal d = i - lane of this 32x32 sub-block: rel = k - q
      const int idx0 = (kbase - (qt * 128 + quarter * 32)) + (Lmax - 1);
      // Bias gradient = diagonal sums (k - q constant) of this 32x32 sub-block, as one systolic pass over the warp,
      // exact in fp32: with ascending columns and the partial sum moving one lane UP per column, lane 0 starts a
      // fresh diagonal each step (keep = 0 there), lane 31 holds the COMPLETE diagonal d = i - 31 (<= 0) after
      // column i (parked in lane i's Rdone), and after the last column lane l holds the complete d = 31 - l (>= 0).
      float Rup = 0.f;
      // 64-key chunk (cq >> 1) of the K-major P / dS tiles; this warp's 32 keys are 16-byte units (cq & 1)*4..+4
      const uint32_t prow = smem_u32(smem) + EB_P_OFF + (cq >> 1) * 16384 + r * 128;
      const uint32_t dsrow = smem_u32(smem) + EB_DS_OFF + (cq >> 1) * 16384 + r * 128;
      // bins of this sub-block's diagonals: column i completes d = i - 31 in lane 31 (parked in the warp's scratch row for
      // lane i); after the last column lane l holds d = 31 - l
      const uint32_t dl_a = smem_u32(s_dlut) + (EB_DLUT_PAD + idx0 - 31) * 4;
      const uint32_t scr_a = smem_u32(smem) + EB_SCR_OFF + warp * 128;  // this warp's 32 parked diagonal sums
      const float nD = -D_q;
      const float dscale = kDrop ? p.drop.scale : 1.f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // two 16-column halves keep the live register set small
        uint32_t ppk[8], dpk[8];        // bf16x2-packed P_drop and dS of 16 columns
        if (!live) {
#pragma unroll
          for (int i = 0; i < 8; ++i) ppk[i] = dpk[i] = 0u;
          if (hh == 1) mbar_arrive(bar_sfree);
        } else {
          uint32_t sv[16], dpv[16];
          tmem_ld_32x16(t_row + TM_S + cq * 32 + hh * 16, sv);
          tmem_ld_32x16(t_row + TM_DP + cq * 32 + hh * 16, dpv);
          tmem_ld_wait();
          if (hh == 1) {  // this thread's last read of the block's S / dP
            tc_fence_before();
            mbar_arrive(bar_sfree);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // 4 keys at a time (one bias window); packed fp32 pair arithmetic
            const int col = hh * 16 + g * 4;  // first of the 4 keys within this warp's 32-key chunk
            float s0 = __uint_as_float(sv[g * 4 + 0]), s1 = __uint_as_float(sv[g * 4 + 1]);
            float s2 = __uint_as_float(sv[g * 4 + 2]), s3 = __uint_as_float(sv[g * 4 + 3]);
            if (need_km) {  // warp-uniform: a key mask was given, or the chunk straddles the end of the passage
              const float4 km = lds_f4(smem_u32(s_kmask) + (kbase + col) * 4);
              fadd2(s0, s1, km.x, km.y);
              fadd2(s2, s3, km.z, km.w);
            }
            const float4 lb = lds_f4(lut_a + (kbase + col) * 16);  // biases * log2(e)
            ffma2v(s0, s1, s0, s1, LOG2E_F, LOG2E_F, lb.x, lb.y);
            ffma2v(s2, s3, s2, s3, LOG2E_F, LOG2E_F, lb.z, lb.w);
            fadd2(s0, s1, nlse_q, nlse_q);
            fadd2(s2, s3, nlse_q, nlse_q);
            const float p0 = fast_exp2(s0), p1 = fast_exp2(s1), p2 = fast_exp2(s2), p3 = fast_exp2(s3);  // P = exp(s - lse)
            float d0 = __uint_as_float(dpv[g * 4 + 0]), d1 = __uint_as_float(dpv[g * 4 + 1]);
            float d2 = __uint_as_float(dpv[g * 4 + 2]), d3 = __uint_as_float(dpv[g * 4 + 3]);
            if (kDrop) {  // dP_drop: dropped keys pass no gradient
              d0 = ((kword >> sliced_bit(col + 0)) & 1u) ? d0 : 0.f;
              d1 = ((kword >> sliced_bit(col + 1)) & 1u) ? d1 : 0.f;
              d2 = ((kword >> sliced_bit(col + 2)) & 1u) ? d2 : 0.f;
              d3 = ((kword >> sliced_bit(col + 3)) & 1u) ? d3 : 0.f;
            }
            // dS = P (dP_drop / (1 - p) - D)
            ffma2v(d0, d1, d0, d1, dscale, dscale, nD, nD);
            ffma2v(d2, d3, d2, d3, dscale, dscale, nD, nD);
            float ds[4];
            fmul2v(ds[0], ds[1], p0, p1, d0, d1);
            fmul2v(ds[2], ds[3], p2, p3, d2, d3);
            // the dV = P_drop^T dO operand: P / (1 - p) with the dropped keys zeroed (pair masks on the packed words)
            uint32_t w0, w1;
            if (kDrop) {
              float q0, q1, q2, q3;
              fmul2v(q0, q1, p0, p1, dscale, dscale);
              fmul2v(q2, q3, p2, p3, dscale, dscale);
              w0 = pack_bf16x2(q0, q1) & sliced_pair_mask(kword, (col >> 1) + 0);
              w1 = pack_bf16x2(q2, q3) & sliced_pair_mask(kword, (col >> 1) + 1);
            } else {
              w0 = pack_bf16x2(p0, p1);
              w1 = pack_bf16x2(p2, p3);
            }
            ppk[g * 2 + 0] = w0;
            ppk[g * 2 + 1] = w1;
            dpk[g * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
            dpk[g * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
            // bias gradient: systolic diagonal sums, lane 31 retires one complete diagonal per column
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              Rup = fmaf(__shfl_up_sync(0xffffffffu, Rup, 1), keep, ds[i]);
              if (lane == 31) sts_f(scr_a + (col + i) * 4, Rup);  // parked for lane (col + i)
            }
          }
        }
        // P / dS of the previous block may still be feeding its gradient MMAs (issued together with this block's
        // S / dP, so they are all but finished by now): wait before overwriting them
        if (hh == 0 && g_pending) {
          mbar_wait(bar_grad, g_phase);
          g_phase ^= 1;
          g_pending = false;
        }
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
          const int unit = (((cq & 1) * 4 + hh * 2 + uu) ^ (r & 7)) << 4;
          sts_u4(prow + unit, make_uint4(ppk[uu * 4], ppk[uu * 4 + 1], ppk[uu * 4 + 2], ppk[uu * 4 + 3]));
          sts_u4(dsrow + unit, make_uint4(dpk[uu * 4], dpk[uu * 4 + 1], dpk[uu * 4 + 2], dpk[uu * 4 + 3]));
        }
      }
      if (live) {
        // lane l: the parked diagonal d = l - 31 and its own d = 31 - l (d = 0 once: lane 31's Rup)
        __syncwarp();
        const float Rdone = lds_f(scr_a + lane * 4);
        red_shared_add_f(dl_a + (62 - lane) * 4, Rup);
        if (lane != 31) red_shared_add_f(dl_a + lane * 4, Rdone);
        __syncwarp();  // the scratch row is rewritten in the next block
      }
      EB_TRACE(8 + (kt * 2 + qt) * 6 + 2);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_ready);
      EB_TRACE(8 + (kt * 2 + qt) * 6 + 3);
      EB_TRACE(8 + (kt * 2 + qt) * 6 + 4);
      g_pending = true;
      if (qt == nt - 1) {
        // dK[kt], dV[kt] are final once the gradient MMAs just issued complete.  (Reading them out one block later instead,
        // without this drain, measured SLOWER -- 447 vs 391 us per layer at cfg4 shapes: every block's softmax then runs
        // against the previous block's gradient MMAs, and tcgen05.ld under MMA traffic to the same tensor memory is slow.)
        mbar_wait(bar_kv, kv_phase);
        kv_phase ^= 1;
        store_dk_dv(kt);
        if (kt == nt - 1) {  // last block of the passage: dQ is read out next, everything has to be complete
          mbar_wait(bar_grad, g_phase);
          g_phase ^= 1;
          g_pending = false;
        }
      }
      EB_TRACE(8 + (kt * 2 + qt) * 6 + 5);
    }
  }
  EB_TRACE(4);
  // every MMA has completed (the last iteration waited for the final commit)
  {
    tc_fence_after();
#pragma unroll 1
    for (int qt = 0; qt < nt; ++qt) {
      const int q = qt * 128 + r;
      uint32_t a[16];
      tmem_ld_32x16(t_row + TM_DQ + qt * 64 + cq * 16, a);
      tmem_ld_wait();
      if (q < L) {
        __nv_bfloat16* dq = p.dqkv + static_cast<size_t>(row0 + q) * 3 * p.inner + h * 64 + cq * 16;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
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
  }  // passages
  }  // softmax warps
  __syncthreads();
  EB_TRACE(5);
  if (p.dlut) {
    for (int i = tid; i < nlut; i += EB_THREADS) {
      const float v = s_dlut[EB_DLUT_PAD + i];
      if (v != 0.f) atomicAdd(&p.dlut[h * nlut + i], v);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
  EB_TRACE(6);
}

#ifdef LAKO_TRACE
extern "C" int lako_debug_a2_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, g_a2_trace, sizeof(long long) * 256) == cudaSuccess ? 0 : 1;
}
extern "C" int lako_debug_eb_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, g_eb_trace, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif

// D[row][h] = sum_d dO[row][h*64 + d] * O[row][h*64 + d], written IN PLACE over the first H floats of O's row (the
// attention output is dead once its o-projection weight gradient has been formed): one warp per row, every lane holds
// its 16-byte pieces of the row before any lane writes.
__global__ void __launch_bounds__(256) attn_rowdot_inplace_kernel(__nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                                  int rows, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int inner = H * 64;
  const uint4* o4 = reinterpret_cast<const uint4*>(o + static_cast<size_t>(warp) * inner);
  const uint4* d4 = reinterpret_cast<const uint4*>(dout + static_cast<size_t>(warp) * inner);
  float acc[4];
  const int nit = (H * 8 + 31) / 32;  // 8 lanes of 16 bytes per head; at most 4 passes (H <= 16 per pass of 4 heads)
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    acc[it] = 0.f;
    const int u = it * 32 + lane;
    if (it < nit && u < H * 8) {
      const uint4 a = __ldg(o4 + u), b = __ldg(d4 + u);
      acc[it] = bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) +
                bf16_hi(a.y) * bf16_hi(b.y) + bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) +
                bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
    }
    acc[it] += __shfl_xor_sync(0xffffffffu, acc[it], 1);
    acc[it] += __shfl_xor_sync(0xffffffffu, acc[it], 2);
    acc[it] += __shfl_xor_sync(0xffffffffu, acc[it], 4);
  }
  __syncwarp();  // every lane has read its part of the row
  float* dst = reinterpret_cast<float*>(o + static_cast<size_t>(warp) * inner);
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int hh = it * 4 + (lane >> 3);
    if (it < nit && (lane & 7) == 0 && hh < H) dst[hh] = acc[it];
  }
}
int attn_rowdot_inplace(__nv_bfloat16* o, const __nv_bfloat16* dout, int rows, int H, cudaStream_t s) {
  LAKO_REQUIRE(rows > 0 && H > 0 && H <= 16, LAKO_ERR_UNSUPPORTED);
  ProfScope prof(PROF_ENC_ATTN_BWD, 0.0, s);
  launch_pdl(attn_rowdot_inplace_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, o, dout, rows, H);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

int enc_attn_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                 const unsigned char* mask, const float* bias_lut, __nv_bfloat16* dqkv, float* dlut, int P, int L,
                 int H, cudaStream_t s, DropSpec drop, const int* cu, int packed_rows, const uint32_t* keepbits,
                 bool d_in_out) {
  int rc = check_arch();
  if (rc) return rc;
  LAKO_REQUIRE(P > 0 && H > 0 && L > 0, LAKO_ERR_BAD_ARG);
  if (L > 256 || enc_attn_impl() == 3)
    return enc_attn_long_bwd(qkv, out, dout, lse, mask, bias_lut, dqkv, dlut, P, L, H, s, drop, cu, d_in_out);
  const int inner = H * 64;
  CUtensorMap tm_qkv, tm_do;
  const uint64_t total_rows = cu ? static_cast<uint64_t>(packed_rows) : static_cast<uint64_t>(P) * L;
  LAKO_REQUIRE(total_rows > 0, LAKO_ERR_BAD_ARG);
  rc = get_tensor_map_2d(&tm_qkv, qkv, 3 * inner, total_rows, 3 * inner, 64, 128);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tm_do, dout, inner, total_rows, inner, 64, 128);
  if (rc) return rc;
  using BwdKern = void (*)(CUtensorMap, CUtensorMap, EncAttnBwdParams, int);
  static const BwdKern bwd_kerns[2] = {enc_attn_bwd_kernel<false>, enc_attn_bwd_kernel<true>};
  static bool configured = false;
  if (!configured) {
    for (BwdKern kf : bwd_kerns) {
      cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, EB_SMEM_BYTES);
      if (e != cudaSuccess) {
        set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
        return LAKO_ERR_LAUNCH;
      }
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
  p.cu = cu;
  p.drop = drop;
  p.keepbits = keepbits;
  p.d_in_out = d_in_out ? 1 : 0;
  ProfScope prof(PROF_ENC_ATTN_BWD, 10.0 * P * H * static_cast<double>(L) * L * 64, s);
  if (H > device_sm_count()) {
    set_last_error("enc_attn_bwd: more heads than SMs", __FILE__, __LINE__);
    return LAKO_ERR_UNSUPPORTED;
  }
  // every CTA serves one head: (SMs / H) CTAs per head (at most one per passage)
  const int per_head = std::max(1, std::min(P, device_sm_count() / H));
  launch_pdl(bwd_kerns[drop.thresh ? 1 : 0], dim3(per_head * H), dim3(EB_THREADS), EB_SMEM_BYTES, s, tm_qkv, tm_do, p, P);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// dtable[bucket][h] += sum over rel of dlut[h][rel] (fold the per-offset bias gradient into the 32 buckets)
__global__ void fold_dlut_kernel(const float* __restrict__ dlut, const int* __restrict__ buckets,
                                 float* __restrict__ dtable, int H, int L, int bucket_center, int nb) {
  pdl_launch_dependents();
  pdl_wait();
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
  launch_pdl(fold_dlut_kernel, dim3(H), dim3(32), 0, s, dlut, buckets, dtable, H, L, bucket_center, nb);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
