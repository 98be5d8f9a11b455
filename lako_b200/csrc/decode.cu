// KV-cached single-token decode kernels for greedy generate (north-star kernel 5) and the few-row Linears of the
// decoder: everything here is HBM/L2-streaming work (M = B <= 64 rows against hundreds of MB of decoder weights and
// the cross-attention K/V of every memory token), so the kernels are shaped for bytes in flight, not for the 128-row
// tcgen05 tiles (which would be > 90 % padding at M = B).
//
// Reference: FiDT5.generate /root/reference/src/model.py:54-60 -> HF greedy search with use_cache (one decoder pass per
// generated token over T5Block HF:modeling_t5.py:411-498); cross_attention_forward src/model.py:286-349 with the
// step-0 score capture (:327-329).
//
//  * smallm_linear_kernel : y[M][N] = epi(alpha * norm?(x)[M][K] W[N][K]^T), M <= 64.  Warp-level mma.sync m16n8k16 with
//    the WEIGHT rows as the 16-row A operand loaded straight from global memory (two 16-byte loads per lane per 32 k;
//    the k index inside a 32-block is permuted identically for A and B, so no shuffles or shared-memory staging of the
//    weights are needed) and the M activation rows as the 8-column B operand from shared memory.  The 8 warps of a CTA
//    split K and reduce through shared memory.  Fused: T5 RMSNorm of the fp32 residual stream as the prologue
//    (HF:55-68), ReLU / residual-add / fp32-logit epilogues.
//  * xattn_decode_kernel  : one query token against the packed memory of its question, 8 lanes per key row, online
//    softmax per lane group, split over 256-key chunks (flash-decoding) + the existing combine kernel; accumulates the
//    step-0 pre-softmax scores (summed over heads) into tok_score on the fly.
//  * dself_step / greedy_pick variants that read the step index from device memory, so that ONE captured CUDA graph
//    serves every decode step.
#include "ops.cuh"

namespace lako {

namespace {

constexpr float DC_MASK_MIN = -3.4028234663852886e38f;
constexpr float DC_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct SmallLinParams {
  const float* x_f32;          // [M][K] fp32 residual stream -> RMSNorm prologue (ln_w != null), else plain cast
  const __nv_bfloat16* x_b16;  // [M][ldx] bf16 input (used when x_f32 == null)
  const float* ln_w;           // [K]
  const __nv_bfloat16* W;      // [N][K]
  void* y;                     // see epi
  int M, N, K, ldx, ldy;
  float eps, alpha;
  int epi;  // 0 bf16 store, 1 bf16 ReLU store, 2 fp32 store, 3 fp32 accumulate (y += : residual add)
  int kc;   // K chunk staged per pass (multiple of 32)
};

constexpr int SL_WARPS = 8;
constexpr int SL_ROWS = 16;   // weight rows (output features) per CTA
constexpr int SL_UNROLL = 4;  // 32-wide k blocks in flight per warp

// dynamic smem: xs bf16 [NT*8][kc + 32]  (row pitch = 2*kc + 64 bytes: the 8 lanes of a quarter-warp LDS.128 hit all
// 32 banks once) | red fp32 [SL_WARPS][16][NT*8] | rstd fp32 [NT*8]
template <int NT>
__global__ void __launch_bounds__(SL_WARPS * 32) smallm_linear_kernel(const SmallLinParams p) {
  pdl_launch_dependents();
  extern __shared__ __align__(16) uint8_t sl_smem[];
  constexpr int MP = NT * 8;
  const int pitch = p.kc * 2 + 64;  // bytes
  uint8_t* xs = sl_smem;
  float* red = reinterpret_cast<float*>(sl_smem + static_cast<size_t>(MP) * pitch);
  float* s_rstd = red + SL_WARPS * 16 * MP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int n0 = blockIdx.x * SL_ROWS;
  // the two weight rows of this lane (clamped: rows past N are computed on a valid row and never stored)
  const int r_lo = min(n0 + g, p.N - 1), r_hi = min(n0 + g + 8, p.N - 1);
  const __nv_bfloat16* w_lo = p.W + static_cast<size_t>(r_lo) * p.K + t4 * 8;
  const __nv_bfloat16* w_hi = p.W + static_cast<size_t>(r_hi) * p.K + t4 * 8;
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;

  // weights do not depend on the predecessor kernel: start streaming the first blocks before the grid dependency
  const int kblocks = p.K >> 5;
  uint4 pa_lo[SL_UNROLL], pa_hi[SL_UNROLL];
  int kb_next = warp;  // this warp owns k blocks warp, warp + 8, ...
#pragma unroll
  for (int u = 0; u < SL_UNROLL; ++u) {
    const int kb = kb_next + u * SL_WARPS;
    if (kb < kblocks) {
      pa_lo[u] = ldg_nc_v4(w_lo + kb * 32);
      pa_hi[u] = ldg_nc_v4(w_hi + kb * 32);
    }
  }
  pdl_wait();

  if (p.x_f32 && p.ln_w) {  // T5 RMSNorm statistics: one warp per row, same summation order as rmsnorm_fwd_kernel
    for (int m = warp; m < MP; m += SL_WARPS) {
      float ss = 0.f;
      if (m < p.M) {
        const float* xr = p.x_f32 + static_cast<size_t>(m) * p.K;
        for (int c = lane * 4; c < p.K; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(xr + c);
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
      }
      ss = warp_sum(ss);
      if (lane == 0) s_rstd[m] = rsqrtf(ss / static_cast<float>(p.K) + p.eps);
    }
    __syncthreads();
  }

  for (int kc0 = 0; kc0 < p.K; kc0 += p.kc) {
    const int kcn = min(p.kc, p.K - kc0);
    if (kc0 > 0) __syncthreads();  // every warp is done with the previous chunk
    // ---- stage x[:, kc0 : kc0 + kcn] as bf16 (normalised when the prologue is on); rows >= M are zero
    for (int i = tid; i < MP * (kcn >> 2); i += SL_WARPS * 32) {
      const int m = i / (kcn >> 2), c = (i - m * (kcn >> 2)) * 4;
      uint2 o = make_uint2(0u, 0u);
      if (m < p.M) {
        if (p.x_f32) {
          float4 v = *reinterpret_cast<const float4*>(p.x_f32 + static_cast<size_t>(m) * p.K + kc0 + c);
          if (p.ln_w) {
            const float rs = s_rstd[m];
            const float4 wv = *reinterpret_cast<const float4*>(p.ln_w + kc0 + c);
            v = make_float4(wv.x * (v.x * rs), wv.y * (v.y * rs), wv.z * (v.z * rs), wv.w * (v.w * rs));
          }
          o.x = pack_bf16x2(v.x, v.y);
          o.y = pack_bf16x2(v.z, v.w);
        } else {
          o = *reinterpret_cast<const uint2*>(p.x_b16 + static_cast<size_t>(m) * p.ldx + kc0 + c);
        }
      }
      *reinterpret_cast<uint2*>(xs + static_cast<size_t>(m) * pitch + c * 2) = o;
    }
    __syncthreads();
    // ---- this warp's k blocks inside the chunk, SL_UNROLL at a time.  Chunks are whole multiples of
    //      32 * SL_WARPS * SL_UNROLL k (or the single chunk is all of K), so a group never straddles a chunk boundary
    //      and the prefetch queue simply runs ahead into the next chunk.
    const int kb_end = (kc0 + kcn) >> 5;
    while (kb_next < kb_end) {
      uint4 a_lo[SL_UNROLL], a_hi[SL_UNROLL];
#pragma unroll
      for (int u = 0; u < SL_UNROLL; ++u) {
        a_lo[u] = pa_lo[u];
        a_hi[u] = pa_hi[u];
      }
#pragma unroll
      for (int u = 0; u < SL_UNROLL; ++u) {
        const int kb = kb_next + (SL_UNROLL + u) * SL_WARPS;
        if (kb < kblocks) {
          pa_lo[u] = ldg_nc_v4(w_lo + kb * 32);
          pa_hi[u] = ldg_nc_v4(w_hi + kb * 32);
        }
      }
#pragma unroll
      for (int u = 0; u < SL_UNROLL; ++u) {
        const int kb = kb_next + u * SL_WARPS;
        if (kb < kblocks) {
          const uint8_t* xb = xs + ((kb << 5) - kc0 + t4 * 8) * 2;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint4 xv = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(nt * 8 + g) * pitch);
            mma_16816(acc[nt], a_lo[u].x, a_hi[u].x, a_lo[u].y, a_hi[u].y, xv.x, xv.y);
            mma_16816(acc[nt], a_lo[u].z, a_hi[u].z, a_lo[u].w, a_hi[u].w, xv.z, xv.w);
          }
        }
      }
      kb_next += SL_UNROLL * SL_WARPS;
    }
  }
  // ---- reduce the 8 warps' partial sums: red[warp][row 0..15][m 0..MP)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float* r0 = red + (warp * 16 + g) * MP + nt * 8 + t4 * 2;
    float* r1 = red + (warp * 16 + g + 8) * MP + nt * 8 + t4 * 2;
    *reinterpret_cast<float2*>(r0) = make_float2(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<float2*>(r1) = make_float2(acc[nt][2], acc[nt][3]);
  }
  __syncthreads();
  for (int i = tid; i < MP * 16; i += SL_WARPS * 32) {
    const int m = i >> 4, r = i & 15;  // 16 consecutive output features of one activation row are contiguous in y
    const int n = n0 + r;
    if (m >= p.M || n >= p.N) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < SL_WARPS; ++w) v += red[(w * 16 + r) * MP + m];
    v *= p.alpha;
    const size_t o = static_cast<size_t>(m) * p.ldy + n;
    if (p.epi == 0) {
      reinterpret_cast<__nv_bfloat16*>(p.y)[o] = __float2bfloat16(v);
    } else if (p.epi == 1) {
      reinterpret_cast<__nv_bfloat16*>(p.y)[o] = __float2bfloat16(fmaxf(v, 0.f));
    } else if (p.epi == 2) {
      reinterpret_cast<float*>(p.y)[o] = v;
    } else {
      reinterpret_cast<float*>(p.y)[o] += v;
    }
  }
}

template <int NT>
int launch_smallm(const SmallLinParams& p, cudaStream_t s) {
  constexpr int MP = NT * 8;
  const size_t smem = static_cast<size_t>(MP) * (p.kc * 2 + 64) + sizeof(float) * (SL_WARPS * 16 * MP + MP);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(smallm_linear_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return LAKO_ERR_LAUNCH;
    }
    configured = smem;
  }
  launch_pdl(smallm_linear_kernel<NT>, dim3((p.N + SL_ROWS - 1) / SL_ROWS), dim3(SL_WARPS * 32), smem, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace

// y = epi(alpha * norm?(x) W^T) for M <= 64 activation rows.  x_f32 != null: the input is the fp32 residual stream
// [M][K]; with ln_w it is RMS-normalised (eps) on the fly.  Otherwise x_b16 [M][ldx] bf16.
int smallm_linear(const float* x_f32, const __nv_bfloat16* x_b16, int ldx, const float* ln_w, float eps,
                  const __nv_bfloat16* W, void* y, int ldy, int M, int N, int K, int epi, float alpha, cudaStream_t s) {
  LAKO_REQUIRE(M > 0 && N > 0 && K > 0 && K % 32 == 0, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE((x_f32 != nullptr) != (x_b16 != nullptr) && W && y && epi >= 0 && epi <= 3, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(x_b16 == nullptr || ldx % 4 == 0, LAKO_ERR_BAD_ARG);
  if (M > 64) {  // more rows than one CTA stages: row blocks of 64 (each re-streams the weights out of L2)
    const size_t ysz = (epi == SL_BF16 || epi == SL_BF16_RELU) ? 2 : 4;
    for (int m0 = 0; m0 < M; m0 += 64) {
      const int rc = smallm_linear(x_f32 ? x_f32 + static_cast<size_t>(m0) * K : nullptr,
                                   x_b16 ? x_b16 + static_cast<size_t>(m0) * ldx : nullptr, ldx, ln_w, eps, W,
                                   static_cast<uint8_t*>(y) + static_cast<size_t>(m0) * ldy * ysz, ldy, min(64, M - m0), N, K,
                                   epi, alpha, s);
      if (rc) return rc;
    }
    return LAKO_OK;
  }
  SmallLinParams p{x_f32, x_b16, ln_w, W, y, M, N, K, ldx, ldy, eps, alpha, epi, 0};
  const int nt = (M + 7) / 8;
  const int NT = nt <= 1 ? 1 : nt <= 2 ? 2 : nt <= 4 ? 4 : 8;
  // x chunk staged per pass: ~128 KB of bf16 activations at most, a whole multiple of 32 * SL_WARPS * SL_UNROLL k
  constexpr int KGROUP = 32 * SL_WARPS * SL_UNROLL;
  const int kc = (131072 / (NT * 8 * 2)) / KGROUP * KGROUP;
  p.kc = K <= kc ? K : kc;
  LAKO_REQUIRE(K <= kc || K % KGROUP == 0, LAKO_ERR_UNSUPPORTED);
  ProfScope prof(PROF_GEMM_SMALL, 2.0 * M * N * static_cast<double>(K), s);
  switch (NT) {
    case 1: return launch_smallm<1>(p, s);
    case 2: return launch_smallm<2>(p, s);
    case 4: return launch_smallm<4>(p, s);
    default: return launch_smallm<8>(p, s);
  }
}

// ------------------------------------------------------------------------------------------------
// cross-attention of ONE query token per question over the fused memory (greedy decode step)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int XD_CH = 256;  // keys per CTA (4 warps x 64)

struct XDecParams {
  const __nv_bfloat16* q;    // [B][ldq]
  const __nv_bfloat16* kv;   // rows = memory tokens, row pitch ldkv
  const unsigned char* mask; // padded layout only: [B][M]
  float* part;               // [B][H][splits][66]
  float* tok_score;          // [B][M] fp32 accumulated when *step == 0 (may be null)
  const int* step;           // device step index (null: always capture when tok_score != null)
  const int* cu;
  const int* tok_index;
  int M, H, ldq, ldkv, koff, voff, splits, npsg;
};

__global__ void __launch_bounds__(128) xattn_decode_kernel(const XDecParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_m[4], s_l[4], s_o[4][64];
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, kg = lane >> 3, c = lane & 7;
  int kstart = b * p.M, Mb = p.M;
  if (p.cu) {
    kstart = p.cu[b * p.npsg];
    Mb = p.cu[(b + 1) * p.npsg] - kstart;
  }
  const int k0 = split * XD_CH;
  if (k0 >= Mb) return;
  const bool capture = p.tok_score != nullptr && (p.step == nullptr || *p.step == 0);
  // this lane's 8 of the 64 query dims
  float qv[8];
  {
    const uint4 qq = *reinterpret_cast<const uint4*>(p.q + static_cast<size_t>(b) * p.ldq + h * 64 + c * 8);
    qv[0] = bf16_lo(qq.x); qv[1] = bf16_hi(qq.x); qv[2] = bf16_lo(qq.y); qv[3] = bf16_hi(qq.y);
    qv[4] = bf16_lo(qq.z); qv[5] = bf16_hi(qq.z); qv[6] = bf16_lo(qq.w); qv[7] = bf16_hi(qq.w);
  }
  float m_run = -INFINITY, l_run = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  const int wk0 = k0 + warp * 64;  // this warp's 64 keys; lane group kg takes keys wk0 + j*4 + kg
  const __nv_bfloat16* kbase = p.kv + (static_cast<size_t>(kstart) + wk0 + kg) * p.ldkv + h * 64 + c * 8;
  const size_t step4 = static_cast<size_t>(4) * p.ldkv;
#pragma unroll 1
  for (int j0 = 0; j0 < 16; j0 += 8) {
    if (wk0 + j0 * 4 >= Mb) break;  // warp-uniform
    uint4 kr[8], vr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = wk0 + (j0 + j) * 4 + kg;
      if (key < Mb) {
        kr[j] = ldg_nc_v4(kbase + (j0 + j) * step4 + p.koff);
        vr[j] = ldg_nc_v4(kbase + (j0 + j) * step4 + p.voff);
      } else {
        kr[j] = make_uint4(0, 0, 0, 0);
        vr[j] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = wk0 + (j0 + j) * 4 + kg;
      float s = qv[0] * bf16_lo(kr[j].x) + qv[1] * bf16_hi(kr[j].x) + qv[2] * bf16_lo(kr[j].y) + qv[3] * bf16_hi(kr[j].y) +
                qv[4] * bf16_lo(kr[j].z) + qv[5] * bf16_hi(kr[j].z) + qv[6] * bf16_lo(kr[j].w) + qv[7] * bf16_hi(kr[j].w);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (key >= Mb) continue;  // uniform within the 8-lane group; the shuffles above ran on all lanes
      bool valid = true;
      if (p.mask) valid = p.mask[static_cast<size_t>(b) * p.M + key] != 0;
      if (capture && valid && c == 0) {
        const size_t ti = p.tok_index ? static_cast<size_t>(p.tok_index[kstart + key]) : static_cast<size_t>(b) * p.M + key;
        atomicAdd(p.tok_score + ti, s);
      }
      if (!valid) s += DC_MASK_MIN;  // HF's additive mask (finfo.min): a fully masked question degrades to uniform
      const float m_new = fmaxf(m_run, s);
      const float sc = (m_run == -INFINITY) ? 0.f : fast_exp2((m_run - m_new) * DC_LOG2E);
      const float pj = fast_exp2((s - m_new) * DC_LOG2E);
      l_run = l_run * sc + pj;
      const float vv[8] = {bf16_lo(vr[j].x), bf16_hi(vr[j].x), bf16_lo(vr[j].y), bf16_hi(vr[j].y),
                           bf16_lo(vr[j].z), bf16_hi(vr[j].z), bf16_lo(vr[j].w), bf16_hi(vr[j].w)};
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = o[i] * sc + pj * vv[i];
      m_run = m_new;
    }
  }
  // ---- merge the 4 lane groups of the warp (same dims c, different keys)
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float m_o = __shfl_xor_sync(0xffffffffu, m_run, off);
    const float l_o = __shfl_xor_sync(0xffffffffu, l_run, off);
    const float m_new = fmaxf(m_run, m_o);
    const float sa = (m_run == -INFINITY) ? 0.f : fast_exp2((m_run - m_new) * DC_LOG2E);
    const float sb = (m_o == -INFINITY) ? 0.f : fast_exp2((m_o - m_new) * DC_LOG2E);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float o_o = __shfl_xor_sync(0xffffffffu, o[i], off);
      o[i] = o[i] * sa + o_o * sb;
    }
    l_run = l_run * sa + l_o * sb;
    m_run = m_new;
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_o[warp][c * 8 + i] = o[i];
    if (c == 0) {
      s_m[warp] = m_run;
      s_l[warp] = l_run;
    }
  }
  __syncthreads();
  // ---- merge the 4 warps, write the chunk partial (o[64], m, l)
  if (tid < 64) {
    const float m = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
    float ov = 0.f, l = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float sc = (s_m[w] == -INFINITY) ? 0.f : fast_exp2((s_m[w] - m) * DC_LOG2E);
      ov += s_o[w][tid] * sc;
      l += s_l[w] * sc;
    }
    float* part = p.part + ((static_cast<size_t>(b) * p.H + h) * p.splits + split) * 66;
    part[tid] = ov;
    if (tid == 0) {
      part[64] = m;
      part[65] = l;
    }
  }
}

// out[b][h*64+d] = sum_s w_s o_s / sum_s w_s l_s over the chunk partials of (b, h); 64 threads
__global__ void __launch_bounds__(64) xattn_decode_combine_kernel(const float* __restrict__ part,
                                                                  __nv_bfloat16* __restrict__ out, int H, int splits,
                                                                  int ldo, const int* __restrict__ cu, int npsg) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  int ns = splits;
  if (cu) ns = (cu[(b + 1) * npsg] - cu[b * npsg] + XD_CH - 1) / XD_CH;
  const float* pp = part + (static_cast<size_t>(b) * H + h) * splits * 66;
  float m = -INFINITY;
  for (int s = 0; s < ns; ++s) m = fmaxf(m, pp[s * 66 + 64]);
  float l = 0.f, o = 0.f;
  for (int s = 0; s < ns; ++s) {
    const float w = fast_exp2((pp[s * 66 + 64] - m) * DC_LOG2E);
    l += w * pp[s * 66 + 65];
    o += w * pp[s * 66 + d];
  }
  out[static_cast<size_t>(b) * ldo + h * 64 + d] = __float2bfloat16(o / l);
}

}  // namespace

size_t xattn_decode_workspace_bytes(int B, int H, int M) {
  return static_cast<size_t>(B) * H * ((M + XD_CH - 1) / XD_CH) * 66 * sizeof(float);
}

int xattn_decode(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, __nv_bfloat16* out,
                 float* tok_score, const int* step, float* workspace, int B, int M, int H, int ldq, int ldkv, int koff,
                 int voff, int ldo, cudaStream_t s, const int* cu, int npsg, const int* tok_index) {
  LAKO_REQUIRE(B > 0 && M > 0 && H > 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && koff % 8 == 0 && voff % 8 == 0, LAKO_ERR_BAD_ARG);
  XDecParams p{q, kv, mask, workspace, tok_score, step, cu, tok_index, M, H, ldq, ldkv, koff, voff, (M + XD_CH - 1) / XD_CH,
               npsg};
  // algorithmic bytes: K and V rows of every memory token of the batch for this layer
  ProfScope prof(PROF_XATTN, 4.0 * B * H * static_cast<double>(M) * 64, s);
  launch_pdl(xattn_decode_kernel, dim3(p.splits, H, B), dim3(128), 0, s, p);
  LAKO_CHECK_LAUNCH();
  launch_pdl(xattn_decode_combine_kernel, dim3(H, B), dim3(64), 0, s, workspace, out, H, p.splits, ldo, cu, npsg);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
