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
#include <algorithm>

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

__device__ __forceinline__ uint32_t dc_swz(int row, int chunk) {  // byte offset in a [rows][128 B] 128B-swizzled tile
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void dc_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void dc_ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

struct SmallLinParams {
  const float* x_f32;          // [M][K] fp32 residual stream -> RMSNorm prologue (ln_w != null), else plain cast
  const __nv_bfloat16* x_b16;  // [M][ldx] bf16 input (used when x_f32 == null)
  const float* ln_w;           // [K]
  const __nv_bfloat16* W;      // [N][K]
  void* y;                     // see epi
  int M, N, K, ldx, ldy;
  float eps, alpha;
  int epi;     // 0 bf16 store, 1 bf16 ReLU store, 2 fp32 store, 3 fp32 accumulate (y += : residual add; atomic when K is split)
  int kc;      // K chunk staged per pass (multiple of 32)
  int kb_per;  // 32-wide k blocks per K split (gridDim.y splits; epi 3 only when > 1 split)
  int ntiles;  // 16-row output tiles; a CTA loops over tiles blockIdx.x, blockIdx.x + gridDim.x, ...
  const uint8_t* pf;  // optional: bytes the NEXT launches will stream (their weights), prefetched to L2 from here
  unsigned pf_lines;  // in 128-byte lines
};

constexpr int SL_WARPS = 8;
constexpr int SL_ROWS = 16;   // weight rows (output features) per tile
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
  if (p.pf) {  // weights of the following launches -> L2 (they do not depend on any predecessor)
    const unsigned nthr = gridDim.x * gridDim.y * blockDim.x;
    for (unsigned i = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + tid; i < p.pf_lines; i += nthr)
      prefetch_l2_line(p.pf + static_cast<size_t>(i) * 128);
  }
  // this CTA's K range (in 32-wide blocks) and whether the staged activations can be kept across output tiles
  const int kblocks_all = p.K >> 5;
  const int kb_lo = blockIdx.y * p.kb_per, kb_hi = min(kblocks_all, kb_lo + p.kb_per);
  const int k_lo = kb_lo << 5, k_hi = kb_hi << 5;
  const bool single_chunk = (k_hi - k_lo) <= p.kc;

  auto row_ptrs = [&](int tile, const __nv_bfloat16*& w_lo, const __nv_bfloat16*& w_hi) {
    const int n0 = tile * SL_ROWS;
    // the two weight rows of this lane (clamped: rows past N are computed on a valid row and never stored)
    const int r_lo = min(n0 + g, p.N - 1), r_hi = min(n0 + g + 8, p.N - 1);
    w_lo = p.W + static_cast<size_t>(r_lo) * p.K + t4 * 8;
    w_hi = p.W + static_cast<size_t>(r_hi) * p.K + t4 * 8;
  };
  // weights do not depend on the predecessor kernel: start streaming the first tile's blocks before the grid dependency
  uint4 pa_lo[SL_UNROLL], pa_hi[SL_UNROLL];
  const __nv_bfloat16 *w_lo, *w_hi;
  int tile = blockIdx.x;
  row_ptrs(tile, w_lo, w_hi);
#pragma unroll
  for (int u = 0; u < SL_UNROLL; ++u) {
    const int kb = kb_lo + warp + u * SL_WARPS;
    if (kb < kb_hi) {
      pa_lo[u] = ldg_nc_v4(w_lo + kb * 32);
      pa_hi[u] = ldg_nc_v4(w_hi + kb * 32);
    }
  }
  pdl_wait();

  if (p.x_f32 && p.ln_w) {  // T5 RMSNorm statistics over the WHOLE row: one warp per row, order of rmsnorm_fwd_kernel
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

  bool staged = false;
  for (; tile < p.ntiles; tile += gridDim.x) {
    const int n0 = tile * SL_ROWS;
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
    int kb_next = kb_lo + warp;  // this warp owns k blocks kb_lo + warp, + 8, ...
    for (int kc0 = k_lo; kc0 < k_hi; kc0 += p.kc) {
      const int kcn = min(p.kc, k_hi - kc0);
      if (!(single_chunk && staged)) {
        if (kc0 > k_lo || staged) __syncthreads();  // every warp is done with what the buffer held
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
        staged = true;
      }
      // ---- this warp's k blocks inside the chunk, SL_UNROLL at a time.  Chunks are whole multiples of
      //      32 * SL_WARPS * SL_UNROLL k (or one chunk covers the CTA's whole K range), so a group never straddles a
      //      chunk boundary and the prefetch queue simply runs ahead into the next chunk.
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
          if (kb < kb_hi) {
            pa_lo[u] = ldg_nc_v4(w_lo + kb * 32);
            pa_hi[u] = ldg_nc_v4(w_hi + kb * 32);
          }
        }
#pragma unroll
        for (int u = 0; u < SL_UNROLL; ++u) {
          const int kb = kb_next + u * SL_WARPS;
          if (kb < kb_hi) {
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
    // the next tile's first weight blocks: in flight while this tile is reduced and stored
    const int next_tile = tile + gridDim.x;
    if (next_tile < p.ntiles) {
      row_ptrs(next_tile, w_lo, w_hi);
#pragma unroll
      for (int u = 0; u < SL_UNROLL; ++u) {
        const int kb = kb_lo + warp + u * SL_WARPS;
        if (kb < kb_hi) {
          pa_lo[u] = ldg_nc_v4(w_lo + kb * 32);
          pa_hi[u] = ldg_nc_v4(w_hi + kb * 32);
        }
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
      } else if (gridDim.y == 1) {
        reinterpret_cast<float*>(p.y)[o] += v;
      } else {
        atomicAdd(reinterpret_cast<float*>(p.y) + o, v);
      }
    }
    if (next_tile < p.ntiles) __syncthreads();  // red is reused by the next tile
  }
}

template <int NT>
int launch_smallm(const SmallLinParams& p, int grid_x, int ksplit, cudaStream_t s) {
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
  launch_pdl(smallm_linear_kernel<NT>, dim3(grid_x, ksplit), dim3(SL_WARPS * 32), smem, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace

// y = epi(alpha * norm?(x) W^T) for few activation rows.  x_f32 != null: the input is the fp32 residual stream [M][K];
// with ln_w it is RMS-normalised (eps) on the fly.  Otherwise x_b16 [M][ldx] bf16.  With the accumulate epilogue
// (SL_F32_ADD) and few output tiles, K is split over gridDim.y CTAs that add their partial sums atomically.
int smallm_linear(const float* x_f32, const __nv_bfloat16* x_b16, int ldx, const float* ln_w, float eps,
                  const __nv_bfloat16* W, void* y, int ldy, int M, int N, int K, int epi, float alpha, cudaStream_t s,
                  const void* prefetch, size_t prefetch_bytes) {
  LAKO_REQUIRE(M > 0 && N > 0 && K > 0 && K % 32 == 0, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE((x_f32 != nullptr) != (x_b16 != nullptr) && W && y && epi >= 0 && epi <= 3, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(x_b16 == nullptr || ldx % 4 == 0, LAKO_ERR_BAD_ARG);
  if (M > 64) {  // more rows than one CTA stages: row blocks of 64 (each re-streams the weights out of L2)
    const size_t ysz = (epi == SL_BF16 || epi == SL_BF16_RELU) ? 2 : 4;
    for (int m0 = 0; m0 < M; m0 += 64) {
      const int rc = smallm_linear(x_f32 ? x_f32 + static_cast<size_t>(m0) * K : nullptr,
                                   x_b16 ? x_b16 + static_cast<size_t>(m0) * ldx : nullptr, ldx, ln_w, eps, W,
                                   static_cast<uint8_t*>(y) + static_cast<size_t>(m0) * ldy * ysz, ldy, min(64, M - m0), N, K,
                                   epi, alpha, s, m0 == 0 ? prefetch : nullptr, prefetch_bytes);
      if (rc) return rc;
    }
    return LAKO_OK;
  }
  SmallLinParams p{x_f32, x_b16, ln_w, W, y, M, N, K, ldx, ldy, eps, alpha, epi, 0, 0, 0,
                   static_cast<const uint8_t*>(prefetch), prefetch ? static_cast<unsigned>(prefetch_bytes / 128) : 0u};
  const int nt = (M + 7) / 8;
  const int NT = nt <= 1 ? 1 : nt <= 2 ? 2 : nt <= 4 ? 4 : 8;
  constexpr int KGROUP = 32 * SL_WARPS * SL_UNROLL;
  p.ntiles = (N + SL_ROWS - 1) / SL_ROWS;
  // K split: only the accumulate epilogue can combine partial sums; used when the output alone gives too few CTAs to
  // stream the weights (N = d_model: 48-64 tiles for 148 SMs).  Every split keeps >= 2 k blocks per warp.
  const int kblocks = K >> 5;
  int ksplit = 1;
  if (epi == SL_F32_ADD && p.ntiles < 2 * device_sm_count()) {
    const int want = (2 * device_sm_count() + p.ntiles - 1) / p.ntiles;
    for (int c = 1; c <= want; ++c)
      if (kblocks % (c * 2 * SL_WARPS) == 0) ksplit = c;
  }
  p.kb_per = kblocks / ksplit;
  // x chunk staged per pass: ~128 KB of bf16 activations at most, a whole multiple of the prefetch group
  const int krange = p.kb_per * 32;
  const int kc = (131072 / (NT * 8 * 2)) / KGROUP * KGROUP;
  p.kc = krange <= kc ? krange : kc;
  LAKO_REQUIRE(krange <= kc || (ksplit == 1 && krange % KGROUP == 0), LAKO_ERR_UNSUPPORTED);
  // persistent over output tiles when the staged activations can be kept (single chunk): ~2 CTAs per SM
  int grid_x = p.ntiles;
  if (krange <= kc) grid_x = min(p.ntiles, 2 * device_sm_count());
  ProfScope prof(PROF_GEMM_SMALL, 2.0 * M * N * static_cast<double>(K), s);
  switch (NT) {
    case 1: return launch_smallm<1>(p, grid_x, ksplit, s);
    case 2: return launch_smallm<2>(p, grid_x, ksplit, s);
    case 4: return launch_smallm<4>(p, grid_x, ksplit, s);
    default: return launch_smallm<8>(p, grid_x, ksplit, s);
  }
}

// ------------------------------------------------------------------------------------------------
// cross-attention of ONE query token per question over the fused memory (greedy decode step)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int XD_CH = 128;            // keys per work item (4 warps x 32)
constexpr int XD_TILE = XD_CH * 128;  // bytes of one 128-key x 64-dim bf16 tile
#ifndef LAKO_XD_STAGES
#define LAKO_XD_STAGES 3
#define LAKO_XD_CTAS 2
#endif
constexpr int XD_STAGES = LAKO_XD_STAGES;  // K+V tile pairs in flight per CTA
constexpr int XD_CTAS = LAKO_XD_CTAS;      // resident CTAs per SM (XD_CTAS x XD_STAGES x 32 KB of tile ring)
constexpr int XD_SMEM = XD_STAGES * 2 * XD_TILE + 1024 + 64;
constexpr int XD_MAXB = 1024;         // questions per launch (prefix table in shared memory)

struct XDecParams {
  const float* q;            // [B][ldq] fp32 (accumulated by the K-split projection; zeroed again here after use)
  float* q_zero;             // same buffer (non-const) or null: zero q[b][h*64 .. +64) once (b, h) is finished
  const unsigned char* mask; // padded layout only: [B][M]
  float* part;               // [B][H][splits][66]
  int* counters;             // [B][H] arrival counters of the chunk items (zero between launches)
  __nv_bfloat16* out;        // [B][ldo]
  float* tok_score;          // [B][M] fp32 accumulated when *step == 0 (may be null)
  const int* step;           // device step index (null: always capture when tok_score != null)
  const int* cu;
  const int* tok_index;
  int B, M, H, ldq, ldo, koff, voff, splits, npsg;
  long long hm_rows;  // > 0: K/V live head-major, [plane = column / 64][hm_rows tokens][64] (contiguous 16 KB per tile)
};

// Persistent CTAs (2 per SM) each take a CONTIGUOUS range of work items (question, head, 128-key chunk) and stream
// their K / V tiles through a 3-stage TMA ring: while an item is being reduced, the tiles of the next two are in
// flight, so every SM keeps loads outstanding for the whole launch (a one-shot CTA per chunk left the memory system
// idle during each CTA's compute phase and paid a query load + an atomic ticket per 32 KB: 3.1 TB/s).  Consecutive
// items of a CTA are consecutive chunks of the same (question, head): the online-softmax state (m, l, o) stays in
// registers across them and only one partial per segment is published.  The last segment of a (question, head) to
// arrive merges the partials (flash-decoding combine) and writes the bf16 context row -- no separate combine launch.
// The first tiles are requested before the grid dependency (the K/V cache does not depend on the preceding kernel).
// Lane group kg (8 lanes) of warp w owns keys w*32 + j*4 + kg of an item, each lane 8 of the 64 dims.
__global__ void __launch_bounds__(128) xattn_decode_kernel(const __grid_constant__ CUtensorMap tm_kv, const XDecParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t xd_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xd_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + XD_STAGES * 2 * XD_TILE);
  __shared__ float s_m[4], s_l[4], s_o[4][64];
  __shared__ int s_last;
  __shared__ int s_pref[XD_MAXB + 1];  // s_pref[b] = number of work items of questions < b
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- item table: question b has H * ceil(len_b / 128) items (cu is older than the immediate predecessor: safe here)
  for (int b = tid; b < p.B; b += 128) {
    int Mb = p.M;
    if (p.cu) Mb = p.cu[(b + 1) * p.npsg] - p.cu[b * p.npsg];
    s_pref[b + 1] = p.H * ((Mb + XD_CH - 1) / XD_CH);
  }
  if (tid == 0) {
    s_pref[0] = 0;
    tma_prefetch_desc(&tm_kv);
    for (int i = 0; i < XD_STAGES; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int b = 0; b < p.B; ++b) s_pref[b + 1] += s_pref[b];
  __syncthreads();
  const int total = s_pref[p.B];
  const int per = (total + gridDim.x - 1) / gridDim.x;  // items per CTA
  const int g_lo = blockIdx.x * per, g_hi = min(total, g_lo + per);
  if (g_lo >= g_hi) return;

  // walking cursor over the items: (b, h, chunk) and the question's geometry
  struct Cur { int b, h, ch, ns, kstart, Mb; };
  auto seek = [&](int g) {
    Cur q;
    int lo = 0, hi = p.B - 1;  // largest b with s_pref[b] <= g
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_pref[mid] <= g) lo = mid; else hi = mid - 1;
    }
    q.b = lo;
    q.kstart = q.b * p.M;
    q.Mb = p.M;
    if (p.cu) {
      q.kstart = p.cu[q.b * p.npsg];
      q.Mb = p.cu[(q.b + 1) * p.npsg] - q.kstart;
    }
    q.ns = (q.Mb + XD_CH - 1) / XD_CH;
    const int r = g - s_pref[q.b];
    q.h = r / q.ns;  // items of one (b, h) are consecutive
    q.ch = r - q.h * q.ns;
    return q;
  };
  auto advance = [&](Cur& q) {
    if (++q.ch < q.ns) return;
    q.ch = 0;
    if (++q.h < p.H) return;
    q.h = 0;
    ++q.b;
    if (q.b >= p.B) return;
    q.kstart = q.b * p.M;
    q.Mb = p.M;
    if (p.cu) {
      q.kstart = p.cu[q.b * p.npsg];
      q.Mb = p.cu[(q.b + 1) * p.npsg] - q.kstart;
    }
    q.ns = (q.Mb + XD_CH - 1) / XD_CH;
  };
  const uint64_t kv_policy = l2_policy_evict_first();
  auto issue = [&](const Cur& q, int stage) {  // thread 0 only
    int kc0 = p.koff + q.h * 64, kr0 = q.kstart + q.ch * XD_CH, vc0 = p.voff + q.h * 64, vr0 = kr0;
    if (p.hm_rows > 0) {  // row of the [planes * tokens][64] view
      kr0 += static_cast<int>((kc0 >> 6) * p.hm_rows);
      vr0 += static_cast<int>((vc0 >> 6) * p.hm_rows);
      kc0 = vc0 = 0;
    }
    uint8_t* st = smem + stage * 2 * XD_TILE;
    mbar_arrive_expect_tx(&bar[stage], 2 * XD_TILE);
    // K / V of a layer are read exactly once per token: evict_first, so that the decoder weights the neighbouring
    // Linears prefetch to L2 survive this ~130 MB stream
    tma_load_2d_hint(st, &tm_kv, &bar[stage], kc0, kr0, kv_policy);
    tma_load_2d_hint(st + XD_TILE, &tm_kv, &bar[stage], vc0, vr0, kv_policy);
  };
  Cur cur = seek(g_lo);
  Cur nxt = cur;  // producer cursor (thread 0): the next item to request
  int g_issued = g_lo;
  if (tid == 0)
    for (int i = 0; i < XD_STAGES - 1 && g_issued < g_hi; ++i) {
      issue(nxt, i);
      advance(nxt);
      ++g_issued;
    }
  pdl_wait();
  const bool capture = p.tok_score != nullptr && (p.step == nullptr || *p.step == 0);
  uint32_t phase_bits = 0;  // parity of each stage's barrier
  // Online-softmax state of this warp over the current segment (consecutive chunks of one (b, h)), in mma.sync
  // m16n8k16 fragment layout with the query as row 0 of the 16-row A operand (rows 1..15 are zero): lanes 0..3 (g == 0)
  // carry the real row -- m / l, and in oacc[nt][0..1] the output dims nt*8 + 2*t4 (+1).  Tensor-core dot products keep
  // the instruction count per key ~4x below a SIMT reduction, which is what bounded the first versions of this kernel
  // (issue-latency bound at 1 instruction/clock/SM, not memory bound).
  const int fr = lane >> 2, t4 = lane & 3;  // fragment row (query row fr / fr + 8) and column pair
  float m_run = -INFINITY, l_run = 0.f, oacc[8][4];
  uint32_t qa[4][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) qa[i][e] = 0u;
  int seg_g0 = g_lo;  // first item of the current segment

  // publish the finished segment of (sb, sh): merge the 4 warps, store the partial, count the arrival; the last
  // segment of the (question, head) merges all of its partials
  auto flush = [&](int sb, int sh, int sns, int g_end) {
    (void)g_end;
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    if (fr == 0) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s_o[warp][nt * 8 + t4 * 2] = oacc[nt][0];
        s_o[warp][nt * 8 + t4 * 2 + 1] = oacc[nt][1];
      }
      if (t4 == 0) {
        s_m[warp] = m_run;
        s_l[warp] = l_run;
      }
    }
    __syncthreads();
    // segments of (sb, sh) are cut at multiples of `per` in the global item order
    const int gq0 = s_pref[sb] + sh * sns;  // first item of this (question, head)
    const int first_cta = gq0 / per;
    const int nseg = (gq0 + sns - 1) / per - first_cta + 1;
    const int ord = seg_g0 / per - first_cta;
    const size_t pbase = (static_cast<size_t>(sb) * p.H + sh) * p.splits;
    if (tid < 64) {
      const float m = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
      float ov = 0.f, l = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float scw = (s_m[w] == -INFINITY) ? 0.f : fast_exp2((s_m[w] - m) * DC_LOG2E);
        ov += s_o[w][tid] * scw;
        l += s_l[w] * scw;
      }
      float* part = p.part + (pbase + ord) * 66;
      part[tid] = ov;
      if (tid == 0) {
        part[64] = m;
        part[65] = l;
      }
    }
    __syncthreads();  // partial stored by this CTA's threads
    if (tid == 0) {
      __threadfence();  // cumulative: the CTA's partial is visible device-wide before the arrival is counted
      const int prev = atomicAdd(&p.counters[sb * p.H + sh], 1);
      s_last = (prev == nseg - 1) ? 1 : 0;
      if (prev == nseg - 1) p.counters[sb * p.H + sh] = 0;  // ready for the next launch
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (tid < 64) {
        const float* pp = p.part + pbase * 66;
        float m = -INFINITY;
        for (int s2 = 0; s2 < nseg; ++s2) m = fmaxf(m, __ldcg(pp + s2 * 66 + 64));
        float l = 0.f, ov = 0.f;
        for (int s2 = 0; s2 < nseg; ++s2) {
          const float w = fast_exp2((__ldcg(pp + s2 * 66 + 64) - m) * DC_LOG2E);
          l += w * __ldcg(pp + s2 * 66 + 65);
          ov += w * __ldcg(pp + s2 * 66 + tid);
        }
        p.out[static_cast<size_t>(sb) * p.ldo + sh * 64 + tid] = __float2bfloat16(ov / l);
        if (p.q_zero) p.q_zero[static_cast<size_t>(sb) * p.ldq + sh * 64 + tid] = 0.f;
      }
    }
    m_run = -INFINITY;
    l_run = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
  };

  int seg_b = -1, seg_h = -1, seg_ns = 0;
  for (int g = g_lo; g < g_hi; ++g) {
    const int stage = (g - g_lo) % XD_STAGES;
    if (cur.b != seg_b || cur.h != seg_h) {  // CTA-uniform: a new (question, head) starts
      if (seg_b >= 0) flush(seg_b, seg_h, seg_ns, g);
      seg_b = cur.b;
      seg_h = cur.h;
      seg_ns = cur.ns;
      seg_g0 = g;
      if (fr == 0) {  // A-operand fragments of the query row: a0 = dims ks*16 + 2*t4 (+1), a2 = the same + 8
        const float* qp = p.q + static_cast<size_t>(cur.b) * p.ldq + cur.h * 64 + t4 * 2;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const float2 lo = *reinterpret_cast<const float2*>(qp + ks * 16);
          const float2 hi = *reinterpret_cast<const float2*>(qp + ks * 16 + 8);
          qa[ks][0] = pack_bf16x2(lo.x, lo.y);
          qa[ks][2] = pack_bf16x2(hi.x, hi.y);
        }
      }
    }
    // keep two items in flight: the stage that item g + 2 will use was released by the barrier at the end of item g - 1
    if (tid == 0 && g_issued < g_hi) {
      issue(nxt, (g_issued - g_lo) % XD_STAGES);
      advance(nxt);
      ++g_issued;
    }
    const uint32_t aK = smem_u32(smem + stage * 2 * XD_TILE), aV = aK + XD_TILE;
    const int k0 = cur.ch * XD_CH;
    mbar_wait(&bar[stage], (phase_bits >> stage) & 1u);
    phase_bits ^= 1u << stage;
    const int kb = warp * 32;  // this warp's keys within the item
    // ---- S = q K^T for 32 keys (4 n-tiles of 8): row 0 in lanes g == 0, keys kb + nt*8 + 2*t4 (+1)
    float sacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t kb4[4];
        dc_ldsm_x4(kb4, aK + dc_swz(kb + np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, ks * 2 + ((lane >> 3) & 1)));
        mma_16816(sacc[np * 2], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], kb4[0], kb4[1]);
        mma_16816(sacc[np * 2 + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], kb4[2], kb4[3]);
      }
    }
    // ---- validity, step-0 score capture, running softmax (row 0 only; the other 15 rows are zero padding)
    float sv[8];
    float m_it = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = k0 + kb + nt * 8 + t4 * 2 + e;
        float sx = sacc[nt][e];
        if (fr != 0 || key >= cur.Mb) {
          sx = -INFINITY;  // padding rows; keys past this question's memory (other questions' rows / zeroed slab tail)
        } else {
          bool valid = true;
          if (p.mask) valid = p.mask[static_cast<size_t>(cur.b) * p.M + key] != 0;
          if (capture && valid) {
            const size_t ti = p.tok_index ? static_cast<size_t>(p.tok_index[cur.kstart + key])
                                          : static_cast<size_t>(cur.b) * p.M + key;
            atomicAdd(p.tok_score + ti, sx);
          }
          if (!valid) sx += DC_MASK_MIN;  // HF's additive mask (finfo.min): a fully masked question degrades to uniform
        }
        sv[nt * 2 + e] = sx;
        m_it = fmaxf(m_it, sx);
      }
    m_it = fmaxf(m_it, __shfl_xor_sync(0xffffffffu, m_it, 1));
    m_it = fmaxf(m_it, __shfl_xor_sync(0xffffffffu, m_it, 2));
    const float m_new = fmaxf(m_run, m_it);
    const float sa = (m_run == -INFINITY || m_new == -INFINITY) ? 0.f : fast_exp2((m_run - m_new) * DC_LOG2E);
    uint32_t pa[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float e0 = (sv[nt * 2] == -INFINITY) ? 0.f : fast_exp2((sv[nt * 2] - m_new) * DC_LOG2E);
      const float e1 = (sv[nt * 2 + 1] == -INFINITY) ? 0.f : fast_exp2((sv[nt * 2 + 1] - m_new) * DC_LOG2E);
      l_run = (nt == 0 ? l_run * sa : l_run) + e0 + e1;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(e0, e1);  // row g (only g == 0 is non-zero)
      pa[nt >> 1][(nt & 1) * 2 + 1] = 0u;                   // row g + 8: padding
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      oacc[i][0] *= sa;
      oacc[i][1] *= sa;
    }
    m_run = m_new;
    // ---- O += P V over this warp's 32 keys (V as the k16 x n8 B operand via ldmatrix.trans)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        dc_ldsm_x4_t(vb, aV + dc_swz(kb + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
        mma_16816(oacc[dp * 2], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], vb[0], vb[1]);
        mma_16816(oacc[dp * 2 + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], vb[2], vb[3]);
      }
    }
    advance(cur);
    __syncthreads();  // every thread is done reading this stage's tiles: thread 0 may refill it
  }
  flush(seg_b, seg_h, seg_ns, g_hi);
}

}  // namespace

size_t xattn_decode_workspace_bytes(int B, int H, int M) {
  // chunk partials + one arrival counter per (question, head)
  return static_cast<size_t>(B) * H * ((M + XD_CH - 1) / XD_CH) * 66 * sizeof(float) + static_cast<size_t>(B) * H * sizeof(int) + 256;
}

// q fp32 [B][ldq]; kv bf16 rows of pitch ldkv (kv_rows of them); workspace must be zero before the FIRST call (the
// kernel leaves its counters zeroed again).  zero_q: clear each question's query slice after use (the K-split
// projection that produces q accumulates into it).
int xattn_decode(const float* q, const __nv_bfloat16* kv, size_t kv_rows, const unsigned char* mask, __nv_bfloat16* out,
                 float* tok_score, const int* step, float* workspace, int B, int M, int H, int ldq, int ldkv, int koff,
                 int voff, int ldo, cudaStream_t s, const int* cu, int npsg, const int* tok_index, bool zero_q,
                 bool head_major) {
  LAKO_REQUIRE(B > 0 && M > 0 && H > 0 && kv_rows > 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(ldq % 4 == 0 && ldkv % 8 == 0 && koff % 8 == 0 && voff % 8 == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(!head_major || (ldkv % 64 == 0 && koff % 64 == 0 && voff % 64 == 0 &&
                               static_cast<uint64_t>(ldkv / 64) * kv_rows < (1ull << 31)), LAKO_ERR_UNSUPPORTED);
  const int splits = (M + XD_CH - 1) / XD_CH;
  int* counters = reinterpret_cast<int*>(workspace + static_cast<size_t>(B) * H * splits * 66);
  LAKO_REQUIRE(B <= XD_MAXB, LAKO_ERR_UNSUPPORTED);
  XDecParams p{q, zero_q ? const_cast<float*>(q) : nullptr, mask, workspace, counters, out, tok_score, step, cu, tok_index,
               B, M, H, ldq, ldo, koff, voff, splits, npsg, head_major ? static_cast<long long>(kv_rows) : 0};
  CUtensorMap tm;
  int rc;
  if (head_major)  // [ldkv/64 planes][kv_rows][64] seen as one tall [planes * kv_rows][64] matrix
    rc = get_tensor_map_2d(&tm, kv, 64, static_cast<uint64_t>(ldkv / 64) * kv_rows, 64, 64, XD_CH);
  else
    rc = get_tensor_map_2d(&tm, kv, static_cast<uint64_t>(ldkv), kv_rows, static_cast<uint64_t>(ldkv), 64, XD_CH);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(xattn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XD_SMEM);
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return LAKO_ERR_LAUNCH;
    }
    configured = true;
  }
  // algorithmic bytes: K and V rows of every memory token of the batch for this layer
  ProfScope prof(PROF_XATTN, 4.0 * B * H * static_cast<double>(M) * 64, s);
  // persistent: two CTAs per SM (2 x 96 KB of tile ring), never more CTAs than work items
  const long long max_items = static_cast<long long>(B) * H * splits;
  const int grid = static_cast<int>(std::min<long long>(max_items, static_cast<long long>(XD_CTAS) * device_sm_count()));
  launch_pdl(xattn_decode_kernel, dim3(grid), dim3(128), XD_SMEM, s, tm, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
