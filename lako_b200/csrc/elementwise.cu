// HBM-bound kernels of the FiDT5 path: fused residual-add + T5 RMSNorm (fwd/bwd), embedding
// gather / scatter-add, fused log-softmax cross-entropy (fwd+bwd in one pass over the logits),
// gradient-norm reduction and the fused clip + HF-AdamW(no bias correction) + bf16 re-cast step.
//
// Reference semantics:
//   T5LayerNorm            HF:modeling_t5.py:55-68   x * rsqrt(mean(x^2) + eps) * w, fp32 accumulate
//   embedding              HF:modeling_t5.py:682
//   CrossEntropyLoss       HF:modeling_t5.py:1112-1117 (ignore_index=-100, mean over valid tokens)
//   clip_grad_norm_        /root/reference/train_reader.py:76
//   AdamW(correct_bias=0)  /root/reference/src/util.py:225
// All are 128-bit vectorised and use warp-shuffle reductions; algorithmic bytes per element are
// listed in DESIGN.md.
#include "ops.cuh"

namespace lako {

// --------------------------------------------------------------------------------------------
// residual-add + RMSNorm forward.  One warp per row, row cached in registers (d <= 1024, d % 128 == 0).
//   x      fp32 [rows][d]   residual stream in
//   delta  bf16 [rows][d]   sub-layer output to add (may be null)
//   x_out  fp32 [rows][d]   x + delta (may alias x, may be null when delta is null)
//   out    bf16 [rows][d]   w * x * rstd * out_scale
//   rstd   fp32 [rows]      saved for backward (may be null)
// --------------------------------------------------------------------------------------------
constexpr int NORM_MAX_VEC = 8;  // float4 per lane -> d <= 1024

// 4 keep bits for the float4 at element offset e (multiple of 4) of a dropout site
__device__ __forceinline__ uint32_t keep4(const DropSpec& ds, size_t e) {
  const uint32_t m = drop_keep8(ds, e >> 3);
  return (e & 4) ? (m >> 4) : (m & 15u);
}
// Keep-nibbles for a whole row handled as float4 slots c = (i*32 + lane)*4, i < nvec (all 32 lanes must call).
// Lanes 2j / 2j+1 hold the two halves of one 8-element Philox group in every slot; for each PAIR of slots the
// even lane draws slot i's group and the odd lane draws slot i+1's group at the same time, then they swap halves
// by shuffle: one Philox call per lane per two slots, no divergent idle lanes.
template <int MAXV>
__device__ __forceinline__ void keep_row(const DropSpec& ds, size_t row_elem0, int lane, int nvec, uint32_t (&kk)[MAXV]) {
#pragma unroll
  for (int i = 0; i < MAXV; i += 2) {
    if (i < nvec) {
      int it = i + (lane & 1);
      if (it >= nvec) it = i;
      const size_t e = row_elem0 + static_cast<size_t>((it * 32 + (lane & ~1)) * 4);
      const uint32_t m = drop_keep8(ds, e >> 3);
      const uint32_t mo = __shfl_xor_sync(0xffffffffu, m, 1);
      const uint32_t mi = (lane & 1) ? mo : m;                              // slot i   (drawn by the even lane)
      const uint32_t mi1 = (i + 1 < nvec) ? ((lane & 1) ? m : mo) : 0xffu;  // slot i+1 (drawn by the odd lane)
      kk[i] = (lane & 1) ? (mi >> 4) : (mi & 15u);
      if (i + 1 < MAXV) kk[i + 1] = (lane & 1) ? (mi1 >> 4) : (mi1 & 15u);
    }
  }
}

// Same, for warps whose lanes 2j / 2j+1 hold the two float4 halves of one 8-element group (all 32 lanes must
// call it): the even lane runs Philox, the odd lane takes its half by shuffle.
__device__ __forceinline__ uint32_t keep4_pair(const DropSpec& ds, size_t e, int lane) {
  uint32_t m = 0;
  if ((lane & 1) == 0) m = drop_keep8(ds, e >> 3);
  m = __shfl_sync(0xffffffffu, m, lane & ~1);
  return (lane & 1) ? (m >> 4) : (m & 15u);
}

__global__ void __launch_bounds__(256) rmsnorm_fwd_kernel(const float* x, const __nv_bfloat16* __restrict__ delta,
                                                           float* x_out, const float* __restrict__ w,
                                                           __nv_bfloat16* __restrict__ out,
                                                           float* __restrict__ rstd_out, int rows, int d, float eps,
                                                           float out_scale, const DropSpec dd, const DropSpec od) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = d >> 7;
  float4 xv[NORM_MAX_VEC];
  const float* xr = x + static_cast<size_t>(warp) * d;
  float* xo = x_out ? x_out + static_cast<size_t>(warp) * d : nullptr;
  const __nv_bfloat16* dr = delta ? delta + static_cast<size_t>(warp) * d : nullptr;
  // every load of the row is issued before the Philox work of the dropout masks, which then runs under their latency
  uint2 dvv[NORM_MAX_VEC];
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    if (i < nvec) {
      const int c = (i * 32 + lane) * 4;
      xv[i] = *reinterpret_cast<const float4*>(xr + c);
      if (dr) dvv[i] = *reinterpret_cast<const uint2*>(dr + c);
    }
  }
  uint32_t kdd[NORM_MAX_VEC], kod[NORM_MAX_VEC];
  if (dr && dd.thresh) keep_row<NORM_MAX_VEC>(dd, static_cast<size_t>(warp) * d, lane, nvec, kdd);
  if (out && od.thresh) keep_row<NORM_MAX_VEC>(od, static_cast<size_t>(warp) * d, lane, nvec, kod);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_MAX_VEC; ++i) {
    if (i < nvec) {
      const int c = (i * 32 + lane) * 4;
      float4 v = xv[i];
      if (dr) {
        const uint2 dv = dvv[i];
        float4 a = make_float4(bf16_lo(dv.x), bf16_hi(dv.x), bf16_lo(dv.y), bf16_hi(dv.y));
        if (dd.thresh) {  // dropout on the sub-layer output before the residual add (HF:375, :149)
          const uint32_t k = kdd[i];
          a.x = (k & 1) ? a.x * dd.scale : 0.f;
          a.y = (k & 2) ? a.y * dd.scale : 0.f;
          a.z = (k & 4) ? a.z * dd.scale : 0.f;
          a.w = (k & 8) ? a.w * dd.scale : 0.f;
        }
        v.x += a.x;
        v.y += a.y;
        v.z += a.z;
        v.w += a.w;
      }
      if (xo && (dr || xo != xr)) *reinterpret_cast<float4*>(xo + c) = v;
      xv[i] = v;
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / static_cast<float>(d) + eps);
  if (rstd_out && lane == 0) rstd_out[warp] = r;
  if (out) {
    __nv_bfloat16* orow = out + static_cast<size_t>(warp) * d;
    const float rs = r * out_scale;
#pragma unroll
    for (int i = 0; i < NORM_MAX_VEC; ++i) {
      if (i < nvec) {
        const int c = (i * 32 + lane) * 4;
        const float4 wv = *reinterpret_cast<const float4*>(w + c);
        float4 y = make_float4(wv.x * (xv[i].x * rs), wv.y * (xv[i].y * rs), wv.z * (xv[i].z * rs), wv.w * (xv[i].w * rs));
        if (od.thresh) {  // dropout after the final layer norm of a stack (HF:767-768)
          const uint32_t k = kod[i];
          y.x = (k & 1) ? y.x * od.scale : 0.f;
          y.y = (k & 2) ? y.y * od.scale : 0.f;
          y.z = (k & 4) ? y.z * od.scale : 0.f;
          y.w = (k & 8) ? y.w * od.scale : 0.f;
        }
        uint2 o;
        o.x = pack_bf16x2(y.x, y.y);
        o.y = pack_bf16x2(y.z, y.w);
        *reinterpret_cast<uint2*>(orow + c) = o;
      }
    }
  }
}

int rmsnorm_fwd(const float* x, const __nv_bfloat16* delta, float* x_out, const float* w, __nv_bfloat16* out,
                float* rstd, int rows, int d, float eps, float out_scale, cudaStream_t s, DropSpec dd, DropSpec od) {
  LAKO_REQUIRE(rows > 0 && d > 0 && d % 128 == 0 && d <= 128 * NORM_MAX_VEC, LAKO_ERR_UNSUPPORTED);
  // algorithmic bytes: read x (4) [+ delta (2) + write x (4)] + write normed (2)
  ProfScope prof(PROF_NORM, static_cast<double>(rows) * d * (delta ? 12.0 : 6.0), s);
  const int warps_per_block = 8;
  const int grid = (rows + warps_per_block - 1) / warps_per_block;
  launch_pdl(rmsnorm_fwd_kernel, dim3(grid), dim3(warps_per_block * 32), 0, s, x, delta, x_out, w, out, rstd, rows, d, eps, out_scale, dd, od);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// --------------------------------------------------------------------------------------------
// RMSNorm backward, fused with the residual-gradient accumulation.
//   dy     bf16 [rows][d]  gradient w.r.t. the normalised output (from the dgrad GEMM); scaled by dy_scale
//   x      fp32 [rows][d]  the (post-residual) input that was normalised; rstd fp32 [rows]
//   dres   fp32 [rows][d]  running gradient of the residual stream: dres = (accumulate ? dres : 0) + dx
//   dres_bf16 (optional)   bf16 copy of the updated dres (operand of the next dgrad/wgrad GEMMs)
//   dw     fp32 [d]        += sum_rows dy * x * rstd   (block-reduced, one atomic per column per block)
// --------------------------------------------------------------------------------------------
// NV = d / 128 float4 slots per lane.  The per-column weight-gradient partial sums live in a per-warp
// private slice of shared memory (no atomics, 128-bit conflict-free accesses) instead of registers, which
// keeps the kernel at ~70 registers so three 8-warp CTAs fit per SM and enough loads are in flight.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int NV>
__global__ void __launch_bounds__(256, 2) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                              const float* __restrict__ x,
                                                              const float* __restrict__ rstd,
                                                              const float* __restrict__ w, float* __restrict__ dres,
                                                              __nv_bfloat16* __restrict__ dres_bf16,
                                                              float* __restrict__ dw, int rows, int accumulate,
                                                              float dy_scale, const DropSpec yd, const DropSpec bd) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int d = NV * 128;
  extern __shared__ float4 s_norm[];
  float4* s_w = s_norm;                // [d/4]
  float4* s_dw = s_norm + d / 4;       // [8 warps][d/4]
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < d / 4; i += 256) s_w[i] = reinterpret_cast<const float4*>(w)[i];
  float4* my_dw = s_dw + wib * (d / 4);
#pragma unroll
  for (int i = 0; i < NV; ++i) my_dw[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int warps_total = gridDim.x * 8;
  for (int row = blockIdx.x * 8 + wib; row < rows; row += warps_total) {
    const float r = rstd[row];
    const float* xr = x + static_cast<size_t>(row) * d;
    const __nv_bfloat16* dyr = dy + static_cast<size_t>(row) * d;
    float* drr = dres + static_cast<size_t>(row) * d;
    // L2 prefetch of what this warp touches next: the residual gradient of this row (read only after the
    // row reduction below) and the operands of its next row -- more bytes in flight per SM than the two
    // resident CTAs' register loads alone provide
    {
      constexpr int L32 = d * 4 / 128;  // 128-byte lines of an fp32 row
      if (accumulate && lane < L32) prefetch_l2(drr + lane * 32);
      const int nrow = row + warps_total;
      if (nrow < rows) {
        if (lane < L32) prefetch_l2(x + static_cast<size_t>(nrow) * d + lane * 32);
        if (lane < L32 / 2) prefetch_l2(dy + static_cast<size_t>(nrow) * d + lane * 64);
      }
    }
    float4 xv[NV];
    uint2 dyp[NV];
    uint32_t kyd[NV], kbd[NV];
    // the row's loads go out first; the Philox work of the dropout masks runs under their latency
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      xv[i] = *reinterpret_cast<const float4*>(xr + c);
      dyp[i] = *reinterpret_cast<const uint2*>(dyr + c);
    }
    if (yd.thresh) {  // gradient of the post-norm dropout: zero the dropped elements (the 1/(1-p) is in dy_scale)
      keep_row<NV>(yd, static_cast<size_t>(row) * d, lane, NV, kyd);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint32_t k = kyd[i];
        dyp[i].x &= ((k & 1) ? 0x0000ffffu : 0u) | ((k & 2) ? 0xffff0000u : 0u);
        dyp[i].y &= ((k & 4) ? 0x0000ffffu : 0u) | ((k & 8) ? 0xffff0000u : 0u);
      }
    }
    if (dres_bf16 && bd.thresh) keep_row<NV>(bd, static_cast<size_t>(row) * d, lane, NV, kbd);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 wv = s_w[i * 32 + lane];
      const float4 dyv = make_float4(bf16_lo(dyp[i].x) * dy_scale, bf16_hi(dyp[i].x) * dy_scale,
                                     bf16_lo(dyp[i].y) * dy_scale, bf16_hi(dyp[i].y) * dy_scale);
      s += dyv.x * wv.x * xv[i].x + dyv.y * wv.y * xv[i].y + dyv.z * wv.z * xv[i].z + dyv.w * wv.w * xv[i].w;
      float4 a = my_dw[i * 32 + lane];
      a.x += dyv.x * xv[i].x * r;
      a.y += dyv.y * xv[i].y * r;
      a.z += dyv.z * xv[i].z * r;
      a.w += dyv.w * xv[i].w * r;
      my_dw[i * 32 + lane] = a;
    }
    s = warp_sum(s);
    const float coef = r * r * r * s / static_cast<float>(d);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 wv = s_w[i * 32 + lane];
      const float4 dyv = make_float4(bf16_lo(dyp[i].x) * dy_scale, bf16_hi(dyp[i].x) * dy_scale,
                                     bf16_lo(dyp[i].y) * dy_scale, bf16_hi(dyp[i].y) * dy_scale);
      float4 o = make_float4(r * dyv.x * wv.x - coef * xv[i].x, r * dyv.y * wv.y - coef * xv[i].y,
                             r * dyv.z * wv.z - coef * xv[i].z, r * dyv.w * wv.w - coef * xv[i].w);
      if (accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(drr + c);
        o.x += old.x;
        o.y += old.y;
        o.z += old.z;
        o.w += old.w;
      }
      *reinterpret_cast<float4*>(drr + c) = o;
      if (dres_bf16) {
        if (bd.thresh) {  // the bf16 copy feeds the previous sub-layer's output GEMMs: apply ITS dropout mask
          const uint32_t k = kbd[i];
          o.x = (k & 1) ? o.x * bd.scale : 0.f;
          o.y = (k & 2) ? o.y * bd.scale : 0.f;
          o.z = (k & 4) ? o.z * bd.scale : 0.f;
          o.w = (k & 8) ? o.w * bd.scale : 0.f;
        }
        uint2 ob;
        ob.x = pack_bf16x2(o.x, o.y);
        ob.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(dres_bf16 + static_cast<size_t>(row) * d + c) = ob;
      }
    }
  }
  if (dw) {
    __syncthreads();
    const float* f_dw = reinterpret_cast<const float*>(s_dw);
    for (int c = threadIdx.x; c < d; c += 256) {
      float acc = 0.f;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) acc += f_dw[ww * d + c];
      atomicAdd(dw + c, acc);
    }
  }
}

template <int NV>
static int launch_rmsnorm_bwd(const __nv_bfloat16* dy, const float* x, const float* rstd, const float* w, float* dres,
                              __nv_bfloat16* dres_bf16, float* dw, int rows, int accumulate, float dy_scale,
                              cudaStream_t s, DropSpec yd, DropSpec bd) {
  constexpr int smem = (NV * 128 * 9) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(rmsnorm_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = true;
  }
  int grid = (rows + 7) / 8;
  grid = min(grid, device_sm_count() * 2);
  launch_pdl(rmsnorm_bwd_kernel<NV>, dim3(grid), dim3(256), smem, s, dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, yd, bd);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

int rmsnorm_bwd(const __nv_bfloat16* dy, const float* x, const float* rstd, const float* w, float* dres,
                __nv_bfloat16* dres_bf16, float* dw, int rows, int d, int accumulate, float dy_scale,
                cudaStream_t s, DropSpec yd, DropSpec bd) {
  LAKO_REQUIRE(rows > 0 && d > 0 && d % 128 == 0 && d <= 128 * NORM_MAX_VEC, LAKO_ERR_UNSUPPORTED);
  // algorithmic bytes: read dy (2) + x (4) + dres (4 if accumulating); write dres (4) + bf16 copy (2)
  ProfScope prof(PROF_NORM, static_cast<double>(rows) * d * (accumulate ? 16.0 : 12.0), s);
  switch (d / 128) {
    case 1: return launch_rmsnorm_bwd<1>(dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, s, yd, bd);
    case 2: return launch_rmsnorm_bwd<2>(dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, s, yd, bd);
    case 3: return launch_rmsnorm_bwd<3>(dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, s, yd, bd);
    case 4: return launch_rmsnorm_bwd<4>(dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, s, yd, bd);
    case 5: return launch_rmsnorm_bwd<5>(dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, s, yd, bd);
    case 6: return launch_rmsnorm_bwd<6>(dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, s, yd, bd);
    case 7: return launch_rmsnorm_bwd<7>(dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, s, yd, bd);
    default: return launch_rmsnorm_bwd<8>(dy, x, rstd, w, dres, dres_bf16, dw, rows, accumulate, dy_scale, s, yd, bd);
  }
}

// --------------------------------------------------------------------------------------------
// embedding gather (fp32 table -> fp32 residual stream) and its scatter-add backward
// --------------------------------------------------------------------------------------------
// tok_index[cu[p] + pos] = p * L + pos : packed (un-padded) row -> flat index into the padded [P][L] id / score arrays
__global__ void build_tok_index_kernel(const int* __restrict__ cu, int* __restrict__ tok_index, int P, int L) {
  pdl_launch_dependents();
  pdl_wait();
  const int psg = blockIdx.x;
  const int r0 = cu[psg], n = cu[psg + 1] - r0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tok_index[r0 + i] = psg * L + i;
}
int build_tok_index(const int* cu, int* tok_index, int P, int L, cudaStream_t s) {
  launch_pdl(build_tok_index_kernel, dim3(P), dim3(128), 0, s, cu, tok_index, P, L);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

__global__ void embed_gather_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                    float* __restrict__ out, int rows, int d, int vocab, const DropSpec ds,
                                    const int* __restrict__ index) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  long long id = ids[index ? index[warp] : warp];
  if (id < 0 || id >= vocab) id = 0;
  const float4* src = reinterpret_cast<const float4*>(table + static_cast<size_t>(id) * d);
  float4* dst = reinterpret_cast<float4*>(out + static_cast<size_t>(warp) * d);
  for (int c = lane; c < (d >> 2); c += 32) {
    float4 v = __ldg(src + c);
    if (ds.thresh) {  // dropout on the embedded inputs (HF:734)
      const uint32_t k = keep4(ds, static_cast<size_t>(warp) * d + c * 4);
      v.x = (k & 1) ? v.x * ds.scale : 0.f;
      v.y = (k & 2) ? v.y * ds.scale : 0.f;
      v.z = (k & 4) ? v.z * ds.scale : 0.f;
      v.w = (k & 8) ? v.w * ds.scale : 0.f;
    }
    dst[c] = v;
  }
}

int embed_gather(const long long* ids, const float* table, float* out, int rows, int d, int vocab, cudaStream_t s,
                 DropSpec ds, const int* index) {
  LAKO_REQUIRE(rows > 0 && d % 8 == 0, LAKO_ERR_BAD_ARG);
  launch_pdl(embed_gather_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, ids, table, out, rows, d, vocab, ds, index);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

__global__ void embed_scatter_add_kernel(const long long* __restrict__ ids, const float* __restrict__ dres,
                                         float* __restrict__ dtable, int rows, int d, int vocab, const DropSpec ds,
                                         const int* __restrict__ index) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  long long id = ids[index ? index[warp] : warp];
  if (id < 0 || id >= vocab) id = 0;
  const float4* src = reinterpret_cast<const float4*>(dres + static_cast<size_t>(warp) * d);
  float* dst = dtable + static_cast<size_t>(id) * d;
  for (int c = lane; c < (d >> 2); c += 32) {
    float4 v = src[c];
    if (ds.thresh) {
      const uint32_t k = keep4(ds, static_cast<size_t>(warp) * d + c * 4);
      v.x = (k & 1) ? v.x * ds.scale : 0.f;
      v.y = (k & 2) ? v.y * ds.scale : 0.f;
      v.z = (k & 4) ? v.z * ds.scale : 0.f;
      v.w = (k & 8) ? v.w * ds.scale : 0.f;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 4), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
  }
}

int embed_scatter_add(const long long* ids, const float* dres, float* dtable, int rows, int d, int vocab,
                      cudaStream_t s, DropSpec ds, const int* index) {
  LAKO_REQUIRE(rows > 0 && d % 8 == 0, LAKO_ERR_BAD_ARG);
  launch_pdl(embed_scatter_add_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, ids, dres, dtable, rows, d, vocab, ds,
             index);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// decoder input ids = shift_right(labels) (HF:modeling_t5.py:595-614): start token 0, -100 -> 0
__global__ void shift_right_kernel(const long long* __restrict__ labels, long long* __restrict__ dec_ids, int B,
                                   int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int t = i % T;
  long long v = (t == 0) ? 0 : labels[i - 1];
  if (v == -100) v = 0;
  dec_ids[i] = v;
}
int shift_right(const long long* labels, long long* dec_ids, int B, int T, cudaStream_t s) {
  launch_pdl(shift_right_kernel, dim3((B * T + 255) / 256), dim3(256), 0, s, labels, dec_ids, B, T);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// --------------------------------------------------------------------------------------------
// fused log-softmax cross-entropy, forward + backward in one launch.
//   logits fp32 [rows][ld]; labels int64 [rows] (-100 = ignore)
//   stats[0] += sum of row losses, stats[1] = number of valid rows (computed by count_valid_kernel,
//   possibly all-reduced across ranks before this launch)
//   dlogits bf16 [rows][ld_d] = (softmax - onehot) / stats[1]  (0 for ignored rows); may be null
// One CTA per row; pass 1 = online max/sum, pass 2 re-reads the row (L2 resident) and writes dlogits.
// --------------------------------------------------------------------------------------------
__global__ void count_valid_kernel(const long long* __restrict__ labels, int n, float* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += (labels[i] != -100) ? 1 : 0;
  atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    stats[0] = 0.f;
    stats[1] = static_cast<float>(s_cnt);
  }
}

__global__ void __launch_bounds__(512) ce_fwd_bwd_kernel(const float* __restrict__ logits,
                                                          const long long* __restrict__ labels,
                                                          float* __restrict__ stats, float* __restrict__ row_loss,
                                                          __nv_bfloat16* __restrict__ dlogits, int V, int ld, int ld_d) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_m[16], s_s[16];
  __shared__ float s_bm, s_bs;
  const int row = blockIdx.x;
  const long long label = labels[row];
  const float* lr = logits + static_cast<size_t>(row) * ld;
  const bool valid = (label >= 0 && label < V);  // -100 = ignore_index; any other out-of-range label is ignored too
  float m = -INFINITY, ssum = 0.f;
  if (valid) {
    for (int c = threadIdx.x * 4; c < V; c += blockDim.x * 4) {
      if (c + 4 <= V) {
        const float4 v = *reinterpret_cast<const float4*>(lr + c);
        const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (mx > m) {
          ssum *= __expf(m - mx);
          m = mx;
        }
        ssum += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
      } else {
        for (int e = c; e < V; ++e) {
          const float v = lr[e];
          if (v > m) {
            ssum *= __expf(m - v);
            m = v;
          }
          ssum += __expf(v - m);
        }
      }
    }
    // warp then block reduction of (m, ssum)
    const float wm = warp_max(m);
    ssum *= (m == -INFINITY) ? 0.f : __expf(m - wm);  // threads/warps that saw no element carry (-inf, 0)
    ssum = warp_sum(ssum);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
      s_m[w] = wm;
      s_s[w] = ssum;
    }
    __syncthreads();
    if (w == 0) {
      const int nw = blockDim.x >> 5;
      float mm = (l < nw) ? s_m[l] : -INFINITY;
      float sv = (l < nw) ? s_s[l] : 0.f;
      const float bm = warp_max(mm);
      sv *= (mm == -INFINITY) ? 0.f : __expf(mm - bm);
      sv = warp_sum(sv);
      if (l == 0) {
        s_bm = bm;
        s_bs = sv;
      }
    }
    __syncthreads();
  }
  const float inv_count = 1.f / fmaxf(stats[1], 1.f);
  float bm = 0.f, inv_sum = 0.f;
  if (valid) {
    bm = s_bm;
    inv_sum = 1.f / s_bs;
    if (threadIdx.x == 0) {
      const float loss = logf(s_bs) + bm - lr[label];
      if (row_loss) row_loss[row] = loss;
      atomicAdd(&stats[0], loss);
    }
  } else if (threadIdx.x == 0 && row_loss) {
    row_loss[row] = 0.f;
  }
  if (dlogits) {
    __nv_bfloat16* dr = dlogits + static_cast<size_t>(row) * ld_d;
    for (int c = threadIdx.x * 8; c < V; c += blockDim.x * 8) {
      float g[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        g[e] = 0.f;
        if (valid && c + e < V) {
          g[e] = __expf(lr[c + e] - bm) * inv_sum;
          if (c + e == label) g[e] -= 1.f;
          g[e] *= inv_count;
        }
      }
      if (c + 8 <= V) {
        uint4 o;
        o.x = pack_bf16x2(g[0], g[1]);
        o.y = pack_bf16x2(g[2], g[3]);
        o.z = pack_bf16x2(g[4], g[5]);
        o.w = pack_bf16x2(g[6], g[7]);
        *reinterpret_cast<uint4*>(dr + c) = o;
      } else {
        for (int e = 0; c + e < V; ++e) dr[c + e] = __float2bfloat16(g[e]);
      }
    }
  }
}

int count_valid(const long long* labels, int n, float* stats, cudaStream_t s) {
  launch_pdl(count_valid_kernel, dim3(1), dim3(256), 0, s, labels, n, stats);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

int ce_fwd_bwd(const float* logits, const long long* labels, float* stats, float* row_loss, __nv_bfloat16* dlogits,
               int rows, int V, int ld, int ld_d, cudaStream_t s) {
  LAKO_REQUIRE(rows > 0 && V > 0 && ld % 4 == 0 && (dlogits == nullptr || ld_d % 8 == 0), LAKO_ERR_BAD_ARG);
  launch_pdl(ce_fwd_bwd_kernel, dim3(rows), dim3(512), 0, s, logits, labels, stats, row_loss, dlogits, V, ld, ld_d);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// --------------------------------------------------------------------------------------------
// casts / transposes / fills
// --------------------------------------------------------------------------------------------
// rezero: the source is a split-K accumulation scratch; leave it zeroed for its next user
__global__ void cast_f32_to_bf16_kernel(float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n, int rezero) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(in + i);
    const float4 b = *reinterpret_cast<const float4*>(in + i + 4);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y);
    o.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(out + i) = o;
    if (rezero) {
      *reinterpret_cast<float4*>(in + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(in + i + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (size_t e = i; e < n; ++e) {
      out[e] = __float2bfloat16(in[e]);
      if (rezero) in[e] = 0.f;
    }
  }
}
int cast_f32_to_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t s, bool rezero) {
  if (n == 0) return LAKO_OK;
  const size_t blocks = (n + 2047) / 2048;
  launch_pdl(cast_f32_to_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, const_cast<float*>(in), out, n,
             rezero ? 1 : 0);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// keep[e] = dropout keep decision of element e of a site (audit / test entry for the Philox mask)
__global__ void dropout_mask_kernel(unsigned char* __restrict__ keep, size_t n, const DropSpec ds) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t grp = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (grp * 8 >= n) return;
  const uint32_t m = ds.thresh ? drop_keep8(ds, grp) : 0xffu;
  for (int j = 0; j < 8 && grp * 8 + j < n; ++j) keep[grp * 8 + j] = (m >> j) & 1u;
}
int dropout_mask(unsigned char* keep, size_t n, unsigned long long seed, unsigned step, unsigned site, float p,
                 cudaStream_t s) {
  if (n == 0) return LAKO_OK;
  const size_t groups = (n + 7) / 8;
  launch_pdl(dropout_mask_kernel, dim3(static_cast<unsigned>((groups + 255) / 256)), dim3(256), 0, s, keep, n, make_drop(seed, step, site, p));
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// x *= *scalar (device scalar); used to apply d(loss) != 1 to the stored dlogits
__global__ void scale_bf16_dev_kernel(__nv_bfloat16* __restrict__ x, size_t n, const float* __restrict__ scalar) {
  pdl_launch_dependents();
  pdl_wait();
  const float sc = *scalar;
  if (sc == 1.0f) return;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] = __float2bfloat16(__bfloat162float(x[i]) * sc);
}
int scale_bf16_dev(__nv_bfloat16* x, size_t n, const float* scalar, cudaStream_t s) {
  if (n == 0) return LAKO_OK;
  launch_pdl(scale_bf16_dev_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, s, x, n, scalar);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// out[c][r] = in[r][c], bf16, 32x32 tiles through padded smem
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows,
                                      int cols, int ld_in, int ld_out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int r = r0 + i;
    for (int j = threadIdx.x; j < 64; j += blockDim.x) {
      const int c = c0 + j;
      tile[i][j] = (r < rows && c < cols) ? in[static_cast<size_t>(r) * ld_in + c] : __float2bfloat16(0.f);
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int c = c0 + i;
    for (int j = threadIdx.x; j < 64; j += blockDim.x) {
      const int r = r0 + j;
      if (r < rows && c < cols) out[static_cast<size_t>(c) * ld_out + r] = tile[j][i];
    }
  }
}
int transpose_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int rows, int cols, int ld_in, int ld_out,
                   cudaStream_t s) {
  dim3 grid((cols + 63) / 64, (rows + 63) / 64), block(32, 8);
  launch_pdl(transpose_bf16_kernel, dim3(grid), dim3(block), 0, s, in, out, rows, cols, ld_in, ld_out);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// --------------------------------------------------------------------------------------------
// gradient global-norm + fused clip / AdamW / bf16 re-cast over the flat parameter buffer
// --------------------------------------------------------------------------------------------
// Deterministic: block partials are combined in index order by whichever block finishes last (an atomicAdd per block
// made ||g||^2 -- hence the clip coefficient and, one ulp at a time, the updated weights -- depend on block scheduling:
// data-parallel replicas that must stay bit-identical drifted apart).
constexpr int SUMSQ_MAX_BLOCKS = 4096;
__device__ float g_sumsq_part[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_ticket;
__global__ void __launch_bounds__(256, 8) sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_w[8];
  __shared__ bool s_last;
  float acc = 0.f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (size_t e = i; e < n; ++e) acc += g[e] * g[e];
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < 8) ? s_w[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      g_sumsq_part[blockIdx.x] = v;
      __threadfence();
      s_last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float tot = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) tot += __ldcg(&g_sumsq_part[i]);
  tot = warp_sum(tot);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += s_w[w];
    *out += v;  // the stream serialises the calls that accumulate into *out
    g_sumsq_ticket = 0;
  }
}
int grad_sumsq(const float* g, size_t n, float* out, cudaStream_t s) {
  // caller zeroes *out
  ProfScope prof(PROF_OTHER, static_cast<double>(n) * 4.0, s);
  const int grid = std::min(device_sm_count() * 8, SUMSQ_MAX_BLOCKS);
  launch_pdl(sumsq_kernel, dim3(grid), dim3(256), 0, s, g, n, out);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// p, m, v, g fp32 flat; w_bf16 = bf16 copy of the updated p.  sumsq points at the device scalar with
// the (already all-reduced) sum of squared gradients; grad_scale pre-multiplies g (1/world for DP mean).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                     const float* __restrict__ g, __nv_bfloat16* __restrict__ w_bf16,
                                                     size_t n, const float* __restrict__ sumsq, float grad_scale,
                                                     float clip, float lr, float beta1, float beta2, float eps,
                                                     float wd) {
  pdl_launch_dependents();
  pdl_wait();
  float coef = grad_scale;
  if (clip > 0.f) {
    const float norm = sqrtf(*sumsq) * grad_scale;
    coef *= fminf(1.f, clip / (norm + 1e-6f));
  }
  const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 4 <= n) {
    float4 pv = *reinterpret_cast<const float4*>(p + i);
    float4 mv = *reinterpret_cast<const float4*>(m + i);
    float4 vv = *reinterpret_cast<const float4*>(v + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float* pp = reinterpret_cast<float*>(&pv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
    const float* gp = reinterpret_cast<const float*>(&gv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = gp[e] * coef;
      mp[e] = beta1 * mp[e] + (1.f - beta1) * gg;
      vp[e] = beta2 * vp[e] + (1.f - beta2) * gg * gg;
      float np = pp[e] - lr * (mp[e] / (sqrtf(vp[e]) + eps));
      np = np - lr * wd * np;
      pp[e] = np;
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    uint2 o;
    o.x = pack_bf16x2(pp[0], pp[1]);
    o.y = pack_bf16x2(pp[2], pp[3]);
    *reinterpret_cast<uint2*>(w_bf16 + i) = o;
  } else {
    for (size_t e = i; e < n; ++e) {
      const float gg = g[e] * coef;
      m[e] = beta1 * m[e] + (1.f - beta1) * gg;
      v[e] = beta2 * v[e] + (1.f - beta2) * gg * gg;
      float np = p[e] - lr * (m[e] / (sqrtf(v[e]) + eps));
      np = np - lr * wd * np;
      p[e] = np;
      w_bf16[e] = __float2bfloat16(np);
    }
  }
}
int adamw_step(float* p, float* m, float* v, const float* g, __nv_bfloat16* w_bf16, size_t n, const float* sumsq,
               float grad_scale, float clip, float lr, float beta1, float beta2, float eps, float wd,
               cudaStream_t s) {
  LAKO_REQUIRE(n > 0 && n % 4 == 0, LAKO_ERR_BAD_ARG);
  ProfScope prof(PROF_OTHER, static_cast<double>(n) * 30.0, s);
  const size_t blocks = (n / 4 + 255) / 256;
  launch_pdl(adamw_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, p, m, v, g, w_bf16, n, sumsq, grad_scale, clip, lr, beta1,
                                                             beta2, eps, wd);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
