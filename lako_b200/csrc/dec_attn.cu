// Decoder-side attention of the FiDT5 reader (< 1 % of the FLOPs; the cost is streaming the fused
// encoder memory's K/V once per layer).  The cross-attention chunk kernels themselves are the
// mma.sync kernels in xattn_mma.cu; this file holds their combine step, the host wrappers and the
// SIMT decoder self-attention:
//
//  * cross-attention over the fused memory (M = n_passages * text_maxlength keys per question),
//    split over key chunks (flash-decoding style partial softmax + combine), forward and backward.
//    Reference: cross_attention_forward /root/reference/src/model.py:286-349 (q k^T + mask, fp32
//    softmax, probs v), and the score capture at :327-329 consumed by get_crossattention_scores
//    (:143-167): instead of keeping (B,H,1,M) pre-softmax scores per layer alive, the step-0 scores
//    are summed over heads (and layers, by accumulation) into tok_score[B][M] on the fly.
//  * causal decoder self-attention with the causal-bucket relative bias (HF:modeling_t5.py:253-344,
//    438-445), forward/backward for teacher-forced training (T <= 64) and the KV-cached
//    single-token step used by greedy generate.
#include "ops.cuh"

namespace lako {

constexpr float MASK_MIN_F = -3.4028234663852886e38f;
constexpr int XA_CH = 128;      // keys per CTA
constexpr int XA_TMAX = 64;     // max decoder length handled by the training kernels

// combine the split partials: out[b,t,h*64+d] = sum_s exp(m_s - m) o_s / sum_s exp(m_s - m) l_s
__global__ void xattn_combine_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ out,
                                     float* __restrict__ lse, int T, int H, int splits, int ldo,
                                     const int* __restrict__ cu, int npsg) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int stride_s = splits;  // workspace pitch (max chunks); a packed question may fill fewer
  if (cu) splits = (cu[(b + 1) * npsg] - cu[b * npsg] + XA_CH - 1) / XA_CH;
  const int d = threadIdx.x;  // 64 threads
  const float* pp = part + ((static_cast<size_t>(b) * H + h) * stride_s) * T * 66 + t * 66;
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, pp[static_cast<size_t>(s) * T * 66 + 64]);
  float l = 0.f, o = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* ps = pp + static_cast<size_t>(s) * T * 66;
    const float w = __expf(ps[64] - m);
    l += w * ps[65];
    o += w * ps[d];
  }
  out[(static_cast<size_t>(b) * T + t) * ldo + h * 64 + d] = __float2bfloat16(o / l);
  if (lse && d == 0) lse[(static_cast<size_t>(b) * H + h) * T + t] = m + logf(l);
}

size_t xattn_workspace_bytes(int B, int H, int T, int M) {
  const int splits = (M + XA_CH - 1) / XA_CH;
  return static_cast<size_t>(B) * H * splits * T * 66 * sizeof(float);
}

int xattn_fwd(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, __nv_bfloat16* out, float* lse,
              float* tok_score, float* workspace, int B, int T, int M, int H, int ldq, int ldkv, int koff, int voff,
              int ldo, cudaStream_t s, DropSpec drop, const int* cu, int npsg, const int* tok_index) {
  LAKO_REQUIRE(B > 0 && T > 0 && T <= XA_TMAX && M > 0 && H > 0, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE(ldkv % 8 == 0 && koff % 8 == 0 && voff % 8 == 0, LAKO_ERR_BAD_ARG);
  const int splits = (M + XA_CH - 1) / XA_CH;
  ProfScope prof(PROF_XATTN, 4.0 * B * H * static_cast<double>(T) * M * 64, s);
  int rc = xattn_fwd_mma(q, kv, mask, tok_score, workspace, B, T, M, H, ldq, ldkv, koff, voff, s, drop, cu, npsg,
                         tok_index);
  if (rc) return rc;
  launch_pdl(xattn_combine_kernel, dim3(T, H, B), dim3(64), 0, s, workspace, out, lse, T, H, splits, ldo, cu, npsg);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// dq[b*T+t][h*64+d] = sum over the question's key chunks, in chunk order, of the partials xattn_bwd_mma left
__global__ void xattn_dq_reduce_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ dq, int T, int H, int splits,
                                       const int* __restrict__ cu, int npsg) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z, d = threadIdx.x;  // 64 threads
  const int stride_s = splits;
  if (cu) splits = (cu[(b + 1) * npsg] - cu[b * npsg] + XA_CH - 1) / XA_CH;
  const float* pp = part + (((static_cast<size_t>(b) * H + h) * stride_s) * T + t) * 64 + d;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += pp[static_cast<size_t>(s) * T * 64];
  dq[(static_cast<size_t>(b) * T + t) * (H * 64) + h * 64 + d] = __float2bfloat16(acc);
}

int xattn_bwd(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, const __nv_bfloat16* out,
              const __nv_bfloat16* dout, const float* lse, __nv_bfloat16* dkv, float* dq_f32, int B, int T, int M,
              int H, int ldq, int ldkv, int koff, int voff, int ldo, cudaStream_t s, DropSpec drop, const int* cu,
              int npsg, float* dq_part, __nv_bfloat16* dq_bf16) {
  LAKO_REQUIRE(B > 0 && T > 0 && T <= XA_TMAX && M > 0 && H > 0, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE((dq_part == nullptr) == (dq_bf16 == nullptr) && (dq_part || dq_f32), LAKO_ERR_BAD_ARG);
  ProfScope prof(PROF_XATTN, 8.0 * B * H * static_cast<double>(T) * M * 64, s);
  int rc = xattn_bwd_mma(q, kv, mask, out, dout, lse, dkv, dq_f32, B, T, M, H, ldq, ldkv, koff, voff, ldo, s, drop, cu,
                         npsg, dq_part);
  if (rc || !dq_part) return rc;
  const int splits = (M + XA_CH - 1) / XA_CH;
  launch_pdl(xattn_dq_reduce_kernel, dim3(T, H, B), dim3(64), 0, s, static_cast<const float*>(dq_part), dq_bf16, T, H, splits,
             cu, npsg);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// ------------------------------------------------------------------------------------------------
// decoder self-attention, teacher-forced (training): one CTA per (head, question), T <= 64.
// qkv bf16 [B*T][3*inner]; lut fp32 [H][2T-1] (causal buckets; only rel <= 0 is used).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const __nv_bfloat16& v) { return __bfloat162float(v); }
__device__ __forceinline__ float ldf(const float& v) { return v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, __nv_bfloat16*, float v) { *p = __float2bfloat16(v); }
// fp32 mode: the result feeds a 3xTF32 GEMM, so it is stored as its (hi, lo) TF32 pair
__device__ __forceinline__ void stf(float* hi, float* lo, float v) {
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  *hi = h;
  if (lo) *lo = __uint_as_float(__float_as_uint(v - h) & 0xffffe000u);
}

template <typename E>
struct DSelfParams {
  const E* qkv;
  const float* lut;
  E* out;                   // [B*T][inner]  (fp32: hi part)
  E* out_lo;                // fp32 only: lo part
  float* probs;             // [B][H][T][T] fp32 saved for backward (un-dropped)
  int T, H, inner;
  DropSpec drop;            // group = ((b*H+h)*T + t)*8 + j/8
};

template <typename E>
__global__ void __launch_bounds__(128) dself_fwd_kernel(const DSelfParams<E> p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float dyn[];
  float (*sQ)[65] = reinterpret_cast<float (*)[65]>(dyn);
  float (*sK)[65] = reinterpret_cast<float (*)[65]>(dyn + XA_TMAX * 65);
  float (*sV)[65] = reinterpret_cast<float (*)[65]>(dyn + 2 * XA_TMAX * 65);
  float (*sP)[XA_TMAX] = reinterpret_cast<float (*)[XA_TMAX]>(dyn + 3 * XA_TMAX * 65);
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T;
  for (int i = tid; i < T * 64; i += 128) {
    const int t = i >> 6, d = i & 63;
    const E* row = p.qkv + (static_cast<size_t>(b) * T + t) * 3 * p.inner + h * 64 + d;
    sQ[t][d] = ldf(row[0]);
    sK[t][d] = ldf(row[p.inner]);
    sV[t][d] = ldf(row[2 * p.inner]);
  }
  __syncthreads();
  const float* lut = p.lut + h * (2 * T - 1) + (T - 1);
  for (int t = warp; t < T; t += 4) {
    float s[2], m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = lane + i * 32;
      s[i] = -INFINITY;
      if (j < T) {
        float acc = 0.f;
        for (int d = 0; d < 64; ++d) acc += sQ[t][d] * sK[j][d];
        s[i] = (j <= t) ? acc + lut[j - t] : MASK_MIN_F;
      }
      m = fmaxf(m, s[i]);
    }
    m = warp_max(m);
    float e[2], l = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      e[i] = (s[i] == -INFINITY) ? 0.f : __expf(s[i] - m);
      l += e[i];
    }
    l = warp_sum(l);
    const float inv = 1.f / l;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = lane + i * 32;
      if (j < T) {
        float pv = e[i] * inv;
        if (p.probs) p.probs[((static_cast<size_t>(b) * p.H + h) * T + t) * T + j] = pv;
        if (p.drop.thresh) {
          const uint32_t k = drop_keep8(p.drop, (static_cast<unsigned long long>(b * p.H + h) * T + t) * 8 + (j >> 3));
          pv = ((k >> (j & 7)) & 1u) ? pv * p.drop.scale : 0.f;
        }
        sP[warp][j] = pv;
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int d = lane + i * 32;
      float acc = 0.f;
      for (int j = 0; j <= t; ++j) acc += sP[warp][j] * sV[j][d];
      const size_t oi = (static_cast<size_t>(b) * T + t) * p.inner + h * 64 + d;
      stf(p.out + oi, p.out_lo ? p.out_lo + oi : nullptr, acc);
    }
    __syncwarp();
  }
}

int dself_fwd(const __nv_bfloat16* qkv, const float* lut, __nv_bfloat16* out, float* probs, int B, int T, int H,
              cudaStream_t s, DropSpec drop) {
  LAKO_REQUIRE(T > 0 && T <= XA_TMAX, LAKO_ERR_UNSUPPORTED);
  DSelfParams<__nv_bfloat16> p{qkv, lut, out, nullptr, probs, T, H, H * 64, drop};
  constexpr int smem = (3 * XA_TMAX * 65 + 4 * XA_TMAX) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(dself_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = true;
  }
  launch_pdl(dself_fwd_kernel<__nv_bfloat16>, dim3(dim3(H, B)), dim3(128), smem, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}
// fp32 mode: qkv fp32, result as its (hi, lo) TF32 pair
int dself_fwd_f32(const float* qkv, const float* lut, float* out_hi, float* out_lo, int B, int T, int H, cudaStream_t s) {
  LAKO_REQUIRE(T > 0 && T <= XA_TMAX, LAKO_ERR_UNSUPPORTED);
  DSelfParams<float> p{qkv, lut, out_hi, out_lo, nullptr, T, H, H * 64, DropSpec{}};
  constexpr int smem = (3 * XA_TMAX * 65 + 4 * XA_TMAX) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(dself_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = true;
  }
  launch_pdl(dself_fwd_kernel<float>, dim3(dim3(H, B)), dim3(128), smem, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

struct DSelfBwdParams {
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* dout;  // [B*T][inner]
  const float* probs;         // [B][H][T][T]
  __nv_bfloat16* dqkv;        // [B*T][3*inner]
  float* dlut;                // [H][2T-1] += dS by offset
  int T, H, inner;
  DropSpec drop;
};

__global__ void __launch_bounds__(128) dself_bwd_kernel(const DSelfBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float dyn[];
  float (*sQ)[65] = reinterpret_cast<float (*)[65]>(dyn);
  float (*sK)[65] = reinterpret_cast<float (*)[65]>(dyn + XA_TMAX * 65);
  float (*sV)[65] = reinterpret_cast<float (*)[65]>(dyn + 2 * XA_TMAX * 65);
  float (*sDO)[65] = reinterpret_cast<float (*)[65]>(dyn + 3 * XA_TMAX * 65);
  float (*sP)[XA_TMAX + 1] = reinterpret_cast<float (*)[XA_TMAX + 1]>(dyn + 4 * XA_TMAX * 65);
  float (*sDS)[XA_TMAX + 1] = reinterpret_cast<float (*)[XA_TMAX + 1]>(dyn + 5 * XA_TMAX * 65);
  float (*sPD)[XA_TMAX + 1] = reinterpret_cast<float (*)[XA_TMAX + 1]>(dyn + 6 * XA_TMAX * 65);  // dropped P
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T;
  for (int i = tid; i < T * 64; i += 128) {
    const int t = i >> 6, d = i & 63;
    const size_t r = static_cast<size_t>(b) * T + t;
    const __nv_bfloat16* row = p.qkv + r * 3 * p.inner + h * 64 + d;
    sQ[t][d] = __bfloat162float(row[0]);
    sK[t][d] = __bfloat162float(row[p.inner]);
    sV[t][d] = __bfloat162float(row[2 * p.inner]);
    sDO[t][d] = __bfloat162float(p.dout[r * p.inner + h * 64 + d]);
  }
  for (int i = tid; i < T * T; i += 128) sP[i / T][i % T] = p.probs[(static_cast<size_t>(b) * p.H + h) * T * T + i];
  __syncthreads();
  // dS = P * (dP - rowsum(P * dP)), dP[t][j] = dO[t] . V[j]
  for (int t = warp; t < T; t += 4) {
    float dp[2], dsum = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = lane + i * 32;
      dp[i] = 0.f;
      if (j < T && j <= t) {
        float acc = 0.f;
        for (int d = 0; d < 64; ++d) acc += sDO[t][d] * sV[j][d];
        float pd = sP[t][j];
        if (p.drop.thresh) {
          const uint32_t k = drop_keep8(p.drop, (static_cast<unsigned long long>(b * p.H + h) * T + t) * 8 + (j >> 3));
          const bool keep = (k >> (j & 7)) & 1u;
          acc = keep ? acc * p.drop.scale : 0.f;
          pd = keep ? pd * p.drop.scale : 0.f;
        }
        sPD[t][j] = pd;
        dp[i] = acc;
        dsum += sP[t][j] * acc;
      }
    }
    dsum = warp_sum(dsum);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = lane + i * 32;
      if (j < T) {
        const float ds = (j <= t) ? sP[t][j] * (dp[i] - dsum) : 0.f;
        sDS[t][j] = ds;
        if (j <= t && p.dlut && ds != 0.f) atomicAdd(&p.dlut[h * (2 * T - 1) + (T - 1) + (j - t)], ds);
      }
    }
  }
  __syncthreads();
  // dQ[t] = sum_j dS[t][j] K[j];  dK[j] = sum_t dS[t][j] Q[t];  dV[j] = sum_t P[t][j] dO[t]
  for (int i = tid; i < T * 64; i += 128) {
    const int t = i >> 6, d = i & 63;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j <= t; ++j) dq += sDS[t][j] * sK[j][d];
    for (int tt = t; tt < T; ++tt) {
      dk += sDS[tt][t] * sQ[tt][d];
      dv += sPD[tt][t] * sDO[tt][d];
    }
    __nv_bfloat16* row = p.dqkv + (static_cast<size_t>(b) * T + t) * 3 * p.inner + h * 64 + d;
    row[0] = __float2bfloat16(dq);
    row[p.inner] = __float2bfloat16(dk);
    row[2 * p.inner] = __float2bfloat16(dv);
  }
}

int dself_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* dout, const float* probs, __nv_bfloat16* dqkv, float* dlut,
              int B, int T, int H, cudaStream_t s, DropSpec drop) {
  LAKO_REQUIRE(T > 0 && T <= XA_TMAX, LAKO_ERR_UNSUPPORTED);
  DSelfBwdParams p{qkv, dout, probs, dqkv, dlut, T, H, H * 64, drop};
  constexpr int smem = 7 * XA_TMAX * 65 * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(dself_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = true;
  }
  launch_pdl(dself_bwd_kernel, dim3(dim3(H, B)), dim3(128), smem, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// ------------------------------------------------------------------------------------------------
// KV-cached single-token decoder self-attention (greedy generate step `pos`):
// appends this step's k,v to the cache and attends over positions 0..pos.
// qkv bf16 [B][3*inner]; cache bf16 [B][H][Tmax][64] x2; lut fp32 [H][2*Tmax-1] (causal buckets)
// ------------------------------------------------------------------------------------------------
template <typename E>
__global__ void __launch_bounds__(64) dself_step_kernel(const E* __restrict__ qkv, E* __restrict__ kcache,
                                                        E* __restrict__ vcache, const float* __restrict__ lut,
                                                        E* __restrict__ out, E* __restrict__ out_lo, int pos,
                                                        const int* __restrict__ pos_dev, int Tmax, int H, int inner) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sQ[64], sS[1024];
  if (pos_dev) pos = *pos_dev;  // step index kept on the device: one captured graph serves every decode step
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const E* row = qkv + static_cast<size_t>(b) * 3 * inner + h * 64;
  E* kc = kcache + (static_cast<size_t>(b) * H + h) * Tmax * 64;
  E* vc = vcache + (static_cast<size_t>(b) * H + h) * Tmax * 64;
  sQ[tid] = ldf(row[tid]);
  kc[pos * 64 + tid] = row[inner + tid];
  vc[pos * 64 + tid] = row[2 * inner + tid];
  __syncthreads();
  const float* l = lut + h * (2 * Tmax - 1) + (Tmax - 1);
  for (int j = tid; j <= pos; j += 64) {
    float acc = 0.f;
    for (int d = 0; d < 64; ++d) acc += sQ[d] * ldf(kc[j * 64 + d]);
    sS[j] = acc + l[j - pos];
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = 0; j <= pos; ++j) m = fmaxf(m, sS[j]);
  float lsum = 0.f, acc = 0.f;
  for (int j = 0; j <= pos; ++j) {
    const float e = __expf(sS[j] - m);
    lsum += e;
    acc += e * ldf(vc[j * 64 + tid]);
  }
  const size_t oi = static_cast<size_t>(b) * inner + h * 64 + tid;
  stf(out + oi, out_lo ? out_lo + oi : nullptr, acc / lsum);
}

int dself_step(const __nv_bfloat16* qkv, __nv_bfloat16* kcache, __nv_bfloat16* vcache, const float* lut,
               __nv_bfloat16* out, int B, int pos, int Tmax, int H, cudaStream_t s, const int* pos_dev) {
  LAKO_REQUIRE(pos >= 0 && pos < Tmax && Tmax <= 1024, LAKO_ERR_UNSUPPORTED);
  launch_pdl(dself_step_kernel<__nv_bfloat16>, dim3(dim3(H, B)), dim3(64), 0, s, qkv, kcache, vcache, lut, out,
             static_cast<__nv_bfloat16*>(nullptr), pos, pos_dev, Tmax, H, H * 64);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}
int dself_step_f32(const float* qkv, float* kcache, float* vcache, const float* lut, float* out_hi, float* out_lo, int B,
                   int pos, int Tmax, int H, cudaStream_t s, const int* pos_dev) {
  LAKO_REQUIRE(pos >= 0 && pos < Tmax && Tmax <= 1024, LAKO_ERR_UNSUPPORTED);
  launch_pdl(dself_step_kernel<float>, dim3(dim3(H, B)), dim3(64), 0, s, qkv, kcache, vcache, lut, out_hi, out_lo, pos,
             pos_dev, Tmax, H, H * 64);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

// greedy token selection: argmax over fp32 logits [B][ld], EOS/pad bookkeeping as HF greedy search
// (SURVEY 3.3): finished rows emit pad (0); a row finishes when it emits EOS (1).
__global__ void __launch_bounds__(1024) greedy_pick_kernel(const float* __restrict__ logits, int V, int ld,
                                                          long long* __restrict__ out_ids, int out_ld, int step,
                                                          int* __restrict__ unfinished, int* __restrict__ n_unfinished,
                                                          long long* __restrict__ next_ids, int ban_eos,
                                                          int* __restrict__ step_dev, int* __restrict__ done_dev,
                                                          int min_length) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (step_dev) {  // device-resident step counter (decode step k writes output column k + 1)
    step = *step_dev + 1;
    ban_eos = step < min_length;
  }
  const float* lr = logits + static_cast<size_t>(b) * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c4 = tid * 4; c4 < V; c4 += 4096) {  // ld % 4 == 0: rows are 16-byte aligned
    float vv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (c4 + 4 <= V) {
      const float4 t = *reinterpret_cast<const float4*>(lr + c4);
      vv[0] = t.x; vv[1] = t.y; vv[2] = t.z; vv[3] = t.w;
    } else {
      for (int e = 0; c4 + e < V; ++e) vv[e] = lr[c4 + e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c4 + e;
      if (c >= V || (ban_eos && c == 1)) continue;  // HF MinLengthLogitsProcessor: EOS logit = -inf below min_length
      const float v = vv[e];
      if (v > best || (v == best && c < bi)) {
        best = v;
        bi = c;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if ((tid & 31) == 0) {
    s_v[tid >> 5] = best;
    s_i[tid >> 5] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) {
        best = s_v[w];
        bi = s_i[w];
      }
    if (bi < 0 || bi >= V) bi = 0;  // all-NaN row: no comparison succeeded
    const int unf = unfinished[b];
    const long long tok = unf ? bi : 0;
    out_ids[static_cast<size_t>(b) * out_ld + step] = tok;
    next_ids[b] = tok;
    if (unf && tok == 1) {
      unfinished[b] = 0;
      atomicSub(n_unfinished, 1);
    }
    if (step_dev) {  // the last CTA to finish advances the counter (every CTA has read it by then)
      __threadfence();
      if (atomicAdd(done_dev, 1) == static_cast<int>(gridDim.x) - 1) {
        *done_dev = 0;
        *step_dev = step;
      }
    }
  }
}

int greedy_pick(const float* logits, int B, int V, int ld, long long* out_ids, int out_ld, int step, int* unfinished,
                int* n_unfinished, long long* next_ids, cudaStream_t s, int ban_eos, int* step_dev, int* done_dev,
                int min_length) {
  LAKO_REQUIRE(ld % 4 == 0, LAKO_ERR_BAD_ARG);
  launch_pdl(greedy_pick_kernel, dim3(B), dim3(1024), 0, s, logits, V, ld, out_ids, out_ld, step, unfinished, n_unfinished,
             next_ids, ban_eos, step_dev, done_dev, min_length);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
