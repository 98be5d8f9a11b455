// fp32 attention and normalisation kernels of the "precision = fp32" evaluation path (forward / generate / score
// extraction): plain SIMT fp32 arithmetic on fp32 activations, no tensor cores -- attention is 3.4 % of the reader's FLOPs
// and this mode exists for the reference's own numerics (it runs fp32 end to end, train_reader.py:67), not for speed.
// The dense projections of this mode are the 3xTF32 tcgen05 GEMMs of gemm_tf32.cu; every kernel here that feeds one
// writes its result already split into the (hi, lo) TF32 pair that GEMM consumes.
//
//   enc_attn_f32   T5 encoder self-attention per passage (HF:modeling_t5.py:253-344 via src/model.py:227-234):
//                  q k^T (no 1/sqrt(dk)) + bucketed relative bias + key mask, online softmax over 64-key tiles, any L.
//   xattn_f32      cross_attention_forward (src/model.py:286-349) for T >= 1 query rows per question against fp32 K/V,
//                  split over 128-key chunks + combine; step-0 score capture for query row 0.
//   rmsnorm_f32    T5LayerNorm (HF:55-68) + residual add, fp32 in / fp32 (hi, lo) out.
#include "ops.cuh"

namespace lako {

namespace {

constexpr float F32_MASK_MIN = -3.4028234663852886e38f;

__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// ------------------------------------------------------------------------------------------------ encoder attention
struct EAF32Params {
  const float* qkv;           // [rows][3*inner]
  const unsigned char* mask;  // [P][L] (padded layout) or null
  const float* lut;           // [H][2L-1]
  float *out, *out_hi, *out_lo;  // [rows][inner]; out may be null, hi/lo may be null
  int L, H, inner;
  const int* cu;
};

constexpr int EF_T = 64;      // tile edge (queries and keys)
constexpr int EF_P = 68;      // padded pitch of the transposed tiles (floats)

__global__ void __launch_bounds__(256) enc_attn_f32_kernel(const EAF32Params p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float ef_smem[];
  float* sQt = ef_smem;                 // [64 k][EF_P]   sQt[k][row]
  float* sKt = sQt + 64 * EF_P;         // [64 k][EF_P]   sKt[k][key]
  float* sPt = sKt + 64 * EF_P;         // [64 key][EF_P] sPt[key][row]
  float* sV = sPt + 64 * EF_P;          // [64 key][64]
  float* sKm = sV + 64 * 64;            // [64]
  float* sLut = sKm + 64;               // [2L-1]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int q0 = blockIdx.x * EF_T, h = blockIdx.y, psg = blockIdx.z;
  const int Lmax = p.L;
  int row0 = psg * Lmax, Lp = Lmax;
  if (p.cu) {
    row0 = p.cu[psg];
    Lp = p.cu[psg + 1] - row0;
  }
  if (q0 >= Lp) return;
  const int ld = 3 * p.inner;
  for (int i = tid; i < 2 * Lmax - 1; i += 256) sLut[i] = p.lut[h * (2 * Lmax - 1) + i];
  for (int idx = tid; idx < 64 * 16; idx += 256) {
    const int r = idx >> 4, c4 = idx & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Lp) v = *reinterpret_cast<const float4*>(p.qkv + static_cast<size_t>(row0 + q0 + r) * ld + h * 64 + c4 * 4);
    sQt[(c4 * 4 + 0) * EF_P + r] = v.x;
    sQt[(c4 * 4 + 1) * EF_P + r] = v.y;
    sQt[(c4 * 4 + 2) * EF_P + r] = v.z;
    sQt[(c4 * 4 + 3) * EF_P + r] = v.w;
  }
  float m[4], l[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  const int ktiles = (Lp + EF_T - 1) / EF_T;
  for (int kt = 0; kt < ktiles; ++kt) {
    __syncthreads();  // previous tile fully consumed (also orders the Q / LUT staging before first use)
    for (int idx = tid; idx < 64 * 16; idx += 256) {
      const int r = idx >> 4, c4 = idx & 15;
      const int key = kt * EF_T + r;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (key < Lp) {
        const float* base = p.qkv + static_cast<size_t>(row0 + key) * ld + h * 64 + c4 * 4;
        kv = *reinterpret_cast<const float4*>(base + p.inner);
        vv = *reinterpret_cast<const float4*>(base + 2 * p.inner);
      }
      sKt[(c4 * 4 + 0) * EF_P + r] = kv.x;
      sKt[(c4 * 4 + 1) * EF_P + r] = kv.y;
      sKt[(c4 * 4 + 2) * EF_P + r] = kv.z;
      sKt[(c4 * 4 + 3) * EF_P + r] = kv.w;
      *reinterpret_cast<float4*>(sV + r * 64 + c4 * 4) = vv;
    }
    if (tid < 64) {
      const int key = kt * EF_T + tid;
      float km = -INFINITY;
      if (key < Lp) km = (p.mask == nullptr || p.mask[static_cast<size_t>(psg) * Lmax + key]) ? 0.f : F32_MASK_MIN;
      sKm[tid] = km;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(sQt + k * EF_P + ty * 4);
      const float4 b = *reinterpret_cast<const float4*>(sKt + k * EF_P + tx * 4);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }
    float mt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qpos = min(q0 + ty * 4 + i, Lp - 1);
      mt[i] = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kl = tx * 4 + j;
        const int kpos = min(kt * EF_T + kl, Lmax - 1);
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
        pr[i][j] = expf(s[i][j] - m_new);  // -inf (beyond the passage) -> 0
        rs += pr[i][j];
        o[i][j] *= sc;
      }
      l[i] = l[i] * sc + rs;
      m[i] = m_new;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(sPt + (tx * 4 + j) * EF_P + ty * 4) = make_float4(pr[0][j], pr[1][j], pr[2][j], pr[3][j]);
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < 64; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(sPt + j * EF_P + ty * 4);
      const float4 b = *reinterpret_cast<const float4*>(sV + j * 64 + tx * 4);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < 4; ++d) o[i][d] = fmaf(av[i], bv[d], o[i][d]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) l[i] += __shfl_xor_sync(0xffffffffu, l[i], off);
    const int q = q0 + ty * 4 + i;
    if (q >= Lp) continue;
    const float inv = 1.f / l[i];
    const float4 v = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    const size_t off = static_cast<size_t>(row0 + q) * p.inner + h * 64 + tx * 4;
    if (p.out) *reinterpret_cast<float4*>(p.out + off) = v;
    if (p.out_hi) {
      const float4 hi = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
      *reinterpret_cast<float4*>(p.out_hi + off) = hi;
      *reinterpret_cast<float4*>(p.out_lo + off) =
          make_float4(tf32_trunc(v.x - hi.x), tf32_trunc(v.y - hi.y), tf32_trunc(v.z - hi.z), tf32_trunc(v.w - hi.w));
    }
  }
}

// ------------------------------------------------------------------------------------------------ cross-attention
constexpr int XF_CH = 128;  // keys per CTA (4 warps x 32)

struct XF32Params {
  const float* q;            // [B*T][ldq]
  const float* kv;           // rows = memory tokens
  const unsigned char* mask; // padded layout: [B][M]
  float* part;               // [B*T][H][splits][66]
  float* tok_score;          // [B][M] (may be null): += scores of query row t == 0 when *step == 0
  const int* step;
  const int* cu;
  const int* tok_index;
  int T, M, H, ldq, ldkv, koff, voff, splits, npsg;
};

__global__ void __launch_bounds__(128) xattn_f32_kernel(const XF32Params p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_m[4], s_l[4], s_o[4][64];
  const int split = blockIdx.x, h = blockIdx.y, bt = blockIdx.z;
  const int b = bt / p.T, t = bt - b * p.T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, kg = lane >> 4, c = lane & 15;
  int kstart = b * p.M, Mb = p.M;
  if (p.cu) {
    kstart = p.cu[b * p.npsg];
    Mb = p.cu[(b + 1) * p.npsg] - kstart;
  }
  const int k0 = split * XF_CH;
  if (k0 >= Mb) return;
  const bool capture = p.tok_score != nullptr && t == 0 && (p.step == nullptr || *p.step == 0);
  const float4 qv = *reinterpret_cast<const float4*>(p.q + static_cast<size_t>(bt) * p.ldq + h * 64 + c * 4);
  float m_run = -INFINITY, l_run = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
  const int wk0 = k0 + warp * 32;
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const int key = wk0 + j * 2 + kg;
    float4 kr = make_float4(0.f, 0.f, 0.f, 0.f), vr = kr;
    if (key < Mb) {
      const float* base = p.kv + (static_cast<size_t>(kstart) + key) * p.ldkv + h * 64 + c * 4;
      kr = *reinterpret_cast<const float4*>(base + p.koff);
      vr = *reinterpret_cast<const float4*>(base + p.voff);
    }
    float s = qv.x * kr.x + qv.y * kr.y + qv.z * kr.z + qv.w * kr.w;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (key >= Mb) continue;
    bool valid = true;
    if (p.mask) valid = p.mask[static_cast<size_t>(b) * p.M + key] != 0;
    if (capture && valid && c == 0) {
      const size_t ti = p.tok_index ? static_cast<size_t>(p.tok_index[kstart + key]) : static_cast<size_t>(b) * p.M + key;
      atomicAdd(p.tok_score + ti, s);
    }
    if (!valid) s += F32_MASK_MIN;
    const float m_new = fmaxf(m_run, s);
    const float sc = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
    const float pj = expf(s - m_new);
    l_run = l_run * sc + pj;
    o0 = o0 * sc + pj * vr.x;
    o1 = o1 * sc + pj * vr.y;
    o2 = o2 * sc + pj * vr.z;
    o3 = o3 * sc + pj * vr.w;
    m_run = m_new;
  }
  {  // merge the two 16-lane key groups
    const float m_o = __shfl_xor_sync(0xffffffffu, m_run, 16);
    const float l_o = __shfl_xor_sync(0xffffffffu, l_run, 16);
    const float m_new = fmaxf(m_run, m_o);
    const float sa = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
    const float sb = (m_o == -INFINITY) ? 0.f : expf(m_o - m_new);
    o0 = o0 * sa + __shfl_xor_sync(0xffffffffu, o0, 16) * sb;
    o1 = o1 * sa + __shfl_xor_sync(0xffffffffu, o1, 16) * sb;
    o2 = o2 * sa + __shfl_xor_sync(0xffffffffu, o2, 16) * sb;
    o3 = o3 * sa + __shfl_xor_sync(0xffffffffu, o3, 16) * sb;
    l_run = l_run * sa + l_o * sb;
    m_run = m_new;
  }
  if (kg == 0) {
    s_o[warp][c * 4 + 0] = o0;
    s_o[warp][c * 4 + 1] = o1;
    s_o[warp][c * 4 + 2] = o2;
    s_o[warp][c * 4 + 3] = o3;
    if (c == 0) {
      s_m[warp] = m_run;
      s_l[warp] = l_run;
    }
  }
  __syncthreads();
  if (tid < 64) {
    const float mm = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
    float ov = 0.f, ll = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float sc = (s_m[w] == -INFINITY) ? 0.f : expf(s_m[w] - mm);
      ov += s_o[w][tid] * sc;
      ll += s_l[w] * sc;
    }
    float* part = p.part + ((static_cast<size_t>(bt) * p.H + h) * p.splits + split) * 66;
    part[tid] = ov;
    if (tid == 0) {
      part[64] = mm;
      part[65] = ll;
    }
  }
}

__global__ void __launch_bounds__(64) xattn_f32_combine_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                               float* __restrict__ out_hi, float* __restrict__ out_lo,
                                                               int T, int H, int splits, int ldo,
                                                               const int* __restrict__ cu, int npsg) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, bt = blockIdx.y, d = threadIdx.x;
  const int b = bt / T;
  int ns = splits;
  if (cu) ns = (cu[(b + 1) * npsg] - cu[b * npsg] + XF_CH - 1) / XF_CH;
  const float* pp = part + (static_cast<size_t>(bt) * H + h) * splits * 66;
  float m = -INFINITY;
  for (int s = 0; s < ns; ++s) m = fmaxf(m, pp[s * 66 + 64]);
  float l = 0.f, o = 0.f;
  for (int s = 0; s < ns; ++s) {
    const float w = expf(pp[s * 66 + 64] - m);
    l += w * pp[s * 66 + 65];
    o += w * pp[s * 66 + d];
  }
  const float v = o / l;
  const size_t off = static_cast<size_t>(bt) * ldo + h * 64 + d;
  if (out) out[off] = v;
  if (out_hi) {
    const float hi = tf32_trunc(v);
    out_hi[off] = hi;
    out_lo[off] = tf32_trunc(v - hi);
  }
}

// ------------------------------------------------------------------------------------------------ RMSNorm (+ residual)
__global__ void __launch_bounds__(256) rmsnorm_f32_kernel(const float* x, const float* __restrict__ delta, float* x_out,
                                                           const float* __restrict__ w, float* __restrict__ out,
                                                           float* __restrict__ out_hi, float* __restrict__ out_lo,
                                                           int rows, int d, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<size_t>(row) * d;
  const float* dr = delta ? delta + static_cast<size_t>(row) * d : nullptr;
  float* xo = x_out ? x_out + static_cast<size_t>(row) * d : nullptr;
  float ss = 0.f;
  for (int c = lane * 4; c < d; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + c);
    if (dr) {
      const float4 a = *reinterpret_cast<const float4*>(dr + c);
      v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
    }
    if (xo && (dr || xo != xr)) *reinterpret_cast<float4*>(xo + c) = v;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / static_cast<float>(d) + eps);
  const bool wrote = xo && (dr || xo != xr);
  const float* src = wrote ? xo : xr;  // each lane re-reads exactly the elements it wrote itself
  for (int c = lane * 4; c < d; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(src + c);
    if (dr && !wrote) {
      const float4 a = *reinterpret_cast<const float4*>(dr + c);
      v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
    }
    const float4 wv = *reinterpret_cast<const float4*>(w + c);
    const float4 y = make_float4(wv.x * (v.x * r), wv.y * (v.y * r), wv.z * (v.z * r), wv.w * (v.w * r));
    const size_t off = static_cast<size_t>(row) * d + c;
    if (out) *reinterpret_cast<float4*>(out + off) = y;
    if (out_hi) {
      const float4 hi = make_float4(tf32_trunc(y.x), tf32_trunc(y.y), tf32_trunc(y.z), tf32_trunc(y.w));
      *reinterpret_cast<float4*>(out_hi + off) = hi;
      *reinterpret_cast<float4*>(out_lo + off) =
          make_float4(tf32_trunc(y.x - hi.x), tf32_trunc(y.y - hi.y), tf32_trunc(y.z - hi.z), tf32_trunc(y.w - hi.w));
    }
  }
}

}  // namespace

int enc_attn_f32(const float* qkv, const unsigned char* mask, const float* lut, float* out, float* out_hi, float* out_lo,
                 int P, int L, int H, cudaStream_t s, const int* cu) {
  LAKO_REQUIRE(P > 0 && H > 0 && L > 0 && L <= 4096 && qkv && lut && (out || out_hi), LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), LAKO_ERR_BAD_ARG);
  EAF32Params p{qkv, mask, lut, out, out_hi, out_lo, L, H, H * 64, cu};
  const size_t smem = sizeof(float) * (3 * 64 * EF_P + 64 * 64 + 64 + 2 * L);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(enc_attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return LAKO_ERR_LAUNCH;
    }
    configured = smem;
  }
  ProfScope prof(PROF_ENC_ATTN_FWD, 4.0 * P * H * static_cast<double>(L) * L * 64, s);
  launch_pdl(enc_attn_f32_kernel, dim3((L + EF_T - 1) / EF_T, H, P), dim3(256), smem, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

size_t xattn_f32_workspace_bytes(int B, int T, int H, int M) {
  return static_cast<size_t>(B) * T * H * ((M + XF_CH - 1) / XF_CH) * 66 * sizeof(float);
}

int xattn_f32(const float* q, const float* kv, const unsigned char* mask, float* out, float* out_hi, float* out_lo,
              float* tok_score, const int* step, float* workspace, int B, int T, int M, int H, int ldq, int ldkv, int koff,
              int voff, int ldo, cudaStream_t s, const int* cu, int npsg, const int* tok_index) {
  LAKO_REQUIRE(B > 0 && T > 0 && M > 0 && H > 0 && q && kv && workspace && (out || out_hi), LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(ldq % 4 == 0 && ldkv % 4 == 0 && koff % 4 == 0 && voff % 4 == 0, LAKO_ERR_BAD_ARG);
  XF32Params p{q, kv, mask, workspace, tok_score, step, cu, tok_index, T, M, H, ldq, ldkv, koff, voff,
               (M + XF_CH - 1) / XF_CH, npsg};
  ProfScope prof(PROF_XATTN, 4.0 * B * H * static_cast<double>(T) * M * 64, s);
  launch_pdl(xattn_f32_kernel, dim3(p.splits, H, B * T), dim3(128), 0, s, p);
  LAKO_CHECK_LAUNCH();
  launch_pdl(xattn_f32_combine_kernel, dim3(H, B * T), dim3(64), 0, s, workspace, out, out_hi, out_lo, T, H, p.splits, ldo,
             cu, npsg);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

int rmsnorm_f32(const float* x, const float* delta, float* x_out, const float* w, float* out, float* out_hi, float* out_lo,
                int rows, int d, float eps, cudaStream_t s) {
  LAKO_REQUIRE(rows > 0 && d > 0 && d % 4 == 0 && x && w, LAKO_ERR_BAD_ARG);
  ProfScope prof(PROF_NORM, static_cast<double>(rows) * d * (delta ? 20.0 : 12.0), s);
  launch_pdl(rmsnorm_f32_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, delta, x_out, w, out, out_hi, out_lo, rows, d, eps);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
