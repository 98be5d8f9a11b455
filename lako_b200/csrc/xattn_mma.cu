// Cross-attention over the fused FiD memory on the legacy-path tensor cores (mma.sync m16n8k16, bf16):
// the op is HBM-bound (it streams K/V of N*L memory tokens for T ~ 10 queries), so the MMA shape that
// fits is the small warp-level one; tcgen05's 128-row tiles would be >90 % padding here.
//
// Reference: cross_attention_forward /root/reference/src/model.py:286-349.
// One CTA = 128 keys of one (question, head); warp w owns keys [32w, 32w+32).
//   forward : S = Q K^T (+mask) -> chunk-local softmax -> O_partial = P V, partial (o, m, l) per chunk;
//             xattn_combine_kernel (dec_attn.cu) merges the chunks.
//   backward: S^T, dP^T per 16-key tile, P^T = exp(S^T - lse), dS^T = P^T (dP^T - D);
//             dV = P^T dO, dK = dS^T Q written in place (bf16); dq = dS K of the chunk via shared memory, left as a
//             per-chunk partial that xattn_dq_reduce_kernel (dec_attn.cu) sums in chunk order (the model's path), or
//             added to a pre-zeroed fp32 buffer with atomics (the stand-alone entry point).
#include "mma_sync.cuh"
#include "ops.cuh"

namespace lako {

constexpr float XM_MASK_MIN = -3.4028234663852886e38f;
constexpr float XM_LOG2E = 1.4426950408889634f;
constexpr int XM_CH = 128;

struct XMParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* kv;
  const unsigned char* mask;
  float* part;       // fwd: [B][H][splits][T][66]
  float* tok_score;  // fwd: [B][M] (may be null)
  const __nv_bfloat16* out;   // bwd
  const __nv_bfloat16* dout;  // bwd
  const float* lse;           // bwd [B][H][T]
  __nv_bfloat16* dkv;         // bwd
  float* dq;                  // bwd [B*T][H*64] fp32 (pre-zeroed), accumulated with atomics -- or, when dq_part is set:
  float* dq_part;             // bwd [B][H][splits][T][64] fp32 per-chunk partials (summed in chunk order by xattn_dq_reduce)
  int T, TP, M, H, ldq, ldkv, koff, voff, ldo, splits;
  // packed (un-padded) memory: question b owns kv rows [cu[b*npsg], cu[(b+1)*npsg]); null = padded [B][M]
  const int* cu;
  int npsg;
  const int* tok_index;  // packed row -> index into the padded [B][M] score buffer (packed layout only)
  DropSpec drop;  // attention-probability dropout; group = ((b*H+h)*T + t) * ceil(M/8) + key/8
};

// dynamic smem: sK 16K | sV 16K | sQ TP*128 | sX TP*128 (bwd: dO) | f32 scratch
// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(128) xattn_fwd_mma_kernel(const XMParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t xs[];
  uint8_t* sK = xs;
  uint8_t* sV = xs + 16384;
  uint8_t* sQ = xs + 32768;
  float* sMask = reinterpret_cast<float*>(sQ + p.TP * 128);  // [128]
  float* sM = sMask + 128;                                   // [4][16]
  float* sL = sM + 64;                                       // [4][16]
  // [4][16][68] partial outputs: with a single 16-query tile (T <= 16, the reader's case) K is dead by the time the
  // partials are exchanged, so they alias the K tile and the CTA needs 35 KB instead of 52 KB (6 CTAs/SM, not 4)
  float* sO = (p.TP == 16) ? reinterpret_cast<float*>(sK) : sL + 64;
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int k0 = split * XM_CH;
  int kstart = b * p.M, Mb = p.M;
  if (p.cu) {
    kstart = p.cu[b * p.npsg];
    Mb = p.cu[(b + 1) * p.npsg] - kstart;
  }
  if (k0 >= Mb) return;  // chunk beyond this question's memory (CTA-uniform)
  const int nk = min(XM_CH, Mb - k0);
  const size_t kvrow0 = static_cast<size_t>(kstart) + k0;
  stage_sw(sK, p.kv + kvrow0 * p.ldkv + p.koff + h * 64, p.ldkv, 128, nk, tid, 128);
  stage_sw(sV, p.kv + kvrow0 * p.ldkv + p.voff + h * 64, p.ldkv, 128, nk, tid, 128);
  stage_sw(sQ, p.q + static_cast<size_t>(b) * p.T * p.ldq + h * 64, p.ldq, p.TP, p.T, tid, 128);
  {
    const int j = tid;
    const bool valid = j >= nk || (p.mask == nullptr) || p.mask[static_cast<size_t>(b) * p.M + k0 + j];
    sMask[j] = (j < nk) ? (valid ? 0.f : XM_MASK_MIN) : -INFINITY;
  }
  cp_async_wait_all();
  __syncthreads();
  const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ);
  const int kb = warp * 32;  // this warp's first key within the chunk
  float* part = p.part + ((static_cast<size_t>(b) * p.H + h) * p.splits + split) * p.T * 66;

  for (int mt = 0; mt < p.TP / 16; ++mt) {
    // ---- S = Q K^T for 16 queries x 32 keys
    float sacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t qa[4];
      ldsm_x4(qa, aQ + swz(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t kb4[4];
        ldsm_x4(kb4, aK + swz(kb + np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, ks * 2 + ((lane >> 3) & 1)));
        mma16816(sacc[np * 2], qa, kb4[0], kb4[1]);
        mma16816(sacc[np * 2 + 1], qa, kb4[2], kb4[3]);
      }
    }
    // ---- mask, optional step-0 score capture, chunk-local softmax
    float mrow[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int key = kb + nt * 8 + t4 * 2;
      const float m0 = sMask[key], m1 = sMask[key + 1];
      if (p.tok_score && mt == 0 && g == 0) {  // query 0 lives in row g == 0
        if (p.tok_index) {
          if (m0 == 0.f) atomicAdd(p.tok_score + p.tok_index[kvrow0 + key], sacc[nt][0]);
          if (m1 == 0.f) atomicAdd(p.tok_score + p.tok_index[kvrow0 + key + 1], sacc[nt][1]);
        } else {
          float* ts = p.tok_score + static_cast<size_t>(b) * p.M + k0 + key;
          if (m0 == 0.f) atomicAdd(ts, sacc[nt][0]);
          if (m1 == 0.f) atomicAdd(ts + 1, sacc[nt][1]);
        }
      }
      sacc[nt][0] += m0;
      sacc[nt][1] += m1;
      sacc[nt][2] += m0;
      sacc[nt][3] += m1;
      mrow[0] = fmaxf(mrow[0], fmaxf(sacc[nt][0], sacc[nt][1]));
      mrow[1] = fmaxf(mrow[1], fmaxf(sacc[nt][2], sacc[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mrow[r] = fmaxf(mrow[r], __shfl_xor_sync(0xffffffffu, mrow[r], 1));
      mrow[r] = fmaxf(mrow[r], __shfl_xor_sync(0xffffffffu, mrow[r], 2));
    }
    float lrow[2] = {0.f, 0.f};
    uint32_t pa[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float e[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float mm = mrow[i >> 1];
        e[i] = (mm == -INFINITY) ? 0.f : fast_exp2((sacc[nt][i] - mm) * XM_LOG2E);
      }
      lrow[0] += e[0] + e[1];
      lrow[1] += e[2] + e[3];
      if (p.drop.thresh) {
        // one Philox call covers the 8 keys of this n-tile for one query row; the quad (t4 = 0..3) shares it:
        // lane t4 == 0 draws row g, lane t4 == 1 draws row g + 8
        const unsigned long long G = (p.M + 7) >> 3, kg = (k0 + kb + nt * 8) >> 3;
        uint32_t mine = 0xffu;
        if (t4 < 2) {
          const int t = mt * 16 + g + t4 * 8;
          if (t < p.T) mine = drop_keep8(p.drop, (static_cast<unsigned long long>(b * p.H + h) * p.T + t) * G + kg);
        }
        const uint32_t kr0 = __shfl_sync(0xffffffffu, mine, lane & ~3);
        const uint32_t kr1 = __shfl_sync(0xffffffffu, mine, (lane & ~3) + 1);
        e[0] = ((kr0 >> (t4 * 2)) & 1u) ? e[0] * p.drop.scale : 0.f;
        e[1] = ((kr0 >> (t4 * 2 + 1)) & 1u) ? e[1] * p.drop.scale : 0.f;
        e[2] = ((kr1 >> (t4 * 2)) & 1u) ? e[2] * p.drop.scale : 0.f;
        e[3] = ((kr1 >> (t4 * 2 + 1)) & 1u) ? e[3] * p.drop.scale : 0.f;
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(e[0], e[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(e[2], e[3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
      lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
    }
    // ---- O_w = P V (16 x 64) over this warp's 32 keys
    float oacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        ldsm_x4_t(vb, aV + swz(kb + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
        mma16816(oacc[dp * 2], pa[kk], vb[0], vb[1]);
        mma16816(oacc[dp * 2 + 1], pa[kk], vb[2], vb[3]);
      }
    }
    // ---- combine the 4 warps of this chunk
    if (t4 == 0) {
      sM[warp * 16 + g] = mrow[0];
      sM[warp * 16 + g + 8] = mrow[1];
      sL[warp * 16 + g] = lrow[0];
      sL[warp * 16 + g + 8] = lrow[1];
    }
    __syncthreads();
    float mg[2], sc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = g + r * 8;
      mg[r] = fmaxf(fmaxf(sM[row], sM[16 + row]), fmaxf(sM[32 + row], sM[48 + row]));
      sc[r] = (mrow[r] == -INFINITY) ? 0.f : fast_exp2((mrow[r] - mg[r]) * XM_LOG2E);
    }
    float* so = sO + warp * 16 * 68;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<float2*>(so + g * 68 + nt * 8 + t4 * 2) = make_float2(oacc[nt][0] * sc[0], oacc[nt][1] * sc[0]);
      *reinterpret_cast<float2*>(so + (g + 8) * 68 + nt * 8 + t4 * 2) =
          make_float2(oacc[nt][2] * sc[1], oacc[nt][3] * sc[1]);
    }
    __syncthreads();
    for (int i = tid; i < 16 * 64; i += 128) {
      const int row = i >> 6, d = i & 63;
      const int t = mt * 16 + row;
      if (t < p.T)
        part[t * 66 + d] = sO[row * 68 + d] + sO[(16 + row) * 68 + d] + sO[(32 + row) * 68 + d] + sO[(48 + row) * 68 + d];
    }
    if (tid < 16) {
      const int t = mt * 16 + tid;
      if (t < p.T) {
        const float m = fmaxf(fmaxf(sM[tid], sM[16 + tid]), fmaxf(sM[32 + tid], sM[48 + tid]));
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float mw = sM[w * 16 + tid];
          l += (mw == -INFINITY) ? 0.f : sL[w * 16 + tid] * fast_exp2((mw - m) * XM_LOG2E);
        }
        part[t * 66 + 64] = m;
        part[t * 66 + 65] = l;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ backward
// TPC = decoder length padded to 16 (compile time: the query-side register tiles scale with it; the training
// shapes use 16, which keeps the kernel near 100 registers and five CTAs per SM).
template <int TPC>
__global__ void __launch_bounds__(128, TPC == 16 ? 5 : 3) xattn_bwd_mma_kernel(const XMParams p) {
  constexpr int NTT = TPC / 8;   // n-tiles over queries (<= 8)
  constexpr int KST = TPC / 16;  // k-steps over queries (<= 4)
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t xs[];
  uint8_t* sK = xs;
  uint8_t* sV = xs + 16384;
  uint8_t* sQ = xs + 32768;
  uint8_t* sDO = sQ + TPC * 128;
  float* sMask = reinterpret_cast<float*>(sDO + TPC * 128);  // [128]
  float* sLse = sMask + 128;                                  // [64]
  float* sD = sLse + 64;                                      // [64]
  float* sO = (TPC == 16) ? reinterpret_cast<float*>(sV) : sD + 64;  // [4][16][68] dq partials (alias: V is dead then)
  uint8_t* sDS = reinterpret_cast<uint8_t*>(sD + 64 + ((TPC == 16) ? 0 : 4 * 16 * 68));  // [128 keys][TP*2+16 B]
  const int dsp = TPC * 2 + 16;
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int k0 = split * XM_CH;
  int kstart = b * p.M, Mb = p.M;
  if (p.cu) {
    kstart = p.cu[b * p.npsg];
    Mb = p.cu[(b + 1) * p.npsg] - kstart;
  }
  if (k0 >= Mb) return;
  const int nk = min(XM_CH, Mb - k0);
  const size_t kvrow0 = static_cast<size_t>(kstart) + k0;
  {
    // K and V chunk tiles: thread (r0 = tid / 8, c = tid % 8) copies 16 B column c of rows r0, r0 + 16, ...; the
    // swizzle term (r & 7) does not change along that walk, so the loop is two adds per copy
    const int r0 = tid >> 3, c = tid & 7;
    const __nv_bfloat16* gk = p.kv + (kvrow0 + r0) * p.ldkv + p.koff + h * 64 + c * 8;
    const ptrdiff_t vdelta = static_cast<ptrdiff_t>(p.voff) - p.koff;
    const size_t gstep = static_cast<size_t>(16) * p.ldkv;
    const uint32_t so = swz(r0, c);
    const uint32_t bK = smem_u32(sK) + so, bV = smem_u32(sV) + so;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = r0 + j * 16 < nk;
      const __nv_bfloat16* src = ok ? gk + j * gstep : p.kv;
      cp_async16(bK + j * 2048, src, ok);
      cp_async16(bV + j * 2048, ok ? src + vdelta : p.kv, ok);
    }
  }
  stage_sw(sQ, p.q + static_cast<size_t>(b) * p.T * p.ldq + h * 64, p.ldq, TPC, p.T, tid, 128);
  stage_sw(sDO, p.dout + static_cast<size_t>(b) * p.T * p.ldo + h * 64, p.ldo, TPC, p.T, tid, 128);
  {
    const bool valid = tid >= nk || (p.mask == nullptr) || p.mask[static_cast<size_t>(b) * p.M + k0 + tid];
    sMask[tid] = (tid < nk) ? (valid ? 0.f : XM_MASK_MIN) : -INFINITY;
  }
#pragma unroll
  for (int t = warp; t < TPC; t += 4) {  // D[t] = sum_d dO*O ; lse = +inf beyond T -> P = 0
    float acc = 0.f;
    if (t < p.T) {
      const size_t row = static_cast<size_t>(b) * p.T + t;
      const __nv_bfloat16* o = p.out + row * p.ldo + h * 64;
      const __nv_bfloat16* d = p.dout + row * p.ldo + h * 64;
      acc = __bfloat162float(o[lane]) * __bfloat162float(d[lane]) +
            __bfloat162float(o[lane + 32]) * __bfloat162float(d[lane + 32]);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      sD[t] = acc;
      sLse[t] = (t < p.T) ? p.lse[(static_cast<size_t>(b) * p.H + h) * p.T + t] : INFINITY;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aDO = smem_u32(sDO), aDS = smem_u32(sDS);
  const int kb = warp * 32;

#pragma unroll 1
  for (int km = 0; km < 2; ++km) {  // 16-key tiles of this warp
    const int key0 = kb + km * 16;
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) st[i][e] = dpt[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ka[4], va[4];
      const uint32_t arow = swz(key0 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4));
      ldsm_x4(ka, aK + arow);
      ldsm_x4(va, aV + arow);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np * 2 < NTT) {
          uint32_t qb[4], db[4];
          const uint32_t brow = swz(np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, ks * 2 + ((lane >> 3) & 1));
          ldsm_x4(qb, aQ + brow);
          ldsm_x4(db, aDO + brow);
          mma16816(st[np * 2], ka, qb[0], qb[1]);
          mma16816(st[np * 2 + 1], ka, qb[2], qb[3]);
          mma16816(dpt[np * 2], va, db[0], db[1]);
          mma16816(dpt[np * 2 + 1], va, db[2], db[3]);
        }
      }
    }
    // P^T, dS^T (rows: keys g / g+8 of this tile; cols: queries)
    const float mk0 = sMask[key0 + g], mk1 = sMask[key0 + g + 8];
    uint32_t pa[4][4], dsa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < NTT) {
        const int t = nt * 8 + t4 * 2;
        const float l0 = sLse[t], l1 = sLse[t + 1], d0 = sD[t], d1 = sD[t + 1];
        float p00 = fast_exp2(((st[nt][0] + mk0) - l0) * XM_LOG2E);
        float p01 = fast_exp2(((st[nt][1] + mk0) - l1) * XM_LOG2E);
        float p10 = fast_exp2(((st[nt][2] + mk1) - l0) * XM_LOG2E);
        float p11 = fast_exp2(((st[nt][3] + mk1) - l1) * XM_LOG2E);
        float e00 = dpt[nt][0], e01 = dpt[nt][1], e10 = dpt[nt][2], e11 = dpt[nt][3];
        float q00 = p00, q01 = p01, q10 = p10, q11 = p11;  // dropped copies feed dV = P_drop^T dO
        if (p.drop.thresh) {
          // masks of (query t | t+1) x (key group kg | kg+1): four Philox calls shared by the 8 lanes that hold
          // the same query columns (same t4); lane g = 0..3 draws one each
          const unsigned long long G = (p.M + 7) >> 3, kg = (k0 + key0) >> 3, rb = static_cast<unsigned long long>(b * p.H + h) * p.T;
          const float sc = p.drop.scale;
          uint32_t mine = 0xffu;
          if (g < 4) {
            const int tt = t + (g >> 1);
            if (tt < p.T) mine = drop_keep8(p.drop, (rb + tt) * G + kg + (g & 1));
          }
          const uint32_t m00 = __shfl_sync(0xffffffffu, mine, 0 * 4 + t4), m01 = __shfl_sync(0xffffffffu, mine, 1 * 4 + t4);
          const uint32_t m10 = __shfl_sync(0xffffffffu, mine, 2 * 4 + t4), m11 = __shfl_sync(0xffffffffu, mine, 3 * 4 + t4);
          const bool a00 = (m00 >> g) & 1u, a10 = (m01 >> g) & 1u;  // query t  : keys g, g+8
          const bool a01 = (m10 >> g) & 1u, a11 = (m11 >> g) & 1u;  // query t+1: keys g, g+8
          e00 = a00 ? e00 * sc : 0.f; q00 = a00 ? q00 * sc : 0.f;
          e10 = a10 ? e10 * sc : 0.f; q10 = a10 ? q10 * sc : 0.f;
          e01 = a01 ? e01 * sc : 0.f; q01 = a01 ? q01 * sc : 0.f;
          e11 = a11 ? e11 * sc : 0.f; q11 = a11 ? q11 * sc : 0.f;
        }
        const float s00 = p00 * (e00 - d0), s01 = p01 * (e01 - d1);
        const float s10 = p10 * (e10 - d0), s11 = p11 * (e11 - d1);
        p00 = q00; p01 = q01; p10 = q10; p11 = q11;
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p00, p01);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p10, p11);
        const uint32_t ds0 = pack_bf16x2(s00, s01), ds1 = pack_bf16x2(s10, s11);
        dsa[nt >> 1][(nt & 1) * 2 + 0] = ds0;
        dsa[nt >> 1][(nt & 1) * 2 + 1] = ds1;
        *reinterpret_cast<uint32_t*>(sDS + (key0 + g) * dsp + t * 2) = ds0;
        *reinterpret_cast<uint32_t*>(sDS + (key0 + g + 8) * dsp + t * 2) = ds1;
      }
    }
    // dV = P^T dO, dK = dS^T Q   (16 keys x 64)
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) dv[i][e] = dk[i][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk < KST) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t ob[4], qb[4];
          const uint32_t brow = swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4));
          ldsm_x4_t(ob, aDO + brow);
          ldsm_x4_t(qb, aQ + brow);
          mma16816(dv[dp * 2], pa[kk], ob[0], ob[1]);
          mma16816(dv[dp * 2 + 1], pa[kk], ob[2], ob[3]);
          mma16816(dk[dp * 2], dsa[kk], qb[0], qb[1]);
          mma16816(dk[dp * 2 + 1], dsa[kk], qb[2], qb[3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = key0 + g + r * 8;
      if (key < nk) {
        __nv_bfloat16* dkrow = p.dkv + (kvrow0 + key) * p.ldkv + p.koff + h * 64;
        __nv_bfloat16* dvrow = p.dkv + (kvrow0 + key) * p.ldkv + p.voff + h * 64;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(dkrow + nt * 8 + t4 * 2) = pack_bf16x2(dk[nt][r * 2], dk[nt][r * 2 + 1]);
          *reinterpret_cast<uint32_t*>(dvrow + nt * 8 + t4 * 2) = pack_bf16x2(dv[nt][r * 2], dv[nt][r * 2 + 1]);
        }
      }
    }
  }
  __syncthreads();  // every warp is done with V / Q / dO: the dq partial exchange may alias them
  // dq[t][d] += sum over this warp's 32 keys of dS[t][key] K[key][d]; A = dS from the [key][t] smem copy (.trans)
  for (int mt = 0; mt < KST; ++mt) {
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[i][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t a[4];
      {
        const int j = lane >> 3;
        const int row = kb + kk * 16 + (lane & 7) + ((j >> 1) & 1) * 8;  // key (k index)
        const int col = mt * 16 + (j & 1) * 8;                           // query (m index)
        ldsm_x4_t(a, aDS + row * dsp + col * 2);
      }
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t kb4[4];
        ldsm_x4_t(kb4, aK + swz(kb + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
        mma16816(dq[dp * 2], a, kb4[0], kb4[1]);
        mma16816(dq[dp * 2 + 1], a, kb4[2], kb4[3]);
      }
    }
    float* so = sO + warp * 16 * 68;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<float2*>(so + g * 68 + nt * 8 + t4 * 2) = make_float2(dq[nt][0], dq[nt][1]);
      *reinterpret_cast<float2*>(so + (g + 8) * 68 + nt * 8 + t4 * 2) = make_float2(dq[nt][2], dq[nt][3]);
    }
    __syncthreads();
    for (int i = tid; i < 16 * 64; i += 128) {
      const int row = i >> 6, d = i & 63;
      const int t = mt * 16 + row;
      if (t < p.T) {
        const float v = sO[row * 68 + d] + sO[(16 + row) * 68 + d] + sO[(32 + row) * 68 + d] + sO[(48 + row) * 68 + d];
        if (p.dq_part)
          p.dq_part[(((static_cast<size_t>(b) * p.H + h) * p.splits + split) * p.T + t) * 64 + d] = v;
        else
          atomicAdd(&p.dq[(static_cast<size_t>(b) * p.T + t) * (p.H * 64) + h * 64 + d], v);
      }
    }
    __syncthreads();
  }
}

static size_t xm_smem_bytes(int TP, bool bwd) {
  size_t n = 32768 + static_cast<size_t>(TP) * 128;
  const size_t scratch = (TP == 16) ? 0 : 4 * 16 * 68;  // partial-output exchange aliases a dead K / V tile when T <= 16
  if (!bwd) return n + (128 + 64 + 64 + scratch) * sizeof(float);
  n += static_cast<size_t>(TP) * 128;
  n += (128 + 64 + 64 + scratch) * sizeof(float);
  n += 128 * (static_cast<size_t>(TP) * 2 + 16);
  return n;
}

int xattn_fwd_mma(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, float* tok_score,
                  float* workspace, int B, int T, int M, int H, int ldq, int ldkv, int koff, int voff,
                  cudaStream_t s, DropSpec drop, const int* cu, int npsg, const int* tok_index) {
  LAKO_REQUIRE(B > 0 && T > 0 && T <= 64 && M > 0 && H > 0, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && koff % 8 == 0 && voff % 8 == 0, LAKO_ERR_BAD_ARG);
  XMParams p{};
  p.q = q; p.kv = kv; p.mask = mask; p.part = workspace; p.tok_score = tok_score;
  p.T = T; p.TP = (T + 15) / 16 * 16; p.M = M; p.H = H; p.ldq = ldq; p.ldkv = ldkv; p.koff = koff; p.voff = voff;
  p.splits = (M + XM_CH - 1) / XM_CH;
  p.drop = drop;
  p.cu = cu;
  p.npsg = npsg;
  p.tok_index = tok_index;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(xattn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xm_smem_bytes(64, false));
    configured = true;
  }
  launch_pdl(xattn_fwd_mma_kernel, dim3(dim3(p.splits, H, B)), dim3(128), xm_smem_bytes(p.TP, false), s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

int xattn_bwd_mma(const __nv_bfloat16* q, const __nv_bfloat16* kv, const unsigned char* mask, const __nv_bfloat16* out,
                  const __nv_bfloat16* dout, const float* lse, __nv_bfloat16* dkv, float* dq_f32, int B, int T, int M,
                  int H, int ldq, int ldkv, int koff, int voff, int ldo, cudaStream_t s, DropSpec drop, const int* cu,
                  int npsg, float* dq_part) {
  LAKO_REQUIRE(B > 0 && T > 0 && T <= 64 && M > 0 && H > 0, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && koff % 8 == 0 && voff % 8 == 0 && ldo % 8 == 0, LAKO_ERR_BAD_ARG);
  XMParams p{};
  p.q = q; p.kv = kv; p.mask = mask; p.out = out; p.dout = dout; p.lse = lse; p.dkv = dkv; p.dq = dq_f32;
  p.dq_part = dq_part;
  p.T = T; p.TP = (T + 15) / 16 * 16; p.M = M; p.H = H; p.ldq = ldq; p.ldkv = ldkv; p.koff = koff; p.voff = voff;
  p.ldo = ldo;
  p.splits = (M + XM_CH - 1) / XM_CH;
  p.drop = drop;
  p.cu = cu;
  p.npsg = npsg;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(xattn_bwd_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xm_smem_bytes(16, true));
    cudaFuncSetAttribute(xattn_bwd_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xm_smem_bytes(32, true));
    cudaFuncSetAttribute(xattn_bwd_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xm_smem_bytes(64, true));
    configured = true;
  }
  const dim3 grid(p.splits, H, B);
  if (p.TP == 16) {
    launch_pdl(xattn_bwd_mma_kernel<16>, grid, dim3(128), xm_smem_bytes(16, true), s, p);
  } else if (p.TP == 32) {
    launch_pdl(xattn_bwd_mma_kernel<32>, grid, dim3(128), xm_smem_bytes(32, true), s, p);
  } else {
    p.TP = 64;
    launch_pdl(xattn_bwd_mma_kernel<64>, grid, dim3(128), xm_smem_bytes(64, true), s, p);
  }
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako
