// bf16 GEMM on the sm_100a tensor cores: C[M,N] = alpha * A[M,K] * B[N,K]^T (+ epilogue).
//
// Replaces the cuBLAS calls behind every nn.Linear of the T5 blocks the reference inherits
// (HF:modeling_t5.py:92-103 FFN, :277-299/:338 q,k,v,o projections, :1107-1110 LM head) and their
// autograd backward (dgrad / wgrad).
//
// Design: persistent, warp-specialised CTA (one per SM):
//   warp 0   : TMA producer  (cp.async.bulk.tensor, 128B-swizzled tiles, mbarrier expect-tx)
//   warp 1   : MMA issuer    (one thread issues tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      (tcgen05.ld -> registers -> fused epilogue -> global)
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile i overlap the main loop of
// tile i+1.  Operands may be K-major (row-major [rows][K]) or MN-major (row-major [K][rows]), so
// forward (X W^T), dgrad (dY W) and wgrad (dY^T X) all run without materialising a transpose.
// Split-K (wgrad: small output, very deep K) reduces with fp32 vector red.global into the
// caller-zeroed gradient buffer.
#include <cstdlib>
#include "ops.cuh"

namespace lako {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 256;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int STAGING_BYTES = 4 * 4096;  // per epilogue warp: 32 rows x 64 bf16, 16B-unit XOR swizzle
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
  void* C;
  const void* aux;
  int M, N, K;
  int ldc, ld_aux;
  int tiles_m, tiles_n, splitk, kb_total, kb_per_split;
  float alpha;
  DropSpec drop;  // EPI_BF16_RELU only: dropout after the activation (HF:95), element index = row * ldc + col
  int c_head_major;  // bf16 TMA epilogue: C is stored as [N/64][plane_rows >= M][64] (one contiguous slab per 64-column
                     // head) through a 3D tensor map -- the layout the decode cross-attention streams
};

// Epilogue of one 128-row x BN-column accumulator tile for a warp owning TMEM lane quarter q: its 64-column
// groups [g_begin, g_end) (several warps may share a lane quarter), staged through 4 KB slot `slot` of sStage.
// tmC != nullptr: the staged 32 x 64 bf16 sub-tile leaves shared memory through one TMA store (which also clips the
// M / N edges) instead of 8 LDS + 8 STG per lane.
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t t_acc, uint8_t* sStage, int m0, int n0,
                                              int q, int lane, int g_begin = 0, int g_end = BN / 64, int slot = -1,
                                              const CUtensorMap* tmC = nullptr) {
  if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_RELU || EPI == EPI_BF16_MASK || EPI == EPI_BF16_BITMASK) {
    // bf16 output: TMEM (thread = row) -> registers -> per-warp swizzled smem -> coalesced 16 B stores
    // (each store instruction writes four full 128 B row segments)
    uint8_t* stg = sStage + (slot < 0 ? q : slot) * 4096;
    __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(p.C);
    const __nv_bfloat16* AUX = reinterpret_cast<const __nv_bfloat16*>(p.aux);
#pragma unroll 1
    for (int g = g_begin; g < g_end; ++g) {
      const int col0 = n0 + g * 64;
      if (col0 >= p.N) break;  // warp-uniform
      // ReLU-mask operand: fetch this lane's eight 16 B pieces up front so their latency overlaps the TMEM load
      uint4 auxv[8];
      if constexpr (EPI == EPI_BF16_MASK) {
        const int pcol = col0 + (lane & 7) * 8;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int prow = m0 + q * 32 + it * 4 + (lane >> 3);
          auxv[it] = make_uint4(0, 0, 0, 0);
          if (prow < p.M && pcol < p.N)
            auxv[it] = __ldg(reinterpret_cast<const uint4*>(AUX + static_cast<size_t>(prow) * p.ld_aux + pcol));
        }
      }
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(t_acc + g * 64, r0);
      tmem_ld_32x32(t_acc + g * 64 + 32, r1);
      if constexpr (EPI != EPI_BF16_MASK) {
        if (tmC) {  // the previous group's TMA store must have finished reading the staging slot
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
      }
      // 1 bit / element ReLU-activity mask, 8 bytes per thread-row per 64-column group: written by the ReLU
      // epilogue (aux = uint8 [M][ld_aux bytes]) and consumed by the dgrad epilogue instead of re-reading h
      uint2 bits = make_uint2(0u, 0u);
      if constexpr (EPI == EPI_BF16_BITMASK) {
        const int brow = m0 + q * 32 + lane;
        if (brow < p.M)
          bits = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(p.aux) +
                                                      static_cast<size_t>(brow) * p.ld_aux + (col0 >> 3)));
      }
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t* rr = (u < 4) ? r0 : r1;
        const int o = (u & 3) * 8;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = __uint_as_float(rr[o + i]) * p.alpha;
          if constexpr (EPI == EPI_BF16_RELU) v[i] = fmaxf(v[i], 0.f);
        }
        if constexpr (EPI == EPI_BF16_RELU) {
          if (p.drop.thresh) {
            const size_t e = static_cast<size_t>(m0 + q * 32 + lane) * p.ldc + col0 + u * 8;
            const uint32_t k = drop_keep8(p.drop, e >> 3);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ((k >> i) & 1u) ? v[i] * p.drop.scale : 0.f;
          }
          if (p.aux != nullptr) {
            uint32_t byte = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) byte |= (v[i] > 0.f ? 1u : 0u) << i;
            if (u < 4) bits.x |= byte << (8 * u);
            else bits.y |= byte << (8 * (u - 4));
          }
        }
        if constexpr (EPI == EPI_BF16_BITMASK) {
          const uint32_t byte = ((u < 4) ? (bits.x >> (8 * u)) : (bits.y >> (8 * (u - 4)))) & 0xffu;
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = ((byte >> i) & 1u) ? v[i] : 0.f;
        }
        uint4 w;
        w.x = pack_bf16x2(v[0], v[1]);
        w.y = pack_bf16x2(v[2], v[3]);
        w.z = pack_bf16x2(v[4], v[5]);
        w.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) = w;
      }
      if constexpr (EPI == EPI_BF16_RELU) {
        const int brow = m0 + q * 32 + lane;
        if (p.aux != nullptr && brow < p.M && col0 + 64 <= p.N)
          *reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(const_cast<void*>(p.aux)) +
                                    static_cast<size_t>(brow) * p.ld_aux + (col0 >> 3)) = bits;
      }
      if constexpr (EPI != EPI_BF16_MASK) {
        if (tmC) {
          fence_proxy_async_smem();  // generic-proxy writes of the slot -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            if (p.c_head_major)
              tma_store_3d(tmC, stg, 0, m0 + q * 32, col0 >> 6);
            else
              tma_store_2d(tmC, stg, col0, m0 + q * 32);
            tma_store_commit();
          }
          continue;
        }
      }
      __syncwarp();
      const int unit = lane & 7;
      const int col = col0 + unit * 8;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rl = it * 4 + (lane >> 3);
        const int row = m0 + q * 32 + rl;
        uint4 w = *reinterpret_cast<const uint4*>(stg + rl * 128 + ((unit ^ (rl & 7)) << 4));
        if (row < p.M && col < p.N) {
          if constexpr (EPI == EPI_BF16_MASK) {
            const uint4 a = auxv[it];
            // keep dY where the forward activation was > 0 (sign bit clear and non-zero)
            auto keep2 = [](uint32_t val, uint32_t act) -> uint32_t {
              const uint32_t lo = ((act & 0x7fffu) != 0 && (act & 0x8000u) == 0) ? 0x0000ffffu : 0u;
              const uint32_t hi = ((act & 0x7fff0000u) != 0 && (act & 0x80000000u) == 0) ? 0xffff0000u : 0u;
              return val & (lo | hi);
            };
            w.x = keep2(w.x, a.x);
            w.y = keep2(w.y, a.y);
            w.z = keep2(w.z, a.z);
            w.w = keep2(w.w, a.w);
          }
          *reinterpret_cast<uint4*>(C + static_cast<size_t>(row) * p.ldc + col) = w;
        }
      }
      __syncwarp();
    }
  } else {
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
#pragma unroll 1
    for (int c = g_begin * 2; c < g_end * 2; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= p.N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32(t_acc + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
      if constexpr (EPI == EPI_F32_ATOMIC) {
        if (tmC) {
          // Stage the 32 x 32 fp32 chunk (128 B rows, 16 B units XOR-swizzled like the TMA's 128B mode; rows past M
          // hold zeros and are clipped) and let ONE TMA reduce-add store accumulate it: whole-line updates by the
          // memory system instead of 256 16-byte red operations per warp.
          uint8_t* stg = sStage + (slot < 0 ? q : slot) * 4096;
          if (lane == 0) tma_store_wait_read<0>();  // the previous chunk's store has read the slot
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(tmC, stg, col0, m0 + q * 32);
            tma_store_commit();
          }
          continue;
        }
      }
      if (!row_ok) continue;
      const bool full = (col0 + 32 <= p.N);
      float* crow = reinterpret_cast<float*>(p.C) + static_cast<size_t>(row) * p.ldc + col0;
      if constexpr (EPI == EPI_F32) {
        if (full) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(crow + j * 4) = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.N) crow[i] = v[i];
        }
      } else {  // EPI_F32_ATOMIC
        if (full) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + j * 4), "f"(v[j * 4]),
                         "f"(v[j * 4 + 1]), "f"(v[j * 4 + 2]), "f"(v[j * 4 + 3])
                         : "memory");
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.N) atomicAdd(crow + i, v[i]);
        }
      }
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sStage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int total_tiles = tiles_mn * p.splitk;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ks = t / tiles_mn;
        const int mn = t - ks * tiles_mn;
        const int m0 = (mn / p.tiles_n) * BM;
        const int n0 = (mn % p.tiles_n) * BN;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* a = sA + stage * Cfg::A_BYTES;
          uint8_t* b = sB + stage * Cfg::B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a, &tmA, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(a + c * 8192, &tmA, &full_bar[stage], m0 + c * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(b, &tmB, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(b + c * 8192, &tmB, &full_bar[stage], n0 + c * 64, kb * BK);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ks = t / tiles_mn;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int ks = t / tiles_mn;
      const int mn = t - ks * tiles_mn;
      const int m0 = (mn / p.tiles_n) * BM;
      const int n0 = (mn % p.tiles_n) * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN, EPI>(p, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN, sStage, m0, n0, q, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant: a 2-CTA cluster computes a 256 x 256 tile with tcgen05.mma.cta_group::2 (UMMA_M = 256).
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N rows); the tensor cores
// of both SMs read the B halves from both CTAs' shared memory.  Per k-block the pair pulls 64 KB from L2
// for 256x256x64 MACs, vs 96 KB for two independent 128x256 CTAs: the single-CTA kernel saturates the
// L2->SM fabric (~6.3 KB/clk chip-wide) at ~50 % of the tensor peak, this one does not.
// Roles per CTA as above; only the leader CTA (cluster rank 0) issues MMAs and its commits are multicast
// to the mbarriers of both CTAs.
// ------------------------------------------------------------------------------------------------
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = 128 * BK * 2;       // this CTA's 128 rows of A
  static constexpr int B_BYTES = 128 * BK * 2;       // this CTA's half of the 256-row B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quarter, 128 of the 256 columns each: with K as short
                                       // as d_model the epilogue (not the MMA) bounds a tile when 4 warps drain it
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
  static constexpr int STAGING_BYTES = EPI_WARPS * 4096;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256;
};

// Body shared by the single-problem kernel and the batched one: `nbatch` problems of identical shape, problem z
// with its own tensor maps tmAs[z] / tmBs[z] and output Cs[z] (Cs == nullptr: one problem writing p.C).
template <bool A_MN, bool B_MN, int EPI>
__device__ __forceinline__ void gemm2_body(const CUtensorMap* tmAs, const CUtensorMap* tmBs, void* const* Cs, int nbatch,
                                           const GemmParams& p, const CUtensorMap* tmC = nullptr) {
  using Cfg = Gemm2Cfg;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sStage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmAs);
    tma_prefetch_desc(tmBs);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 2);   // one arrive.expect_tx per CTA of the pair (only the leader's copy is used)
      mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);   // multicast tcgen05.commit
      mbar_init(&tempty_bar[i], 2 * Cfg::EPI_WARPS);  // epilogue warps x 2 CTAs (only the leader's copy is used)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int tiles_mn = p.tiles_m * p.tiles_n;  // tiles_m counts 256-row tiles here
  const int per_problem = tiles_mn * p.splitk;
  const int total_tiles = per_problem * nbatch;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters) {
        const int z = t / per_problem;
        const int tz = t - z * per_problem;
        const CUtensorMap& tmA = tmAs[z];
        const CUtensorMap& tmB = tmBs[z];
        const int ks = tz / tiles_mn;
        const int mn = tz - ks * tiles_mn;
        const int m0 = (mn / p.tiles_n) * 256 + static_cast<int>(rank) * 128;
        const int n0 = (mn % p.tiles_n) * BN + static_cast<int>(rank) * 128;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t fb = mapa_u32(&full_bar[stage], 0);  // the leader's barrier
          mbar_arrive_expect_tx_cluster(fb, Cfg::STAGE_BYTES);
          uint8_t* a = sA + stage * Cfg::A_BYTES;
          uint8_t* b = sB + stage * Cfg::B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d_2sm(a, &tmA, fb, kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(a + c * 8192, &tmA, fb, m0 + c * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2sm(b, &tmB, fb, kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(b + c * 8192, &tmB, fb, n0 + c * 64, kb * BK);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters) {
        const int ks = (t % per_problem) / tiles_mn;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    GemmParams pz = p;
    for (int t = cluster_id; t < total_tiles; t += n_clusters) {
      const int z = t / per_problem;
      const int tz = t - z * per_problem;
      if (Cs) pz.C = Cs[z];
      const int mn = tz % tiles_mn;
      const int m0 = (mn / p.tiles_n) * 256 + static_cast<int>(rank) * 128;
      const int n0 = (mn % p.tiles_n) * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      constexpr int G_PER_WARP = (BN / 64) / (Cfg::EPI_WARPS / 4);
      const int part = (warp - 4) >> 2;
      epilogue_tile<BN, EPI>(pz, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN, sStage, m0, n0, q, lane,
                             part * G_PER_WARP, (part + 1) * G_PER_WARP, warp - 4, tmC ? tmC + z : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  if (tmC && warp >= 4 && lane == 0) tma_store_wait_all<0>();  // this warp's TMA stores are complete
  tc_fence_before();
  cluster_sync_all();  // the peer's smem / TMEM / barriers must outlive every remote access
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg::THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const GemmParams p, const int use_tmc) {
  constexpr bool kTmaOut = EPI == EPI_BF16 || EPI == EPI_BF16_RELU || EPI == EPI_BF16_BITMASK || EPI == EPI_F32_ATOMIC;
  gemm2_body<A_MN, B_MN, EPI>(&tmA, &tmB, nullptr, 1, p, (kTmaOut && use_tmc) ? &tmC : nullptr);
}

// Up to GEMM_MAX_BATCH same-shape problems in one launch (the decoder's per-layer weight gradients: each is a
// few-microsecond problem whose launch + pipeline fill costs more than its work).
constexpr int GEMM_MAX_BATCH = 24;
struct Gemm2Batch {
  CUtensorMap a[GEMM_MAX_BATCH];
  CUtensorMap b[GEMM_MAX_BATCH];
  CUtensorMap cmap[GEMM_MAX_BATCH];  // fp32 outputs as TMA reduce-add destinations
  void* c[GEMM_MAX_BATCH];
  int n;
};
template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg::THREADS, 1)
gemm2_batched_kernel(const __grid_constant__ Gemm2Batch bt, const GemmParams p) {
  gemm2_body<A_MN, B_MN, EPI>(bt.a, bt.b, bt.c, bt.n, p, EPI == EPI_F32_ATOMIC ? bt.cmap : nullptr);
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, int use_tmc,
                        const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg;
  auto kern = gemm2_bf16_kernel<A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return LAKO_ERR_LAUNCH;
    }
    configured = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.splitk;
  const int clusters = min(total, device_sm_count() / 2);
  launch_pdl(kern, dim3(clusters * 2), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, p, use_tmc);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

template <bool A_MN, bool B_MN>
static int dispatch_epi2(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, int use_tmc,
                         const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case EPI_BF16: return launch_gemm2<A_MN, B_MN, EPI_BF16>(tmA, tmB, tmC, use_tmc, p, s);
    case EPI_BF16_RELU: return launch_gemm2<A_MN, B_MN, EPI_BF16_RELU>(tmA, tmB, tmC, use_tmc, p, s);
    case EPI_F32: return launch_gemm2<A_MN, B_MN, EPI_F32>(tmA, tmB, tmC, use_tmc, p, s);
    case EPI_F32_ATOMIC: return launch_gemm2<A_MN, B_MN, EPI_F32_ATOMIC>(tmA, tmB, tmC, use_tmc, p, s);
    case EPI_BF16_MASK: return launch_gemm2<A_MN, B_MN, EPI_BF16_MASK>(tmA, tmB, tmC, use_tmc, p, s);
    case EPI_BF16_BITMASK: return launch_gemm2<A_MN, B_MN, EPI_BF16_BITMASK>(tmA, tmB, tmC, use_tmc, p, s);
  }
  set_last_error("unknown epilogue", __FILE__, __LINE__);
  return LAKO_ERR_BAD_ARG;
}

static int dispatch_major2(int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                           const CUtensorMap& tmC, int use_tmc, const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return dispatch_epi2<false, false>(epi, tmA, tmB, tmC, use_tmc, p, s);
  if (!a_mn && b_mn) return dispatch_epi2<false, true>(epi, tmA, tmB, tmC, use_tmc, p, s);
  if (a_mn && !b_mn) return dispatch_epi2<true, false>(epi, tmA, tmB, tmC, use_tmc, p, s);
  return dispatch_epi2<true, true>(epi, tmA, tmB, tmC, use_tmc, p, s);
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return LAKO_ERR_LAUNCH;
    }
    configured = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.splitk;
  const int grid = min(total, device_sm_count());
  launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                        cudaStream_t s) {
  switch (epi) {
    case EPI_BF16: return launch_gemm<BN, A_MN, B_MN, EPI_BF16>(tmA, tmB, p, s);
    case EPI_BF16_RELU: return launch_gemm<BN, A_MN, B_MN, EPI_BF16_RELU>(tmA, tmB, p, s);
    case EPI_F32: return launch_gemm<BN, A_MN, B_MN, EPI_F32>(tmA, tmB, p, s);
    case EPI_F32_ATOMIC: return launch_gemm<BN, A_MN, B_MN, EPI_F32_ATOMIC>(tmA, tmB, p, s);
    case EPI_BF16_MASK: return launch_gemm<BN, A_MN, B_MN, EPI_BF16_MASK>(tmA, tmB, p, s);
    case EPI_BF16_BITMASK: return launch_gemm<BN, A_MN, B_MN, EPI_BF16_BITMASK>(tmA, tmB, p, s);
  }
  set_last_error("unknown epilogue", __FILE__, __LINE__);
  return LAKO_ERR_BAD_ARG;
}

template <int BN>
static int dispatch_major(int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return dispatch_epi<BN, false, false>(epi, tmA, tmB, p, s);
  if (!a_mn && b_mn) return dispatch_epi<BN, false, true>(epi, tmA, tmB, p, s);
  if (a_mn && !b_mn) return dispatch_epi<BN, true, false>(epi, tmA, tmB, p, s);
  return dispatch_epi<BN, true, true>(epi, tmA, tmB, p, s);
}

// Host entry used by the C ABI and by the model orchestrator.
//  A: a_mn==0 -> row-major [M][lda>=K];   a_mn==1 -> row-major [K][lda>=M]
//  B: b_mn==0 -> row-major [N][ldb>=K];   b_mn==1 -> row-major [K][ldb>=N]
//  splitk<=0 -> chosen automatically (only for EPI_F32_ATOMIC); bn: 0 auto, 128 or 256.
int gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
              int b_mn, int epi, float alpha, const void* aux, int ld_aux, int splitk, int bn, cudaStream_t stream,
              DropSpec drop, int c_head_major) {
  int rc = check_arch();
  if (rc) return rc;
  LAKO_REQUIRE(M > 0 && N > 0 && K > 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(epi >= 0 && epi <= 5, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE((epi != EPI_BF16_MASK && epi != EPI_BF16_BITMASK) || aux != nullptr, LAKO_ERR_BAD_ARG);
  const bool out_bf16 = (epi == EPI_BF16 || epi == EPI_BF16_RELU || epi == EPI_BF16_MASK || epi == EPI_BF16_BITMASK);
  if (epi == EPI_BF16_BITMASK || (epi == EPI_BF16_RELU && aux != nullptr))
    LAKO_REQUIRE(N % 64 == 0 && ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 7) == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % (out_bf16 ? 8 : 4) == 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(!out_bf16 || N % 8 == 0, LAKO_ERR_UNSUPPORTED);  // bf16 epilogue stores 8-element (16 B) units
  if (epi == EPI_BF16_MASK)
    LAKO_REQUIRE((reinterpret_cast<uintptr_t>(aux) & 15) == 0 && ld_aux % 8 == 0, LAKO_ERR_BAD_ARG);

  const int sms = device_sm_count();
  // bn: 0 = auto, 128 / 256 = single-CTA kernel with that tile width, 512 = CTA-pair kernel (256x256 per cluster)
  if (c_head_major) {  // [N/64][M][64] output: only the CTA-pair kernel's TMA-store epilogue writes it
    LAKO_REQUIRE(epi == EPI_BF16 && N % 64 == 0, LAKO_ERR_UNSUPPORTED);
    bn = 512;
  }
  bool use_pair = (bn == 512);
  if (bn == 0 && M >= 512 && N >= 256) {
    // the pair kernel needs ~1/3 less L2->SM operand traffic per FLOP; use it whenever the problem fills the chip
    const long t_pair = static_cast<long>((M + 255) / 256) * ((N + 255) / 256);
    use_pair = (t_pair >= sms / 4) || epi == EPI_F32_ATOMIC;
  }
  if (use_pair) bn = 256;
  if (bn != 128 && bn != 256) {
    // prefer 128x256 tiles; fall back to 128x128 when that leaves most of the chip idle
    const long t256 = static_cast<long>((M + BM - 1) / BM) * ((N + 255) / 256);
    bn = (N <= 128 || (t256 < sms && epi != EPI_F32_ATOMIC)) ? 128 : 256;
  }
  GemmParams p;
  p.C = C;
  p.aux = aux;
  p.M = M;
  p.N = N;
  p.K = K;
  p.ldc = ldc;
  p.ld_aux = ld_aux;
  p.alpha = alpha;
  p.drop = drop;
  p.c_head_major = c_head_major;
  p.tiles_m = use_pair ? (M + 255) / 256 : (M + BM - 1) / BM;
  p.tiles_n = (N + bn - 1) / bn;
  p.kb_total = (K + BK - 1) / BK;
  if (epi != EPI_F32_ATOMIC) {
    splitk = 1;
  } else if (splitk <= 0) {
    // wgrad: few output tiles, very deep K.  Pick the split count whose work-item count best fills whole
    // waves of CTAs (or CTA pairs), preferring fewer splits (less atomic traffic) among near-equal fills.
    const int tiles = p.tiles_m * p.tiles_n;
    const int slots = use_pair ? sms / 2 : sms;
    const int max_split = max(1, min(p.kb_total / 8, 64));  // keep >= 8 k-blocks per split
    float best = -1.f;
    splitk = 1;
    for (int sk = 1; sk <= max_split; ++sk) {
      const long items = static_cast<long>(tiles) * sk;
      const long waves = (items + slots - 1) / slots;
      const float fill = static_cast<float>(items) / static_cast<float>(waves * slots);
      if (fill > best + 0.03f) {
        best = fill;
        splitk = sk;
      }
    }
  }
  splitk = max(1, min(splitk, p.kb_total));
  p.kb_per_split = (p.kb_total + splitk - 1) / splitk;
  p.splitk = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;

  CUtensorMap tmA, tmB;
  if (!a_mn)
    rc = get_tensor_map_2d(&tmA, A, K, M, lda, 64, BM);
  else
    rc = get_tensor_map_2d(&tmA, A, M, K, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn)
    rc = get_tensor_map_2d(&tmB, B, K, N, ldb, 64, use_pair ? 128 : bn);
  else
    rc = get_tensor_map_2d(&tmB, B, N, K, ldb, 64, BK);
  if (rc) return rc;
  const bool small = (M <= 128) || (a_mn && b_mn && K <= 128);
  const int pcat = small ? PROF_GEMM_SMALL : (a_mn ? PROF_GEMM_WGRAD : (b_mn ? PROF_GEMM_DGRAD : PROF_GEMM));
  ProfScope prof(pcat, 2.0 * M * N * static_cast<double>(K), stream);
  if (use_pair) {
    // bf16 results leave through TMA stores of 32 x 64 boxes (128B swizzle, the staging layout of the epilogue)
    CUtensorMap tmC = tmA;
    int use_tmc = 0;
    if (out_bf16 && epi != EPI_BF16_MASK) {
      static const bool tma_epi = [] {
        const char* e = getenv("LAKO_TMA_EPILOGUE");
        return !(e && e[0] == '0');
      }();
      if (c_head_major) {
        LAKO_REQUIRE(c_head_major >= M, LAKO_ERR_BAD_ARG);  // the value is the row capacity of one head's slab
        rc = get_tensor_map_3d(&tmC, C, 64, M, N / 64, 64, static_cast<uint64_t>(c_head_major) * 64, 64, 32, 1);
        if (rc) return rc;
        use_tmc = 1;
      } else if (tma_epi) {
        rc = get_tensor_map_2d(&tmC, C, N, M, ldc, 64, 32);
        if (rc) return rc;
        use_tmc = 1;
      }
    } else if (epi == EPI_F32_ATOMIC) {
      // gradient accumulation / split-K partial sums: TMA reduce-add stores of 32 x 32 fp32 boxes
      static const bool tma_red = [] {
        const char* e = getenv("LAKO_TMA_REDUCE");
        return !(e && e[0] == '0');
      }();
      if (tma_red) {
        rc = get_tensor_map_2d_f32(&tmC, C, N, M, ldc, 32, 32);
        if (rc) return rc;
        use_tmc = 1;
      }
    }
    return dispatch_major2(a_mn, b_mn, epi, tmA, tmB, tmC, use_tmc, p, stream);
  }
  if (bn == 256) return dispatch_major<256>(a_mn, b_mn, epi, tmA, tmB, p, stream);
  return dispatch_major<128>(a_mn, b_mn, epi, tmA, tmB, p, stream);
}

// dW_z[M][N] += alpha * A_z^T B_z for z < nbatch: A_z row-major [K][lda >= M], B_z row-major [K][ldb >= N] (both
// MN-major operands), C_z fp32 [M][ldc] accumulated with red.global.add -- the weight-gradient form, for many small
// same-shape problems (one per decoder layer) in ONE launch of the CTA-pair kernel.
int gemm_bf16_wgrad_batched(const void* const* A, const void* const* B, float* const* C, int nbatch, int M, int N, int K,
                            int lda, int ldb, int ldc, float alpha, cudaStream_t stream) {
  int rc = check_arch();
  if (rc) return rc;
  LAKO_REQUIRE(nbatch > 0 && M > 0 && N > 0 && K > 0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0, LAKO_ERR_BAD_ARG);
  GemmParams p;
  p.C = nullptr;
  p.aux = nullptr;
  p.M = M;
  p.N = N;
  p.K = K;
  p.ldc = ldc;
  p.ld_aux = 0;
  p.alpha = alpha;
  p.drop = DropSpec{};
  p.c_head_major = 0;
  p.tiles_m = (M + 255) / 256;
  p.tiles_n = (N + 255) / 256;
  p.kb_total = (K + BK - 1) / BK;
  p.splitk = 1;
  p.kb_per_split = p.kb_total;
  auto kern = gemm2_batched_kernel<true, true, EPI_F32_ATOMIC>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error(cudaGetErrorString(e), __FILE__, __LINE__);
      return LAKO_ERR_LAUNCH;
    }
    configured = true;
  }
  for (int z0 = 0; z0 < nbatch; z0 += GEMM_MAX_BATCH) {
    Gemm2Batch bt;
    bt.n = min(GEMM_MAX_BATCH, nbatch - z0);
    for (int z = 0; z < bt.n; ++z) {
      LAKO_REQUIRE((reinterpret_cast<uintptr_t>(C[z0 + z]) & 15) == 0, LAKO_ERR_BAD_ARG);
      rc = get_tensor_map_2d(&bt.a[z], A[z0 + z], M, K, lda, 64, BK);
      if (rc) return rc;
      rc = get_tensor_map_2d(&bt.b[z], B[z0 + z], N, K, ldb, 64, BK);
      if (rc) return rc;
      rc = get_tensor_map_2d_f32(&bt.cmap[z], C[z0 + z], N, M, ldc, 32, 32);
      if (rc) return rc;
      bt.c[z] = C[z0 + z];
    }
    for (int z = bt.n; z < GEMM_MAX_BATCH; ++z) {  // keep the unused slots valid descriptors
      bt.a[z] = bt.a[0];
      bt.b[z] = bt.b[0];
      bt.cmap[z] = bt.cmap[0];
      bt.c[z] = bt.c[0];
    }
    ProfScope prof(PROF_GEMM_SMALL, 2.0 * M * N * static_cast<double>(K) * bt.n, stream);
    const int total = p.tiles_m * p.tiles_n * bt.n;
    const int clusters = min(total, device_sm_count() / 2);
    launch_pdl(kern, dim3(clusters * 2), dim3(Gemm2Cfg::THREADS), Gemm2Cfg::SMEM_BYTES, stream, bt, p);
    LAKO_CHECK_LAUNCH();
  }
  return LAKO_OK;
}

}  // namespace lako

extern "C" int lako_gemm_bf16_wgrad_batched(const void* const* A, const void* const* B, float* const* C, int nbatch, int M,
                                            int N, int K, int lda, int ldb, int ldc, float alpha, void* stream) {
  return lako::gemm_bf16_wgrad_batched(A, B, C, nbatch, M, N, K, lda, ldb, ldc, alpha, static_cast<cudaStream_t>(stream));
}

extern "C" int lako_gemm_bf16_head_major(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb,
                                         int plane_rows, float alpha, void* stream) {
  return lako::gemm_bf16(A, B, C, M, N, K, lda, ldb, N, 0, 0, lako::EPI_BF16, alpha, nullptr, 0, 1, 0,
                         static_cast<cudaStream_t>(stream), lako::DropSpec{}, plane_rows);
}

extern "C" int lako_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                              int a_mn, int b_mn, int epi, float alpha, const void* aux, int ld_aux, int splitk,
                              int bn, void* stream) {
  return lako::gemm_bf16(A, B, C, M, N, K, lda, ldb, ldc, a_mn, b_mn, epi, alpha, aux, ld_aux, splitk, bn,
                         static_cast<cudaStream_t>(stream), lako::DropSpec{});
}
