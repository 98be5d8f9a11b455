// Warp-level tensor-core helpers (mma.sync m16n8k16 bf16, ldmatrix, cp.async staging into 128B-swizzled [rows][64] tiles)
// shared by the kernels whose tiles are too small for tcgen05's 128-row MMAs: the cross-attention of T ~ 10 query rows
// (xattn_mma.cu) and the any-length encoder attention (enc_attn_long.cu).
#pragma once
#include "common.cuh"

namespace lako {

__device__ __forceinline__ uint32_t swz(int row, int chunk) {  // byte offset in a [rows][128 B] swizzled tile
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// stage a [nrows_valid of `rows`][64] bf16 tile (row pitch ld elements) into a swizzled [rows][128 B] smem tile
__device__ __forceinline__ void stage_sw(uint8_t* s, const __nv_bfloat16* g, size_t ld, int rows, int nvalid, int tid,
                                         int nthreads) {
  const uint32_t base = smem_u32(s);
  for (int i = tid; i < rows * 8; i += nthreads) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < nvalid;
    cp_async16(base + swz(r, c), g + (ok ? static_cast<size_t>(r) * ld + c * 8 : 0), ok);
  }
}

}  // namespace lako
