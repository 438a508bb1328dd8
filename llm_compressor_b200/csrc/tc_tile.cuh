// Building blocks of the persistent tensor-core tile kernels (chol.cu: tile-task Cholesky-inverse; solver.cu: GPTQ
// super-block chain): cp.async staging of [128][32]-float chunks in the K-major SWIZZLE_128B layout the tensor core reads,
// 3xTF32 tcgen05.mma issue (M = 128, N <= 128 per instruction), release / acquire flags, named barriers.
#pragma once

#include "umma.cuh"

namespace lcb {

constexpr int TT_ROWS = 128;
constexpr int TT_PLANE = TT_ROWS * 32 * 4;   // 16 KB: one [128][32] fp32 chunk

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const uint32_t* p) {
  while (ld_acquire_u32(p) == 0u) __nanosleep(40);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void umma_tf32_128(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  // c_format F32, a/b TF32, K-major both, N = 128, M = 128
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_n(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t n, uint32_t accumulate) {
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t tt_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// x = hi + lo exactly, hi = tf32(x) (round to nearest)
__device__ __forceinline__ void tt_split4(const float4 v, float4& h, float4& l) {
  uint32_t b;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.x)); h.x = __uint_as_float(b); l.x = __fsub_rn(v.x, h.x);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.y)); h.y = __uint_as_float(b); l.y = __fsub_rn(v.y, h.y);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.z)); h.z = __uint_as_float(b); l.z = __fsub_rn(v.z, h.z);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.w)); h.w = __uint_as_float(b); l.w = __fsub_rn(v.w, h.w);
}
// the 12 MMAs of one 32-wide chunk: D (+)= (Ah + Al)(Bh + Bl)^T without the lo * lo term, small terms first
__device__ __forceinline__ void tt_mma_chunk(uint32_t tmem_d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, bool first) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t ko = k * 32;
    const uint64_t dah = tt_desc(ah + ko), dal = tt_desc(al + ko), dbh = tt_desc(bh + ko), dbl = tt_desc(bl + ko);
    umma_tf32_128(tmem_d, dal, dbh, (first && k == 0) ? 0u : 1u);
    umma_tf32_128(tmem_d, dah, dbl, 1u);
    umma_tf32_128(tmem_d, dah, dbh, 1u);
  }
}
// rows [row0, 128) of one [128][32]-float chunk of a row-major tile -> dst (row - row0 at 128 bytes each, 16-byte pieces
// XOR-swizzled by row & 7; row0 % 8 == 0); rows >= rows_valid are zero-filled.  NT = number of loading threads.
template <int NT>
__device__ __forceinline__ void tt_load_plane(uint8_t* dst, const float* tile, int64_t ld, int kofs, int rows_valid, int tid,
                                              int row0 = 0) {
  const uint32_t d0 = smem_u32(dst);
  const int pieces = (TT_ROWS - row0) * 8;
  for (int p = tid; p < pieces; p += NT) {
    const int rl = p >> 3, kq = p & 7, row = row0 + rl;
    const bool ok = row < rows_valid;
    const float* src = tile + (int64_t)(ok ? row : 0) * ld + kofs + 4 * kq;
    cp_async16(d0 + (uint32_t)(rl * 128 + ((kq ^ (row & 7)) << 4)), src, ok ? 16 : 0);
  }
}


}  // namespace lcb
