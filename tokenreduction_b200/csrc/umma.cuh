// Minimal hand-written tcgen05 / TMEM / mbarrier layer for sm_100a (single CTA, cta_group::1).
//
// Every tensor-core use in this library has the same shape: the operands are PRODUCED IN-KERNEL (normalised /
// rounded tokens), so threads write them straight into shared memory in the canonical K-major, no-swizzle UMMA
// layout (no TMA: there is no global tile to fetch as-is), one elected thread issues tcgen05.mma with the
// accumulator in TMEM, completion is signalled through tcgen05.commit -> mbarrier, and the epilogue reads the
// accumulator with tcgen05.ld 32x32b (thread = accumulator row), which is exactly the access pattern the per-row
// arg-reductions need.
//
// Canonical K-major / SWIZZLE_NONE operand layout (16-byte units; cute::UMMA "INTERLEAVE"):
//   a "core matrix" is 8 rows x 16 bytes stored contiguously (128 B, row stride 16 B);
//   core matrices adjacent along K are LBO bytes apart, adjacent along M/N (next 8 rows) SBO bytes apart.
//   byte offset of element (row r, k) with E-byte elements:  (r/8)*SBO + (r%8)*16 + (k/(16/E))*LBO + (k%(16/E))*E
// We use LBO = 128 (K-chunks of a row group back to back) and SBO = (#K-chunks)*128.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tokred {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ------------------------------------------------------------------------------------------ proxies / fences
// generic-proxy shared-memory writes -> visible to the async proxy the tensor core reads operands through
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ TMEM
// ncols: power of two in [32, 512].  Must be executed by ONE full warp; the same warp deallocates.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__host__ __device__ constexpr uint32_t tmem_cols_pow2(uint32_t n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

// accumulator -> registers: warp w may only touch TMEM lanes 32*(w%4) .. +31; lane l of the warp receives row
// 32*(w%4)+l, registers = consecutive fp32 columns starting at the column in taddr.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

// ------------------------------------------------------------------------------------------ descriptors
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version for sm_100
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// 128-byte swizzled operand tiles (what a TMA box with CU_TENSOR_MAP_SWIZZLE_128B writes): rows of 128 bytes, the 16-byte
// chunk index XORed with (row % 8); the tile base must be 1024-byte aligned.  K-major: a row is one M/N index with 64
// bf16 along K, SBO = 1024 (next 8 rows), LBO unused, a K step of 16 elements advances the start address by 32 bytes.
// MN-major: a row is one K index with 64 bf16 along M/N, SBO = 1024 (next 8 K indices), LBO = stride to the next 64
// M/N elements (unused for N = 64).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version for sm_100
  d |= (uint64_t)2 << 61;   // layout_type SWIZZLE_128B
  return d;
}
enum { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, both operands K-major, dense
__host__ __device__ constexpr uint32_t instr_desc(uint32_t fmt, uint32_t m, uint32_t n) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of (row, k) in the canonical layout described at the top; esize = element bytes
__device__ __forceinline__ uint32_t kmajor_offset(uint32_t row, uint32_t k, uint32_t esize, uint32_t sbo_bytes) {
  const uint32_t per_chunk = 16u / esize;
  return (row >> 3) * sbo_bytes + (row & 7u) * 16u + (k / per_chunk) * 128u + (k % per_chunk) * esize;
}

}  // namespace umma
}  // namespace tokred
