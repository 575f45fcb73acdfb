// Shared host/device helpers for libtokred_sm100a.so (sm_100a only; no torch headers, C ABI in include/tokred.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tokred.h"

namespace tokred {

// ---------------------------------------------------------------------------------------- host side
void set_error(const char* fmt, ...);   // thread-local message returned by tokred_last_error()
void count_launch(int n);               // feeds tokred_launch_count()

#define TOKRED_REQUIRE(cond, ...)                         \
  do {                                                    \
    if (!(cond)) {                                        \
      ::tokred::set_error(__VA_ARGS__);                   \
      return TOKRED_ERR_ARGUMENT;                         \
    }                                                     \
  } while (0)

// Called after every launch: enqueue-only API, so only launch-configuration errors surface here.
inline int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  count_launch(1);
  return TOKRED_OK;
}

// Opt a kernel in to > 48 KB dynamic shared memory (cheap; the driver caches the attribute).
template <typename K>
inline int allow_smem(K kernel, size_t bytes, const char* what) {
  if (bytes > 227 * 1024) {
    set_error("%s: needs %zu B of shared memory per CTA (> 227 KB)", what, bytes);
    return TOKRED_ERR_UNSUPPORTED;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu): %s", what, bytes, cudaGetErrorString(e));
      return (int)e;
    }
  }
  return TOKRED_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int dtype_size(int dt) { return dt == TOKRED_BF16 ? 2 : 4; }
inline bool valid_float_dtype(int dt) { return dt == TOKRED_F32 || dt == TOKRED_BF16; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
constexpr int kNumSMs = 148;   // B200

// ---------------------------------------------------------------------------------------- device side
#ifdef __CUDACC__

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
// round-trip through T: identity for fp32, round-to-nearest-even for bf16 (emulates elementwise bf16 arithmetic)
template <typename T> __device__ __forceinline__ float round_as(float v) { return to_f32(from_f32<T>(v)); }
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// 16-byte streaming copies: rows are read once and written once, keep them out of L1.
__device__ __forceinline__ int4 ld_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One warp copies `bytes` (multiple of 16, both pointers 16-B aligned) with up to 4 requests in flight per lane.
__device__ __forceinline__ void warp_copy_row16(void* dst, const void* src, int bytes, int lane) {
  const char* s = reinterpret_cast<const char*>(src);
  char* d = reinterpret_cast<char*>(dst);
  int off = lane * 16;
  for (; off + 3 * 512 < bytes; off += 4 * 512) {
    int4 a = ld_stream16(s + off), b = ld_stream16(s + off + 512), c = ld_stream16(s + off + 1024),
         e = ld_stream16(s + off + 1536);
    st_stream16(d + off, a); st_stream16(d + off + 512, b); st_stream16(d + off + 1024, c);
    st_stream16(d + off + 1536, e);
  }
  for (; off < bytes; off += 512) st_stream16(d + off, ld_stream16(s + off));
}
// One warp copies up to 4 rows of `bytes` each (multiple of 16, 16-B aligned): the loads of ALL rows are issued before
// the first store -- 8 requests (4 KB per warp) in flight instead of 3, which is what the gather kernels needed to
// keep HBM busy with few resident warps (small batches).
__device__ __forceinline__ void warp_copy_rows16x4(char* dbase, const int (&doff)[4], const char* sbase, const int (&soff)[4],
                                                   int nrows, int bytes, int lane) {
  for (int off = lane * 16; off < bytes; off += 1024) {
    int4 a[4], b[4];
    const bool two = off + 512 < bytes;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nrows) {
        a[u] = ld_stream16(sbase + soff[u] + off);
        if (two) b[u] = ld_stream16(sbase + soff[u] + off + 512);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nrows) {
        st_stream16(dbase + doff[u] + off, a[u]);
        if (two) st_stream16(dbase + doff[u] + off + 512, b[u]);
      }
  }
}
// Generic fallback (any alignment / size), element type T.
template <typename T>
__device__ __forceinline__ void warp_copy_row_elems(T* __restrict__ dst, const T* __restrict__ src, int n, int lane) {
  int i = lane;
  for (; i + 96 < n; i += 128) {          // 4 independent loads in flight per lane
    const T a = src[i], b = src[i + 32], c = src[i + 64], d = src[i + 96];
    dst[i] = a; dst[i + 32] = b; dst[i + 64] = c; dst[i + 96] = d;
  }
  for (; i < n; i += 32) dst[i] = src[i];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ATen's max / argmax treat NaN as the greatest value (it propagates, index = first NaN)
__device__ __forceinline__ bool nan_gt(float a, float b) { return a > b || (a != a && b == b); }
__device__ __forceinline__ bool nan_eq(float a, float b) { return a == b || (a != a && b != b); }
// (value, index) arg-reductions with lowest-index tie-break (ATen max/min semantics).
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (nan_gt(ov, v) || (nan_eq(ov, v) && oi < i)) { v = ov; i = oi; }
  }
}
__device__ __forceinline__ void warp_argmin(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// Descending rank of element i among n keys held in shared memory: number of keys that sort before it under
// (value desc, index asc).  O(n) broadcast reads per caller; n <= a few hundred tokens on this path, so the
// whole ordering is O(n^2) conflict-free LDS with no barriers — cheaper than a bitonic network at this size.
// The order is TOTAL: NaN sorts as the largest value (what ATen's topk / sort(descending) do) and NaNs tie among
// themselves by index, so every rank in [0, n) is produced exactly once even for NaN keys (a zero-norm ToMe
// metric row, overflowed activations) and no consumer ever reads an unwritten index slot.
__device__ __forceinline__ int rank_desc(const float* keys, int n, int i) {
  const float ki = keys[i];
  int r = 0;
  if (ki != ki) {
    for (int j = 0; j < i; ++j) { const float kj = keys[j]; r += (kj != kj); }
  } else {
    int r1 = 0, r2 = 0, r3 = 0, j = 0;            // four counters: the add chain was the critical path
    for (; j + 3 < n; j += 4) {
      const float k0 = keys[j], k1 = keys[j + 1], k2 = keys[j + 2], k3 = keys[j + 3];
      r += !(k0 <= ki) || (k0 == ki && j < i);     // !(kj <= ki): kj > ki, or kj is NaN
      r1 += !(k1 <= ki) || (k1 == ki && j + 1 < i);
      r2 += !(k2 <= ki) || (k2 == ki && j + 2 < i);
      r3 += !(k3 <= ki) || (k3 == ki && j + 3 < i);
    }
    for (; j < n; ++j) { const float kj = keys[j]; r += !(kj <= ki) || (kj == ki && j < i); }
    r += r1 + r2 + r3;
  }
  return r;
}

// rank_desc for a warp whose lanes rank the elements of one window [ilo, ihi) (warp-uniform bounds, any lane's i
// inside it).  Keys before the window only need "kj >= ki or NaN" (they win ties), keys after it "kj > ki or NaN":
// one compare + one add per key and 16-byte broadcast reads, against three compares and the index test of the
// generic form -- which is kept inside the window.  Same total order, same result as rank_desc.
template <bool GE>
__device__ __forceinline__ int count_before(const float* keys, int a, int b, float ki) {
  int r0 = 0, r1 = 0, r2 = 0, r3 = 0, j = a;
  for (; j < b && (reinterpret_cast<uintptr_t>(keys + j) & 15u); ++j) { const float kj = keys[j]; r0 += GE ? !(kj < ki) : !(kj <= ki); }
  for (; j + 3 < b; j += 4) {
    const float4 k = *reinterpret_cast<const float4*>(keys + j);
    r0 += GE ? !(k.x < ki) : !(k.x <= ki);
    r1 += GE ? !(k.y < ki) : !(k.y <= ki);
    r2 += GE ? !(k.z < ki) : !(k.z <= ki);
    r3 += GE ? !(k.w < ki) : !(k.w <= ki);
  }
  for (; j < b; ++j) { const float kj = keys[j]; r0 += GE ? !(kj < ki) : !(kj <= ki); }
  return (r0 + r1) + (r2 + r3);
}
__device__ __forceinline__ int rank_desc_window(const float* keys, int n, int i, int ilo, int ihi) {
  const float ki = keys[i];
  if (ki != ki) return rank_desc(keys, n, i);
  ilo = ilo < 0 ? 0 : ilo;
  ihi = ihi > n ? n : ihi;
  int r = count_before<true>(keys, 0, ilo, ki);
  for (int j = ilo; j < ihi; ++j) { const float kj = keys[j]; r += !(kj <= ki) || (kj == ki && j < i); }
  return r + count_before<false>(keys, ihi, n, ki);
}

// the same count restricted to keys [j0, j1): two lanes split a rank and add their halves
__device__ __forceinline__ int rank_desc_range(const float* keys, int j0, int j1, int i) {
  const float ki = keys[i];
  int r = 0, r1 = 0, r2 = 0, r3 = 0, j = j0;
  if (ki != ki) {
    for (; j < j1 && j < i; ++j) { const float kj = keys[j]; r += (kj != kj); }
    return r;
  }
  for (; j + 3 < j1; j += 4) {
    const float k0 = keys[j], k1 = keys[j + 1], k2 = keys[j + 2], k3 = keys[j + 3];
    r += !(k0 <= ki) || (k0 == ki && j < i);
    r1 += !(k1 <= ki) || (k1 == ki && j + 1 < i);
    r2 += !(k2 <= ki) || (k2 == ki && j + 2 < i);
    r3 += !(k3 <= ki) || (k3 == ki && j + 3 < i);
  }
  for (; j < j1; ++j) { const float kj = keys[j]; r += !(kj <= ki) || (kj == ki && j < i); }
  return r + r1 + r2 + r3;
}

__device__ __forceinline__ int clamp_idx(long long v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : (int)v); }

// Phase stamps for kernel bring-up (tools/diag/*): compiled ONLY into the debug library (python -m
// tokenreduction_b200.build --stamps -> libtokred_sm100a_dbg.so); the release build contains no stamp code at all.
#ifdef TOKRED_STAMPS
static __device__ unsigned long long* g_tokred_stamps = nullptr;      // one per translation unit: [CTA][8 images][32 slots]
__device__ __forceinline__ void tokred_stamp(int img, int slot) {
  if (g_tokred_stamps) g_tokred_stamps[((size_t)blockIdx.x * 8 + (img & 7)) * 32 + slot] = (unsigned long long)clock64();
}
#define TOKRED_STAMP(cond, img, slot) do { if (cond) ::tokred::tokred_stamp((img), (slot)); } while (0)
// each .cu that stamps exports tokred_debug_set_stamps_<unit>(device buffer or NULL)
#define TOKRED_STAMP_SETTER(unit)                                                                            \
  extern "C" __attribute__((visibility("default"))) int tokred_debug_set_stamps_##unit(void* device_buffer) { \
    unsigned long long* p = (unsigned long long*)device_buffer;                                              \
    return (int)cudaMemcpyToSymbol(::tokred::g_tokred_stamps, &p, sizeof(p));                                \
  }
#else
#define TOKRED_STAMP(cond, img, slot) do { } while (0)
#define TOKRED_STAMP_SETTER(unit)
#endif

#endif  // __CUDACC__
}  // namespace tokred
