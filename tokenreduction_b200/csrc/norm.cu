// Residual add + LayerNorm + bf16 cast in one pass (the callers either side of the attention producer):
//     x = x + drop_path(branch)           fp32 residual stream (bf16 branch promoted)      e.g. models/topk.py:87
//     y = norm(x)                         autocast runs layer_norm in fp32                  e.g. models/topk.py:94
//     linear(y)                           autocast casts y to bf16 (a separate copy kernel)
// Under bf16 autocast the reference spends 24 bytes per element on this (add 4+2+4, LayerNorm 4+4, cast 4+2); here the row
// is read once, the new residual row written once and the normalised row written once in bf16: 12 bytes per element, or 6
// when there is no branch to add.  Mean and variance are two-pass in registers (the row lives in registers), fp32.
#include "common.cuh"

namespace tokred {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float4 ld4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one warp per row; V = float4 chunks per lane (C = 128 * V): 3 for DeiT-S, 6 for DeiT-B
template <int V, typename TB>
__global__ void __launch_bounds__(kThreads) add_layernorm_kernel(const float* __restrict__ x, const TB* __restrict__ branch,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float eps, long long rows, float* __restrict__ x_out,
                                                                  __nv_bfloat16* __restrict__ y) {
  constexpr int C = 128 * V;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  float4 g[V], bt[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    g[j] = *reinterpret_cast<const float4*>(gamma + (j * 32 + lane) * 4);
    bt[j] = *reinterpret_cast<const float4*>(beta + (j * 32 + lane) * 4);
  }
  for (long long row = warp0; row < rows; row += (long long)gridDim.x * kWarps) {
    float4 v[V];
    const float* xr = x + row * C;
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = ld4(xr + (j * 32 + lane) * 4);
    if (branch) {
      const TB* br = branch + row * C;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float4 a;
        if (sizeof(TB) == 2) {
          const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(br) + (j * 32 + lane) * 4);
          a.x = __uint_as_float(raw.x << 16); a.y = __uint_as_float(raw.x & 0xffff0000u);
          a.z = __uint_as_float(raw.y << 16); a.w = __uint_as_float(raw.y & 0xffff0000u);
        } else {
          a = ld4(reinterpret_cast<const float*>(br) + (j * 32 + lane) * 4);
        }
        v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
      }
      float* xo = x_out + row * C;
#pragma unroll
      for (int j = 0; j < V; ++j) st4(xo + (j * 32 + lane) * 4, v[j]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
    __nv_bfloat16* yr = y + row * C;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = (v[j].x - mean) * rstd * g[j].x + bt[j].x, b = (v[j].y - mean) * rstd * g[j].y + bt[j].y;
      const float c = (v[j].z - mean) * rstd * g[j].z + bt[j].z, d = (v[j].w - mean) * rstd * g[j].w + bt[j].w;
      const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(yr + (j * 32 + lane) * 4) = o;
    }
  }
}

// x + branch alone (fp32 stream, bf16 branch): what a consumer other than a LayerNorm reads when a block's residual sum was
// deferred (modules.value(): the slices in front of a cluster layer, the DynamicViT predictor).  ATen's mixed-dtype add runs
// its unrolled, non-vectorised kernel at ~3.3 TB/s; this is the same fp32 addition with 16-byte accesses.
__global__ void __launch_bounds__(kThreads) residual_add_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ branch,
                                                                 long long n4, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i0 = (long long)blockIdx.x * kThreads + threadIdx.x; i0 < n4; i0 += 4 * stride) {
    float4 v[4];
    uint2 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        v[u] = ld4(x + i * 4);
        r[u] = *reinterpret_cast<const uint2*>(branch + i * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        v[u].x += __uint_as_float(r[u].x << 16); v[u].y += __uint_as_float(r[u].x & 0xffff0000u);
        v[u].z += __uint_as_float(r[u].y << 16); v[u].w += __uint_as_float(r[u].y & 0xffff0000u);
        st4(out + i * 4, v[u]);
      }
    }
  }
}

}  // namespace
}  // namespace tokred

using namespace tokred;

extern "C" int tokred_residual_add(const float* x, const void* branch, int64_t n, float* out, void* stream) {
  const char* what = "tokred_residual_add";
  if (n == 0) return TOKRED_OK;
  TOKRED_REQUIRE(x && branch && out, "%s: null tensor", what);
  TOKRED_REQUIRE(n > 0, "%s: n=%lld", what, (long long)n);
  if (n % 4 != 0 || !aligned16(x) || !aligned16(out) || (reinterpret_cast<uintptr_t>(branch) & 7u)) {
    set_error("%s: needs a multiple of 4 elements and 16-byte aligned tensors", what);
    return TOKRED_ERR_UNSUPPORTED;
  }
  const long long n4 = n / 4, want = (n4 + 4LL * kThreads - 1) / (4LL * kThreads);
  const int grid = (int)(want < (long long)kNumSMs * 8 ? want : (long long)kNumSMs * 8);
  residual_add_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, (const __nv_bfloat16*)branch, n4, out);
  return finish_launch(what);
}

extern "C" int tokred_add_layernorm(const float* x, const void* branch, int branch_dtype, const float* gamma, const float* beta,
                                    float eps, int64_t rows, int C, float* x_out, void* y, void* stream) {
  const char* what = "tokred_add_layernorm";
  if (rows == 0) return TOKRED_OK;
  TOKRED_REQUIRE(x && gamma && beta && y, "%s: null tensor", what);
  TOKRED_REQUIRE(!branch || x_out, "%s: branch without x_out", what);
  TOKRED_REQUIRE(!branch || valid_float_dtype(branch_dtype), "%s: bad branch dtype %d", what, branch_dtype);
  TOKRED_REQUIRE(rows > 0 && C > 0, "%s: rows=%lld C=%d", what, (long long)rows, C);
  if (C % 128 != 0 || C > 1024) {
    set_error("%s: C=%d (needs a multiple of 128 up to 1024)", what, C);
    return TOKRED_ERR_UNSUPPORTED;
  }
  TOKRED_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(beta) && (!branch || aligned16(branch)) && (!x_out || aligned16(x_out)) &&
                     (reinterpret_cast<uintptr_t>(y) & 7u) == 0,
                 "%s: tensors must be 16-byte aligned", what);
  const long long want = (rows + kWarps - 1) / kWarps;
  const int grid = (int)(want < (long long)kNumSMs * 8 ? want : (long long)kNumSMs * 8);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(V)                                                                                                             \
  do {                                                                                                                        \
    if (branch && branch_dtype == TOKRED_F32)                                                                                 \
      add_layernorm_kernel<V, float><<<grid, kThreads, 0, st>>>(x, (const float*)branch, gamma, beta, eps, rows, x_out,       \
                                                                (__nv_bfloat16*)y);                                          \
    else                                                                                                                      \
      add_layernorm_kernel<V, __nv_bfloat16><<<grid, kThreads, 0, st>>>(x, (const __nv_bfloat16*)branch, gamma, beta, eps,    \
                                                                        rows, x_out, (__nv_bfloat16*)y);                     \
  } while (0)
  switch (C / 128) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    case 5: LAUNCH(5); break;
    case 6: LAUNCH(6); break;
    case 7: LAUNCH(7); break;
    default: LAUNCH(8); break;
  }
#undef LAUNCH
  return finish_launch(what);
}
