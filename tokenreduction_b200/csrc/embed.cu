// The data formats in front of the first block (the callers either side of the path), bf16 autocast only:
//   patchify          image [B,Cin,H,W] fp32 -> [B, gh*gw, Cin*ph*pw] bf16: the operand of the patch-embedding GEMM.  The
//                     reference's stride-16 convolution (models/deit_viz.py PatchEmbed) is that GEMM; ATen forms the operand
//                     with a cast kernel and a permuting copy (33 + 117 us at B=256: the permuted copy moves 2-byte elements
//                     one by one), here it is one pass at HBM speed.
//   embed_layernorm   x = cat(cls [, dist], patches) + pos_embed  (fp32 residual stream, models/deit_viz.py forward_features)
//                     and y = blocks[0].norm1(x) rounded to bf16 in the same pass: cat, add, LayerNorm and cast were four
//                     launches and 44 bytes per element, now 10.
#include "common.cuh"

namespace tokred {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// grid = (gh, B): one CTA turns the ph image rows x Cin channels of one patch row into gw output rows (contiguous in out).
// Loads are 16-byte coalesced along the image row; the bf16 patch rows are assembled in shared memory (patch stride padded
// by 32 bytes: the 8-byte stores of a warp spread over all banks) and leave as 16-byte coalesced stores.
__global__ void __launch_bounds__(kThreads)
patchify_kernel(const float* __restrict__ img, int Cin, int H, int W, int ph, int pw, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int gy = blockIdx.x, b = blockIdx.y, gw = W / pw, W4 = W / 4;
  const int D = Cin * ph * pw;                       // elements per output row
  const int SP = D * 2 + 32;                         // bytes between patches in shared memory (+32: a warp's 8-byte stores tile all banks)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* ib = img + (long long)b * Cin * H * W + (long long)gy * ph * W;
  // A warp walks image rows r = (c, py); a lane owns the same (up to four) 16-byte column groups q = lane + 32 j of
  // every row, so the patch / pixel split of a column is computed once per thread, not once per element (the first
  // version spent its time in integer divisions: 53 % of the HBM rate).
  constexpr int QMAX = 4;
  int soff[QMAX];
#pragma unroll
  for (int j = 0; j < QMAX; ++j) {
    const int q = lane + 32 * j;
    soff[j] = q < W4 ? ((q * 4) / pw) * SP + ((q * 4) % pw) * 2 : -1;
  }
  const int nrows = Cin * ph;
  if (W4 <= 32 * QMAX) {
    for (int r0 = warp; r0 < nrows; r0 += 2 * kWarps) {          // two rows = up to eight 16-byte loads in flight per lane
      float4 v[2][QMAX];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = r0 + u * kWarps;
        if (r < nrows) {
          const float* rowp = ib + ((long long)(r / ph) * H + (r % ph)) * W + lane * 4;
#pragma unroll
          for (int j = 0; j < QMAX; ++j)
            if (soff[j] >= 0) v[u][j] = *reinterpret_cast<const float4*>(rowp + 128 * j);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = r0 + u * kWarps;
        if (r < nrows) {
          unsigned char* dst = sm + (size_t)r * pw * 2;           // (c*ph + py) * pw elements into the patch row
#pragma unroll
          for (int j = 0; j < QMAX; ++j)
            if (soff[j] >= 0) {
              const __nv_bfloat162 lo = __floats2bfloat162_rn(v[u][j].x, v[u][j].y), hi = __floats2bfloat162_rn(v[u][j].z, v[u][j].w);
              uint2 o;
              o.x = *reinterpret_cast<const uint32_t*>(&lo);
              o.y = *reinterpret_cast<const uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(dst + soff[j]) = o;
            }
        }
      }
    }
  } else {                                                         // very wide images: generic column loop
    for (int r = warp; r < nrows; r += kWarps) {
      const float* rowp = ib + ((long long)(r / ph) * H + (r % ph)) * W;
      for (int q = lane; q < W4; q += 32) {
        const float4 v = *reinterpret_cast<const float4*>(rowp + q * 4);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(sm + (size_t)((q * 4) / pw) * SP + ((size_t)r * pw + (q * 4) % pw) * 2) = o;
      }
    }
  }
  __syncthreads();
  const int row16 = D / 8;                           // 16-byte units per output row
  unsigned char* ob = reinterpret_cast<unsigned char*>(out + ((long long)b * gridDim.x + gy) * gw * D);
  int gx = 0, w = threadIdx.x;
  while (w >= row16) { w -= row16; ++gx; }
  for (int i = threadIdx.x; i < gw * row16; i += kThreads) {
    st_stream16(ob + (size_t)i * 16, *reinterpret_cast<const int4*>(sm + (size_t)gx * SP + (size_t)w * 16));
    w += kThreads;
    while (w >= row16) { w -= row16; ++gx; }
  }
}

__device__ __forceinline__ float4 ld4s(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// one warp per token row; V = float4 chunks per lane (C = 128 * V).  The LayerNorm is norm.cu's, operation for operation.
// TP: dtype of the patch rows (bf16: GEMM / soft-merge outputs; fp32: merged cluster tokens).  tok_bs: elements between the
// images' token rows (0: one shared set -- cls / dist parameters; N*C: the class rows x[:, :T] of a [B,N,C] stream, read in
// place); pos may be null (the re-concatenation after a cluster layer, e.g. models/sinkhorn.py:168, adds nothing).
template <int V, typename TP>
__global__ void __launch_bounds__(kThreads)
embed_layernorm_kernel(const TP* __restrict__ patches, const float* __restrict__ tokens, long long tok_bs,
                       const float* __restrict__ pos, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       int B, int P, int T, float* __restrict__ x_out, __nv_bfloat16* __restrict__ y) {
  constexpr int C = 128 * V;
  const int lane = threadIdx.x & 31, N = T + P;
  const long long rows = (long long)B * N;
  float4 g[V], bt[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    g[j] = *reinterpret_cast<const float4*>(gamma + (j * 32 + lane) * 4);
    bt[j] = *reinterpret_cast<const float4*>(beta + (j * 32 + lane) * 4);
  }
  for (long long row = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * kWarps) {
    const int t = (int)(row % N);
    const long long b = row / N;
    float4 v[V];
    if (t < T) {
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = *reinterpret_cast<const float4*>(tokens + b * tok_bs + (long long)t * C + (j * 32 + lane) * 4);
    } else {
      const TP* pr = patches + (b * P + (t - T)) * C;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if constexpr (sizeof(TP) == 2) {
          const uint2 raw = *reinterpret_cast<const uint2*>(pr + (j * 32 + lane) * 4);
          v[j].x = __uint_as_float(raw.x << 16); v[j].y = __uint_as_float(raw.x & 0xffff0000u);
          v[j].z = __uint_as_float(raw.y << 16); v[j].w = __uint_as_float(raw.y & 0xffff0000u);
        } else {
          v[j] = ld4s(reinterpret_cast<const float*>(pr) + (j * 32 + lane) * 4);
        }
      }
    }
    const float* pe = pos ? pos + (long long)t * C : nullptr;
    float* xo = x_out + row * C;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (pe) {
        const float4 a = *reinterpret_cast<const float4*>(pe + (j * 32 + lane) * 4);
        v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
      }
      asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(xo + (j * 32 + lane) * 4), "f"(v[j].x), "f"(v[j].y),
                   "f"(v[j].z), "f"(v[j].w)
                   : "memory");
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = v[j].x - mean, bb = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
    __nv_bfloat16* yr = y + row * C;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = (v[j].x - mean) * rstd * g[j].x + bt[j].x, bb = (v[j].y - mean) * rstd * g[j].y + bt[j].y;
      const float c = (v[j].z - mean) * rstd * g[j].z + bt[j].z, d = (v[j].w - mean) * rstd * g[j].w + bt[j].w;
      const __nv_bfloat162 lo = __floats2bfloat162_rn(a, bb), hi = __floats2bfloat162_rn(c, d);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(yr + (j * 32 + lane) * 4) = o;
    }
  }
}

}  // namespace
}  // namespace tokred

using namespace tokred;

extern "C" int tokred_patchify(const float* img, int B, int Cin, int H, int W, int ph, int pw, void* out, void* stream) {
  const char* what = "tokred_patchify";
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(img && out, "%s: null tensor", what);
  TOKRED_REQUIRE(B > 0 && Cin >= 1 && ph >= 1 && pw >= 1 && H >= ph && W >= pw && H % ph == 0 && W % pw == 0,
                 "%s: bad shape B=%d Cin=%d H=%d W=%d patch %dx%d", what, B, Cin, H, W, ph, pw);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (pw % 4 != 0 || (Cin * ph * pw) % 8 != 0 || !aligned16(img) || !aligned16(out)) {
    set_error("%s: needs a patch width that is a multiple of 4, rows of whole 16-byte units and 16-byte aligned tensors", what);
    return TOKRED_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)(W / pw) * ((size_t)Cin * ph * pw * 2 + 32);
  if (int e = allow_smem(patchify_kernel, smem, what)) return e;
  patchify_kernel<<<dim3(H / ph, B), kThreads, smem, (cudaStream_t)stream>>>(img, Cin, H, W, ph, pw, (__nv_bfloat16*)out);
  return finish_launch(what);
}

extern "C" int tokred_embed_layernorm(const void* patches, int patch_dtype, const float* tokens, int64_t tokens_batch_stride,
                                      const float* pos, const float* gamma, const float* beta, float eps, int B, int P, int T,
                                      int C, float* x_out, void* y, void* stream) {
  const char* what = "tokred_embed_layernorm";
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(patches && (tokens || T == 0) && gamma && beta && x_out && y, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(patch_dtype), "%s: bad patch dtype %d", what, patch_dtype);
  TOKRED_REQUIRE(tokens_batch_stride >= 0 && tokens_batch_stride % 4 == 0, "%s: tokens_batch_stride=%lld", what,
                 (long long)tokens_batch_stride);
  TOKRED_REQUIRE(B > 0 && P >= 1 && T >= 0 && C > 0, "%s: bad shape B=%d P=%d T=%d C=%d", what, B, P, T, C);
  if (C % 128 != 0 || C > 1024) {
    set_error("%s: C=%d (needs a multiple of 128 up to 1024)", what, C);
    return TOKRED_ERR_UNSUPPORTED;
  }
  TOKRED_REQUIRE(aligned16(tokens) && aligned16(pos) && aligned16(gamma) && aligned16(beta) && aligned16(x_out) &&
                     (reinterpret_cast<uintptr_t>(patches) & (patch_dtype == TOKRED_F32 ? 15u : 7u)) == 0 &&
                     (reinterpret_cast<uintptr_t>(y) & 7u) == 0,
                 "%s: tensors must be 16-byte aligned", what);
  const long long rows = (long long)B * (T + P);
  const long long want = (rows + kWarps - 1) / kWarps;
  const int grid = (int)(want < (long long)kNumSMs * 8 ? want : (long long)kNumSMs * 8);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(V)                                                                                                       \
  do {                                                                                                                  \
    if (patch_dtype == TOKRED_F32)                                                                                      \
      embed_layernorm_kernel<V, float><<<grid, kThreads, 0, st>>>((const float*)patches, tokens, (long long)tokens_batch_stride, \
                                                                  pos, gamma, beta, eps, B, P, T, x_out, (__nv_bfloat16*)y); \
    else                                                                                                                \
      embed_layernorm_kernel<V, __nv_bfloat16><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)patches, tokens,          \
                                                                          (long long)tokens_batch_stride, pos, gamma, beta, \
                                                                          eps, B, P, T, x_out, (__nv_bfloat16*)y);         \
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
