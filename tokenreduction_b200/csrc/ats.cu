// ATS: adaptive token sampling (inverse-CDF sampling of the CLS-attention significance score).
// Reference: models/ats.py:44-89.
//
// One CTA per image replaces 37 ATen launches plus one torch.unique (a host sync) PER IMAGE: value norms and
// the significance score are reduced in shared memory, the CDF is a sequential fp32 scan (the order the CPU
// reference uses), every sampling step does an exact argmin over the <= 196 CDF entries with ATen's cdist
// arithmetic, and the per-image sorted unique ids come from a flag + count compaction — no sort, no sync.
// The batch-wide padded width max_b #unique is produced with one atomicMax so the caller needs at most ONE
// scalar read (exact-shape mode) or none (static mode, width = sample_count).
#include <math_constants.h>

#include "common.cuh"

namespace tokred {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <typename TV>
__global__ void __launch_bounds__(kThreads)
ats_sample_kernel(const TV* __restrict__ v, long long vs_b, long long vs_h, long long vs_n,
                  const float* __restrict__ attn, long long ahs, const uint8_t* __restrict__ mask,
                  const float* __restrict__ steps, int H, int N, int Dh, int n_steps, float eps, int use_mm, int64_t* __restrict__ ids_out,
                  uint8_t* __restrict__ mask_out, int32_t* __restrict__ max_count) {
  extern __shared__ float smem[];
  const int P = N - 1, b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* hp = smem;                                   // [H][P] cls_attn * ||v||
  float* cdf = hp + H * P;                            // [P]
  float* red = cdf + P;                               // [kWarps]
  int* hit = reinterpret_cast<int*>(red + kWarps);    // [N]
  int* count_s = hit + N;                             // [1]

  // significance per (head, patch): one THREAD per pair — the 64-element value row is one or two 128-byte lines,
  // read with 16-byte loads that are all independent (the first version used a warp per pair with a shuffle
  // reduction: 294 dependent global-latency round trips per warp, 370 us at B=128)
  const float* ab = attn + (long long)b * H * ahs;    // ahs: element stride between the heads' CLS rows
  {
    constexpr int VE = 16 / sizeof(TV);
    const bool vec = (Dh % VE == 0) && (vs_n % VE == 0) && (vs_h % VE == 0) && (vs_b % VE == 0) &&
                     ((reinterpret_cast<uintptr_t>(v) & 15u) == 0);
    if (vec && Dh <= 8 * VE) {
      // two (head, patch) pairs per thread and iteration, each value row (<= 8 x 16 bytes) entirely in flight: the
      // phase is a chain of dependent global-latency rounds, 2352 pairs / 256 threads = 9 of them with one pair
      const int nv = Dh / VE;
      for (int e0 = tid; e0 < H * P; e0 += 2 * kThreads) {
        int4 r[2][8];
        float a[2];
        bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = e0 + u * kThreads;
          on[u] = e < H * P;
          if (on[u]) {
            const int h = e / P, p = e - h * P;
            const TV* row = v + (long long)b * vs_b + (long long)h * vs_h + (long long)(1 + p) * vs_n;
#pragma unroll
            for (int w = 0; w < 8; ++w)
              if (w < nv) r[u][w] = *reinterpret_cast<const int4*>(row + w * VE);
            a[u] = ab[(long long)h * ahs + 1 + p];
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (on[u]) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w)
              if (w < nv) {
                const TV* q = reinterpret_cast<const TV*>(&r[u][w]);
#pragma unroll
                for (int i = 0; i < VE; ++i) { const float x = to_f32(q[i]); s = fmaf(x, x, s); }
              }
            hp[e0 + u * kThreads] = a[u] * sqrtf(s);
          }
        }
      }
    } else
    for (int e = tid; e < H * P; e += kThreads) {
      const int h = e / P, p = e % P;
      const TV* row = v + (long long)b * vs_b + (long long)h * vs_h + (long long)(1 + p) * vs_n;
      float s = 0.f;
      if (vec) {
        for (int d = 0; d < Dh; d += 8 * VE) {          // 8 x 16 bytes in flight: a whole 64-element bf16 row
          int4 r[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (d + u * VE < Dh) r[u] = *reinterpret_cast<const int4*>(row + d + u * VE);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (d + u * VE < Dh) {
              const TV* q = reinterpret_cast<const TV*>(&r[u]);
#pragma unroll
              for (int i = 0; i < VE; ++i) { const float x = to_f32(q[i]); s = fmaf(x, x, s); }
            }
        }
      } else {
        for (int d = 0; d < Dh; ++d) { const float x = to_f32(row[d]); s = fmaf(x, x, s); }
      }
      hp[e] = ab[(long long)h * ahs + 1 + p] * sqrtf(s);
    }
  }
  for (int t = tid; t < N; t += kThreads) hit[t] = 0;
  __syncthreads();
  // sum over heads (in head order), then the normaliser
  float part = 0.f;
  for (int p = tid; p < P; p += kThreads) {
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += hp[h * P + p];
    cdf[p] = s;
    part += s;
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  {
    float total = 0.f;
    for (int w = 0; w < kWarps; ++w) total += red[w];
    const float denom = total + eps;
    for (int p = tid; p < P; p += kThreads) cdf[p] = cdf[p] / denom;     // the divisions in parallel ...
  }
  __syncthreads();
  if (tid == 0) {
    float run = 0.f;
#pragma unroll 4
    for (int p = 0; p < P; ++p) {            // ... the inclusive scan sequential (fp32 order of the CPU reference)
      run += cdf[p];
      cdf[p] = run;
    }
  }
  __syncthreads();
  for (int p = tid; p < P; p += kThreads)
    if (!mask[(long long)b * N + 1 + p]) cdf[p] += 0.1f;
  __syncthreads();

  // inverse-CDF sampling: nearest CDF entry to every step (lowest index on ties)
  for (int q = tid; q < n_steps; q += kThreads) {
    const float t = steps[q];
    const float m2t = -2.0f * t, tt = __fmul_rn(t, t);
    int bp = 0;
    if (use_mm) {
      // ATen cdist matmul expansion on 1-d points: d = sqrt(max(((-2t)*c + t^2) + c^2, 1e-30)); argmin with the lowest
      // index on ties.  sqrtf is monotone, so the winners are exactly the entries whose radicand lies in [xmin, X], X the
      // largest float with sqrtf(X) == sqrtf(xmin) (at most three floats share a square root): one pass for xmin, a few
      // ulp steps for X, one pass for the first index -- the same index as the sqrt-per-entry loop (IEEE sqrtf with its
      // slow-path branches was a third of the kernel's stall samples) without a square root per entry.
      float xmin = CUDART_INF_F;
#pragma unroll 4
      for (int p = 0; p < P; ++p) {
        const float c = cdf[p];
        xmin = fminf(xmin, fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(m2t, c), tt), __fmul_rn(c, c)), 1e-30f));
      }
      const float dmin = sqrtf(xmin);
      float X = xmin;
      for (int s = 0; s < 8; ++s) {
        const float nx = __uint_as_float(__float_as_uint(X) + 1u);          // next float up (X > 0, finite)
        if (!(sqrtf(nx) == dmin)) break;
        X = nx;
      }
      bp = P;
      for (int p = P - 1; p >= 0; --p) {
        const float c = cdf[p];
        const float x = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(m2t, c), tt), __fmul_rn(c, c)), 1e-30f);
        bp = x <= X ? p : bp;
      }
      bp = (bp == P || !(xmin < CUDART_INF_F)) ? 0 : bp;                    // nothing below +inf: the strict "<" loop keeps index 0
    } else {
      float best = CUDART_INF_F;
      for (int p = 0; p < P; ++p) {
        const float d = fabsf(t - cdf[p]);
        if (d < best) { best = d; bp = p; }
      }
    }
    hit[1 + bp] = 1;
  }
  __syncthreads();

  // sorted unique ids by ballot compaction: position = hits in earlier chunks + earlier warps + lower lanes (the first
  // version counted with one thread and ranked every hit by a scan over all earlier flags)
  const int W = n_steps + 1;
  int64_t* ids_b = ids_out + (long long)b * W;
  uint8_t* m_b = mask_out + (long long)b * W;
  int* wcnt = count_s + 1;                             // [kWarps]
  int base = 0;
  for (int t0 = 1; t0 < N; t0 += kThreads) {
    const int t = t0 + tid;
    const bool h = t < N && hit[t] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, h);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    int pos = base + __popc(bal & ((1u << lane) - 1u)), tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int c = wcnt[w];
      pos += w < warp ? c : 0;
      tot += c;
    }
    if (h) {
      ids_b[1 + pos] = t;
      m_b[1 + pos] = 1;
    }
    base += tot;
    __syncthreads();
  }
  const int cnt = base;
  if (tid == 0) {
    atomicMax(max_count, cnt);
    ids_b[0] = 0;
    m_b[0] = 1;
  }
  for (int j = 1 + cnt + tid; j < W; j += kThreads) { ids_b[j] = 0; m_b[j] = 0; }
}

}  // namespace
}  // namespace tokred

using namespace tokred;

extern "C" int tokred_ats_sample(const void* v, int v_dtype, int64_t v_stride_b, int64_t v_stride_h,
                                 int64_t v_stride_n, const float* attn, int64_t attn_head_stride, const uint8_t* mask,
                                 const float* steps, int B,
                                 int H, int N, int Dh, int n_steps, float eps, int64_t* ids_out, uint8_t* mask_out,
                                 int32_t* max_count, void* stream) {
  const char* what = "tokred_ats_sample";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(v && attn && mask && steps && ids_out && mask_out && max_count, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(v_dtype), "%s: bad v dtype %d", what, v_dtype);
  TOKRED_REQUIRE(B >= 0 && H >= 1 && N >= 2 && Dh >= 1, "%s: bad shape B=%d H=%d N=%d Dh=%d", what, B, H, N, Dh);
  // n_steps may exceed N-1: after the first ATS stage N = 1 + max_b #unique shrinks with peaked attention while the
  // per-stage sample_count is fixed by keep_rate (models/ats.py:204-205).  hit[] is sized N, #unique <= min(n_steps, N-1)
  // and the id / mask rows are n_steps + 1 wide, so nothing in the kernel needs an upper bound.
  TOKRED_REQUIRE(n_steps >= 1 && n_steps <= 65535, "%s: n_steps=%d outside [1, 65535]", what, n_steps);
  if (B == 0) return TOKRED_OK;
  const long long ahs = attn_head_stride > 0 ? attn_head_stride : (long long)N * N;
  TOKRED_REQUIRE(ahs >= N, "%s: attn_head_stride %lld < N", what, ahs);
  const int P = N - 1;
  const size_t smem = ((size_t)H * P + P + kWarps + N + 1 + kWarps) * 4;
  const int use_mm = (n_steps > 25 || P > 25) ? 1 : 0;     // ATen: matmul expansion when either side has > 25 points
  cudaStream_t st = (cudaStream_t)stream;
  if (v_dtype == TOKRED_F32) {
    if (int e = allow_smem(ats_sample_kernel<float>, smem, what)) return e;
    ats_sample_kernel<float><<<B, kThreads, smem, st>>>((const float*)v, v_stride_b, v_stride_h, v_stride_n, attn, ahs, mask,
                                                        steps, H, N, Dh, n_steps, eps, use_mm, ids_out, mask_out,
                                                        max_count);
  } else {
    if (int e = allow_smem(ats_sample_kernel<__nv_bfloat16>, smem, what)) return e;
    ats_sample_kernel<__nv_bfloat16><<<B, kThreads, smem, st>>>((const __nv_bfloat16*)v, v_stride_b, v_stride_h,
                                                                v_stride_n, attn, ahs, mask, steps, H, N, Dh, n_steps, eps,
                                                                use_mm, ids_out, mask_out, max_count);
  }
  return finish_launch(what);
}
