// Soft (dense-assignment) merges: Sinkhorn, PatchMerger, SiT TokenSlimmingModule.
// Reference: models/sinkhorn.py:25-86, models/patchmerger.py:35-39, models/sit.py:36-40.
//
// All three are  out[K,C] = W[K,P] . Xn[P,C]  where W comes from a score matrix Z[K,P] that is itself a
// contraction  Z = Q[K,C] . Xn[P,C]^T  (Sinkhorn: Q = normalised centres, Xn = L2-normalised tokens;
// PatchMerger: Q = queries, Xn = LayerNorm(x)) or is given (SiT: logits^T).  The reference runs them as
// GEMM -> [B,K,P] round trip -> 2*iters logsumexp launches / softmax -> GEMM (40 / 6 launches).  Here ONE
// persistent CTA per image keeps Z/W ([K][P] fp32, <= 138 KB) in shared memory between the two contractions,
// applies the token normalisation while staging operand tiles, and runs the Sinkhorn iterations / softmax as
// shared-memory row and column passes.  Global traffic = x (re-read from L2 by the second contraction), W out,
// out out.  lowp = 1 reproduces CUDA autocast: operands and results of both contractions rounded to bf16,
// everything between them in fp32.
#include <math_constants.h>

#include "gemm_nt.cuh"

namespace tokred {
namespace {

constexpr int kThreads = kGemmThreads;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 208, kMaxP = 208;
constexpr int PT = 32;     // tokens per staged tile of the second contraction
constexpr int CT = 128;    // output columns per pass of the second contraction

enum { MODE_SINKHORN = 0, MODE_PATCHMERGER = 1, MODE_SIT = 2 };

struct SoftParams {
  const void* x;            // [B,P,C], images xbs elements apart
  long long xbs;
  const float* q;           // [K,C]   (sinkhorn: v_hat, patchmerger: queries)
  const void* logits;       // [B,P,K] (sit)
  int logits_dtype;
  const float* ln_w;
  const float* ln_b;
  const float* scale_ptr;   // sit: device scalar
  float scale;              // patchmerger: sim * scale ; sinkhorn: 1/eps
  float log_norm;           // sinkhorn: -log(K+P) as the reference computes it
  float ln_eps;
  int iters, lowp;
  int P, C, K;
  void* out;                // [B,K,C]
  float* weights;           // [B,K,P]
};

template <typename T, int MODE>
struct TokenXform {
  const float* s0;   // sinkhorn: denom ; LN: mean
  const float* s1;   // LN: rstd
  const float* g;
  const float* bta;
  int lowp;
  __device__ __forceinline__ float operator()(int p, int c, float v) const {
    float r;
    if (MODE == MODE_SINKHORN) r = v / s0[p];
    else if (MODE == MODE_PATCHMERGER) r = g[c] * (s1[p] * (v - s0[p])) + bta[c];
    else r = v;
    return lowp ? bf16_round(r) : r;
  }
};

// out[k][c0 + ...] = sum_p W[k][p] * Xn[p][c]; thread (ty=warp, tx=lane): rows ty+8a (a<RK), cols c0+tx+32cc (cc<4)
template <int RK, typename T, typename TO, typename XF>
__device__ __forceinline__ void second_contraction(const T* __restrict__ xb, const float* W, int PS, float* xt, int P, int C,
                                                   int K, const XF& xf, TO* __restrict__ ob) {
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
  const bool vec_ok = stage_vec_ok(xb, C);
  for (int rbase = 0; rbase < K; rbase += 8 * RK) {
    int wrow[RK];
#pragma unroll
    for (int a = 0; a < RK; ++a) wrow[a] = min(rbase + ty + 8 * a, K - 1) * PS;
    for (int c0 = 0; c0 < C; c0 += CT) {
      float acc[RK][4];
#pragma unroll
      for (int a = 0; a < RK; ++a)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) acc[a][cc] = 0.f;
      for (int p0 = 0; p0 < P; p0 += PT) {
        __syncthreads();
        // stage Xn[p0 .. p0+PT) x [c0 .. c0+CT) (zero-filled outside), 4 columns per thread-step
        for (int e = tid; e < PT * (CT / 4); e += kThreads) {
          const int pp = e / (CT / 4), c4 = (e % (CT / 4)) * 4;
          const int p = p0 + pp, c = c0 + c4;
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          if (p < P) {
            const T* g = xb + (long long)p * C + c;
            if (vec_ok && c + 3 < C) {
              if (sizeof(T) == 4) {
                const float4 q4 = *reinterpret_cast<const float4*>(g);
                v[0] = q4.x; v[1] = q4.y; v[2] = q4.z; v[3] = q4.w;
              } else {
                const uint2 q2 = *reinterpret_cast<const uint2*>(g);
                const T* h = reinterpret_cast<const T*>(&q2);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = to_f32(h[i]);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = xf(p, c + i, v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = (c + i < C) ? xf(p, c + i, to_f32(g[i])) : 0.f;
            }
          }
          *reinterpret_cast<float4*>(xt + pp * CT + c4) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
        const int plim = min(PT, PS - p0);
#pragma unroll 1
        for (int pp = 0; pp < plim; pp += 4) {
          float xv[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) xv[i][cc] = xt[(pp + i) * CT + tx + 32 * cc];
#pragma unroll
          for (int a = 0; a < RK; ++a) {
            const float4 w4 = *reinterpret_cast<const float4*>(W + wrow[a] + p0 + pp);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              acc[a][cc] = fmaf(w4.x, xv[0][cc], acc[a][cc]);
              acc[a][cc] = fmaf(w4.y, xv[1][cc], acc[a][cc]);
              acc[a][cc] = fmaf(w4.z, xv[2][cc], acc[a][cc]);
              acc[a][cc] = fmaf(w4.w, xv[3][cc], acc[a][cc]);
            }
          }
        }
      }
#pragma unroll
      for (int a = 0; a < RK; ++a) {
        const int k = rbase + ty + 8 * a;
        if (k >= K) continue;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = c0 + tx + 32 * cc;
          if (c < C) ob[(long long)k * C + c] = from_f32<TO>(acc[a][cc]);
        }
      }
    }
  }
}

template <typename T, typename TO, int MODE>
__global__ void __launch_bounds__(kThreads, 1) soft_merge_kernel(SoftParams prm) {
  extern __shared__ __align__(16) float smem[];
  const int P = prm.P, C = prm.C, K = prm.K, lowp = prm.lowp;
  const int PS = (P + 3) & ~3;
  float* Z = smem;                       // [K][PS]
  float* at = Z + K * PS;                // [K][XS]   | second contraction: xt [PT][CT]
  float* bt = at + K * XS;               // [P][XS]
  float* tiles_end = at + max((K + P) * XS, PT * CT);
  float* s0 = tiles_end;                 // [P]
  float* s1 = s0 + P;                    // [P]
  float* uvec = s1 + P;                  // [K]
  float* vvec = uvec + K;                // [P]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* xb = reinterpret_cast<const T*>(prm.x) + (long long)b * prm.xbs;

  // ---- 1. per-token statistics
  if (MODE == MODE_SINKHORN) {
    for (int p = warp; p < P; p += kWarps) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) { float v = to_f32(xb[(long long)p * C + c]); s = fmaf(v, v, s); }
      s = warp_sum(s);
      if (lane == 0) s0[p] = fmaxf(sqrtf(s), 1e-12f);     // F.normalize: x / max(||x||, eps)
    }
  } else if (MODE == MODE_PATCHMERGER) {
    for (int p = warp; p < P; p += kWarps) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s += to_f32(xb[(long long)p * C + c]);
      const float mean = warp_sum(s) / (float)C;
      float q = 0.f;
      for (int c = lane; c < C; c += 32) { float d = to_f32(xb[(long long)p * C + c]) - mean; q = fmaf(d, d, q); }
      const float var = warp_sum(q) / (float)C;
      if (lane == 0) { s0[p] = mean; s1[p] = rsqrtf(var + prm.ln_eps); }
    }
  }
  // zero the pad columns of Z once (they feed the 4-wide W reads of the second contraction)
  for (int e = tid; e < K * (PS - P); e += kThreads) Z[(e / (PS - P)) * PS + P + e % (PS - P)] = 0.f;
  __syncthreads();
  const TokenXform<T, MODE> xf{s0, s1, prm.ln_w, prm.ln_b, lowp};

  // ---- 2. scores Z[k][p]
  if (MODE == MODE_SIT) {
    const float scale = prm.scale_ptr[0];
    const long long base = (long long)b * P * K;
    for (int e = tid; e < P * K; e += kThreads) {
      const int p = e / K, k = e % K;
      const float l = prm.logits_dtype == TOKRED_BF16
                          ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(prm.logits)[base + e])
                          : reinterpret_cast<const float*>(prm.logits)[base + e];
      Z[k * PS + p] = l * scale;
    }
    __syncthreads();
  } else {
    const bool qvec = stage_vec_ok(prm.q, C), xvec = stage_vec_ok(xb, C);
    const float post = prm.scale;
    gemm_nt(K, P, C, at, bt,
            [&](int k0) {
              stage_rows(prm.q, K, C, C, k0, at, qvec, [&](int, int, float v) { return lowp ? bf16_round(v) : v; });
              stage_rows(xb, P, C, C, k0, bt, xvec, xf);
            },
            [&](int k, int p, float acc) {
              float z = lowp ? bf16_round(acc) : acc;
              z *= post;
              Z[k * PS + p] = lowp ? bf16_round(z) : z;
            });
  }

  // ---- 3. W from Z
  if (MODE == MODE_SINKHORN) {
    const float nrm = prm.log_norm;
    for (int k = tid; k < K; k += kThreads) uvec[k] = 0.f;
    for (int p = tid; p < P; p += kThreads) vvec[p] = 0.f;
    __syncthreads();
    for (int it = 0; it < prm.iters; ++it) {
      // u_k = norm - logsumexp_p(Z_kp + v_p): one warp per row
      for (int k = warp; k < K; k += kWarps) {
        float m = -CUDART_INF_F;
        for (int p = lane; p < P; p += 32) m = fmaxf(m, Z[k * PS + p] + vvec[p]);
        m = warp_max(m);
        float s = 0.f;
        for (int p = lane; p < P; p += 32) s += expf(Z[k * PS + p] + vvec[p] - m);
        s = warp_sum(s);
        if (lane == 0) uvec[k] = nrm - (logf(s) + m);
      }
      __syncthreads();
      // v_p = norm - logsumexp_k(Z_kp + u_k): one thread per column (lanes on consecutive p: conflict-free)
      for (int p = tid; p < P; p += kThreads) {
        float m = -CUDART_INF_F;
        for (int k = 0; k < K; ++k) m = fmaxf(m, Z[k * PS + p] + uvec[k]);
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += expf(Z[k * PS + p] + uvec[k] - m);
        vvec[p] = nrm - (logf(s) + m);
      }
      __syncthreads();
    }
    float* wout = prm.weights + (long long)b * K * P;
    for (int k = warp; k < K; k += kWarps) {
      for (int p = lane; p < P; p += 32) {
        const float w = expf(((Z[k * PS + p] + uvec[k]) + vvec[p]) - nrm);
        wout[(long long)k * P + p] = w;
        Z[k * PS + p] = lowp ? bf16_round(w) : w;
      }
    }
  } else {
    float* wout = prm.weights + (long long)b * K * P;
    for (int k = warp; k < K; k += kWarps) {
      float m = -CUDART_INF_F;
      for (int p = lane; p < P; p += 32) m = fmaxf(m, Z[k * PS + p]);
      m = warp_max(m);
      float s = 0.f;
      for (int p = lane; p < P; p += 32) {
        const float e = expf(Z[k * PS + p] - m);
        Z[k * PS + p] = e;
        s += e;
      }
      s = warp_sum(s);
      for (int p = lane; p < P; p += 32) {
        const float w = Z[k * PS + p] / s;
        wout[(long long)k * P + p] = w;
        Z[k * PS + p] = lowp ? bf16_round(w) : w;
      }
    }
  }
  __syncthreads();

  // ---- 4. out = W . Xn
  TO* ob = reinterpret_cast<TO*>(prm.out) + (long long)b * K * C;
  if (K <= 64) second_contraction<8>(xb, Z, PS, at, P, C, K, xf, ob);
  else if (K <= 176) second_contraction<22>(xb, Z, PS, at, P, C, K, xf, ob);
  else second_contraction<26>(xb, Z, PS, at, P, C, K, xf, ob);
}

size_t soft_smem_bytes(int P, int C, int K) {
  const int PS = (P + 3) & ~3;
  const size_t tiles = (size_t)max((K + P) * XS, PT * CT);
  return ((size_t)K * PS + tiles + 3 * (size_t)P + K) * 4;
}

template <int MODE>
int launch_soft(const SoftParams& prm, int B, int x_dtype, int out_dtype, const char* what, void* stream) {
  const size_t smem = soft_smem_bytes(prm.P, prm.C, prm.K);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(T, TO)                                                              \
  do {                                                                             \
    if (int e = allow_smem(soft_merge_kernel<T, TO, MODE>, smem, what)) return e;  \
    soft_merge_kernel<T, TO, MODE><<<B, kThreads, smem, st>>>(prm);                \
  } while (0)
  if (x_dtype == TOKRED_F32 && out_dtype == TOKRED_F32) LAUNCH(float, float);
  else if (x_dtype == TOKRED_F32 && out_dtype == TOKRED_BF16) LAUNCH(float, __nv_bfloat16);
  else if (x_dtype == TOKRED_BF16 && out_dtype == TOKRED_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else LAUNCH(__nv_bfloat16, float);
#undef LAUNCH
  return finish_launch(what);
}

int check_soft(const char* what, int B, int P, int C, int K, int x_dtype, int out_dtype) {
  TOKRED_REQUIRE(valid_float_dtype(x_dtype) && valid_float_dtype(out_dtype), "%s: bad dtype", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1 && K >= 1, "%s: bad shape B=%d P=%d C=%d K=%d", what, B, P, C, K);
  if (P > kMaxP || K > kMaxK) {
    set_error("%s: P=%d / K=%d above %d is not supported (score matrix is kept in shared memory)", what, P, K, kMaxP);
    return TOKRED_ERR_UNSUPPORTED;
  }
  return TOKRED_OK;
}

}  // namespace
}  // namespace tokred

namespace tokred {
int launch_soft_merge_tc(int mode, const void* x, int x_dtype, const float* q, const float* ln_w, const float* ln_b,
                         int B, int P, int C, int K, float scale, float log_norm, float ln_eps, int iters, void* out,
                         float* weights, void* stream, const char* what, const void* logits, const float* scale_ptr,
                         long long xbs);
int launch_soft_merge_tc2(int mode, const void* x, int x_dtype, const float* q, const float* ln_w, const float* ln_b,
                          int B, int P, int C, int K, float scale, float log_norm, float ln_eps, int iters, void* out,
                          float* weights, void* stream, const char* what, const void* logits, const float* scale_ptr,
                          void* workspace, size_t workspace_bytes, long long xbs);
size_t soft_merge_tc2_workspace_bytes(int B, int P, int C, int K);
}
using namespace tokred;

extern "C" size_t tokred_soft_merge_workspace_bytes(int B, int P, int C, int K) {
  if (B <= 0 || P <= 0 || C <= 0 || K <= 0) return 0;
  return soft_merge_tc2_workspace_bytes(B, P, C, K);
}

extern "C" int tokred_sinkhorn_merge(const void* x, int x_dtype, int64_t x_batch_stride, const float* v_hat, int B, int P, int C, int K,
                                     float eps, float log_norm, int iters, int lowp, void* out, int out_dtype,
                                     float* weights, void* workspace, size_t workspace_bytes, void* stream) {
  const char* what = "tokred_sinkhorn_merge";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && v_hat && out && weights, "%s: null tensor", what);
  if (int e = check_soft(what, B, P, C, K, x_dtype, out_dtype)) return e;
  TOKRED_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (int64_t)P * C, "%s: x_batch_stride=%lld < P*C", what, (long long)x_batch_stride);
  const long long xbs = x_batch_stride ? (long long)x_batch_stride : (long long)P * C;
  TOKRED_REQUIRE(eps > 0.f && iters >= 0, "%s: eps=%g iters=%d", what, (double)eps, iters);
  if (B == 0) return TOKRED_OK;
  if (lowp == 1 && out_dtype == TOKRED_BF16) {     // bf16 autocast semantics on tcgen05 tensor cores
    const int rc2 = launch_soft_merge_tc2(MODE_SINKHORN, x, x_dtype, v_hat, nullptr, nullptr, B, P, C, K, 1.0f / eps, log_norm,
                                          0.f, iters, out, weights, stream, what, nullptr, nullptr, workspace, workspace_bytes, xbs);
    if (rc2 != 1) return rc2;
    const int rc = launch_soft_merge_tc(MODE_SINKHORN, x, x_dtype, v_hat, nullptr, nullptr, B, P, C, K, 1.0f / eps, log_norm,
                                        0.f, iters, out, weights, stream, what, nullptr, nullptr, xbs);
    if (rc != 1) return rc;
  }
  SoftParams prm{};
  prm.x = x; prm.xbs = xbs; prm.q = v_hat; prm.scale = 1.0f / eps; prm.log_norm = log_norm; prm.iters = iters; prm.lowp = lowp ? 1 : 0;
  prm.P = P; prm.C = C; prm.K = K; prm.out = out; prm.weights = weights;
  return launch_soft<MODE_SINKHORN>(prm, B, x_dtype, out_dtype, what, stream);
}

extern "C" int tokred_patchmerger(const void* x, int x_dtype, int64_t x_batch_stride, const float* ln_weight, const float* ln_bias,
                                  const float* queries, int B, int P, int C, int K, float scale, float ln_eps, int lowp,
                                  void* out, int out_dtype, float* attn, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  const char* what = "tokred_patchmerger";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && ln_weight && ln_bias && queries && out && attn, "%s: null tensor", what);
  if (int e = check_soft(what, B, P, C, K, x_dtype, out_dtype)) return e;
  TOKRED_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (int64_t)P * C, "%s: x_batch_stride=%lld < P*C", what, (long long)x_batch_stride);
  const long long xbs = x_batch_stride ? (long long)x_batch_stride : (long long)P * C;
  if (B == 0) return TOKRED_OK;
  if (lowp == 1 && out_dtype == TOKRED_BF16) {
    const int rc2 = launch_soft_merge_tc2(MODE_PATCHMERGER, x, x_dtype, queries, ln_weight, ln_bias, B, P, C, K, scale, 0.f,
                                          ln_eps, 0, out, attn, stream, what, nullptr, nullptr, workspace, workspace_bytes, xbs);
    if (rc2 != 1) return rc2;
    const int rc = launch_soft_merge_tc(MODE_PATCHMERGER, x, x_dtype, queries, ln_weight, ln_bias, B, P, C, K, scale, 0.f,
                                        ln_eps, 0, out, attn, stream, what, nullptr, nullptr, xbs);
    if (rc != 1) return rc;
  }
  SoftParams prm{};
  prm.x = x; prm.xbs = xbs; prm.q = queries; prm.ln_w = ln_weight; prm.ln_b = ln_bias; prm.scale = scale; prm.ln_eps = ln_eps;
  prm.lowp = lowp ? 1 : 0; prm.P = P; prm.C = C; prm.K = K; prm.out = out; prm.weights = attn;
  return launch_soft<MODE_PATCHMERGER>(prm, B, x_dtype, out_dtype, what, stream);
}

extern "C" int tokred_sit_merge(const void* x, int x_dtype, int64_t x_batch_stride, const void* logits, int logits_dtype, const float* scale,
                                int B, int P, int C, int K, int lowp, void* out, int out_dtype, float* weights,
                                void* workspace, size_t workspace_bytes, void* stream) {
  const char* what = "tokred_sit_merge";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && logits && scale && out && weights, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(logits_dtype), "%s: bad logits dtype", what);
  if (int e = check_soft(what, B, P, C, K, x_dtype, out_dtype)) return e;
  TOKRED_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (int64_t)P * C, "%s: x_batch_stride=%lld < P*C", what, (long long)x_batch_stride);
  const long long xbs = x_batch_stride ? (long long)x_batch_stride : (long long)P * C;
  if (B == 0) return TOKRED_OK;
  if (lowp == 1 && out_dtype == TOKRED_BF16 && logits_dtype == TOKRED_BF16) {
    // bulk-copy fed kernel first (needs the workspace); the scratch-free one covers callers without a workspace
    {
      const int rc2 = launch_soft_merge_tc2(MODE_SIT, x, x_dtype, nullptr, nullptr, nullptr, B, P, C, K, 1.f, 0.f, 0.f, 0, out,
                                            weights, stream, what, logits, scale, workspace, workspace_bytes, xbs);
      if (rc2 != 1) return rc2;
    }
    const int rc = launch_soft_merge_tc(MODE_SIT, x, x_dtype, nullptr, nullptr, nullptr, B, P, C, K, 1.f, 0.f, 0.f, 0, out,
                                        weights, stream, what, logits, scale, xbs);
    if (rc != 1) return rc;
  }
  SoftParams prm{};
  prm.x = x; prm.xbs = xbs; prm.logits = logits; prm.logits_dtype = logits_dtype; prm.scale_ptr = scale; prm.lowp = lowp ? 1 : 0;
  prm.P = P; prm.C = C; prm.K = K; prm.out = out; prm.weights = weights;
  return launch_soft<MODE_SIT>(prm, B, x_dtype, out_dtype, what, stream);
}
