// Clustering kernels: DPC-KNN (cluster + merge), K-Medoids (token weights + fit).
// Reference: models/dpcknn.py:44-140, models/kmedoids.py:62-85,240.
//
// dpcknn_cluster / kmedoids_fit: ONE persistent CTA per image.  The P x P distance matrix (P <= 196 patches)
// never leaves the SM: X is streamed through a [P][32] shared-memory tile, the Gram matrix is accumulated in
// registers (13x7 outputs per thread, LDS.128 operands) and written to shared memory as
// D_ij = sqrt(max(|xi|^2 + |xj|^2 - 2 xi.xj, 1e-30)) — the matmul expansion torch.cdist uses for P > 25
// (direct differences for P <= 25), so that diag(D) and cancellation error behave like the reference
// (SURVEY.md A.7).  D is bit-symmetric by construction, so every later pass reads COLUMNS (conflict-free).
// All the reference's [B,P,P] intermediates (3 for DPC-KNN, K*iters clones for K-Medoids = 895 launches)
// collapse into shared-memory passes of this one kernel; global traffic is x in, two index vectors out.
#include <math_constants.h>

#include "gemm_nt.cuh"

namespace tokred {
namespace {

constexpr int kThreads = kGemmThreads;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxP = 208;    // P*P + staging must fit 227 KB of shared memory

// Fills D[P*P] (shared) with the pairwise distances of the P rows of xb (global, [P][C]) times post_scale.
// xt: [P][XS] staging, sq: [P] squared norms.  All threads of the CTA must call.
__device__ void pairdist_to_smem(const float* __restrict__ xb, int P, int C, float* D, float* xt, float* sq,
                                 float post_scale) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (P <= 25) {
    // direct form (ATen's non-matmul cdist path): sqrt(sum (xi - xj)^2)
    for (int e = tid; e < P * P; e += kThreads) {
      const int i = e / P, j = e % P;
      const float* a = xb + (long long)i * C;
      const float* c = xb + (long long)j * C;
      float s = 0.f;
      for (int k = 0; k < C; ++k) { float d = a[k] - c[k]; s = fmaf(d, d, s); }
      D[e] = sqrtf(s) * post_scale;
    }
    __syncthreads();
    return;
  }
  for (int i = warp; i < P; i += kWarps) {
    const float* row = xb + (long long)i * C;
    float s = 0.f;
    for (int k = lane; k < C; k += 32) { float v = row[k]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) sq[i] = s;
  }
  const bool vec_ok = stage_vec_ok(xb, C);
  gemm_nt(P, P, C, xt, xt,
          [&](int k0) { stage_rows(xb, P, C, C, k0, xt, vec_ok, [](int, int, float v) { return v; }); },
          [&](int i, int j, float g) {
            const float d2 = (sq[i] + sq[j]) - 2.0f * g;
            D[i * P + j] = sqrtf(fmaxf(d2, 1e-30f)) * post_scale;
          });
}

// ------------------------------------------------------------------------------------------ plain cdist(x, x)
__global__ void __launch_bounds__(kThreads, 1)
pairwise_dist_kernel(const float* __restrict__ x, int P, int C, float post_scale, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* D = smem;
  float* xt = D + ((P * P + 3) & ~3);   // keep the tile 16-byte aligned for LDS.128
  float* sq = xt + P * XS;
  pairdist_to_smem(x + (long long)blockIdx.x * P * C, P, C, D, xt, sq, post_scale);
  float* ob = out + (long long)blockIdx.x * P * P;
  for (int e = threadIdx.x; e < P * P; e += kThreads) ob[e] = D[e];
}

// ------------------------------------------------------------------------------------------ DPC-KNN cluster
__global__ void __launch_bounds__(kThreads, 1)
dpcknn_cluster_kernel(const float* __restrict__ x, const float* __restrict__ noise_u, int P, int C, int K, int knn,
                      float inv_sqrt_c, int64_t* __restrict__ idx_cluster, int64_t* __restrict__ index_down) {
  extern __shared__ __align__(16) float smem[];
  float* D = smem;
  float* xt = D + ((P * P + 3) & ~3);   // keep the tile 16-byte aligned for LDS.128
  float* sq = xt + P * XS;
  float* rho = sq + P;
  float* score = rho + P;
  int* centre = reinterpret_cast<int*>(score + P);   // [K]
  float* red = reinterpret_cast<float*>(centre + K);  // [kWarps]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  pairdist_to_smem(x + (long long)b * P * C, P, C, D, xt, sq, inv_sqrt_c);

  // local density from the knn nearest (self included): exp(-mean(d^2)) + 1e-6 * U
  float lmax = 0.f;
  for (int i = tid; i < P; i += kThreads) {
    float prev_v = -1.f;
    int prev_j = -1;
    float sumsq = 0.f;
    for (int t = 0; t < knn; ++t) {
      float best = CUDART_INF_F;
      int bj = -1;
      for (int j = 0; j < P; ++j) {
        const float v = D[j * P + i];
        const bool after = (v > prev_v) || (v == prev_v && j > prev_j);
        if (after && v < best) { best = v; bj = j; }
      }
      sumsq += best * best;
      prev_v = best; prev_j = bj;
    }
    rho[i] = expf(-(sumsq * (1.0f / (float)knn))) + noise_u[(long long)b * P + i] * 1e-6f;
    for (int j = 0; j < P; ++j) lmax = fmaxf(lmax, D[j * P + i]);
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float dmax = red[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) dmax = fmaxf(dmax, red[w]);

  // distance to the nearest denser token (or the global max), centre score
  for (int i = tid; i < P; i += kThreads) {
    const float ri = rho[i];
    float best = dmax;
    for (int j = 0; j < P; ++j) {
      const float v = rho[j] > ri ? D[j * P + i] : dmax;
      best = fminf(best, v);
    }
    score[i] = best * ri;
  }
  __syncthreads();
  for (int i = tid; i < P; i += kThreads) {
    const int rk = rank_desc(score, P, i);
    if (rk < K) { centre[rk] = i; index_down[(long long)b * K + rk] = i; }
  }
  __syncthreads();
  // nearest centre (lowest k on ties); centres belong to their own cluster
  for (int i = tid; i < P; i += kThreads) {
    float best = CUDART_INF_F;
    int bk = 0;
    for (int k = 0; k < K; ++k) {
      const float v = D[centre[k] * P + i];
      if (v < best) { best = v; bk = k; }
    }
    for (int k = 0; k < K; ++k)
      if (centre[k] == i) bk = k;
    idx_cluster[(long long)b * P + i] = bk;
  }
}

// ------------------------------------------------------------------------------------------ K-Medoids fit
__global__ void __launch_bounds__(kThreads, 1)
kmedoids_fit_kernel(const float* __restrict__ x, const float* __restrict__ token_weight, int P, int C, int K, int iters,
                    float* __restrict__ centres, int64_t* __restrict__ cluster_idx, int64_t* __restrict__ assignment) {
  extern __shared__ __align__(16) float smem[];
  float* D = smem;
  float* xt = D + ((P * P + 3) & ~3);   // keep the tile 16-byte aligned for LDS.128
  float* sq = xt + P * XS;
  float* w = sq + P;
  float* S = w + P;
  int* assign = reinterpret_cast<int*>(S + P);   // [P]
  int* centre = assign + P;                       // [K]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xb = x + (long long)b * P * C;

  for (int i = tid; i < P; i += kThreads) w[i] = token_weight[(long long)b * P + i];
  pairdist_to_smem(xb, P, C, D, xt, sq, 1.0f);

  // S_i = sum_j (D_ij * w_i); initial centres = top-K token weights (descending, lowest index on ties)
  for (int i = tid; i < P; i += kThreads) {
    const float wi = w[i];
    float s = 0.f;
    for (int j = 0; j < P; ++j) s += __fmul_rn(D[j * P + i], wi);
    S[i] = s;
    const int rk = rank_desc(w, P, i);
    if (rk < K) centre[rk] = i;
  }
  __syncthreads();
  const float big = 1.0e6f * (float)P;     // P masked columns of 1e6 sum exactly in fp32
  for (int it = 0; it <= iters; ++it) {
    for (int i = tid; i < P; i += kThreads) {
      float best = CUDART_INF_F;
      int bk = 0;
      for (int k = 0; k < K; ++k) {
        const float v = D[centre[k] * P + i];
        if (v < best) { best = v; bk = k; }
      }
      assign[i] = bk;
    }
    __syncthreads();
    if (it == iters) break;
    for (int k = tid; k < K; k += kThreads) {
      float best = CUDART_INF_F;
      int bi = 0;
      for (int i = 0; i < P; ++i) {
        const float v = assign[i] == k ? S[i] : big;
        if (v < best) { best = v; bi = i; }
      }
      centre[k] = bi;
    }
    __syncthreads();
  }
  for (int i = tid; i < P; i += kThreads) assignment[(long long)b * P + i] = assign[i];
  for (int k = tid; k < K; k += kThreads) cluster_idx[(long long)b * K + k] = centre[k];
  // medoid rows verbatim
  float* cb = centres + (long long)b * K * C;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(cb) & 15u) == 0);
  for (int k = warp; k < K; k += kWarps) {
    const float* src = xb + (long long)centre[k] * C;
    if (vec) warp_copy_row16(cb + (long long)k * C, src, C * 4, lane);
    else warp_copy_row_elems(cb + (long long)k * C, src, C, lane);
  }
}

// ------------------------------------------------------------------------------------------ DPC-KNN merge
// grid (splits, B); one warp per cluster row.  Members are accumulated in ascending token order with unfused
// multiply/add: the order and rounding of CPU index_add_ (models/dpcknn.py:122-131).
__global__ void __launch_bounds__(kThreads)
dpcknn_merge_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx_token,
                    const float* __restrict__ agg_weight, const int64_t* __restrict__ idx_cluster,
                    const float* __restrict__ token_weight, int P, int C, int K, int T, float* __restrict__ x_merged,
                    int64_t* __restrict__ idx_token_new, float* __restrict__ agg_weight_new, int vec) {
  extern __shared__ __align__(16) float smem[];
  float* nw = smem;                                   // [P] normalised weights
  float* wsum = nw + P;                               // [K]
  int* cl = reinterpret_cast<int*>(wsum + K);         // [P] cluster of token
  int* cnt = cl + P;                                  // [K]
  int* offs = cnt + K;                                // [K]
  int* member = offs + K;                             // [P] tokens grouped by cluster, ascending
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < P; i += kThreads) {
    int c = (int)idx_cluster[(long long)b * P + i];
    cl[i] = c < 0 ? 0 : (c >= K ? K - 1 : c);
    nw[i] = token_weight ? token_weight[(long long)b * P + i] : 1.f;
  }
  __syncthreads();
  for (int k = tid; k < K; k += kThreads) {
    float s = 0.f;
    int c = 0;
    for (int i = 0; i < P; ++i)
      if (cl[i] == k) { s += nw[i]; ++c; }
    wsum[k] = s + 1e-6f;
    cnt[k] = c;
  }
  __syncthreads();
  for (int k = tid; k < K; k += kThreads) {
    int o = 0;
    for (int q = 0; q < k; ++q) o += cnt[q];
    offs[k] = o;
    for (int i = 0; i < P; ++i)
      if (cl[i] == k) member[o++] = i;
  }
  __syncthreads();
  for (int i = tid; i < P; i += kThreads) nw[i] = nw[i] / wsum[cl[i]];
  __syncthreads();

  const float* xb = x + (long long)b * P * C;
  float* ob = x_merged + (long long)b * K * C;
  for (int k = blockIdx.x * kWarps + warp; k < K; k += gridDim.x * kWarps) {
    const int n = cnt[k];
    const int* mem = member + offs[k];
    if (vec) {
      for (int c = lane * 4; c < C; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int m = 0; m < n; ++m) {
          const int i = mem[m];
          const float wgt = nw[i];
          int4 raw = ld_stream16(xb + (long long)i * C + c);
          const float* v = reinterpret_cast<const float*>(&raw);
          acc.x = __fadd_rn(acc.x, __fmul_rn(v[0], wgt));
          acc.y = __fadd_rn(acc.y, __fmul_rn(v[1], wgt));
          acc.z = __fadd_rn(acc.z, __fmul_rn(v[2], wgt));
          acc.w = __fadd_rn(acc.w, __fmul_rn(v[3], wgt));
        }
        st_stream16(ob + (long long)k * C + c, *reinterpret_cast<const int4*>(&acc));
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
        for (int m = 0; m < n; ++m) {
          const int i = mem[m];
          acc = __fadd_rn(acc, __fmul_rn(xb[(long long)i * C + c], nw[i]));
        }
        ob[(long long)k * C + c] = acc;
      }
    }
  }
  if (blockIdx.x == 0) {
    for (int t = tid; t < T; t += kThreads) {
      long long it = idx_token[(long long)b * T + t];
      it = it < 0 ? 0 : (it >= P ? P - 1 : it);
      idx_token_new[(long long)b * T + t] = cl[it];
      agg_weight_new[(long long)b * T + t] = agg_weight[(long long)b * T + t] * nw[it];
    }
  }
}

// ------------------------------------------------------------------------------------------ attention column sums
// out[b,p] = sum_q ( sum_h attn[b,h,q,nt+p] ): heads first, then query rows (the reference's two torch.sum calls).
// One CTA per image, 1024 threads = 4 row-phases x 256 columns; lanes run along columns -> coalesced 128-byte
// rows; the H loads of one (q, column) are independent -> H requests in flight per thread.
template <typename T>
__global__ void __launch_bounds__(1024)
attn_colsum_kernel(const T* __restrict__ attn, int H, int N, int nt, float* __restrict__ out) {
  __shared__ float part[4][256];
  const int b = blockIdx.y, tid = threadIdx.x, ph = tid >> 8, lc = tid & 255;
  const int col = blockIdx.x * 256 + lc;
  const T* ab = attn + (long long)b * H * N * N;
  float acc = 0.f;
  if (col < N) {
    for (int q = ph; q < N; q += 4) {
      float t = 0.f;
#pragma unroll 4
      for (int h = 0; h < H; ++h) t += to_f32(ab[((long long)h * N + q) * N + col]);
      acc += t;
    }
  }
  part[ph][lc] = acc;
  __syncthreads();
  if (ph == 0 && col < N && col >= nt) {
    const float s = ((part[0][lc] + part[1][lc]) + part[2][lc]) + part[3][lc];
    out[(long long)b * (N - nt) + (col - nt)] = s;
  }
}

}  // namespace
}  // namespace tokred

using namespace tokred;

static size_t dist_smem_bytes(int P, int extra_floats) {
  return ((((size_t)P * P + 3) & ~(size_t)3) + (size_t)P * XS + P + extra_floats) * 4;
}

extern "C" int tokred_pairwise_dist(const float* x, int B, int P, int C, float post_scale, float* out, void* stream) {
  const char* what = "tokred_pairwise_dist";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && out, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  if (P > kMaxP) { set_error("%s: P=%d > %d patches is not supported (distance matrix is kept in shared memory)", what, P, kMaxP); return TOKRED_ERR_UNSUPPORTED; }
  if (B == 0) return TOKRED_OK;
  const size_t smem = dist_smem_bytes(P, 0);
  if (int e = allow_smem(pairwise_dist_kernel, smem, what)) return e;
  pairwise_dist_kernel<<<B, kThreads, smem, (cudaStream_t)stream>>>(x, P, C, post_scale, out);
  return finish_launch(what);
}

extern "C" int tokred_dpcknn_cluster(const float* x, const float* noise_u, int B, int P, int C, int K, int knn,
                                     int64_t* idx_cluster, int64_t* index_down, void* stream) {
  const char* what = "tokred_dpcknn_cluster";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && noise_u && idx_cluster && index_down, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  TOKRED_REQUIRE(K >= 1 && K <= P, "%s: cluster_num=%d outside [1, P=%d]", what, K, P);
  TOKRED_REQUIRE(knn >= 1 && knn <= P, "%s: k=%d outside [1, P=%d]", what, knn, P);
  if (P > kMaxP) { set_error("%s: P=%d > %d patches is not supported (distance matrix is kept in shared memory)", what, P, kMaxP); return TOKRED_ERR_UNSUPPORTED; }
  if (B == 0) return TOKRED_OK;
  const size_t smem = dist_smem_bytes(P, 2 * P + K + kWarps);
  if (int e = allow_smem(dpcknn_cluster_kernel, smem, what)) return e;
  const float inv = 1.0f / (float)sqrt((double)C);     // CUDA tensor / python-scalar = multiply by fp32 reciprocal
  dpcknn_cluster_kernel<<<B, kThreads, smem, (cudaStream_t)stream>>>(x, noise_u, P, C, K, knn, inv, idx_cluster, index_down);
  return finish_launch(what);
}

extern "C" int tokred_kmedoids_fit(const float* x, const float* token_weight, int B, int P, int C, int K, int iters,
                                   float* centres, int64_t* cluster_idx, int64_t* assignment, void* stream) {
  const char* what = "tokred_kmedoids_fit";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && token_weight && centres && cluster_idx && assignment, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  TOKRED_REQUIRE(K >= 1 && K <= P, "%s: cluster_num=%d outside [1, P=%d]", what, K, P);
  TOKRED_REQUIRE(iters >= 0, "%s: iters=%d < 0", what, iters);
  if (P > kMaxP) { set_error("%s: P=%d > %d patches is not supported (distance matrix is kept in shared memory)", what, P, kMaxP); return TOKRED_ERR_UNSUPPORTED; }
  if (B == 0) return TOKRED_OK;
  const size_t smem = dist_smem_bytes(P, 3 * P + K);
  if (int e = allow_smem(kmedoids_fit_kernel, smem, what)) return e;
  kmedoids_fit_kernel<<<B, kThreads, smem, (cudaStream_t)stream>>>(x, token_weight, P, C, K, iters, centres, cluster_idx,
                                                                   assignment);
  return finish_launch(what);
}

extern "C" int tokred_dpcknn_merge(const float* x, const int64_t* idx_token, const float* agg_weight,
                                   const int64_t* idx_cluster, const float* token_weight, int B, int P, int C, int K,
                                   int T, float* x_merged, int64_t* idx_token_new, float* agg_weight_new, void* stream) {
  const char* what = "tokred_dpcknn_merge";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && idx_token && agg_weight && idx_cluster && x_merged && idx_token_new && agg_weight_new,
                 "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1 && K >= 1 && T >= 0, "%s: bad shape", what);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (B == 0) return TOKRED_OK;
  const size_t smem = (size_t)(3 * P + 3 * K) * 4;
  if (int e = allow_smem(dpcknn_merge_kernel, smem, what)) return e;
  const int vec = (C % 4 == 0) && aligned16(x) && aligned16(x_merged);
  int splits = ceil_div(4 * kNumSMs, B);
  splits = max(1, min(splits, ceil_div(K, kWarps)));
  dpcknn_merge_kernel<<<dim3(splits, B), kThreads, smem, (cudaStream_t)stream>>>(
      x, idx_token, agg_weight, idx_cluster, token_weight, P, C, K, T, x_merged, idx_token_new, agg_weight_new, vec);
  return finish_launch(what);
}

extern "C" int tokred_attn_colsum(const void* attn, int attn_dtype, int B, int H, int N, int num_tokens, float* out,
                                  void* stream) {
  const char* what = "tokred_attn_colsum";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(attn && out, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(attn_dtype), "%s: bad dtype %d", what, attn_dtype);
  TOKRED_REQUIRE(B >= 0 && H >= 1 && N >= 1 && num_tokens >= 0 && num_tokens < N, "%s: bad shape", what);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (B == 0) return TOKRED_OK;
  dim3 grid(ceil_div(N, 256), B);
  if (attn_dtype == TOKRED_F32)
    attn_colsum_kernel<float><<<grid, 1024, 0, (cudaStream_t)stream>>>((const float*)attn, H, N, num_tokens, out);
  else
    attn_colsum_kernel<__nv_bfloat16><<<grid, 1024, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)attn, H, N,
                                                                               num_tokens, out);
  return finish_launch(what);
}
