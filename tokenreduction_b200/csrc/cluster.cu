// Clustering kernels: DPC-KNN (cluster + merge), K-Medoids (token weights + fit).
// Reference: models/dpcknn.py:44-140, models/kmedoids.py:62-85,240.
//
// dpcknn_cluster / kmedoids_fit / pairwise_dist keep the P x P distance matrix of an image (P <= 208 patches) in the
// shared memory of one CTA, so all the reference's [B,P,P] intermediates (3 for DPC-KNN, K*iters clones for
// K-Medoids = 895 launches) collapse into shared-memory passes; global traffic is x in, two index vectors out.
// D_ij = sqrt(max(|xi|^2 + |xj|^2 - 2 xi.xj, 1e-30)) is the matmul expansion torch.cdist uses for P > 25 (direct
// differences for P <= 25), so diag(D) and cancellation error behave like the reference.  D is
// bit-symmetric by construction, so every later pass reads COLUMNS (conflict-free with the odd row stride).
//
// Three Gram engines feed the same epilogues:
//   * dist_pipe.cuh  (default, P > 25): persistent warp-specialised tcgen05 pipeline, 3 x fp16-split MMAs, the
//     epilogue of image i overlaps the staging + MMAs of image i+1;
//   * FFMA           (exact_fp32 = 1): 13x7 register tiles, true fp32 products (gemm_nt.cuh), one CTA per image;
//   * direct         (P <= 25): staged rows, sqrt(sum (xi - xj)^2) like ATen's non-matmul path.
#include <math_constants.h>

#include "dist_pipe.cuh"
#include "gemm_nt.cuh"

namespace tokred {
namespace {

constexpr int kThreads = kGemmThreads;      // FFMA variants (gemm_nt.cuh is written for 256 threads)
constexpr int kWarps = kThreads / 32;
constexpr int kLightThreads = 512;           // direct-difference variants (P <= 25): no FFMA register tile
constexpr int kMaxP = pipe::kMaxP;           // P*P + staging must fit 227 KB of shared memory

struct BlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
struct BackSync { __device__ __forceinline__ void operator()() const { pipe::bar_sync(pipe::BAR_BACK, pipe::kBack); } };

// ------------------------------------------------------------------------------------------ legacy (one CTA / image)
struct DistCtx {
  float* D; int DS;
  float* xt;                 // FFMA path: [P][XS] staging tile
  float* sq;                 // [P]
  float* extra;              // kernel-specific vectors
};
__host__ __device__ inline size_t dist_region0_bytes(int P) {
  size_t d = ((size_t)P * (P | 1) * 4 + 15) & ~(size_t)15;
  if (P <= 25) d += (size_t)P * 257 * 4 + 16;      // direct path: staged rows [P][256+1]
  return d;
}
__host__ __device__ inline size_t dist_smem_bytes(int P, int extra_floats, int light) {
  size_t n = dist_region0_bytes(P);
  if (!light) n += (size_t)P * XS * 4;
  n += (((size_t)P + extra_floats) * 4 + 15) & ~(size_t)15;
  return n + 32;
}
__device__ __forceinline__ DistCtx dist_setup(float* smem, int P, int light) {
  DistCtx cx;
  cx.D = smem;
  cx.DS = P | 1;
  float* after = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem) + dist_region0_bytes(P));
  cx.xt = after;
  if (!light) after += P * XS;
  cx.sq = after;
  cx.extra = after + P;
  return cx;
}

__device__ __forceinline__ void row_sqnorms(const float* __restrict__ xb, int P, int C, float* sq) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15u) == 0);
  for (int i = warp; i < P; i += nwarps) {
    const float* row = xb + (long long)i * C;
    float s = 0.f;
    if (vec) {
      for (int k = lane * 4; k < C; k += 128) {
        const float4 v = *reinterpret_cast<const float4*>(row + k);
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
      }
    } else {
      for (int k = lane; k < C; k += 32) { float v = row[k]; s = fmaf(v, v, s); }
    }
    s = warp_sum(s);
    if (lane == 0) sq[i] = s;
  }
}

// Fills cx.D (shared) with the pairwise distances of the P rows of xb (global, [P][C]) times post_scale.
// LIGHT = true: direct differences only (P <= 25, 512 threads); LIGHT = false: FFMA Gram (or direct when P <= 25).
template <bool LIGHT>
__device__ __forceinline__ void pairdist_to_smem(const float* __restrict__ xb, int P, int C, DistCtx& cx, float post_scale) {
  const int tid = threadIdx.x;
  float* D = cx.D;
  const int DS = cx.DS;
  if (P <= 25) {
    // direct form (ATen's non-matmul cdist path): sqrt(sum (xi - xj)^2).  The <= 25 rows are staged through shared
    // memory in 256-column chunks (the D region is free until the end; pair sums live in registers).
    constexpr int CH = 256;
    float* xs = D + (((P * DS + 3) & ~3));          // [P][CH+1] after the (tiny) D matrix, inside region 0
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;               // up to 3 pairs per thread (P*P <= 625 <= 3*256)
    const int nthr = (int)blockDim.x;
    for (int k0 = 0; k0 < C; k0 += CH) {
      const int kn = min(CH, C - k0);
      __syncthreads();
      for (int e = tid; e < P * kn; e += nthr) xs[(e / kn) * (CH + 1) + e % kn] = xb[(long long)(e / kn) * C + k0 + e % kn];
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int e = tid + u * nthr;
        if (e < P * P) {
          const float* a = xs + (e / P) * (CH + 1);
          const float* c = xs + (e % P) * (CH + 1);
          float s = u == 0 ? s0 : (u == 1 ? s1 : s2);
          for (int k = 0; k < kn; ++k) { const float d = a[k] - c[k]; s = fmaf(d, d, s); }
          if (u == 0) s0 = s; else if (u == 1) s1 = s; else s2 = s;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int e = tid + u * nthr;
      if (e < P * P) D[(e / P) * DS + e % P] = sqrtf(u == 0 ? s0 : (u == 1 ? s1 : s2)) * post_scale;
    }
    __syncthreads();
    return;
  }
  if constexpr (!LIGHT) {
    row_sqnorms(xb, P, C, cx.sq);
    const bool vec_ok = stage_vec_ok(xb, C);
    float* xt = cx.xt;
    const float* sq = cx.sq;
    gemm_nt(P, P, C, xt, xt,
            [&](int k0) { stage_rows(xb, P, C, C, k0, xt, vec_ok, [](int, int, float v) { return v; }); },
            [&](int i, int j, float g) {
              const float d2 = (sq[i] + sq[j]) - 2.0f * g;
              D[i * DS + j] = sqrtf(fmaxf(d2, 1e-30f)) * post_scale;
            });
  }
}

// ------------------------------------------------------------------------------------------ epilogues
// Run by a group of NT threads (a whole CTA, or the BACK group of the pipeline) on a finished D in shared memory;
// `tid` is the index inside the group, `sync` its barrier.

template <int KS, int NT>
__device__ __forceinline__ void knn_density(const float* D, int DS, int P, int knn, const float* __restrict__ noise_b, float* rho,
                                            int tid, float& lmax) {
  for (int base = 0; base < 2 * P; base += NT) {
    const int it = base + tid;
    const bool act = it < 2 * P;
    const int i = act ? it >> 1 : 0, h = it & 1;
    const int j0 = h ? (P + 1) / 2 : 0, j1 = act ? (h ? P : (P + 1) / 2) : 0;
    float nb[KS];
#pragma unroll
    for (int t = 0; t < KS; ++t) nb[t] = CUDART_INF_F;
#pragma unroll 8
    for (int j = j0; j < j1; ++j) {
      const float cur = D[j * DS + i];
      lmax = fmaxf(lmax, cur);
      // sorted insert without a dependency chain: new[t] = min(old[t], max(old[t-1], cur)), highest slot first
#pragma unroll
      for (int t = KS - 1; t > 0; --t) nb[t] = fminf(nb[t], fmaxf(nb[t - 1], cur));
      nb[0] = fminf(nb[0], cur);
    }
    float pv[KS];
#pragma unroll
    for (int t = 0; t < KS; ++t) pv[t] = __shfl_xor_sync(0xffffffffu, nb[t], 1);
    if (act && h == 0) {
#pragma unroll
      for (int q = 0; q < KS; ++q) {
        const float cur = pv[q];
#pragma unroll
        for (int t = KS - 1; t > 0; --t) nb[t] = fminf(nb[t], fmaxf(nb[t - 1], cur));
        nb[0] = fminf(nb[0], cur);
      }
      float sumsq = 0.f;
#pragma unroll
      for (int t = 0; t < KS; ++t)
        if (t < knn) sumsq += nb[t] * nb[t];
      rho[i] = expf(-(sumsq * (1.0f / (float)knn))) + noise_b[i] * 1e-6f;
    }
  }
}

// DPC-KNN (models/dpcknn.py:62-98).  extra: rho[P], score[P], centre[K] (int), red[NT/32].
__host__ __device__ constexpr int dpc_extra_floats(int P, int K) { return 2 * P + K + 16; }
template <int NT, class Sync>
__device__ __forceinline__ void dpc_epilogue(const float* D, int DS, int P, int K, int knn, const float* __restrict__ noise_b,
                                             float* extra, int tid, Sync sync, int64_t* __restrict__ idx_cluster_b,
                                             int64_t* __restrict__ index_down_b, int img = 0) {
  float* rho = extra;
  float* score = rho + P;
  int* centre = reinterpret_cast<int*>(score + P);    // [K]
  float* red = reinterpret_cast<float*>(centre + K);   // [NT / 32]
  const int warp = tid >> 5, lane = tid & 31;

  // local density from the knn nearest (self included): exp(-mean(d^2)) + 1e-6 * U
  float lmax = 0.f;
  if (knn <= 8) {
    // Two threads (adjacent lanes) per token, one per half of its column: each keeps its KS smallest values sorted in
    // registers with a BRANCH-FREE min/max insertion (2 FMNMX per slot; the sum of squares only needs the multiset of
    // the knn smallest values, so ties need no index), then the even lane inserts the odd lane's list.  The first
    // version took a divergent branch into an 8-step compare-swap chain: some lane of the warp needs it on almost
    // every row, so every row paid for it (7.3 M warp instructions, 16-20 us of a 40 us epilogue in the stamps).
    if (knn <= 5) knn_density<5, NT>(D, DS, P, knn, noise_b, rho, tid, lmax);
    else knn_density<8, NT>(D, DS, P, knn, noise_b, rho, tid, lmax);
  } else {
    for (int i = tid; i < P; i += NT) {
      float sumsq = 0.f;
      float prev_v = -1.f;
      int prev_j = -1;
      for (int t = 0; t < knn; ++t) {
        float best = CUDART_INF_F;
        int bj = -1;
        for (int j = 0; j < P; ++j) {
          const float v = D[j * DS + i];
          const bool after = (v > prev_v) || (v == prev_v && j > prev_j);
          if (after && v < best) { best = v; bj = j; }
        }
        sumsq += best * best;
        prev_v = best; prev_j = bj;
      }
      for (int j = 0; j < P; ++j) lmax = fmaxf(lmax, D[j * DS + i]);
      rho[i] = expf(-(sumsq * (1.0f / (float)knn))) + noise_b[i] * 1e-6f;
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  sync();
  TOKRED_STAMP(tid == 0, img, 11);
  float dmax = red[0];
#pragma unroll
  for (int w = 1; w < NT / 32; ++w) dmax = fmaxf(dmax, red[w]);

  // The three passes below give every token to TWO adjacent lanes (one half of the scan each, combined with one
  // shuffle): 2P tasks keep all warps of the group busy and halve the serial scan length.
  // distance to the nearest denser token (or the global max), centre score
  for (int base = 0; base < 2 * P; base += NT) {
    const int it = base + tid;
    const bool act = it < 2 * P;
    const int i = act ? it >> 1 : 0, h = it & 1;
    const int j0 = h ? (P + 1) / 2 : 0, j1 = act ? (h ? P : (P + 1) / 2) : 0;
    const float ri = rho[i];
    float b0 = dmax, b1 = dmax, b2 = dmax, b3 = dmax;      // four independent minima: the fmin chain was the critical path
    int j = j0;
    for (; j + 3 < j1; j += 4) {
      b0 = fminf(b0, rho[j] > ri ? D[j * DS + i] : dmax);
      b1 = fminf(b1, rho[j + 1] > ri ? D[(j + 1) * DS + i] : dmax);
      b2 = fminf(b2, rho[j + 2] > ri ? D[(j + 2) * DS + i] : dmax);
      b3 = fminf(b3, rho[j + 3] > ri ? D[(j + 3) * DS + i] : dmax);
    }
    for (; j < j1; ++j) b0 = fminf(b0, rho[j] > ri ? D[j * DS + i] : dmax);
    float best = fminf(fminf(b0, b1), fminf(b2, b3));
    best = fminf(best, __shfl_xor_sync(0xffffffffu, best, 1));
    if (act && h == 0) score[i] = best * ri;
  }
  sync();
  TOKRED_STAMP(tid == 0, img, 12);
  for (int base = 0; base < 2 * P; base += NT) {
    const int it = base + tid;
    const bool act = it < 2 * P;
    const int i = act ? it >> 1 : 0, h = it & 1;
    int rk = rank_desc_range(score, h ? (P + 1) / 2 : 0, act ? (h ? P : (P + 1) / 2) : 0, i);
    rk += __shfl_xor_sync(0xffffffffu, rk, 1);
    if (act && h == 0 && rk < K) { centre[rk] = i; index_down_b[rk] = i; }
  }
  sync();
  TOKRED_STAMP(tid == 0, img, 13);
  // nearest centre (lowest k on ties); centres belong to their own cluster
  for (int base = 0; base < 2 * P; base += NT) {
    const int it = base + tid;
    const bool act = it < 2 * P;
    const int i = act ? it >> 1 : 0, h = it & 1;
    const int k0 = h ? (K + 1) / 2 : 0, k1 = act ? (h ? K : (K + 1) / 2) : 0;
    float best = CUDART_INF_F;
    int bk = 0x7fffffff, own = -1;
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
      const int ck = centre[k];
      const float v = D[ck * DS + i];
      if (v < best) { best = v; bk = k; }
      if (ck == i) own = k;
    }
    const float ob = __shfl_xor_sync(0xffffffffu, best, 1);
    const int ok = __shfl_xor_sync(0xffffffffu, bk, 1), oo = __shfl_xor_sync(0xffffffffu, own, 1);
    if (act && h == 0) {
      if (ob < best || bk == 0x7fffffff) bk = ok;             // the upper half only wins when strictly nearer
      if (bk == 0x7fffffff) bk = 0;
      own = own >= 0 ? own : oo;
      idx_cluster_b[i] = own >= 0 ? own : bk;
    }
  }
}

// K-Medoids (models/kmedoids.py:62-85).  extra: w[P], S[P], assign[P] (int), centre[K] (int), best_key[K] (u64).
__host__ __device__ constexpr int kmed_extra_floats(int P, int K) { return 3 * P + K + 2 + 2 * K + 2; }
template <int NT, class Sync>
__device__ __forceinline__ void kmed_epilogue(const float* D, int DS, int P, int C, int K, int iters, const float* __restrict__ tw_b,
                                              const float* __restrict__ xb, float* extra, int tid, Sync sync,
                                              float* __restrict__ centres_b, int64_t* __restrict__ cidx_b,
                                              int64_t* __restrict__ assign_b, const int64_t* __restrict__ init_b = nullptr) {
  float* w = extra;
  float* S = w + P;
  int* assign = reinterpret_cast<int*>(S + P);   // [P]
  int* centre = assign + P;                       // [K]
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < P; i += NT) w[i] = tw_b ? tw_b[i] : 1.f;      // no weights: the equal_weight variant (:61)
  sync();
  // S_i = sum_j (D_ij * w_i); initial centres = top-K token weights (descending, lowest index on ties)
  for (int i = tid; i < P; i += NT) {
    const float wi = w[i];
    float s = 0.f;
    for (int j = 0; j < P; ++j) s += __fmul_rn(D[j * DS + i], wi);
    S[i] = s;
    const int rk = rank_desc(w, P, i);
    if (rk < K) centre[rk] = i;
  }
  sync();
  if (init_b) {        // caller-supplied initial medoids (the equal_weight farthest-point initialisation, :43-59)
    for (int k = tid; k < K; k += NT) centre[k] = clamp_idx(init_b[k], P);
    sync();
  }
  // Medoid update: c_k = argmin_{i: assign_i = k} S_i with the lowest i on ties, empty cluster -> token 0 (every
  // masked row sums to P * 1e6 > any S).  One shared-memory atomicMin per token on the key (S bits, i):
  // S > 0, so the fp32 bit pattern orders like the value -- replaces K serial scans of P by 49 threads.
  unsigned long long* best_key =                                                                  // [K], 8-byte aligned
      reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(centre + K) + 7) & ~(uintptr_t)7);
  for (int it = 0; it <= iters; ++it) {
    for (int base = 0; base < 2 * P; base += NT) {          // two adjacent lanes per token, half of the centres each
      const int t2 = base + tid;
      const bool act = t2 < 2 * P;
      const int i = act ? t2 >> 1 : 0, h = t2 & 1;
      const int k0 = h ? (K + 1) / 2 : 0, k1 = act ? (h ? K : (K + 1) / 2) : 0;
      float best = CUDART_INF_F;
      int bk = 0x7fffffff;
#pragma unroll 4
      for (int k = k0; k < k1; ++k) {
        const float v = D[centre[k] * DS + i];
        if (v < best) { best = v; bk = k; }
      }
      const float ob = __shfl_xor_sync(0xffffffffu, best, 1);
      const int ok = __shfl_xor_sync(0xffffffffu, bk, 1);
      if (act && h == 0) {
        if (ob < best || bk == 0x7fffffff) bk = ok;
        assign[i] = bk == 0x7fffffff ? 0 : bk;
      }
    }
    for (int k = tid; k < K; k += NT) best_key[k] = ~0ull;
    sync();
    if (it == iters) break;
    for (int i = tid; i < P; i += NT)
      atomicMin(&best_key[assign[i]], ((unsigned long long)__float_as_uint(S[i]) << 32) | (unsigned)i);
    sync();
    for (int k = tid; k < K; k += NT) centre[k] = best_key[k] == ~0ull ? 0 : (int)(best_key[k] & 0xffffffffu);
    sync();
  }
  for (int i = tid; i < P; i += NT) assign_b[i] = assign[i];
  for (int k = tid; k < K; k += NT) cidx_b[k] = centre[k];
  // medoid rows verbatim
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(centres_b) & 15u) == 0);
  for (int k = warp; k < K; k += NT / 32) {
    const float* src = xb + (long long)centre[k] * C;
    if (vec) warp_copy_row16(centres_b + (long long)k * C, src, C * 4, lane);
    else warp_copy_row_elems(centres_b + (long long)k * C, src, C, lane);
  }
}

// ------------------------------------------------------------------------------------------ legacy kernels
template <bool LIGHT>
__global__ void __launch_bounds__(LIGHT ? kLightThreads : kThreads, 1)
pairwise_dist_kernel(const float* __restrict__ x, int P, int C, float post_scale, float* __restrict__ out) {
  constexpr int NT = LIGHT ? kLightThreads : kThreads;
  extern __shared__ __align__(128) float smem[];
  DistCtx cx = dist_setup(smem, P, LIGHT);
  pairdist_to_smem<LIGHT>(x + (long long)blockIdx.x * P * C, P, C, cx, post_scale);
  float* ob = out + (long long)blockIdx.x * P * P;
  for (int i = threadIdx.x >> 5; i < P; i += NT / 32)
    for (int j = threadIdx.x & 31; j < P; j += 32) ob[i * P + j] = cx.D[i * cx.DS + j];
}

template <bool LIGHT>
__global__ void __launch_bounds__(LIGHT ? kLightThreads : kThreads, 1)
dpcknn_cluster_kernel(const float* __restrict__ x, long long xbs, const float* __restrict__ noise_u, int P, int C, int K, int knn,
                      float inv_sqrt_c, int64_t* __restrict__ idx_cluster, int64_t* __restrict__ index_down) {
  constexpr int NT = LIGHT ? kLightThreads : kThreads;
  extern __shared__ __align__(128) float smem[];
  DistCtx cx = dist_setup(smem, P, LIGHT);
  const int b = blockIdx.x;
  pairdist_to_smem<LIGHT>(x + (long long)b * xbs, P, C, cx, inv_sqrt_c);
  __syncthreads();
  dpc_epilogue<NT>(cx.D, cx.DS, P, K, knn, noise_u + (long long)b * P, cx.extra, (int)threadIdx.x, BlockSync(),
                   idx_cluster + (long long)b * P, index_down + (long long)b * K);
}

template <bool LIGHT>
__global__ void __launch_bounds__(LIGHT ? kLightThreads : kThreads, 1)
kmedoids_fit_kernel(const float* __restrict__ x, long long xbs, const float* __restrict__ token_weight,
                    const int64_t* __restrict__ init_idx, int P, int C, int K, int iters,
                    float* __restrict__ centres, int64_t* __restrict__ cluster_idx, int64_t* __restrict__ assignment) {
  constexpr int NT = LIGHT ? kLightThreads : kThreads;
  extern __shared__ __align__(128) float smem[];
  DistCtx cx = dist_setup(smem, P, LIGHT);
  const int b = blockIdx.x;
  const float* xb = x + (long long)b * xbs;
  pairdist_to_smem<LIGHT>(xb, P, C, cx, 1.0f);
  __syncthreads();
  kmed_epilogue<NT>(cx.D, cx.DS, P, C, K, iters, token_weight ? token_weight + (long long)b * P : nullptr, xb, cx.extra,
                    (int)threadIdx.x, BlockSync(), centres + (long long)b * K * C, cluster_idx + (long long)b * K,
                    assignment + (long long)b * P, init_idx ? init_idx + (long long)b * K : nullptr);
}

// ------------------------------------------------------------------------------------------ pipelined kernels
struct PipeParams {
  const float* x; long long xbs; int B, P, C; float post_scale;
  const float* noise; int K, knn; int64_t* idx_cluster; int64_t* index_down;                    // DPC-KNN
  const float* tw; const int64_t* init; int iters; float* centres; int64_t* cidx; int64_t* assign;   // K-Medoids
  float* out;                                                                                   // plain cdist
};
enum { EPI_PLAIN = 0, EPI_DPC = 1, EPI_KMED = 2 };

template <int EPI>
__global__ void __launch_bounds__(pipe::kThreads, 1) dist_pipe_kernel(const PipeParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int extra = EPI == EPI_DPC ? dpc_extra_floats(p.P, p.K) : (EPI == EPI_KMED ? kmed_extra_floats(p.P, p.K) : 0);
  pipe::Ctx cx = pipe::setup(smem_raw, p.B, p.P, p.C, extra);
  if ((int)threadIdx.x < pipe::kLoad) {
    pipe::loader_run(p.x, p.xbs, cx);
  } else if ((int)threadIdx.x < pipe::kFront) {
    pipe::mma_run(cx);
  } else {
    const int bt = (int)threadIdx.x - pipe::kFront;
    for (int it = 0; it < cx.n_img; ++it) {
      const long long b = (long long)blockIdx.x + (long long)it * gridDim.x;
      pipe::back_drain(cx, it, p.post_scale);
      if constexpr (EPI == EPI_PLAIN) {
        float* ob = p.out + b * p.P * p.P;
        for (int i = bt >> 5; i < p.P; i += pipe::kBack / 32)
          for (int j = bt & 31; j < p.P; j += 32) ob[i * p.P + j] = cx.D[i * cx.DS + j];
      } else if constexpr (EPI == EPI_DPC) {
        dpc_epilogue<pipe::kBack>(cx.D, cx.DS, p.P, p.K, p.knn, p.noise + b * p.P, cx.extra, bt, BackSync(),
                                  p.idx_cluster + b * p.P, p.index_down + b * p.K, it);
      } else {
        kmed_epilogue<pipe::kBack>(cx.D, cx.DS, p.P, p.C, p.K, p.iters, p.tw ? p.tw + b * p.P : nullptr, p.x + b * p.xbs, cx.extra,
                                   bt, BackSync(), p.centres + b * (long long)p.K * p.C, p.cidx + b * p.K, p.assign + b * p.P,
                                   p.init ? p.init + b * p.K : nullptr);
      }
      pipe::bar_sync(pipe::BAR_BACK, pipe::kBack);       // D and the epilogue vectors are free for the next image
      TOKRED_STAMP(bt == 0, it, 14);
    }
  }
  pipe::teardown(cx);
}

// ------------------------------------------------------------------------------------------ DPC-KNN merge
// grid (splits, B).  Member lists of ALL clusters are built at once by a stable counting sort (rank of a token inside
// its cluster = number of earlier tokens of the same cluster: ascending token order = the order of CPU index_add_,
// models/dpcknn.py:122-131, so the unfused multiply/add accumulation is bit-identical).  A warp then treats the
// member rows of its clusters (k = gw, gw + W, ...) as ONE stream and keeps MB rows in flight across cluster
// boundaries: the rows of a cluster are spread over the image, and a large cluster (30-50 tokens on random data)
// is otherwise a chain of dependent HBM latencies that sets the kernel time (ncu: SMs active 47 % of the duration).
// Measured alternatives that were slower: (cluster, 128-channel slice) work items (each row is then fetched by three
// warps at different times: 55 us vs 35 us), largest-first assignment of clusters to warps (no gain: at B=256 the
// time is set by the SMs that hold two images, ~31 GB/s per SM).
// CPL = 16-byte chunks per lane (0: scalar fallback).
template <int CPL>
__global__ void __launch_bounds__(kThreads, 2)
dpcknn_merge_kernel(const float* __restrict__ x, long long xbs, const int64_t* __restrict__ idx_token,
                    const float* __restrict__ agg_weight, const int64_t* __restrict__ idx_cluster,
                    const float* __restrict__ token_weight, int P, int C, int K, int T, float* __restrict__ x_merged,
                    int64_t* __restrict__ idx_token_new, float* __restrict__ agg_weight_new) {
  extern __shared__ float smem[];
  float* nw = smem;                                   // [P] token weights, then normalised weights
  float* wsum = nw + P;                               // [K]
  int* cl = reinterpret_cast<int*>(wsum + K);         // [P] cluster of token
  int* members = cl + P;                              // [P] tokens sorted by (cluster, token)
  int* off = members + P;                             // [K + 1] start of each cluster's list
  int* rk = off + K + 1;                              // [P] rank of a token inside its cluster
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < P; i += kThreads) {
    int c = (int)idx_cluster[(long long)b * P + i];
    cl[i] = c < 0 ? 0 : (c >= K ? K - 1 : c);
    nw[i] = token_weight ? token_weight[(long long)b * P + i] : 1.f;
  }
  for (int k = tid; k <= K; k += kThreads) off[k] = 0;
  __syncthreads();
  for (int i = tid; i < P; i += kThreads) {
    const int c = cl[i];
    int before = 0, total = 0;
    for (int j = 0; j < P; ++j) {
      const int same = cl[j] == c;
      total += same;
      before += same & (j < i);
    }
    rk[i] = before;
    if (before == total - 1) off[c + 1] = total;      // the last member publishes the cluster size
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int k = 0; k < K; ++k) { run += off[k + 1]; off[k + 1] = run; }
  }
  __syncthreads();
  for (int i = tid; i < P; i += kThreads) members[off[cl[i]] + rk[i]] = i;
  __syncthreads();
  for (int k = tid; k < K; k += kThreads) {
    float s = 0.f;
    for (int j = off[k]; j < off[k + 1]; ++j) s += nw[members[j]];
    wsum[k] = s + 1e-6f;
  }
  __syncthreads();
  for (int i = tid; i < P; i += kThreads) nw[i] = nw[i] / wsum[cl[i]];
  __syncthreads();

  const float* xb = x + (long long)b * xbs;
  float* ob = x_merged + (long long)b * K * C;
  const int nchunks = C / 4;
  const int gw = blockIdx.x * kWarps + warp, stride = gridDim.x * kWarps;
  if constexpr (CPL > 0) {
    constexpr int MB = 20 / CPL > 8 ? 8 : 20 / CPL;      // rows in flight: MB * CPL 16-byte registers per lane (<= 20)
    float acc[CPL][4];
#pragma unroll
    for (int q = 0; q < CPL; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f; }
    int k = K, knext = gw, j = 0, jend = 0;
    auto advance = [&]() {
      k = knext < K ? knext : K;
      knext += stride;
      if (k < K) { j = off[k]; jend = off[k + 1]; }
    };
    advance();
    while (k < K) {
      int tok[MB], kk[MB];
      bool last[MB];
#pragma unroll
      for (int u = 0; u < MB; ++u) {
        while (k < K && j == jend) {                   // cluster finished, or empty: an empty one still owns a zero row
          if (off[k + 1] == off[k]) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              const int c = lane + 32 * q;
              if (c < nchunks) st_stream16(ob + (long long)k * C + c * 4, make_int4(0, 0, 0, 0));
            }
          }
          advance();
        }
        if (k < K) {
          tok[u] = members[j]; kk[u] = k; ++j; last[u] = j == jend;
        } else {
          tok[u] = -1; kk[u] = 0; last[u] = false;
        }
      }
      int4 raw[MB][CPL];
#pragma unroll
      for (int u = 0; u < MB; ++u) {
        if (tok[u] >= 0) {
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const int c = lane + 32 * q;
            if (c < nchunks) raw[u][q] = ld_stream16(xb + (long long)tok[u] * C + c * 4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < MB; ++u) {
        if (tok[u] >= 0) {
          const float wgt = nw[tok[u]];
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const float* v = reinterpret_cast<const float*>(&raw[u][q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[q][e] = __fadd_rn(acc[q][e], __fmul_rn(v[e], wgt));
          }
          if (last[u]) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              const int c = lane + 32 * q;
              if (c < nchunks) st_stream16(ob + (long long)kk[u] * C + c * 4, *reinterpret_cast<const int4*>(acc[q]));
              acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
            }
          }
        }
      }
    }
  } else {
    for (int k = gw; k < K; k += stride) {
      for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
        for (int j = off[k]; j < off[k + 1]; ++j) {
          const int i = members[j];
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
// out[b,p] = sum_h sum_q attn[b,h,q,nt+p]  (models/kmedoids.py:240).  The [H,N,N] block of an image is one flat run
// of H*N rows of N elements, so the column of flat element e is (e mod N).  Threads stream ALIGNED 16-byte vectors
// of that run; a "group" is Tg = N / gcd(N, VE) threads, whose combined stride VE*Tg is a multiple of N — so a
// thread hits the same VE columns in every iteration and keeps VE register accumulators.  G groups interleave
// iterations.  Combine: G*VE rounds in which the Tg threads of one group add one accumulator each to DISTINCT
// columns of a shared array (plain adds, fixed order => deterministic, no atomics).
template <typename T>
__global__ void __launch_bounds__(1024)
attn_colsum_kernel(const T* __restrict__ attn, int H, int N, int nt, int Tg, int G, float* __restrict__ out) {
  constexpr int VE = 16 / sizeof(T);
  extern __shared__ float colsum[];     // [N]
  const int b = blockIdx.x, tid = threadIdx.x;
  const long long len = (long long)H * N * N;
  const T* base = attn + (long long)b * len;
  // first 16-byte aligned element of this image's run
  const int mis = (int)((reinterpret_cast<uintptr_t>(base) & 15u) / sizeof(T));
  const int head = mis ? VE - mis : 0;            // scalar prologue elements [0, head)
  for (int c = tid; c < N; c += blockDim.x) colsum[c] = 0.f;
  const int grp = tid / Tg, t = tid % Tg;
  float acc[VE];
#pragma unroll
  for (int i = 0; i < VE; ++i) acc[i] = 0.f;
  const bool live = grp < G;
  const long long nvec = (len - head) / VE;       // aligned vectors
  if (live) {
    long long v = (long long)grp * Tg + t;
    const long long step = (long long)G * Tg;
    for (; v + 3 * step < nvec; v += 4 * step) {  // 4 independent 16-byte loads in flight
      int4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = ld_stream16(base + head + (v + u * step) * VE);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const T* e = reinterpret_cast<const T*>(&r[u]);
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[i] += to_f32(e[i]);
      }
    }
    for (; v < nvec; v += step) {
      const int4 r = ld_stream16(base + head + v * VE);
      const T* e = reinterpret_cast<const T*>(&r);
#pragma unroll
      for (int i = 0; i < VE; ++i) acc[i] += to_f32(e[i]);
    }
  }
  __syncthreads();
  // combine (group-major, accumulator-minor order); columns of one (group, i) round are all distinct
  const int col0 = (int)(((long long)head + (long long)t * VE) % N);
  for (int g = 0; g < G; ++g)
    for (int i = 0; i < VE; ++i) {
      if (live && grp == g) colsum[(col0 + i) % N] += acc[i];
      __syncthreads();
    }
  // scalar prologue / epilogue elements (at most 2*(VE-1) per image)
  if (tid == 0) {
    for (int e = 0; e < head && e < len; ++e) colsum[e % N] += to_f32(base[e]);
    for (long long e = head + nvec * VE; e < len; ++e) colsum[(int)(e % N)] += to_f32(base[e]);
  }
  __syncthreads();
  for (int c = nt + tid; c < N; c += blockDim.x) out[(long long)b * (N - nt) + (c - nt)] = colsum[c];
}

}  // namespace
}  // namespace tokred

using namespace tokred;
TOKRED_STAMP_SETTER(cluster)

// tensor cores for every P that takes ATen's matmul form (P > 25), unless the caller asks for the exact-fp32 FFMA path
static int pick_pipe(int P, int exact_fp32) { return (P > 25 && !exact_fp32) ? 1 : 0; }

template <int EPI>
static int launch_pipe(const PipeParams& p, int extra_floats, const char* what, cudaStream_t st) {
  const size_t smem = pipe::smem_bytes(p.P, extra_floats);
  if (int e = allow_smem(dist_pipe_kernel<EPI>, smem, what)) return e;
  const int grid = p.B < kNumSMs ? p.B : kNumSMs;       // persistent: one CTA per SM walks its images
  dist_pipe_kernel<EPI><<<grid, pipe::kThreads, smem, st>>>(p);
  return finish_launch(what);
}

extern "C" int tokred_pairwise_dist(const float* x, int B, int P, int C, float post_scale, int exact_fp32, float* out,
                                    void* stream) {
  const char* what = "tokred_pairwise_dist";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && out, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  if (P > kMaxP) { set_error("%s: P=%d > %d patches is not supported (distance matrix is kept in shared memory)", what, P, kMaxP); return TOKRED_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  if (pick_pipe(P, exact_fp32)) {
    PipeParams p{};
    p.x = x; p.xbs = (long long)P * C; p.B = B; p.P = P; p.C = C; p.post_scale = post_scale; p.out = out;
    return launch_pipe<EPI_PLAIN>(p, 0, what, st);
  }
  const int light = P <= 25;
  const size_t smem = dist_smem_bytes(P, 0, light);
  if (light) {
    if (int e = allow_smem(pairwise_dist_kernel<true>, smem, what)) return e;
    pairwise_dist_kernel<true><<<B, kLightThreads, smem, st>>>(x, P, C, post_scale, out);
  } else {
    if (int e = allow_smem(pairwise_dist_kernel<false>, smem, what)) return e;
    pairwise_dist_kernel<false><<<B, kThreads, smem, st>>>(x, P, C, post_scale, out);
  }
  return finish_launch(what);
}

extern "C" int tokred_dpcknn_cluster(const float* x, int64_t x_batch_stride, const float* noise_u, int B, int P, int C, int K, int knn,
                                     int exact_fp32, int64_t* idx_cluster, int64_t* index_down, void* stream) {
  const char* what = "tokred_dpcknn_cluster";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && noise_u && idx_cluster && index_down, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  TOKRED_REQUIRE(K >= 1 && K <= P, "%s: cluster_num=%d outside [1, P=%d]", what, K, P);
  TOKRED_REQUIRE(knn >= 1 && knn <= P, "%s: k=%d outside [1, P=%d]", what, knn, P);
  if (P > kMaxP) { set_error("%s: P=%d > %d patches is not supported (distance matrix is kept in shared memory)", what, P, kMaxP); return TOKRED_ERR_UNSUPPORTED; }
  TOKRED_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (int64_t)P * C, "%s: x_batch_stride=%lld < P*C", what, (long long)x_batch_stride);
  const long long xbs = x_batch_stride ? (long long)x_batch_stride : (long long)P * C;
  const float inv = 1.0f / (float)sqrt((double)C);     // CUDA tensor / python-scalar = multiply by fp32 reciprocal
  cudaStream_t st = (cudaStream_t)stream;
  if (pick_pipe(P, exact_fp32)) {
    PipeParams p{};
    p.x = x; p.xbs = xbs; p.B = B; p.P = P; p.C = C; p.post_scale = inv;
    p.noise = noise_u; p.K = K; p.knn = knn; p.idx_cluster = idx_cluster; p.index_down = index_down;
    return launch_pipe<EPI_DPC>(p, dpc_extra_floats(P, K), what, st);
  }
  const int light = P <= 25;
  const size_t smem = dist_smem_bytes(P, dpc_extra_floats(P, K), light);
  if (light) {
    if (int e = allow_smem(dpcknn_cluster_kernel<true>, smem, what)) return e;
    dpcknn_cluster_kernel<true><<<B, kLightThreads, smem, st>>>(x, xbs, noise_u, P, C, K, knn, inv, idx_cluster, index_down);
  } else {
    if (int e = allow_smem(dpcknn_cluster_kernel<false>, smem, what)) return e;
    dpcknn_cluster_kernel<false><<<B, kThreads, smem, st>>>(x, xbs, noise_u, P, C, K, knn, inv, idx_cluster, index_down);
  }
  return finish_launch(what);
}

static int launch_kmedoids_fit(const char* what, const float* x, int64_t x_batch_stride, const float* token_weight,
                               const int64_t* init_idx, int B, int P, int C, int K, int iters, int exact_fp32, float* centres,
                               int64_t* cluster_idx, int64_t* assignment, void* stream) {
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && (token_weight || init_idx) && centres && cluster_idx && assignment, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  TOKRED_REQUIRE(K >= 1 && K <= P, "%s: cluster_num=%d outside [1, P=%d]", what, K, P);
  TOKRED_REQUIRE(iters >= 0, "%s: iters=%d < 0", what, iters);
  if (P > kMaxP) { set_error("%s: P=%d > %d patches is not supported (distance matrix is kept in shared memory)", what, P, kMaxP); return TOKRED_ERR_UNSUPPORTED; }
  TOKRED_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (int64_t)P * C, "%s: x_batch_stride=%lld < P*C", what, (long long)x_batch_stride);
  const long long xbs = x_batch_stride ? (long long)x_batch_stride : (long long)P * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (pick_pipe(P, exact_fp32)) {
    PipeParams p{};
    p.x = x; p.xbs = xbs; p.B = B; p.P = P; p.C = C; p.post_scale = 1.0f;
    p.tw = token_weight; p.init = init_idx; p.K = K; p.iters = iters; p.centres = centres; p.cidx = cluster_idx; p.assign = assignment;
    return launch_pipe<EPI_KMED>(p, kmed_extra_floats(P, K), what, st);
  }
  const int light = P <= 25;
  const size_t smem = dist_smem_bytes(P, kmed_extra_floats(P, K), light);
  if (light) {
    if (int e = allow_smem(kmedoids_fit_kernel<true>, smem, what)) return e;
    kmedoids_fit_kernel<true><<<B, kLightThreads, smem, st>>>(x, xbs, token_weight, init_idx, P, C, K, iters, centres, cluster_idx, assignment);
  } else {
    if (int e = allow_smem(kmedoids_fit_kernel<false>, smem, what)) return e;
    kmedoids_fit_kernel<false><<<B, kThreads, smem, st>>>(x, xbs, token_weight, init_idx, P, C, K, iters, centres, cluster_idx, assignment);
  }
  return finish_launch(what);
}

extern "C" int tokred_kmedoids_fit(const float* x, int64_t x_batch_stride, const float* token_weight, int B, int P, int C, int K, int iters,
                                   int exact_fp32, float* centres, int64_t* cluster_idx, int64_t* assignment, void* stream) {
  const char* what = "tokred_kmedoids_fit";
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(token_weight, "%s: null token_weight", what);
  return launch_kmedoids_fit(what, x, x_batch_stride, token_weight, nullptr, B, P, C, K, iters, exact_fp32, centres, cluster_idx,
                             assignment, stream);
}

extern "C" int tokred_kmedoids_fit_init(const float* x, int64_t x_batch_stride, const float* token_weight, const int64_t* init_idx,
                                        int B, int P, int C, int K, int iters, int exact_fp32, float* centres,
                                        int64_t* cluster_idx, int64_t* assignment, void* stream) {
  const char* what = "tokred_kmedoids_fit_init";
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(init_idx, "%s: null init_idx", what);
  return launch_kmedoids_fit(what, x, x_batch_stride, token_weight, init_idx, B, P, C, K, iters, exact_fp32, centres, cluster_idx,
                             assignment, stream);
}

extern "C" int tokred_dpcknn_merge(const float* x, int64_t x_batch_stride, const int64_t* idx_token, const float* agg_weight,
                                   const int64_t* idx_cluster, const float* token_weight, int B, int P, int C, int K,
                                   int T, float* x_merged, int64_t* idx_token_new, float* agg_weight_new, void* stream) {
  const char* what = "tokred_dpcknn_merge";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && idx_token && agg_weight && idx_cluster && x_merged && idx_token_new && agg_weight_new,
                 "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1 && K >= 1 && T >= 0, "%s: bad shape", what);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (int64_t)P * C, "%s: x_batch_stride=%lld < P*C", what, (long long)x_batch_stride);
  const long long xbs = x_batch_stride ? (long long)x_batch_stride : (long long)P * C;
  const size_t smem = (size_t)(4 * P + 2 * K + 1) * 4;
  const bool vec = (C % 4 == 0) && (xbs % 4 == 0) && aligned16(x) && aligned16(x_merged);
  const int cpl = vec ? ceil_div(C / 4, 32) : 0;
  int splits = (2 * kNumSMs) / B;            // one wave of resident CTAs (2 per SM; see tokred_tome_merge)
  splits = max(1, min(splits, ceil_div(K, kWarps)));
  dim3 grid(splits, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(CPL)                                                                                                  \
  do {                                                                                                               \
    if (int e = allow_smem(dpcknn_merge_kernel<CPL>, smem, what)) return e;                                          \
    dpcknn_merge_kernel<CPL><<<grid, kThreads, smem, st>>>(x, xbs, idx_token, agg_weight, idx_cluster, token_weight, P, C, K, \
                                                           T, x_merged, idx_token_new, agg_weight_new);             \
  } while (0)
  switch (cpl) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    case 5: case 6: LAUNCH(6); break;
    default: LAUNCH(0); break;
  }
#undef LAUNCH
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
  // group size: smallest thread count whose combined vector stride is a multiple of N
  const int ve = attn_dtype == TOKRED_F32 ? 4 : 8;
  int gcd = N, a = ve;
  while (a) { const int tmp = gcd % a; gcd = a; a = tmp; }
  const int Tg = N / gcd;
  TOKRED_REQUIRE(Tg <= 1024, "%s: N=%d needs a %d-thread group (> 1024)", what, N, Tg);
  const int G = 1024 / Tg > 8 ? 8 : 1024 / Tg;
  const int threads = ((G * Tg + 31) / 32) * 32;
  const size_t smem = (size_t)N * 4;
  if (attn_dtype == TOKRED_F32)
    attn_colsum_kernel<float><<<B, threads, smem, (cudaStream_t)stream>>>((const float*)attn, H, N, num_tokens, Tg, G, out);
  else
    attn_colsum_kernel<__nv_bfloat16><<<B, threads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)attn, H, N,
                                                                                  num_tokens, Tg, G, out);
  return finish_launch(what);
}
