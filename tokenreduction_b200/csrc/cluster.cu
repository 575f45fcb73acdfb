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
#include "umma.cuh"

namespace tokred {
namespace {

constexpr int kThreads = kGemmThreads;      // FFMA variants (gemm_nt.cuh is written for 256 threads)
constexpr int kWarps = kThreads / 32;
constexpr int kTcThreads = 512;              // tensor-core variants: no FFMA register tile -> 16 warps per CTA
constexpr int kMaxP = 208;    // P*P + staging must fit 227 KB of shared memory
constexpr int KT = 32;        // tf32 path: contraction columns per stage

// Shared-memory plan of the distance kernels.  D has an odd row stride so that both row-per-thread stores (the
// TMEM epilogue) and column reads (every later pass) are bank-conflict free.  On the tensor-core path the two
// operand stages alias D's region: D is only written after the last MMA has completed.
struct DistCtx {
  float* D; int DS;
  float* xt;                 // FFMA path: [P][XS] staging tile
  unsigned char* stage;      // tensor-core path: 2 stages x {hi, lo} canonical tf32 tiles
  float* sq;                 // [P]
  float* extra;              // kernel-specific vectors
  uint64_t* bars;            // [2]
  uint32_t* tmem_slot;
  uint32_t tmem_base, tmem_cols;
  int use_tc;
};

__host__ __device__ inline size_t dist_stage_bytes(int P) { return (size_t)((P + 127) / 128) * 16 * 1024 * 2; }   // hi + lo
__host__ __device__ inline size_t dist_region0_bytes(int P, int use_tc) {
  size_t d = ((size_t)P * (P | 1) * 4 + 15) & ~(size_t)15;
  if (P <= 25) d += (size_t)P * 257 * 4 + 16;      // direct path: staged rows [P][256+1]
  const size_t st = use_tc ? 2 * dist_stage_bytes(P) : 0;
  return d > st ? d : st;
}
__host__ __device__ inline size_t dist_smem_bytes(int P, int extra_floats, int use_tc) {
  size_t n = dist_region0_bytes(P, use_tc);
  if (!use_tc) n += (size_t)P * XS * 4;
  n += (((size_t)P + extra_floats) * 4 + 15) & ~(size_t)15;
  return n + 32;
}

// carve the dynamic shared memory, and (tensor-core path) allocate TMEM + init the mbarriers.  All threads call.
__device__ __forceinline__ DistCtx dist_setup(float* smem, int P, int extra_floats, int use_tc) {
  DistCtx cx;
  cx.use_tc = use_tc;
  cx.D = smem;
  cx.DS = P | 1;
  cx.stage = reinterpret_cast<unsigned char*>(smem);
  float* after = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem) + dist_region0_bytes(P, use_tc));
  cx.xt = after;
  if (!use_tc) after += P * XS;
  cx.sq = after;
  cx.extra = after + P;
  cx.bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(after) + ((((size_t)P + extra_floats) * 4 + 15) & ~(size_t)15));
  cx.tmem_slot = reinterpret_cast<uint32_t*>(cx.bars + 2);
  cx.tmem_base = 0;
  const int Np = (P + 15) & ~15;
  cx.tmem_cols = P > 128 ? 512u : umma::tmem_cols_pow2((uint32_t)Np);
  if (use_tc) {
    if ((threadIdx.x >> 5) == 0) umma::tmem_alloc(cx.tmem_slot, cx.tmem_cols);
    if (threadIdx.x == 0) { umma::mbar_init(&cx.bars[0], 1); umma::mbar_init(&cx.bars[1], 1); umma::fence_mbar_init(); }
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    cx.tmem_base = *cx.tmem_slot;
  }
  return cx;
}
__device__ __forceinline__ void dist_teardown(const DistCtx& cx) {
  if (cx.use_tc) {
    umma::tc_fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) umma::tmem_dealloc(cx.tmem_base, cx.tmem_cols);
  }
}

// round-to-nearest tf32, returned in an fp32 container (low 13 mantissa bits zero)
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
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

// Gram matrix on tcgen05 with 3xTF32 error compensation: x = hi + lo (hi = tf32(x), lo = x - hi exactly),
// G ~= hi.hi^T + hi.lo^T + lo.hi^T accumulated in one fp32 TMEM accumulator — the same accuracy class as the fp32
// matmul torch.cdist runs (plain tf32 would be 1e-3 and flip neighbour decisions).  Both MMA operands are the SAME
// shared-memory tile (A = rows of an M tile, B = rows 0..Np-1), written by the threads in the canonical K-major
// layout (4 tf32 per 16-byte core row); two stages alias the D region.
__device__ void pairdist_tc(const float* __restrict__ xb, int P, int C, DistCtx& cx, float post_scale) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthreads = blockDim.x, nwarps = nthreads >> 5;
  row_sqnorms(xb, P, C, cx.sq);
  const int n_mt = (P + 127) / 128, Np = (P + 15) & ~15;
  const size_t stage_bytes = dist_stage_bytes(P), half = stage_bytes / 2;
  const uint32_t sbo = (KT / 4) * 128;      // 1024
  const uint32_t idesc = umma::instr_desc(umma::FMT_TF32, 128, (uint32_t)Np);
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15u) == 0);
  const int nchunk = (C + KT - 1) / KT;
  const int ng = ((P + 7) / 8) * 64;        // (8 rows) x (8 core columns of 4 floats) per row group
  // software pipeline: the loads of chunk c+1 are issued right after the barrier of chunk c and stay in registers while
  // the MMAs of chunk c run; split + store happen one iteration later
  constexpr int GMAX = 8;                   // float4 groups per thread: (208/8)*64 / 256 threads = 6.5
  float4 v[GMAX];
  auto load_chunk = [&](int c) {
    const int k0 = c * KT;
#pragma unroll
    for (int u = 0; u < GMAX; ++u) {
      const int gI = tid + u * nthreads;
      const int row = (gI & 7) + ((gI >> 6) << 3), k = k0 + ((gI >> 3) & 7) * 4;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gI < ng && row < P) {
        const float* g = xb + (long long)row * C + k;
        if (vec && k + 3 < C) v[u] = *reinterpret_cast<const float4*>(g);
        else { if (k < C) v[u].x = g[0]; if (k + 1 < C) v[u].y = g[1]; if (k + 2 < C) v[u].z = g[2]; if (k + 3 < C) v[u].w = g[3]; }
      }
    }
  };
  load_chunk(0);
  for (int c = 0; c < nchunk; ++c) {
    const int st = c & 1;
    unsigned char* hi = cx.stage + (size_t)st * stage_bytes;
    unsigned char* lo = hi + half;
    if (c >= 2) umma::mbar_wait(&cx.bars[st], (uint32_t)(((c - 2) >> 1) & 1));
#pragma unroll
    for (int u = 0; u < GMAX; ++u) {
      const int gI = tid + u * nthreads;
      const int row = (gI & 7) + ((gI >> 6) << 3), core = (gI >> 3) & 7;
      if (gI < ng) {
        float4 h, l;
        h.x = to_tf32(v[u].x); h.y = to_tf32(v[u].y); h.z = to_tf32(v[u].z); h.w = to_tf32(v[u].w);
        // the remainder is exact in fp32; round it to tf32 ourselves (the MMA would TRUNCATE it: a biased error)
        l.x = to_tf32(v[u].x - h.x); l.y = to_tf32(v[u].y - h.y); l.z = to_tf32(v[u].z - h.z); l.w = to_tf32(v[u].w - h.w);
        const uint32_t off = (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 16u + (uint32_t)core * 128u;
        *reinterpret_cast<float4*>(hi + off) = h;
        *reinterpret_cast<float4*>(lo + off) = l;
      }
    }
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after_sync();
      const uint32_t h0 = umma::smem_u32(hi), l0 = umma::smem_u32(lo);
      for (int mt = 0; mt < n_mt; ++mt)
        for (int ks = 0; ks < KT / 8; ++ks) {
          const uint32_t ao = (uint32_t)mt * 16u * sbo + (uint32_t)ks * 256u, bo = (uint32_t)ks * 256u;
          const uint64_t a_hi = umma::smem_desc_kmajor(h0 + ao, 128, sbo), a_lo = umma::smem_desc_kmajor(l0 + ao, 128, sbo);
          const uint64_t b_hi = umma::smem_desc_kmajor(h0 + bo, 128, sbo), b_lo = umma::smem_desc_kmajor(l0 + bo, 128, sbo);
          const uint32_t acc = cx.tmem_base + (uint32_t)mt * 256u;
          umma::mma_tf32(acc, a_hi, b_hi, idesc, (c > 0 || ks > 0) ? 1u : 0u);
          umma::mma_tf32(acc, a_hi, b_lo, idesc, 1u);
          umma::mma_tf32(acc, a_lo, b_hi, idesc, 1u);
        }
      umma::mma_commit(&cx.bars[st]);
    }
    if (c + 1 < nchunk) load_chunk(c + 1);
  }
  {
    const int last = nchunk - 1;
    umma::mbar_wait(&cx.bars[last & 1], (uint32_t)((last >> 1) & 1));
    if (nchunk >= 2) umma::mbar_wait(&cx.bars[(last - 1) & 1], (uint32_t)(((last - 1) >> 1) & 1));
  }
  umma::tc_fence_after_sync();
  // accumulator row i -> D row i (thread = row; odd row stride => conflict-free).  Warp w owns TMEM lane quarter
  // w % 4; the (w / 4) index enumerates (M tile, column part) pairs so that 8 or 16 warps all take part.
  {
    const int q = warp & 3, slot = warp >> 2, nslots = nwarps >> 2;
    const int parts = nslots / n_mt > 0 ? nslots / n_mt : 1;            // column parts per M tile
    for (int s = slot; s < n_mt * parts; s += nslots) {
      const int mt = s / parts, part = s % parts;
      const int cbeg = ((Np / 16) * part / parts) * 16, cend = ((Np / 16) * (part + 1) / parts) * 16;
      const int i = mt * 128 + q * 32 + lane;
      const float sqi = i < P ? cx.sq[i] : 0.f;
      for (int c0 = cbeg; c0 < cend; c0 += 16) {
        uint32_t v[16];
        umma::tmem_ld16(umma::tmem_addr(cx.tmem_base, (uint32_t)(q * 32), (uint32_t)(mt * 256 + c0)), v);
        umma::tmem_ld_wait();
        if (i < P) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < P) {
              // diagonal: g_ii == |x_i|^2 exactly in exact arithmetic; use it (ATen's diagonal is clamp-level noise too)
              const float d2 = (c0 + j == i) ? 0.f : (sqi + cx.sq[c0 + j]) - 2.0f * __uint_as_float(v[j]);
              cx.D[i * cx.DS + c0 + j] = sqrtf(fmaxf(d2, 1e-30f)) * post_scale;
            }
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  // G_ij and G_ji add the same products in a different order: mirror the upper triangle so D is bit-symmetric
  for (int i = warp; i < P; i += nwarps)
    for (int j = i + 1 + lane; j < P; j += 32) cx.D[j * cx.DS + i] = cx.D[i * cx.DS + j];
  __syncthreads();
}

// Fills cx.D (shared) with the pairwise distances of the P rows of xb (global, [P][C]) times post_scale.
template <bool TC>
__device__ __forceinline__ void pairdist_to_smem(const float* __restrict__ xb, int P, int C, DistCtx& cx, float post_scale) {
  const int tid = threadIdx.x;
  float* D = cx.D;
  const int DS = cx.DS;
  if (P <= 25) {
    // direct form (ATen's non-matmul cdist path): sqrt(sum (xi - xj)^2).  The <= 25 rows are staged through shared
    // memory in 256-column chunks (the D region is free until the end; pair sums live in registers).
    constexpr int CH = 256;
    float* xs = D + (((P * DS + 3) & ~3));          // [P][CH+1] after the (tiny) D matrix, inside region 0 / tile space
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
  if constexpr (TC) {
    pairdist_tc(xb, P, C, cx, post_scale);      // light kernels are only launched with use_tc = 1 when P > 25
    return;
  } else {
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

// ------------------------------------------------------------------------------------------ plain cdist(x, x)
// Every distance kernel exists in two variants: TC = true (tcgen05 Gram, 512 threads, no FFMA register tile) and
// TC = false (exact fp32 FFMA Gram, 256 threads).
template <bool TC>
__global__ void __launch_bounds__(TC ? kTcThreads : kThreads, 1)
pairwise_dist_kernel(const float* __restrict__ x, int P, int C, float post_scale, float* __restrict__ out, int use_tc) {
  constexpr int NT = TC ? kTcThreads : kThreads;
  extern __shared__ __align__(128) float smem[];
  DistCtx cx = dist_setup(smem, P, 0, use_tc);
  pairdist_to_smem<TC>(x + (long long)blockIdx.x * P * C, P, C, cx, post_scale);
  float* ob = out + (long long)blockIdx.x * P * P;
  for (int i = threadIdx.x >> 5; i < P; i += NT / 32)
    for (int j = threadIdx.x & 31; j < P; j += 32) ob[i * P + j] = cx.D[i * cx.DS + j];
  dist_teardown(cx);
}

// ------------------------------------------------------------------------------------------ DPC-KNN cluster
template <bool TC>
__global__ void __launch_bounds__(TC ? kTcThreads : kThreads, 1)
dpcknn_cluster_kernel(const float* __restrict__ x, long long xbs, const float* __restrict__ noise_u, int P, int C, int K, int knn,
                      float inv_sqrt_c, int64_t* __restrict__ idx_cluster, int64_t* __restrict__ index_down, int use_tc) {
  constexpr int NT = TC ? kTcThreads : kThreads;
  extern __shared__ __align__(128) float smem[];
  DistCtx cx = dist_setup(smem, P, 2 * P + K + kTcThreads / 32, use_tc);
  float* D = cx.D;
  const int DS = cx.DS;
  float* rho = cx.extra;
  float* score = rho + P;
  int* centre = reinterpret_cast<int*>(score + P);   // [K]
  float* red = reinterpret_cast<float*>(centre + K);  // [NT / 32]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  pairdist_to_smem<TC>(x + (long long)b * xbs, P, C, cx, inv_sqrt_c);
  dist_teardown(cx);

  // local density from the knn nearest (self included): exp(-mean(d^2)) + 1e-6 * U
  float lmax = 0.f;
  if (knn <= 8) {
    // Two threads (adjacent lanes) per token, one per half of its column: each keeps the 8 smallest of its rows
    // sorted in registers (strict < keeps the lower row first on ties), then the even lane inserts the odd lane's
    // list -- rows above its own, in order -- so the result equals the single ascending scan.  The scan is serial
    // per column and was 19 % of the kernel's stall samples with 196 of 512 threads active.
    for (int base = 0; base < 2 * P; base += NT) {
      const int it = base + tid;
      const bool act = it < 2 * P;
      const int i = act ? it >> 1 : 0, h = it & 1;
      const int j0 = h ? (P + 1) / 2 : 0, j1 = act ? (h ? P : (P + 1) / 2) : 0;
      float nb[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) nb[t] = CUDART_INF_F;
      for (int j = j0; j < j1; ++j) {
        float cur = D[j * DS + i];
        lmax = fmaxf(lmax, cur);
        if (cur < nb[7]) {
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (cur < nb[t]) { const float tmp = nb[t]; nb[t] = cur; cur = tmp; }
        }
      }
      float pv[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) pv[t] = __shfl_xor_sync(0xffffffffu, nb[t], 1);
      if (act && h == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float cur = pv[q];
          if (cur < nb[7]) {
#pragma unroll
            for (int t = 0; t < 8; ++t)
              if (cur < nb[t]) { const float tmp = nb[t]; nb[t] = cur; cur = tmp; }
          }
        }
        float sumsq = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (t < knn) sumsq += nb[t] * nb[t];
        rho[i] = expf(-(sumsq * (1.0f / (float)knn))) + noise_u[(long long)b * P + i] * 1e-6f;
      }
    }
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
      rho[i] = expf(-(sumsq * (1.0f / (float)knn))) + noise_u[(long long)b * P + i] * 1e-6f;
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float dmax = red[0];
#pragma unroll
  for (int w = 1; w < NT / 32; ++w) dmax = fmaxf(dmax, red[w]);

  // distance to the nearest denser token (or the global max), centre score
  for (int i = tid; i < P; i += NT) {
    const float ri = rho[i];
    float best = dmax;
    for (int j = 0; j < P; ++j) {
      const float v = rho[j] > ri ? D[j * DS + i] : dmax;
      best = fminf(best, v);
    }
    score[i] = best * ri;
  }
  __syncthreads();
  for (int i = tid; i < P; i += NT) {
    const int rk = rank_desc(score, P, i);
    if (rk < K) { centre[rk] = i; index_down[(long long)b * K + rk] = i; }
  }
  __syncthreads();
  // nearest centre (lowest k on ties); centres belong to their own cluster
  for (int i = tid; i < P; i += NT) {
    float best = CUDART_INF_F;
    int bk = 0;
    for (int k = 0; k < K; ++k) {
      const float v = D[centre[k] * DS + i];
      if (v < best) { best = v; bk = k; }
    }
    for (int k = 0; k < K; ++k)
      if (centre[k] == i) bk = k;
    idx_cluster[(long long)b * P + i] = bk;
  }
}

// ------------------------------------------------------------------------------------------ K-Medoids fit
template <bool TC>
__global__ void __launch_bounds__(TC ? kTcThreads : kThreads, 1)
kmedoids_fit_kernel(const float* __restrict__ x, long long xbs, const float* __restrict__ token_weight, int P, int C, int K, int iters,
                    float* __restrict__ centres, int64_t* __restrict__ cluster_idx, int64_t* __restrict__ assignment,
                    int use_tc) {
  constexpr int NT = TC ? kTcThreads : kThreads;
  extern __shared__ __align__(128) float smem[];
  DistCtx cx = dist_setup(smem, P, 3 * P + K, use_tc);
  float* D = cx.D;
  const int DS = cx.DS;
  float* w = cx.extra;
  float* S = w + P;
  int* assign = reinterpret_cast<int*>(S + P);   // [P]
  int* centre = assign + P;                       // [K]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xb = x + (long long)b * xbs;

  for (int i = tid; i < P; i += NT) w[i] = token_weight[(long long)b * P + i];
  pairdist_to_smem<TC>(xb, P, C, cx, 1.0f);
  dist_teardown(cx);

  // S_i = sum_j (D_ij * w_i); initial centres = top-K token weights (descending, lowest index on ties)
  for (int i = tid; i < P; i += NT) {
    const float wi = w[i];
    float s = 0.f;
    for (int j = 0; j < P; ++j) s += __fmul_rn(D[j * DS + i], wi);
    S[i] = s;
    const int rk = rank_desc(w, P, i);
    if (rk < K) centre[rk] = i;
  }
  __syncthreads();
  const float big = 1.0e6f * (float)P;     // P masked columns of 1e6 sum exactly in fp32
  for (int it = 0; it <= iters; ++it) {
    for (int i = tid; i < P; i += NT) {
      float best = CUDART_INF_F;
      int bk = 0;
      for (int k = 0; k < K; ++k) {
        const float v = D[centre[k] * DS + i];
        if (v < best) { best = v; bk = k; }
      }
      assign[i] = bk;
    }
    __syncthreads();
    if (it == iters) break;
    for (int k = tid; k < K; k += NT) {
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
  for (int i = tid; i < P; i += NT) assignment[(long long)b * P + i] = assign[i];
  for (int k = tid; k < K; k += NT) cluster_idx[(long long)b * K + k] = centre[k];
  // medoid rows verbatim
  float* cb = centres + (long long)b * K * C;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(cb) & 15u) == 0);
  for (int k = warp; k < K; k += NT / 32) {
    const float* src = xb + (long long)centre[k] * C;
    if (vec) warp_copy_row16(cb + (long long)k * C, src, C * 4, lane);
    else warp_copy_row_elems(cb + (long long)k * C, src, C, lane);
  }
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

// tensor cores for every P that takes ATen's matmul form (P > 25), unless the caller asks for the exact-fp32 FFMA path
static int pick_tc(int P, int exact_fp32) { return (P > 25 && !exact_fp32) ? 1 : 0; }
// the light 512-thread kernels (no FFMA register tile) also serve the direct-difference path of P <= 25
static int pick_light(int P, int use_tc) { return (use_tc || P <= 25) ? 1 : 0; }

extern "C" int tokred_pairwise_dist(const float* x, int B, int P, int C, float post_scale, int exact_fp32, float* out,
                                    void* stream) {
  const char* what = "tokred_pairwise_dist";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && out, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  if (P > kMaxP) { set_error("%s: P=%d > %d patches is not supported (distance matrix is kept in shared memory)", what, P, kMaxP); return TOKRED_ERR_UNSUPPORTED; }
  if (B == 0) return TOKRED_OK;
  const int use_tc = pick_tc(P, exact_fp32);
  const size_t smem = dist_smem_bytes(P, 0, use_tc);
  if (pick_light(P, use_tc)) {
    if (int e = allow_smem(pairwise_dist_kernel<true>, smem, what)) return e;
    pairwise_dist_kernel<true><<<B, kTcThreads, smem, (cudaStream_t)stream>>>(x, P, C, post_scale, out, use_tc);
  } else {
    if (int e = allow_smem(pairwise_dist_kernel<false>, smem, what)) return e;
    pairwise_dist_kernel<false><<<B, kThreads, smem, (cudaStream_t)stream>>>(x, P, C, post_scale, out, 0);
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
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (int64_t)P * C, "%s: x_batch_stride=%lld < P*C", what, (long long)x_batch_stride);
  const long long xbs = x_batch_stride ? (long long)x_batch_stride : (long long)P * C;
  const int use_tc = pick_tc(P, exact_fp32);
  const size_t smem = dist_smem_bytes(P, 2 * P + K + kTcThreads / 32, use_tc);
  const float inv = 1.0f / (float)sqrt((double)C);     // CUDA tensor / python-scalar = multiply by fp32 reciprocal
  if (pick_light(P, use_tc)) {
    if (int e = allow_smem(dpcknn_cluster_kernel<true>, smem, what)) return e;
    dpcknn_cluster_kernel<true><<<B, kTcThreads, smem, (cudaStream_t)stream>>>(x, xbs, noise_u, P, C, K, knn, inv, idx_cluster,
                                                                              index_down, use_tc);
  } else {
    if (int e = allow_smem(dpcknn_cluster_kernel<false>, smem, what)) return e;
    dpcknn_cluster_kernel<false><<<B, kThreads, smem, (cudaStream_t)stream>>>(x, xbs, noise_u, P, C, K, knn, inv, idx_cluster,
                                                                             index_down, 0);
  }
  return finish_launch(what);
}

extern "C" int tokred_kmedoids_fit(const float* x, int64_t x_batch_stride, const float* token_weight, int B, int P, int C, int K, int iters,
                                   int exact_fp32, float* centres, int64_t* cluster_idx, int64_t* assignment, void* stream) {
  const char* what = "tokred_kmedoids_fit";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && token_weight && centres && cluster_idx && assignment, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 1, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  TOKRED_REQUIRE(K >= 1 && K <= P, "%s: cluster_num=%d outside [1, P=%d]", what, K, P);
  TOKRED_REQUIRE(iters >= 0, "%s: iters=%d < 0", what, iters);
  if (P > kMaxP) { set_error("%s: P=%d > %d patches is not supported (distance matrix is kept in shared memory)", what, P, kMaxP); return TOKRED_ERR_UNSUPPORTED; }
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(x_batch_stride == 0 || x_batch_stride >= (int64_t)P * C, "%s: x_batch_stride=%lld < P*C", what, (long long)x_batch_stride);
  const long long xbs = x_batch_stride ? (long long)x_batch_stride : (long long)P * C;
  const int use_tc = pick_tc(P, exact_fp32);
  const size_t smem = dist_smem_bytes(P, 3 * P + K, use_tc);
  if (pick_light(P, use_tc)) {
    if (int e = allow_smem(kmedoids_fit_kernel<true>, smem, what)) return e;
    kmedoids_fit_kernel<true><<<B, kTcThreads, smem, (cudaStream_t)stream>>>(x, xbs, token_weight, P, C, K, iters, centres,
                                                                            cluster_idx, assignment, use_tc);
  } else {
    if (int e = allow_smem(kmedoids_fit_kernel<false>, smem, what)) return e;
    kmedoids_fit_kernel<false><<<B, kThreads, smem, (cudaStream_t)stream>>>(x, xbs, token_weight, P, C, K, iters, centres,
                                                                           cluster_idx, assignment, 0);
  }
  return finish_launch(what);
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
