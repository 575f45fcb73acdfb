// CTA-wide fp32 "NT" contraction with both operands staged through shared memory and the result handed to an
// epilogue functor:  Z[i][j] = sum_c A[i][c] * B[j][c],  i < M,  j < Nn.
//
// This is the exact-fp32 (FFMA) path used where the reference runs a true fp32 matmul (torch.cdist, fp32
// Sinkhorn / PatchMerger): tensor-core tf32/bf16 would break the 1e-5 parity bar and flip near-tie decisions.
// 256 threads as a 16x16 grid; thread (ty,tx) owns rows ty+16a (a<RA) and cols cbase+tx+16c (c<RC) so the lanes
// of a warp read consecutive tile rows.  Tile rows are XS=36 floats apart (== 4 mod 32) so the LDS.128 operand
// reads of 8 consecutive rows cover all 32 banks.  Per 4 k-steps a thread issues RA+RC LDS.128 for 4*RA*RC FFMA.
#pragma once
#include "common.cuh"

namespace tokred {

constexpr int kGemmThreads = 256;
constexpr int KC = 32;   // k-depth of a staged tile
constexpr int XS = 36;   // tile row stride (floats)

// stage(k0): all threads cooperatively fill the A tile [M][XS] and the B tile [Nn][XS] with columns [k0, k0+KC)
//            (zero beyond the contraction length).  epi(i, j, acc) is called once per valid output.
template <int RA, int RC, typename StageFn, typename EpiFn>
__device__ __forceinline__ void gemm_nt_tiles(int M, int Nn, int Cdim, const float* at, const float* bt, StageFn&& stage,
                                              EpiFn&& epi) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  for (int rbase = 0; rbase < M; rbase += 16 * RA) {
    int arow[RA];
#pragma unroll
    for (int a = 0; a < RA; ++a) arow[a] = min(rbase + ty + 16 * a, M - 1) * XS;
    for (int cbase = 0; cbase < Nn; cbase += 16 * RC) {
      float acc[RA][RC];
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int c = 0; c < RC; ++c) acc[a][c] = 0.f;
      int brow[RC];
#pragma unroll
      for (int c = 0; c < RC; ++c) brow[c] = min(cbase + tx + 16 * c, Nn - 1) * XS;

      for (int k0 = 0; k0 < Cdim; k0 += KC) {
        __syncthreads();   // previous tile fully consumed
        stage(k0);
        __syncthreads();
#pragma unroll 2
        for (int kk = 0; kk < KC; kk += 4) {
          float4 b4[RC];
#pragma unroll
          for (int c = 0; c < RC; ++c) b4[c] = *reinterpret_cast<const float4*>(bt + brow[c] + kk);
#pragma unroll
          for (int a = 0; a < RA; ++a) {
            const float4 a4 = *reinterpret_cast<const float4*>(at + arow[a] + kk);
#pragma unroll
            for (int c = 0; c < RC; ++c) {
              acc[a][c] = fmaf(a4.x, b4[c].x, acc[a][c]);
              acc[a][c] = fmaf(a4.y, b4[c].y, acc[a][c]);
              acc[a][c] = fmaf(a4.z, b4[c].z, acc[a][c]);
              acc[a][c] = fmaf(a4.w, b4[c].w, acc[a][c]);
            }
          }
        }
      }
#pragma unroll
      for (int a = 0; a < RA; ++a) {
        const int i = rbase + ty + 16 * a;
        if (i >= M) continue;
#pragma unroll
        for (int c = 0; c < RC; ++c) {
          const int j = cbase + tx + 16 * c;
          if (j < Nn) epi(i, j, acc[a][c]);
        }
      }
    }
  }
  __syncthreads();
}

// Picks a register tile that does not waste most of its work on clamped duplicates for small problems.
template <typename StageFn, typename EpiFn>
__device__ __forceinline__ void gemm_nt(int M, int Nn, int Cdim, const float* at, const float* bt, StageFn&& stage,
                                        EpiFn&& epi) {
  if (M <= 64 && Nn <= 64) gemm_nt_tiles<4, 4>(M, Nn, Cdim, at, bt, stage, epi);
  else gemm_nt_tiles<13, 7>(M, Nn, Cdim, at, bt, stage, epi);
}

// Stage rows [0, rows) x columns [k0, k0+KC) of a row-major matrix of T (leading dimension ld, `cols` valid
// columns) into tile[rows][XS] as fp32, applying f(row, col, value) on the way.  Vector loads when aligned.
template <typename T, typename F>
__device__ __forceinline__ void stage_rows(const T* __restrict__ src, int rows, int cols, int ld, int k0, float* tile,
                                           bool vec_ok, F&& f) {
  for (int e = threadIdx.x; e < rows * (KC / 4); e += kGemmThreads) {
    const int row = e / (KC / 4), ch = e % (KC / 4);
    const int k = k0 + ch * 4;
    const T* g = src + (long long)row * ld + k;
    float v[4];
    if (vec_ok && k + 3 < cols) {
      if (sizeof(T) == 4) {
        const float4 q = *reinterpret_cast<const float4*>(g);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
        const uint2 q = *reinterpret_cast<const uint2*>(g);
        const T* h = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = to_f32(h[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (k + i < cols) ? to_f32(g[i]) : 0.f;
    }
    float4 o;
    o.x = k < cols ? f(row, k, v[0]) : 0.f;
    o.y = k + 1 < cols ? f(row, k + 1, v[1]) : 0.f;
    o.z = k + 2 < cols ? f(row, k + 2, v[2]) : 0.f;
    o.w = k + 3 < cols ? f(row, k + 3, v[3]) : 0.f;
    *reinterpret_cast<float4*>(tile + row * XS + ch * 4) = o;
  }
}

// vec_ok for stage_rows: 4 consecutive elements loadable as one aligned vector from every row start.
template <typename T>
__device__ __forceinline__ bool stage_vec_ok(const T* p, int ld) {
  return (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) % (4 * sizeof(T))) == 0);
}

}  // namespace tokred
