// ToMe: bipartite soft matching (tome_match) and size-weighted merge with source map (tome_merge).
// Reference: models/tome.py:230-337 and Block_ToMe.forward :78-104.
//
// tome_match   one CTA per image.  Rows are L2-normalised into shared memory (fp32, optionally rounded to bf16
//              the way the autocast matmul rounds its operands), the even x odd similarity tile is computed
//              with 4x4 register tiles, per-row (max, argmax) by one warp per row, and the edge ordering by
//              rank-counting.  Emits the three int64 index lists the reference closure captures.
// tome_merge   grid (splits, B), one warp per OUTPUT row, 16-byte streaming loads.  Destination rows add
//              their sources in src-list order (the order CPU scatter_add applies them) with unfused
//              multiply/add so that fp32 results are bit-identical to the reference on CPU; no atomics.
//              Also writes size_out and the float "reduced_cluster_idx" map without the [B,N,N] identity the
//              reference pushes through the merge (models/tome.py:91-99, ~35 % of its op time).
#include <math_constants.h>

#include "common.cuh"
#include "umma.cuh"

namespace tokred {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ------------------------------------------------------------------------------------------ match
template <typename TM>
__global__ void __launch_bounds__(kThreads)
tome_match_kernel(const TM* __restrict__ metric, int N, int D, int r, int class_token, int lowp,
                  int64_t* __restrict__ unm_idx, int64_t* __restrict__ src_idx, int64_t* __restrict__ dst_idx) {
  extern __shared__ float smem[];
  const int na = (N + 1) / 2, nb = N / 2, DP = D + 1, SP = nb + 1;
  float* A = smem;                        // [na][DP]  normalised even tokens
  float* Bm = A + na * DP;                // [nb][DP]  normalised odd tokens
  float* S = Bm + nb * DP;                // [na][SP]  similarity
  float* node_max = S + na * SP;          // [na]
  int* node_idx = reinterpret_cast<int*>(node_max + na);   // [na]
  unsigned char* merged = reinterpret_cast<unsigned char*>(node_idx + na);   // [na]

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TM* mb = metric + (long long)b * N * D;
  // class_token carries the protection flags: bit 0 = class token (row 0 never merges away, models/tome.py:263-264),
  // bit 1 = distillation token (odd token 0 never receives a merge, :265-266)
  const bool dist_token = (class_token & 2) != 0;
  class_token &= 1;

  // 1a. stage the raw metric tile into shared memory with every load of the CTA in flight at once (the first
  //     version walked the rows one warp at a time from global memory and was latency-bound: ncu r01, 55 us)
  {
    constexpr int VE = 16 / sizeof(TM);
    const bool vec = (D % VE == 0) && ((reinterpret_cast<uintptr_t>(mb) & 15u) == 0);
    if (vec) {
      const int per_row = D / VE;
      for (int e = tid; e < N * per_row; e += kThreads) {
        const int t = e / per_row, d0 = (e % per_row) * VE;
        const int4 raw = *reinterpret_cast<const int4*>(mb + (long long)t * D + d0);
        const TM* v = reinterpret_cast<const TM*>(&raw);
        float* dst = ((t & 1) ? (Bm + (t >> 1) * DP) : (A + (t >> 1) * DP)) + d0;
#pragma unroll
        for (int i = 0; i < VE; ++i) dst[i] = to_f32(v[i]);
      }
    } else {
      for (int e = tid; e < N * D; e += kThreads) {
        const int t = e / D, d = e % D;
        ((t & 1) ? (Bm + (t >> 1) * DP) : (A + (t >> 1) * DP))[d] = to_f32(mb[e]);
      }
    }
  }
  __syncthreads();
  // 1b. m / ||m||  (norm accumulated in fp32; true division like the reference), in place.  8 lanes per row, so a
  //     warp keeps 4 rows in flight and the shuffle chain is 3 deep (one row per warp was latency-bound).
  for (int base = warp * 4; base < N; base += kWarps * 4) {      // warp-uniform trip count (full-mask shuffles inside)
    const int t = base + (lane >> 3), sub = lane & 7;
    const bool live = t < N;
    float* row = ((t & 1) ? (Bm + (t >> 1) * DP) : (A + (t >> 1) * DP));
    float ss = 0.f;
    if (live)
      for (int d = sub; d < D; d += 8) { float v = row[d]; ss = fmaf(v, v, ss); }
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    const float nrm = sqrtf(ss);
    if (live)
      for (int d = sub; d < D; d += 8) {
        float v = row[d] / nrm;
        row[d] = lowp ? bf16_round(v) : v;
      }
  }
  __syncthreads();

  // 2. S = A Bm^T with 4x4 register tiles; rows/cols of a tile are strided (ti + TA*a, tj + TB*b) so that the
  //    lanes of a warp read consecutive Bm rows (stride DP = D+1 floats -> conflict-free).
  const int TA = (na + 3) / 4, TB = (nb + 3) / 4;
  for (int t = tid; t < TA * TB; t += kThreads) {
    const int ti = t / TB, tj = t % TB;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    const float* ap[4];
    const float* bp[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ap[a] = A + min(ti + TA * a, na - 1) * DP;
#pragma unroll
    for (int c = 0; c < 4; ++c) bp[c] = Bm + min(tj + TB * c, nb - 1) * DP;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = ap[a][d];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = bp[c][d];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = ti + TA * a;
      if (i >= na) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = tj + TB * c;
        if (j < nb) S[i * SP + j] = lowp ? bf16_round(acc[a][c]) : acc[a][c];
      }
    }
  }
  __syncthreads();

  // 3. per-row (max, argmax), lowest column on ties; CLS row (0) is protected with -inf.
  for (int i = warp; i < na; i += kWarps) {
    float best = -CUDART_INF_F;
    int bj = 0x7fffffff;
    if (!(class_token && i == 0)) {
      for (int j = lane; j < nb; j += 32) {
        float v = (dist_token && j == 0) ? -CUDART_INF_F : S[i * SP + j];
        if (nan_gt(v, best) || bj == 0x7fffffff) { best = v; bj = j; }
      }
    }
    warp_argmax(best, bj);
    if (lane == 0) { node_max[i] = best; node_idx[i] = (bj == 0x7fffffff) ? 0 : bj; }
  }
  __syncthreads();

  // 4. edge order: descending node_max (ties -> lower row); first r rows are merged away.
  const int n_unm = na - r;
  for (int i = tid; i < na; i += kThreads) {
    const int rk = rank_desc(node_max, na, i);
    merged[i] = rk < r;
    if (rk < r) {
      src_idx[(long long)b * r + rk] = i;
      dst_idx[(long long)b * r + rk] = node_idx[i];
    } else if (!class_token) {
      unm_idx[(long long)b * n_unm + (rk - r)] = i;
    }
  }
  if (class_token) {
    __syncthreads();
    for (int i = tid; i < na; i += kThreads) {
      if (!merged[i]) {
        int pos = 0;
        for (int q = 0; q < i; ++q) pos += !merged[q];
        unm_idx[(long long)b * n_unm + pos] = i;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ match, tensor cores
// bf16-autocast path (score_lowp = 1): the reference's `a @ b^T` IS a bf16 tensor-core matmul with fp32 accumulate
// and a bf16-rounded result, so the similarity tile goes to tcgen05: normalised rows are written as bf16 straight
// into the canonical K-major UMMA layout, ONE thread issues D/16 tcgen05.mma (M=128 x N<=256 x K=16) into a TMEM
// accumulator, and the epilogue reads it with tcgen05.ld 32x32b — thread i owns accumulator row i, which is exactly
// what the per-row (max, argmax) needs: no similarity matrix in shared memory, no cross-thread reduction.
constexpr int kTcThreads = 256;
constexpr int kTc2Threads = 512;

template <typename TM>
__global__ void __launch_bounds__(kTcThreads)
tome_match_tc_kernel(const TM* __restrict__ metric, int N, int D, int r, int class_token,
                     int64_t* __restrict__ unm_idx, int64_t* __restrict__ src_idx, int64_t* __restrict__ dst_idx) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int na = (N + 1) / 2, nb = N / 2;
  const int Dp = (D + 15) & ~15, Np = (nb + 15) & ~15;
  const uint32_t sbo = (uint32_t)(Dp / 8) * 128u;            // bytes between 8-row groups
  unsigned char* opA = smem_raw;                               // 128 rows (even tokens), canonical K-major bf16
  unsigned char* opB = opA + 16 * sbo;                         // Np rows (odd tokens)
  TM* raw = reinterpret_cast<TM*>(opB + (Np / 8) * sbo);       // [N][D] staged metric
  float* node_max = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(raw) + (((size_t)N * D * sizeof(TM) + 15) & ~(size_t)15));
  int* node_idx = reinterpret_cast<int*>(node_max + 128);
  unsigned char* merged = reinterpret_cast<unsigned char*>(node_idx + 128);   // [128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(merged + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TM* mb = metric + (long long)b * N * D;
  const uint32_t ncols = umma::tmem_cols_pow2((uint32_t)Np);
  const bool dist_token = (class_token & 2) != 0;      // protection flags, see tome_match_kernel
  class_token &= 1;

  if (warp == 0) umma::tmem_alloc(tmem_slot, ncols);
  if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_mbar_init(); }

  // stage the raw metric (every load of the CTA in flight at once) and clear the operand tiles (padding = 0)
  {
    constexpr int VE = 16 / sizeof(TM);
    const int total = N * D;
    if ((total % VE == 0) && ((reinterpret_cast<uintptr_t>(mb) & 15u) == 0)) {
      for (int e = tid; e < total / VE; e += kTcThreads)
        reinterpret_cast<int4*>(raw)[e] = *reinterpret_cast<const int4*>(mb + (long long)e * VE);
    } else {
      for (int e = tid; e < total; e += kTcThreads) raw[e] = mb[e];
    }
    const int op_int4 = (int)((16 + Np / 8) * sbo / 16);
    for (int e = tid; e < op_int4; e += kTcThreads) reinterpret_cast<int4*>(opA)[e] = make_int4(0, 0, 0, 0);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // m / ||m|| in fp32, rounded to bf16 (what the autocast matmul does to its operands), into the UMMA layout.
  // 8 lanes per row: lane `sub` owns elements sub, sub+8, ... so the 8 lanes of a row write one 16-byte K-chunk.
  for (int base = warp * 4; base < N; base += (kTcThreads / 32) * 4) {   // warp-uniform trip count (full-mask shuffles)
    const int t = base + (lane >> 3), sub = lane & 7;
    const bool live = t < N;
    const TM* row = raw + (size_t)t * D;
    float ss = 0.f;
    if (live)
      for (int d = sub; d < D; d += 8) { float v = to_f32(row[d]); ss = fmaf(v, v, ss); }
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    const float nrm = sqrtf(ss);
    if (live) {
      unsigned char* op = ((t & 1) ? opB : opA) + (uint32_t)(t >> 4) * sbo + (uint32_t)((t >> 1) & 7) * 16u + (uint32_t)sub * 2u;
      for (int d = sub; d < D; d += 8)
        *reinterpret_cast<__nv_bfloat16*>(op + (uint32_t)(d >> 3) * 128u) = __float2bfloat16_rn(to_f32(row[d]) / nrm);
    }
  }
  umma::fence_proxy_async_smem();
  __syncthreads();

  if (tid == 0) {
    umma::tc_fence_after_sync();
    const uint32_t idesc = umma::instr_desc(umma::FMT_BF16, 128, (uint32_t)Np);
    const uint32_t a0 = umma::smem_u32(opA), b0 = umma::smem_u32(opB);
    for (int ks = 0; ks < Dp / 16; ++ks) {
      const uint64_t da = umma::smem_desc_kmajor(a0 + ks * 256, 128, sbo);
      const uint64_t db = umma::smem_desc_kmajor(b0 + ks * 256, 128, sbo);
      umma::mma_bf16(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    umma::mma_commit(bar);
  }
  umma::mbar_wait(bar, 0);
  umma::tc_fence_after_sync();

  // per-row (max, argmax) straight from TMEM: lane quarter q = warp % 4 holds rows 32q..32q+31; warps q and q+4
  // split the columns (a warp may only touch its own lane quarter).  Lowest column wins ties; CLS row protected.
  {
    const int i = (warp & 3) * 32 + lane;
    const int half = ((Np / 16 + 1) / 2) * 16;
    const int cbeg = warp < 4 ? 0 : half, cend = warp < 4 ? half : Np;
    float best = -CUDART_INF_F;
    int bj = 0x7fffffff;
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      uint32_t v[16];
      umma::tmem_ld16(umma::tmem_addr(tmem_base, (uint32_t)((warp & 3) * 32), (uint32_t)c0), v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float sc = (dist_token && c0 + j == 0) ? -CUDART_INF_F : bf16_round(__uint_as_float(v[j]));
        if (c0 + j < nb && (nan_gt(sc, best) || bj == 0x7fffffff)) { best = sc; bj = c0 + j; }
      }
    }
    if (warp >= 4) { node_max[i] = best; node_idx[i] = bj; }
    __syncthreads();
    if (warp < 4) {
      const float ob = node_max[i];
      const int oj = node_idx[i];
      if (oj != 0x7fffffff && (nan_gt(ob, best) || bj == 0x7fffffff)) { best = ob; bj = oj; }   // ties keep the lower column
      if (bj == 0x7fffffff) bj = 0;
      if (class_token && i == 0) { best = -CUDART_INF_F; bj = 0; }
    }
    __syncthreads();
    if (warp < 4) { node_max[i] = best; node_idx[i] = bj; }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, ncols);

  const int n_unm = na - r;
  for (int i = tid; i < na; i += kTcThreads) {
    const int rk = rank_desc(node_max, na, i);
    merged[i] = rk < r;
    if (rk < r) {
      src_idx[(long long)b * r + rk] = i;
      dst_idx[(long long)b * r + rk] = node_idx[i];
    } else if (!class_token) {
      unm_idx[(long long)b * n_unm + (rk - r)] = i;
    }
  }
  if (class_token) {
    __syncthreads();
    for (int i = tid; i < na; i += kTcThreads) {
      if (!merged[i]) {
        int pos = 0;
        for (int q = 0; q < i; ++q) pos += !merged[q];
        unm_idx[(long long)b * n_unm + pos] = i;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ match, tensor cores, v2
// bf16 metric with D = 64 (every DeiT), optionally given as the per-head keys themselves: metric[b,n,h,:] at
// ((b*N + n) * token_stride + h*64) -- the k slice of the qkv Linear's output, so that `metric = k.mean(1)`
// (models/tome.py:58) costs neither a launch nor an HBM round trip: the head mean is taken here in fp32 and rounded to
// bf16 exactly where ATen's mean rounds.  Against the first tensor-core kernel: no raw staging copy, 32 contiguous
// bytes per lane straight from global memory, one conflict-free 16-byte store per K-chunk into the UMMA layout instead
// of sixty-four 2-byte stores per row (721k shared-memory bank conflicts in the round-1 profile), ballot-based
// compaction of the unmerged list.
template <bool HEADS>
__global__ void __launch_bounds__(kTc2Threads, 2)
tome_match_tc2_kernel(const __nv_bfloat16* __restrict__ metric, long long token_stride, int heads, int N, int r, int class_token,
                      int64_t* __restrict__ unm_idx, int64_t* __restrict__ src_idx, int64_t* __restrict__ dst_idx) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int na = (N + 1) / 2, nb = N / 2;
  const int Np = (nb + 15) & ~15;
  unsigned char* opA = smem_raw;                               // 128 rows (even tokens), canonical K-major bf16, 128 B per row
  unsigned char* opB = opA + 16 * 1024;                        // Np rows (odd tokens)
  float* node_max = reinterpret_cast<float*>(opB + (Np / 8) * 1024);
  int* node_idx = reinterpret_cast<int*>(node_max + 128);
  int* wcount = node_idx + 128;                                // [4]
  unsigned char* merged = reinterpret_cast<unsigned char*>(wcount + 4);   // [128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(merged + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* mb = metric + (long long)b * N * token_stride;
  const uint32_t ncols = umma::tmem_cols_pow2((uint32_t)Np);
  const bool dist_token = (class_token & 2) != 0;      // protection flags, see tome_match_kernel
  class_token &= 1;

  if (warp == 0) umma::tmem_alloc(tmem_slot, ncols);
  if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_mbar_init(); }

  // m / ||m|| in fp32, rounded to bf16 (what the autocast matmul does to its operands), into the UMMA layout.
  // job = 8 rows of one operand; lane -> row = lane % 8, K-chunks 2*(lane/8), 2*(lane/8)+1 (16 elements).
  {
    const int r8 = lane & 7, cp = lane >> 3;
    const float inv_h = 1.0f / (float)heads;
    const int njobs = 16 + Np / 8;
    for (int j = warp; j < njobs; j += kTc2Threads / 32) {
      const bool isA = j < 16;
      const int g = isA ? j : j - 16;
      const int i = g * 8 + r8;
      const bool live = i < (isA ? na : nb);
      float v[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = 0.f;
      if (live) {
        const __nv_bfloat16* row = mb + (long long)(2 * i + (isA ? 0 : 1)) * token_stride + cp * 16;
        if (HEADS) {
          // three heads = six independent 16-byte loads in flight per lane.  (Six heads at once cost 111 registers:
          // ONE 512-thread CTA per SM, so B=256 ran as 1.73 waves; at <= 64 registers two CTAs are resident, the batch
          // is a single wave and the SM has the same twelve loads per lane pair in flight.)  Heads are still summed in
          // ascending order.
          for (int h0 = 0; h0 < heads; h0 += 3) {
            int4 a[3], c[3];
#pragma unroll
            for (int u = 0; u < 3; ++u)
              if (h0 + u < heads) { a[u] = ld_stream16(row + (h0 + u) * 64); c[u] = ld_stream16(row + (h0 + u) * 64 + 8); }
#pragma unroll
            for (int u = 0; u < 3; ++u)
              if (h0 + u < heads) {
                const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a[u]);
                const __nv_bfloat16* pc = reinterpret_cast<const __nv_bfloat16*>(&c[u]);
#pragma unroll
                for (int e = 0; e < 8; ++e) { v[e] += __bfloat162float(pa[e]); v[8 + e] += __bfloat162float(pc[e]); }
              }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = bf16_round(v[e] * inv_h);        // ATen mean: fp32 sum * (1/H), one rounding
        } else {
          const int4 a = ld_stream16(row), c = ld_stream16(row + 8);
          const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a);
          const __nv_bfloat16* pc = reinterpret_cast<const __nv_bfloat16*>(&c);
#pragma unroll
          for (int e = 0; e < 8; ++e) { v[e] = __bfloat162float(pa[e]); v[8 + e] = __bfloat162float(pc[e]); }
        }
      }
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) ss = fmaf(v[e], v[e], ss);
      ss += __shfl_xor_sync(0xffffffffu, ss, 8);
      ss += __shfl_xor_sync(0xffffffffu, ss, 16);
      const float nrm = sqrtf(ss);
      __nv_bfloat162 o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        o[e] = live ? __floats2bfloat162_rn(v[2 * e] / nrm, v[2 * e + 1] / nrm) : __floats2bfloat162_rn(0.f, 0.f);
      unsigned char* dst = (isA ? opA : opB) + (size_t)g * 1024 + r8 * 16 + cp * 256;
      *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(&o[0]);
      *reinterpret_cast<int4*>(dst + 128) = *reinterpret_cast<const int4*>(&o[4]);
    }
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = umma::instr_desc(umma::FMT_BF16, 128, (uint32_t)Np);
    const uint32_t a0 = umma::smem_u32(opA), b0 = umma::smem_u32(opB);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma::mma_bf16(tmem_base, umma::smem_desc_kmajor(a0 + ks * 256, 128, 1024), umma::smem_desc_kmajor(b0 + ks * 256, 128, 1024),
                     idesc, ks > 0 ? 1u : 0u);
    umma::mma_commit(bar);
    umma::mbar_wait(bar, 0);
  }
  __syncthreads();
  umma::tc_fence_after_sync();

  // per-row (max, argmax) straight from TMEM: lane quarter q = warp % 4 holds rows 32q..32q+31; warps q and q+4
  // split the columns (a warp may only touch its own lane quarter).  Lowest column wins ties; CLS row protected.
  {
    const int i = (warp & 3) * 32 + lane;
    const int half = ((Np / 16 + 1) / 2) * 16;
    const int cbeg = warp < 4 ? 0 : half, cend = warp < 4 ? half : (warp < 8 ? Np : 0);     // warps 8.. idle here
    float best = -CUDART_INF_F;
    int bj = 0x7fffffff;
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      uint32_t v[16];
      umma::tmem_ld16(umma::tmem_addr(tmem_base, (uint32_t)((warp & 3) * 32), (uint32_t)c0), v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float sc = (dist_token && c0 + j == 0) ? -CUDART_INF_F : bf16_round(__uint_as_float(v[j]));
        if (c0 + j < nb && (nan_gt(sc, best) || bj == 0x7fffffff)) { best = sc; bj = c0 + j; }
      }
    }
    if (warp >= 4 && warp < 8) { node_max[i] = best; node_idx[i] = bj; }
    __syncthreads();
    if (warp < 4) {
      const float ob = node_max[i];
      const int oj = node_idx[i];
      if (oj != 0x7fffffff && (nan_gt(ob, best) || bj == 0x7fffffff)) { best = ob; bj = oj; }   // ties keep the lower column
      if (bj == 0x7fffffff) bj = 0;
      if (class_token && i == 0) { best = -CUDART_INF_F; bj = 0; }
    }
    __syncthreads();
    if (warp < 4) { node_max[i] = best; node_idx[i] = bj; }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, ncols);

  // edge order: four lanes per even token split the rank count (all 512 threads busy)
  const int n_unm = na - r;
  {
    const int i = tid >> 2, part = tid & 3;
    int rk = 0;
    if (i < na) rk = rank_desc_range(node_max, na * part / 4, na * (part + 1) / 4, i);
    rk += __shfl_xor_sync(0xffffffffu, rk, 1);
    rk += __shfl_xor_sync(0xffffffffu, rk, 2);
    if (i < na && part == 0) {
      merged[i] = rk < r;
      if (rk < r) {
        src_idx[(long long)b * r + rk] = i;
        dst_idx[(long long)b * r + rk] = node_idx[i];
      } else if (!class_token) {
        unm_idx[(long long)b * n_unm + (rk - r)] = i;
      }
    }
  }
  if (class_token) {          // unmerged even tokens in ascending order (models/tome.py:272): ballot compaction
    __syncthreads();
    bool keep = false;
    unsigned bal = 0;
    if (warp < 4) {
      keep = tid < na && !merged[tid];
      bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) wcount[warp] = __popc(bal);
    }
    __syncthreads();
    if (warp < 4 && keep) {
      int pos = __popc(bal & ((1u << lane) - 1u));
      for (int w = 0; w < warp; ++w) pos += wcount[w];
      unm_idx[(long long)b * n_unm + pos] = tid;
    }
  }
}

// ------------------------------------------------------------------------------------------ merge
template <typename T> struct Chunk;            // one 16-byte lane chunk
template <> struct Chunk<float> { static constexpr int VE = 4; };
template <> struct Chunk<__nv_bfloat16> { static constexpr int VE = 8; };

// acc = x*z (rounded to T: the reference materialises x*size), then fp32 accumulation of the sources with ONE
// rounding to T at the end — what CPU scatter_add does for bf16 (opmath accumulation); identity for fp32.
// Multiply and add stay UNFUSED so that fp32 results are bit-identical to the reference.
template <typename T>
__device__ __forceinline__ float mul_as(float x, float z) { return round_as<T>(__fmul_rn(x, z)); }
template <typename T>
__device__ __forceinline__ float add_as(float a, float b) { return __fadd_rn(a, b); }

// One warp produces one output row.  CPL = 16-byte chunks per lane: ALL chunks of a source row are requested
// before any is consumed (CPL x 512 B in flight per warp) — the kernel is latency-bound otherwise (ncu r01:
// long-scoreboard stalls, 20 % DRAM utilisation).  CPL == 0 selects the generic element-wise path.
template <typename T, int CPL>
struct RowAcc {
  static constexpr int VE = Chunk<T>::VE;
  float v[CPL > 0 ? CPL : 1][VE];
};

template <typename T, int CPL>
__device__ __forceinline__ void load_row_chunks(const T* __restrict__ row, int nchunks, int lane, int4 (&raw)[CPL]) {
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) raw[i] = ld_stream16(row + c * Chunk<T>::VE);
  }
}

// Fused caller-side work of the bf16-autocast block (models/tome.py:88-104): `x + attn branch` is formed on every row
// read (the fp32 add the reference runs as its own kernel) and the finished row is LayerNorm-ed and rounded to bf16 for
// the MLP's first Linear -- add_layernorm's arithmetic, chunk for chunk, so the fused launch is bit-identical to
// add -> tome_merge -> add_layernorm.
struct MergeLn {
  const __nv_bfloat16* branch;   // [N,C] of this image, or nullptr
  const float* gamma;
  const float* beta;
  float eps;
  __nv_bfloat16* yrow;           // output row of y
};

template <int CPL>
__device__ __forceinline__ void load_branch_chunks(const __nv_bfloat16* __restrict__ row, int nchunks, int lane, uint2 (&rb)[CPL]) {
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) rb[i] = *reinterpret_cast<const uint2*>(row + c * 4);
  }
}
__device__ __forceinline__ float bf16_of(const uint2& u, int e) {
  const uint32_t w = e < 2 ? u.x : u.y;
  return __uint_as_float((e & 1) ? (w & 0xffff0000u) : (w << 16));
}

template <typename T, int CPL, bool FUSED = false>
__device__ __forceinline__ void merge_row_vec(const T* __restrict__ xb, T* __restrict__ orow, int C, int lane, int t0,
                                              int j, const int* __restrict__ src, const int* __restrict__ dst, int r,
                                              const float* __restrict__ zs, bool has_size, bool divide, float& zsum_out,
                                              const MergeLn* ln = nullptr) {
  constexpr int VE = Chunk<T>::VE;
  const int nchunks = C / VE;
  RowAcc<T, CPL> acc;
  int4 raw[CPL];
  uint2 rb[FUSED ? CPL : 1];
  const bool has_br = FUSED && ln->branch != nullptr;
  load_row_chunks<T, CPL>(xb + (long long)t0 * C, nchunks, lane, raw);
  if constexpr (FUSED) { if (has_br) load_branch_chunks<CPL>(ln->branch + (long long)t0 * C, nchunks, lane, rb); }
  const float z0 = zs[t0];
  float zsum = z0;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const T* v = reinterpret_cast<const T*>(&raw[i]);
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      float xv = to_f32(v[e]);
      if constexpr (FUSED) { if (has_br) xv = __fadd_rn(xv, bf16_of(rb[i], e)); }
      acc.v[i][e] = has_size ? mul_as<T>(xv, z0) : xv;
    }
  }
  if (j >= 0) {
    // sources of odd token j, in src-list order: ballot over the list, then walk the set bits (warp-uniform)
    for (int base = 0; base < r; base += 32) {
      const int sidx = base + lane;
      unsigned m = __ballot_sync(0xffffffffu, sidx < r && dst[sidx] == j);
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        const int ts = 2 * src[base + bit];
        load_row_chunks<T, CPL>(xb + (long long)ts * C, nchunks, lane, raw);
        if constexpr (FUSED) { if (has_br) load_branch_chunks<CPL>(ln->branch + (long long)ts * C, nchunks, lane, rb); }
        const float z = zs[ts];
        zsum = add_as<T>(zsum, z);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const T* v = reinterpret_cast<const T*>(&raw[i]);
#pragma unroll
          for (int e = 0; e < VE; ++e) {
            float xv = to_f32(v[e]);
            if constexpr (FUSED) { if (has_br) xv = __fadd_rn(xv, bf16_of(rb[i], e)); }
            acc.v[i][e] = add_as<T>(acc.v[i][e], has_size ? mul_as<T>(xv, z) : xv);
          }
        }
      }
    }
  }
  zsum = round_as<T>(zsum);
  // x / 2^k == x * 2^-k exactly, so power-of-two sizes (all of stage 1: sizes 1 and 2) skip the IEEE division, which
  // was 17 % of the kernel's issue slots in the r01 profile; other sizes keep the true division (bit-exactness)
  const bool pow2 = (__float_as_uint(zsum) & 0x007fffffu) == 0u;
  const float inv = 1.0f / zsum;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      T outv[VE];
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        const float a = round_as<T>(acc.v[i][e]);
        outv[e] = !divide ? from_f32<T>(acc.v[i][e]) : from_f32<T>(pow2 ? __fmul_rn(a, inv) : __fdiv_rn(a, zsum));
        if constexpr (FUSED) acc.v[i][e] = to_f32(outv[e]);       // the finished row stays in registers for the LayerNorm
      }
      st_stream16(orow + c * VE, *reinterpret_cast<const int4*>(outv));
    }
  }
  zsum_out = zsum;
  if constexpr (FUSED) {
    // norm.cu's add_layernorm, operation for operation (C = 128 * CPL: every lane chunk is live)
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) sum += (acc.v[i][0] + acc.v[i][1]) + (acc.v[i][2] + acc.v[i][3]);
    const float mean = warp_sum(sum) * (1.0f / (float)(128 * CPL));
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const float a = acc.v[i][0] - mean, b = acc.v[i][1] - mean, c = acc.v[i][2] - mean, d = acc.v[i][3] - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / (float)(128 * CPL)) + ln->eps);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c4 = (i * 32 + lane) * 4;
      const float4 g = *reinterpret_cast<const float4*>(ln->gamma + c4), bt = *reinterpret_cast<const float4*>(ln->beta + c4);
      const float a = (acc.v[i][0] - mean) * rstd * g.x + bt.x, b = (acc.v[i][1] - mean) * rstd * g.y + bt.y;
      const float c = (acc.v[i][2] - mean) * rstd * g.z + bt.z, d = (acc.v[i][3] - mean) * rstd * g.w + bt.w;
      const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(ln->yrow + c4) = o;
    }
  }
}

// generic path: any C / alignment, one element per lane-step
template <typename T>
__device__ __forceinline__ void merge_row_any(const T* __restrict__ xb, T* __restrict__ orow, int C, int lane, int t0, int j,
                                              const int* __restrict__ src, const int* __restrict__ dst, int r,
                                              const float* __restrict__ zs, bool has_size, bool divide, float& zsum_out) {
  float zsum = zs[t0];
  if (j >= 0)
    for (int s = 0; s < r; ++s)
      if (dst[s] == j) zsum = add_as<T>(zsum, zs[2 * src[s]]);
  zsum = round_as<T>(zsum);
  for (int c = lane; c < C; c += 32) {
    const float x0 = to_f32(xb[(long long)t0 * C + c]);
    float acc = has_size ? mul_as<T>(x0, zs[t0]) : x0;
    if (j >= 0)
      for (int s = 0; s < r; ++s)
        if (dst[s] == j) {
          const int ts = 2 * src[s];
          const float xv = to_f32(xb[(long long)ts * C + c]);
          acc = add_as<T>(acc, has_size ? mul_as<T>(xv, zs[ts]) : xv);
        }
    orow[c] = divide ? from_f32<T>(__fdiv_rn(round_as<T>(acc), zsum)) : from_f32<T>(acc);
  }
  zsum_out = zsum;
}

template <typename T, int CPL>
__global__ void __launch_bounds__(kThreads, CPL <= 3 ? 4 : 2)
tome_merge_kernel(const T* __restrict__ x, const T* __restrict__ size, const int64_t* __restrict__ unm_idx,
                  const int64_t* __restrict__ src_idx, const int64_t* __restrict__ dst_idx, int N, int C, int r,
                  T* __restrict__ x_out, T* __restrict__ size_out, float* __restrict__ rci, int divide) {
  extern __shared__ float smem[];
  const int na = (N + 1) / 2, n_unm = na - r, n_out = N - r;
  float* zs = smem;                                   // [N]   token sizes
  int* unm = reinterpret_cast<int*>(zs + N);          // [n_unm]
  int* src = unm + n_unm;                             // [r]
  int* dst = src + r;                                 // [r]
  int* rowmap = dst + r;                              // [na]  output row of every even token

  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_size = size != nullptr;
  for (int t = tid; t < N; t += kThreads) zs[t] = has_size ? to_f32(size[(long long)b * N + t]) : 1.f;
  // caller-supplied index lists are clamped into range (as dpcknn_merge / gather_rows do): no out-of-bounds row
  for (int q = tid; q < n_unm; q += kThreads) unm[q] = clamp_idx(unm_idx[(long long)b * n_unm + q], na);
  for (int s = tid; s < r; s += kThreads) {
    src[s] = clamp_idx(src_idx[(long long)b * r + s], na);
    dst[s] = clamp_idx(dst_idx[(long long)b * r + s], N / 2);
  }
  __syncthreads();

  const T* xb = x + (long long)b * N * C;
  T* ob = x_out + (long long)b * n_out * C;
  for (int q = blockIdx.x * kWarps + warp; q < n_out; q += gridDim.x * kWarps) {
    const int j = q < n_unm ? -1 : q - n_unm;
    const int t0 = q < n_unm ? 2 * unm[q] : 2 * j + 1;
    float zsum;
    if constexpr (CPL > 0) merge_row_vec<T, CPL>(xb, ob + (long long)q * C, C, lane, t0, j, src, dst, r, zs, has_size, divide != 0, zsum);
    else merge_row_any<T>(xb, ob + (long long)q * C, C, lane, t0, j, src, dst, r, zs, has_size, divide != 0, zsum);
    if (lane == 0) size_out[(long long)b * n_out + q] = from_f32<T>(zsum);
  }

  if (rci != nullptr && blockIdx.x == 0) {
    for (int q = tid; q < n_unm; q += kThreads) rowmap[unm[q]] = q;
    for (int s = tid; s < r; s += kThreads) rowmap[src[s]] = n_unm + dst[s];
    __syncthreads();
    for (int t = 1 + tid; t < N; t += kThreads) {
      const int row = (t & 1) ? n_unm + (t >> 1) : rowmap[t >> 1];
      rci[(long long)b * (N - 1) + (t - 1)] = (float)(row - 1);
    }
  }
}

// add + merge + LayerNorm in one launch: tome_merge_kernel's structure on the fp32 residual stream (C = 128 * CPL)
template <int CPL>
__global__ void __launch_bounds__(kThreads, CPL <= 3 ? 4 : 2)
tome_merge_ln_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ branch, const float* __restrict__ size,
                     const int64_t* __restrict__ unm_idx, const int64_t* __restrict__ src_idx,
                     const int64_t* __restrict__ dst_idx, int N, int r, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float* __restrict__ x_out, float* __restrict__ size_out,
                     float* __restrict__ rci, __nv_bfloat16* __restrict__ y) {
  extern __shared__ float smem[];
  constexpr int C = 128 * CPL;
  const int na = (N + 1) / 2, n_unm = na - r, n_out = N - r;
  float* zs = smem;
  int* unm = reinterpret_cast<int*>(zs + N);
  int* src = unm + n_unm;
  int* dst = src + r;
  int* rowmap = dst + r;
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_size = size != nullptr;
  for (int t = tid; t < N; t += kThreads) zs[t] = has_size ? size[(long long)b * N + t] : 1.f;
  for (int q = tid; q < n_unm; q += kThreads) unm[q] = clamp_idx(unm_idx[(long long)b * n_unm + q], na);
  for (int s = tid; s < r; s += kThreads) {
    src[s] = clamp_idx(src_idx[(long long)b * r + s], na);
    dst[s] = clamp_idx(dst_idx[(long long)b * r + s], N / 2);
  }
  __syncthreads();
  const float* xb = x + (long long)b * N * C;
  float* ob = x_out + (long long)b * n_out * C;
  MergeLn ln{branch ? branch + (long long)b * N * C : nullptr, gamma, beta, eps, nullptr};
  for (int q = blockIdx.x * kWarps + warp; q < n_out; q += gridDim.x * kWarps) {
    const int j = q < n_unm ? -1 : q - n_unm;
    const int t0 = q < n_unm ? 2 * unm[q] : 2 * j + 1;
    float zsum;
    ln.yrow = y + ((long long)b * n_out + q) * C;
    merge_row_vec<float, CPL, true>(xb, ob + (long long)q * C, C, lane, t0, j, src, dst, r, zs, has_size, true, zsum, &ln);
    if (lane == 0) size_out[(long long)b * n_out + q] = zsum;
  }
  if (rci != nullptr && blockIdx.x == 0) {
    for (int q = tid; q < n_unm; q += kThreads) rowmap[unm[q]] = q;
    for (int s = tid; s < r; s += kThreads) rowmap[src[s]] = n_unm + dst[s];
    __syncthreads();
    for (int t = 1 + tid; t < N; t += kThreads) {
      const int row = (t & 1) ? n_unm + (t >> 1) : rowmap[t >> 1];
      rci[(long long)b * (N - 1) + (t - 1)] = (float)(row - 1);
    }
  }
}

}  // namespace
}  // namespace tokred

using namespace tokred;

extern "C" int tokred_tome_effective_r(int N, int r, int class_token) {
  const int cap = (N - ((class_token & 1) ? 1 : 0) - ((class_token & 2) ? 1 : 0)) / 2;
  const int e = r < cap ? r : cap;
  return e > 0 ? e : 0;
}

extern "C" int tokred_tome_match(const void* metric, int metric_dtype, int heads, int64_t token_stride, int B, int N, int D, int r,
                                 int class_token, int score_lowp, int64_t* unm_idx, int64_t* src_idx, int64_t* dst_idx,
                                 void* stream) {
  const char* what = "tokred_tome_match";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(metric && unm_idx && src_idx && dst_idx, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(metric_dtype), "%s: bad metric dtype %d", what, metric_dtype);
  TOKRED_REQUIRE(B >= 0 && N >= 2 && D >= 1, "%s: bad shape B=%d N=%d D=%d", what, B, N, D);
  const int re = tokred_tome_effective_r(N, r, class_token);
  TOKRED_REQUIRE(re >= 1, "%s: effective r = %d (r=%d, N=%d): nothing to merge, caller must skip", what, re, r, N);
  if (B == 0) return TOKRED_OK;
  const int na = (N + 1) / 2, nb = N / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (heads < 1) heads = 1;
  if (token_stride <= 0) token_stride = (int64_t)heads * D;
  TOKRED_REQUIRE(token_stride >= (int64_t)heads * D, "%s: token_stride %lld < heads*D", what, (long long)token_stride);
  // tensor-core path: bf16-rounded similarity (autocast), one 128-row M tile, N <= 256 columns
  const int Dp = (D + 15) & ~15, Np = (nb + 15) & ~15;
  if (score_lowp == 1 && metric_dtype == TOKRED_BF16 && D == 64 && na <= 128 && Np <= 256 && aligned16(metric) &&
      token_stride % 8 == 0) {
    const size_t smem2 = (size_t)(16 + Np / 8) * 1024 + 128 * 4 + 128 * 4 + 16 + 128 + 16;
    if (heads > 1) {
      if (int e = allow_smem(tome_match_tc2_kernel<true>, smem2, what)) return e;
      tome_match_tc2_kernel<true><<<B, kTc2Threads, smem2, st>>>((const __nv_bfloat16*)metric, token_stride, heads, N, re,
                                                                class_token, unm_idx, src_idx, dst_idx);
    } else {
      if (int e = allow_smem(tome_match_tc2_kernel<false>, smem2, what)) return e;
      tome_match_tc2_kernel<false><<<B, kTc2Threads, smem2, st>>>((const __nv_bfloat16*)metric, token_stride, 1, N, re,
                                                                 class_token, unm_idx, src_idx, dst_idx);
    }
    return finish_launch(what);
  }
  if (heads != 1 || token_stride != D) {
    set_error("%s: per-head / strided metric (heads=%d, token_stride=%lld) needs the bf16 tensor-core path (bf16, D=64, score_lowp=1)",
              what, heads, (long long)token_stride);
    return TOKRED_ERR_UNSUPPORTED;
  }
  const size_t tc_smem = (size_t)(16 + Np / 8) * (Dp / 8) * 128 + (((size_t)N * D * dtype_size(metric_dtype) + 15) & ~(size_t)15) +
                         128 * 4 + 128 * 4 + 128 + 16;
  if (score_lowp && (score_lowp & 2) == 0 && na <= 128 && Np <= 256 && tc_smem <= 200 * 1024) {
    if (metric_dtype == TOKRED_F32) {
      if (int e = allow_smem(tome_match_tc_kernel<float>, tc_smem, what)) return e;
      tome_match_tc_kernel<float><<<B, kTcThreads, tc_smem, st>>>((const float*)metric, N, D, re, class_token, unm_idx,
                                                                  src_idx, dst_idx);
    } else {
      if (int e = allow_smem(tome_match_tc_kernel<__nv_bfloat16>, tc_smem, what)) return e;
      tome_match_tc_kernel<__nv_bfloat16><<<B, kTcThreads, tc_smem, st>>>((const __nv_bfloat16*)metric, N, D, re,
                                                                          class_token, unm_idx, src_idx, dst_idx);
    }
    return finish_launch(what);
  }
  score_lowp = score_lowp ? 1 : 0;
  const size_t smem = ((size_t)(na + nb) * (D + 1) + (size_t)na * (nb + 1) + 2 * na) * 4 + na;
  if (metric_dtype == TOKRED_F32) {
    if (int e = allow_smem(tome_match_kernel<float>, smem, what)) return e;
    tome_match_kernel<float><<<B, kThreads, smem, st>>>((const float*)metric, N, D, re, class_token, score_lowp,
                                                        unm_idx, src_idx, dst_idx);
  } else {
    if (int e = allow_smem(tome_match_kernel<__nv_bfloat16>, smem, what)) return e;
    tome_match_kernel<__nv_bfloat16><<<B, kThreads, smem, st>>>((const __nv_bfloat16*)metric, N, D, re, class_token,
                                                                score_lowp, unm_idx, src_idx, dst_idx);
  }
  return finish_launch(what);
}

extern "C" int tokred_tome_merge(const void* x, int x_dtype, const void* size, const int64_t* unm_idx,
                                 const int64_t* src_idx, const int64_t* dst_idx, int B, int N, int C, int r,
                                 void* x_out, void* size_out, float* reduced_cluster_idx, int divide, void* stream) {
  const char* what = "tokred_tome_merge";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && unm_idx && src_idx && dst_idx && x_out && size_out, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(x_dtype), "%s: bad x dtype %d", what, x_dtype);
  TOKRED_REQUIRE(B >= 0 && N >= 2 && C >= 1, "%s: bad shape B=%d N=%d C=%d", what, B, N, C);
  TOKRED_REQUIRE(r >= 1 && r <= N / 2 && r <= (N + 1) / 2, "%s: r=%d outside [1, %d]", what, r, N / 2);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (B == 0) return TOKRED_OK;
  const int na = (N + 1) / 2, n_unm = na - r, n_out = N - r;
  const size_t smem = (size_t)(N + n_unm + 2 * r + na) * 4;
  const int ve = x_dtype == TOKRED_F32 ? 4 : 8;
  const bool vec = (C % ve == 0) && aligned16(x) && aligned16(x_out);
  const int cpl = vec ? ceil_div(C / ve, 32) : 0;
  // ONE wave: as many CTAs as are resident at once (4 per SM up to 3 chunks per lane, else 2), never more -- a
  // second, partly filled wave costs more than larger CTAs do (B=256, N=197: 512 CTAs 28.7 us, 1024 CTAs 30.7 us,
  // 2304 CTAs 32.8 us); batches beyond the resident count simply queue whole images.
  const int resident = kNumSMs * ((cpl >= 1 && cpl <= 3) ? 4 : 2);
  int splits = resident / B;
  splits = max(1, min(splits, ceil_div(n_out, kWarps)));
  dim3 grid(splits, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(T, CPL)                                                                                          \
  do {                                                                                                          \
    if (int e = allow_smem(tome_merge_kernel<T, CPL>, smem, what)) return e;                                    \
    tome_merge_kernel<T, CPL><<<grid, kThreads, smem, st>>>((const T*)x, (const T*)size, unm_idx, src_idx, dst_idx, \
                                                            N, C, r, (T*)x_out, (T*)size_out, reduced_cluster_idx, divide); \
  } while (0)
#define DISPATCH(T)                         \
  switch (cpl) {                            \
    case 1: LAUNCH(T, 1); break;            \
    case 2: LAUNCH(T, 2); break;            \
    case 3: LAUNCH(T, 3); break;            \
    case 4: LAUNCH(T, 4); break;            \
    case 5: case 6: LAUNCH(T, 6); break;    \
    default: LAUNCH(T, 0); break;           \
  }
  if (x_dtype == TOKRED_F32) { DISPATCH(float) } else { DISPATCH(__nv_bfloat16) }
#undef DISPATCH
#undef LAUNCH
  return finish_launch(what);
}

extern "C" int tokred_tome_merge_ln(const float* x, const void* branch, const float* size, const int64_t* unm_idx,
                                    const int64_t* src_idx, const int64_t* dst_idx, int B, int N, int C, int r,
                                    const float* gamma, const float* beta, float eps, float* x_out, float* size_out,
                                    float* reduced_cluster_idx, void* y, void* stream) {
  const char* what = "tokred_tome_merge_ln";
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(x && unm_idx && src_idx && dst_idx && x_out && size_out && gamma && beta && y, "%s: null tensor", what);
  TOKRED_REQUIRE(B >= 0 && N >= 2 && C >= 1, "%s: bad shape B=%d N=%d C=%d", what, B, N, C);
  TOKRED_REQUIRE(r >= 1 && r <= N / 2 && r <= (N + 1) / 2, "%s: r=%d outside [1, %d]", what, r, N / 2);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (C % 128 != 0 || C > 768) {
    set_error("%s: C=%d (needs a multiple of 128 up to 768)", what, C);
    return TOKRED_ERR_UNSUPPORTED;
  }
  TOKRED_REQUIRE(aligned16(x) && aligned16(x_out) && aligned16(gamma) && aligned16(beta) &&
                     (!branch || (reinterpret_cast<uintptr_t>(branch) & 7u) == 0) && (reinterpret_cast<uintptr_t>(y) & 7u) == 0,
                 "%s: tensors must be 16-byte aligned", what);
  const int na = (N + 1) / 2, n_unm = na - r, n_out = N - r;
  const size_t smem = (size_t)(N + n_unm + 2 * r + na) * 4;
  const int cpl = C / 128;
  // ONE wave of resident CTAs (4 per SM at 64 registers up to C = 384, else 2), like tome_merge.  Measured at B=256, N=197:
  // 512 CTAs 43.1 us, 768 CTAs 47.9 us, 1024 CTAs 44.5 us, 1536 CTAs 44.8 us (at 72 registers / 3 CTAs per SM: 54 / 47 us)
  const int resident = kNumSMs * (cpl <= 3 ? 4 : 2);
  int splits = resident / B;
  splits = max(1, min(splits, ceil_div(n_out, kWarps)));
  dim3 grid(splits, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(CPL)                                                                                                     \
  do {                                                                                                                  \
    if (int e = allow_smem(tome_merge_ln_kernel<CPL>, smem, what)) return e;                                            \
    tome_merge_ln_kernel<CPL><<<grid, kThreads, smem, st>>>(x, (const __nv_bfloat16*)branch, size, unm_idx, src_idx,    \
                                                            dst_idx, N, r, gamma, beta, eps, x_out, size_out,           \
                                                            reduced_cluster_idx, (__nv_bfloat16*)y);                    \
  } while (0)
  switch (cpl) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    case 5: LAUNCH(5); break;
    default: LAUNCH(6); break;
  }
#undef LAUNCH
  return finish_launch(what);
}
