// Persistent, warp-specialised pairwise-distance pipeline for the clustering kernels (DPC-KNN, K-Medoids, cdist).
//
//   grid = min(B, 148) CTAs of 512 threads, each CTA walks images b = blockIdx.x, blockIdx.x + gridDim.x, ...
//   LOADER warps 0-13:  stream the image's token rows HBM -> registers (two 32-column chunks ahead), split every
//        fp32 value x into two fp16 terms h = fp16(x), l = fp16(x - h) (|x - h - l| <= 2^-22 |x|) and write them as
//        canonical K-major UMMA tiles into a 2-stage shared-memory ring (stage_full mbarrier, no CTA barrier).
//   MMA warp 14:  one lane issues, per 16-column k-step, three tcgen05.mma kind::f16 (h.h^T, h.l^T, l.h^T) into ONE
//        fp32 TMEM accumulator; tcgen05.commit hands the stage back to the loaders (the first version issued from
//        a loader thread behind a group barrier: 1.4 us per chunk, the stamps showed the loaders waiting on it).
//        Why fp16 and not tf32: both carry an 11-bit significand, but the f16 kind runs at twice the tf32 rate and
//        its operands are half the shared-memory bytes; the dropped l.l^T term is 2^-22 relative.
//        The Gram matrix is symmetric, so the second M tile (rows 128..) only computes columns 128.. (N = Np - 128):
//        69 % of the MMA work of the full matrix at P = 196.
//        The squared row norms |x_i|^2 are accumulated in fp32 from the very registers being staged (the tensor
//        core's truncating accumulator is ~20x less accurate on the diagonal, whose terms are all positive and 20x
//        larger than the off-diagonal dot products: measured 4.9e-5 vs 1.0e-5 on d ~ 27).
//   BACK group (warps 15-22):  waits for the accumulator (tcgen05.commit -> mbarrier) and the row norms, reads the
//        upper triangle from TMEM (thread = accumulator row), forms
//        D_ij = sqrt(max(|x_i|^2 + |x_j|^2 - 2 g_ij, 1e-30)) * scale  — the matmul form torch.cdist uses for P > 25 —
//        and stores it to BOTH D[i][j] and D[j][i] of the shared-memory matrix (bit-symmetric by construction, odd
//        row stride: both stores conflict-free), releases the accumulator, and runs the operator's epilogue
//        (density / ranking / assignment passes) on D while the FRONT group is already staging and multiplying the
//        NEXT image: the tensor pipe, the HBM stream and the shared-memory passes overlap across images instead of
//        running back to back inside one image (r01: 152 us at B=256, tensor pipe 22 %, HBM 6 %).
//
// Synchronisation (all mbarriers): stage_full[2] (count 448) MMA warp<-loaders; stage_free[2] (tcgen05.commit)
// loaders<-tensor core; acc_full (tcgen05.commit) BACK<-tensor core; acc_empty (count 256) MMA warp<-BACK; sq_full
// (count 448) BACK<-loaders (row norms, double buffered per image parity); named barrier 1 inside the BACK group.
#pragma once
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"
#include "umma.cuh"

namespace tokred {
namespace pipe {

// 14 loader warps + 1 MMA warp + 8 BACK warps = 736 threads at 80 registers.
// Measured (tools/diag/pipe_stamps.py, B=256 P=196 C=384): a chunk takes ~1.2 us end to end whatever the data size
// (P=49 too): ~0.5 us stage + arrive of the FIRST loader warp, up to 1 us until the LAST one arrives (the loaders are
// issue-bound: ~20 instructions per float once address arithmetic and predicates are counted), 0.8 us for one lane to
// issue 12 MMAs + 2 commits, 0.55 us MMA execution.  Not the bottleneck (measured, so they are not retried blind):
// more bytes in flight (3 register buffers: slower), L2 prefetches (bulk, per line, 512 B per row: no change), an L2
// warm-up by the idle BACK group (no change), a third of the MMAs (no change), 7 vs 14 loader warps (no change).
constexpr int kLoad = 448;              // loader threads (warps 0-13); warp 14 issues the MMAs
constexpr int kItems = 2;               // (row, core column) items per loader thread and chunk: 2 x 448 >= 208 rows x 4
constexpr int kBackWarps = 8;           // (10 loader + 13 BACK warps measured slower: the front went from 14 to 22 us per image)
constexpr int kFront = kLoad + 32, kBack = kBackWarps * 32, kThreads = kFront + kBack;
constexpr int KC = 32;                  // contraction columns per stage (4 core columns of 8 halves)
constexpr int kStages = 2;
constexpr int kMaxP = 208;
constexpr uint32_t kSBO = (KC / 8) * 128;   // bytes between 8-row groups of an operand tile
enum { BAR_BACK = 1, BAR_FRONT = 2 };

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

// one lane polls, the warp follows: 32x fewer try_wait loops competing for issue slots with the working warps
// (the first version had every thread spin: 0.8 M spin-loop warp instructions per launch, and the BACK phases ran
// 2x slower while the loaders were active)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) umma::mbar_wait(bar, parity);
  __syncwarp();
}

__host__ __device__ inline int round16(int v) { return (v + 15) & ~15; }
__host__ __device__ inline size_t stage_bytes(int P) { return (size_t)(round16(P) / 8) * kSBO * 2; }   // hi + lo
__host__ __device__ inline size_t d_bytes(int P) { return (((size_t)P * (P | 1) * 4) + 15) & ~(size_t)15; }
__host__ __device__ inline size_t smem_bytes(int P, int extra_floats) {
  const size_t used = kStages * stage_bytes(P) + d_bytes(P) + (((size_t)(2 * round16(P) + extra_floats) * 4 + 15) & ~(size_t)15) + 8 * 8 + 16;
  // an M = 128 MMA always reads 16 row groups of its A operand: for small P that runs past the operand tile (the rows
  // feed accumulator lanes nobody reads) and must still land inside this CTA's allocation
  const size_t overread = kStages * stage_bytes(P) - stage_bytes(P) / 2 + (size_t)(P > 128 ? 32 : 16) * kSBO + 512;
  return used > overread ? used : overread;
}

struct Ctx {
  unsigned char* stages;
  float* D; int DS;
  float* sq;             // [2][Np] squared row norms, buffer = image parity
  float* extra;          // epilogue vectors
  uint64_t* bars;        // [0..1] stage_free, [2] acc_full, [3] acc_empty, [4] sq_full, [5..6] stage_full
  uint32_t tmem_base;
  int P, Np, N2, C, nchunk, n_img;
};

// x -> (h, l) for 8 consecutive values: two 16-byte core rows
__device__ __forceinline__ void split8(const float4& a, const float4& b, int4& h, int4& l) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __half2 hh[4], ll[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hh[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 back = __half22float2(hh[i]);
    ll[i] = __floats2half2_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);     // the remainder is exact in fp32
  }
  h = *reinterpret_cast<const int4*>(hh);
  l = *reinterpret_cast<const int4*>(ll);
}

// ------------------------------------------------------------------------------------------------ FRONT
// item t = tid + 448 u (u < 2): r8 = t & 7, kc = (t >> 3) & 3, rg = t >> 5  ->  a warp covers one 8-row group x 4 core
// columns: its global loads are whole 128-byte lines, its 16-byte shared stores are conflict-free (8 consecutive
// rows of one core column per quarter warp).
struct FrontRegs { float4 v[kItems][2]; };
struct FrontNorms {
  float acc[kItems];                     // this thread's partial |x_row|^2 per item (its 8 columns of every chunk)
#ifdef TOKRED_STAMPS
  long long t_free = 0, t_data = 0, t_conv = 0, t_last = 0;
#endif
};

__device__ __forceinline__ void front_load(const float* __restrict__ x, long long xbs, int gidx, const Ctx& cx, bool vec, FrontRegs& r) {
  const int total = cx.n_img * cx.nchunk;
  if (gidx >= total) return;
  const int img = gidx / cx.nchunk, c = gidx - img * cx.nchunk;
  const float* xb = x + (long long)(blockIdx.x + (long long)img * gridDim.x) * xbs;
  const int tid = threadIdx.x;
#pragma unroll
  for (int u = 0; u < kItems; ++u) {
    const int t = tid + kLoad * u;
    const int row = ((t >> 5) << 3) + (t & 7), k = c * KC + ((t >> 3) & 3) * 8;
    r.v[u][0] = r.v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < cx.P) {
      const float* g = xb + (long long)row * cx.C + k;
      if (vec && k + 7 < cx.C) {
        r.v[u][0] = *reinterpret_cast<const float4*>(g);
        r.v[u][1] = *reinterpret_cast<const float4*>(g + 4);
      } else {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = (k + e < cx.C) ? g[e] : 0.f;
        r.v[u][0] = make_float4(f[0], f[1], f[2], f[3]);
        r.v[u][1] = make_float4(f[4], f[5], f[6], f[7]);
      }
    }
  }
}

__device__ __forceinline__ void front_stage(int gidx, const Ctx& cx, const FrontRegs& r, FrontNorms& nm) {
  const int tid = threadIdx.x, s = gidx & 1;
  const int img_of = gidx / cx.nchunk, c_of = gidx - img_of * cx.nchunk;
  if (c_of == 0) {
#pragma unroll
    for (int u = 0; u < kItems; ++u) nm.acc[u] = 0.f;
  }
  TOKRED_STAMP(tid == 0 && c_of == 0, img_of, 0);
#ifdef TOKRED_STAMPS
  const long long c0 = clock64();
#endif
  if (gidx >= 2) mbar_wait_warp(&cx.bars[s], (uint32_t)(((gidx >> 1) - 1) & 1));       // the MMAs that read this stage are done
#ifdef TOKRED_STAMPS
  const long long c1 = clock64();
  TOKRED_STAMP(tid == 0 && img_of == 0 && c_of < 32, 4, c_of);
  {   // first touch of the prefetched registers: time until the chunk's loads have landed
    float probe = r.v[0][0].x + r.v[1][1].w + r.v[1][0].y + r.v[0][1].z;
    asm volatile("" ::"f"(probe));
  }
  const long long c2 = clock64();
#endif
  unsigned char* hi = cx.stages + (size_t)s * stage_bytes(cx.P);
  unsigned char* lo = hi + stage_bytes(cx.P) / 2;
#pragma unroll
  for (int u = 0; u < kItems; ++u) {
    const int t = tid + kLoad * u;
    const int rg = t >> 5, row = (rg << 3) + (t & 7);
    {
      const float4 a = r.v[u][0], b = r.v[u][1];           // rows >= P hold zeros
      float acc = nm.acc[u];
      acc = fmaf(a.x, a.x, acc); acc = fmaf(a.y, a.y, acc); acc = fmaf(a.z, a.z, acc); acc = fmaf(a.w, a.w, acc);
      acc = fmaf(b.x, b.x, acc); acc = fmaf(b.y, b.y, acc); acc = fmaf(b.z, b.z, acc); acc = fmaf(b.w, b.w, acc);
      nm.acc[u] = acc;
    }
    if (row < cx.P) {
      int4 h, l;
      split8(r.v[u][0], r.v[u][1], h, l);
      const uint32_t off = (uint32_t)rg * kSBO + (uint32_t)((t >> 3) & 3) * 128u + (uint32_t)(t & 7) * 16u;
      *reinterpret_cast<int4*>(hi + off) = h;
      *reinterpret_cast<int4*>(lo + off) = l;
    }
  }
  if (c_of == cx.nchunk - 1) {
    // the 4 core-column partials of a row sit in lanes l, l+8, l+16, l+24 of one warp: fixed-order butterfly, then the
    // norms of this image go to the buffer of its parity and the BACK group is told
    float* sq = cx.sq + (img_of & 1) * cx.Np;
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      float v = nm.acc[u];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      const int t = tid + kLoad * u, row = ((t >> 5) << 3) + (t & 7);
      if (((t >> 3) & 3) == 0 && row < cx.P) sq[row] = v;
    }
    mbar_arrive(&cx.bars[4]);
  }
  umma::fence_proxy_async_smem();
  mbar_arrive(&cx.bars[5 + s]);                               // stage_full: this thread's part of chunk gidx is in place
#ifdef TOKRED_STAMPS
  const long long c3 = clock64();
  TOKRED_STAMP(tid == 0 && img_of == 0 && c_of < 32, 5, c_of);
  nm.t_free += c1 - c0; nm.t_data += c2 - c1; nm.t_conv += c3 - c2;
  if (tid == 0 && g_tokred_stamps && gidx == cx.n_img * cx.nchunk - 1) {
    unsigned long long* o = g_tokred_stamps + ((size_t)blockIdx.x * 8 + 7) * 32;
    o[0] = (unsigned long long)nm.t_free; o[1] = (unsigned long long)nm.t_data; o[2] = (unsigned long long)nm.t_conv;
  }
#endif
}

// MMA warp: waits for a full stage, lane 0 issues the chunk's MMAs and commits the stage back to the loaders
__device__ __forceinline__ void mma_run(const Ctx& cx) {
  const int lane = threadIdx.x & 31, total = cx.n_img * cx.nchunk;
  const uint32_t idesc1 = umma::instr_desc(umma::FMT_F16, 128, (uint32_t)cx.Np);
  const uint32_t idesc2 = umma::instr_desc(umma::FMT_F16, 128, (uint32_t)(cx.N2 > 0 ? cx.N2 : 16));
  for (int g = 0; g < total; ++g) {
    const int s = g & 1, img = g / cx.nchunk, c = g - img * cx.nchunk;
    mbar_wait_warp(&cx.bars[5 + s], (uint32_t)((g >> 1) & 1));
    if (c == 0 && img > 0) mbar_wait_warp(&cx.bars[3], (uint32_t)((img - 1) & 1));     // accumulator drained by the BACK group
    umma::tc_fence_after_sync();
    TOKRED_STAMP(lane == 0 && img == 0 && c < 32, 2, c);
    if (lane == 0) {
      const unsigned char* hi = cx.stages + (size_t)s * stage_bytes(cx.P);
      const uint32_t h0 = umma::smem_u32(hi), l0 = h0 + (uint32_t)(stage_bytes(cx.P) / 2);
#pragma unroll
      for (int ks = 0; ks < KC / 16; ++ks) {
        const uint32_t o = (uint32_t)ks * 256u;
        const uint32_t acc = (c > 0 || ks > 0) ? 1u : 0u;
        const uint64_t ah = umma::smem_desc_kmajor(h0 + o, 128, kSBO), al = umma::smem_desc_kmajor(l0 + o, 128, kSBO);
        umma::mma_bf16(cx.tmem_base, ah, ah, idesc1, acc);       // kind::f16 with fp16 operands (format in idesc)
        umma::mma_bf16(cx.tmem_base, ah, al, idesc1, 1u);
        umma::mma_bf16(cx.tmem_base, al, ah, idesc1, 1u);
        if (cx.N2 > 0) {      // rows 128.. x columns 128.. (upper triangle of the second M tile)
          const uint32_t o2 = o + 16u * kSBO;
          const uint64_t bh = umma::smem_desc_kmajor(h0 + o2, 128, kSBO), bl = umma::smem_desc_kmajor(l0 + o2, 128, kSBO);
          umma::mma_bf16(cx.tmem_base + (uint32_t)cx.Np, bh, bh, idesc2, acc);
          umma::mma_bf16(cx.tmem_base + (uint32_t)cx.Np, bh, bl, idesc2, 1u);
          umma::mma_bf16(cx.tmem_base + (uint32_t)cx.Np, bl, bh, idesc2, 1u);
        }
      }
      umma::mma_commit(&cx.bars[s]);
      if (c == cx.nchunk - 1) umma::mma_commit(&cx.bars[2]);
      TOKRED_STAMP(c == 0, img, 1);
      TOKRED_STAMP(c == cx.nchunk - 1, img, 2);
      TOKRED_STAMP(img == 0 && c < 32, 3, c);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void loader_run(const float* __restrict__ x, long long xbs, const Ctx& cx) {
  const bool vec = (cx.C % 4 == 0) && (xbs % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  const int total = cx.n_img * cx.nchunk;
  FrontRegs ra, rb;
  FrontNorms nm;
  front_load(x, xbs, 0, cx, vec, ra);
  front_load(x, xbs, 1, cx, vec, rb);
  for (int g = 0; g < total; g += 2) {
    front_stage(g, cx, ra, nm);
    front_load(x, xbs, g + 2, cx, vec, ra);
    if (g + 1 < total) {
      front_stage(g + 1, cx, rb, nm);
      front_load(x, xbs, g + 3, cx, vec, rb);
    }
  }
}

// ------------------------------------------------------------------------------------------------ BACK: TMEM -> D
__device__ __forceinline__ float sqrt_approx(float v) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));      // MUFU.SQRT, <= 1 ulp: far inside the 3-term split's own error
  return r;
}
// 16 consecutive columns j0.. of accumulator row i: all 16 distances are computed first (independent chains the
// scheduler can overlap: the first version branched per element and ran at 0.08 IPC per warp), then the upper-triangle
// ones are stored to D[i][j] and D[j][i].  gdj = row norms of columns j0..j0+15 (16-byte aligned).
__device__ __forceinline__ void drain16(float* D, int DS, int P, int i, int j0, float gi, const float* gdj, const uint32_t (&v)[16],
                                        float post_scale) {
  float gj[16], d[16];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const float4 t = *reinterpret_cast<const float4*>(gdj + 4 * q4);
    gj[4 * q4] = t.x; gj[4 * q4 + 1] = t.y; gj[4 * q4 + 2] = t.z; gj[4 * q4 + 3] = t.w;
  }
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    const float d2 = (gi + gj[jj]) - 2.0f * __uint_as_float(v[jj]);
    d[jj] = sqrt_approx(fmaxf(d2, 1e-30f)) * post_scale;
  }
  if (i < P) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const int j = j0 + jj;
      if (j > i && j < P) {
        D[i * DS + j] = d[jj];
        D[j * DS + i] = d[jj];
      }
    }
  }
}

// bt = thread index inside the BACK group (0..255).  Fills cx.D for the image whose accumulator is complete.
__device__ __forceinline__ void back_drain(const Ctx& cx, int it, float post_scale) {
  // a warp may only touch the TMEM lane quarter (warp index in the CTA) % 4; the BACK warps of a quarter take turns
  // on PAIRS of 16-column chunks: `half` = this warp's turn, `nq` = warps in its quarter
  const int bt = threadIdx.x - kFront, bw = bt >> 5, lane = bt & 31, q = (threadIdx.x >> 5) & 3, half = bw >> 2;
  const int nq = (kBackWarps - (bw & 3) + 3) >> 2;
  const int P = cx.P, Np = cx.Np, N2 = cx.N2, DS = cx.DS;
  float* D = cx.D;
  TOKRED_STAMP(bt == 0, it, 8);
  mbar_wait_warp(&cx.bars[2], (uint32_t)(it & 1));
  umma::tc_fence_after_sync();
  TOKRED_STAMP(bt == 0, it, 9);
  mbar_wait_warp(&cx.bars[4], (uint32_t)(it & 1));            // row norms of this image (written by the loaders)
  const float* gd = cx.sq + (it & 1) * Np;
  // B. upper triangle: thread = accumulator row i, the two warps of a lane quarter alternate 16-column chunks
  // Two 16-column TMEM loads are issued per tcgen05.wait::ld (the wait covers every outstanding load of the thread,
  // so this is the way to have two in flight); the two warps of a lane quarter alternate PAIRS of chunks.
  if (q * 32 < min(P, 128)) {
    const int i = q * 32 + lane;
    const float gi = i < P ? gd[i] : 0.f;
    const int nch = Np / 16, first = (q * 32) / 16;           // chunks below `first` hold only j <= i for every lane
    for (int ch = first + 2 * half; ch < nch; ch += 2 * nq) {
      const bool two = ch + 1 < nch;
      uint32_t va[16], vb[16];
      umma::tmem_ld16(umma::tmem_addr(cx.tmem_base, (uint32_t)(q * 32), (uint32_t)(ch * 16)), va);
      if (two) umma::tmem_ld16(umma::tmem_addr(cx.tmem_base, (uint32_t)(q * 32), (uint32_t)(ch * 16 + 16)), vb);
      umma::tmem_ld_wait();
      drain16(D, DS, P, i, ch * 16, gi, gd + ch * 16, va, post_scale);
      if (two) drain16(D, DS, P, i, ch * 16 + 16, gi, gd + ch * 16 + 16, vb, post_scale);
    }
  }
  if (N2 > 0 && 128 + q * 32 < P) {
    const int i = 128 + q * 32 + lane;
    const float gi = i < P ? gd[i] : 0.f;
    const int nch = N2 / 16, first = (q * 32) / 16;
    for (int ch = first + 2 * half; ch < nch; ch += 2 * nq) {
      const bool two = ch + 1 < nch;
      uint32_t va[16], vb[16];
      umma::tmem_ld16(umma::tmem_addr(cx.tmem_base, (uint32_t)(q * 32), (uint32_t)(Np + ch * 16)), va);
      if (two) umma::tmem_ld16(umma::tmem_addr(cx.tmem_base, (uint32_t)(q * 32), (uint32_t)(Np + ch * 16 + 16)), vb);
      umma::tmem_ld_wait();
      drain16(D, DS, P, i, 128 + ch * 16, gi, gd + 128 + ch * 16, va, post_scale);
      if (two) drain16(D, DS, P, i, 128 + ch * 16 + 16, gi, gd + 128 + ch * 16 + 16, vb, post_scale);
    }
  }
  // self distances: |x_i|^2 + |x_i|^2 - 2 x_i.x_i = 0 in exact arithmetic -> the clamp value (ATen's own diagonal is
  // cancellation noise up to 1.7e-2)
  for (int i = bt; i < P; i += kBack) D[i * DS + i] = sqrtf(1e-30f) * post_scale;
  umma::tc_fence_before_sync();
  mbar_arrive(&cx.bars[3]);                                   // the FRONT group may overwrite the accumulator
  bar_sync(BAR_BACK, kBack);
  TOKRED_STAMP(bt == 0, it, 10);
}

// carve shared memory, allocate TMEM, init the barriers.  All 512 threads call.
__device__ __forceinline__ Ctx setup(unsigned char* smem, int B, int P, int C, int extra_floats) {
  Ctx cx;
  cx.P = P; cx.C = C; cx.Np = round16(P); cx.N2 = P > 128 ? cx.Np - 128 : 0; cx.DS = P | 1;
  cx.nchunk = (C + KC - 1) / KC;
  cx.n_img = ((int)blockIdx.x < B) ? (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  cx.stages = smem;
  cx.D = reinterpret_cast<float*>(smem + kStages * stage_bytes(P));
  cx.sq = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(cx.D) + d_bytes(P));
  cx.extra = cx.sq + 2 * cx.Np;
  cx.bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(cx.sq) + ((((size_t)2 * cx.Np + extra_floats) * 4 + 15) & ~(size_t)15));
  uint32_t* slot = reinterpret_cast<uint32_t*>(cx.bars + 7);
  const uint32_t ncols = P > 128 ? 512u : umma::tmem_cols_pow2((uint32_t)cx.Np);
  if ((threadIdx.x >> 5) == 0) umma::tmem_alloc(slot, ncols);
  if (threadIdx.x == 0) {
    umma::mbar_init(&cx.bars[0], 1); umma::mbar_init(&cx.bars[1], 1); umma::mbar_init(&cx.bars[2], 1);
    umma::mbar_init(&cx.bars[3], kBack);
    umma::mbar_init(&cx.bars[4], kLoad);
    umma::mbar_init(&cx.bars[5], kLoad); umma::mbar_init(&cx.bars[6], kLoad);
    umma::fence_mbar_init();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  cx.tmem_base = *slot;
  return cx;
}
__device__ __forceinline__ void teardown(const Ctx& cx) {
  umma::tc_fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) umma::tmem_dealloc(cx.tmem_base, cx.P > 128 ? 512u : umma::tmem_cols_pow2((uint32_t)cx.Np));
}

}  // namespace pipe
}  // namespace tokred
