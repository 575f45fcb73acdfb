// Tensor-core Sinkhorn / PatchMerger / SiT, version 2: bulk-copy fed, warp-specialised.
//
// v1 (softmerge_tc.cu) converted every operand element to bf16 on the CUDA cores inside both GEMM loops; phase
// stamps showed it issue-bound there (GEMM 1 staging 41 %, GEMM 2 staging + store 30 % of the kernel, tensor pipe
// 7 % active).  v2 converts each token ONCE:
//   pack kernel   Q [K,C] fp32 -> bf16 in the canonical K-major stage image, one contiguous block per 64-column chunk.
//   phase 0       token rows stream HBM -> smem through a ring of cp.async.bulk copies (one per pair of adjacent rows,
//                 issued by warp 16, 16-32 pairs in flight); a worker warp normalises a pair in registers, rounds it
//                 to bf16 and writes it to a per-image scratch of CORE-MATRIX TILES  [p/8][c/8][8 rows x 16 B]
//                 (32-byte STG.256: two tokens per lane; L2-resident, 2 B/element).  The same 128-byte tile is a
//                 K-major core matrix for GEMM 1 (rows = tokens, 16 B along C) and an MN-major core matrix for GEMM 2
//                 (rows = the contraction index p, 16 B along the N index c).  A ring slot is released only after a
//                 fence.proxy.async (generic reads, then async-proxy refill).  SiT: the raw logits block is
//                 bulk-copied in the same phase.
//   phase 1       warp 16 streams 2-4 stages with cp.async.bulk (1 KB per token group, one lane each, + one block of
//                 Q per chunk) onto mbarriers (expect_tx); warp 17 issues tcgen05.mma and hands stages back with
//                 tcgen05.commit.
//   phase 2       accumulator -> bf16 Z in smem (it IS bf16 in the reference).
//   phase 3       weights: register-row softmax (PatchMerger; SiT after a shared -> shared transpose of its logits)
//                 or log-domain Sinkhorn (row passes in registers, column passes split over K with an online
//                 log-sum-exp, ex2.approx/FMA exponentials); W -> global fp32 + bf16 A operand of GEMM 2.
//   phase 4       same producer / MMA warps: B operand = 2 KB tile segments used MN-major; 16 worker warps drain two
//                 TMEM accumulator sets (tcgen05.ld) and store bf16 rows (STG.256) while the next chunk's MMAs run.
// Measured steps and dead ends: profiles/softmerge_phases_r01.txt, DESIGN.md section 7.
// 576 threads: warps 0-15 workers, warp 16 copy producer, warp 17 MMA issuer.
#include <math_constants.h>

#include "common.cuh"
#include "umma.cuh"

namespace tokred {
extern long long* g_phase_dbg;
namespace {

constexpr int kWorkers = 512;
constexpr int kThreads = kWorkers + 64;
constexpr int kWWarps = kWorkers / 32;
constexpr int kMaxK = 208, kMaxP = 208;
constexpr int S1MAX = 4;   // GEMM 1 stages: as many (2..4) as fit under the Z / GEMM-2 areas, see make_geo
constexpr int S2 = 2;      // GEMM 2 B-operand stages
constexpr int NC2 = 128;   // output columns per accumulator set

enum { MODE_SINKHORN = 0, MODE_PATCHMERGER = 1, MODE_SIT = 2 };

constexpr int kRingMax = 32;      // pair slots of the phase-0 token ring

struct Geo {
  int Np, NG, K8, n_mt, Cc8, Cc16, nchunk1, nchunk2, PSb, Ks, S1;
  uint32_t sbo2;
  size_t q_chunk_bytes, q_bytes, tile_row_bytes, img_bytes;
  size_t stage1A, stage1B, z_off, xt_off, vec_off, total;
};

__host__ __device__ inline Geo make_geo(int P, int C, int K) {
  Geo g;
  g.Np = (P + 15) & ~15;
  g.NG = g.Np / 8;                       // token groups (incl. zero padding)
  g.K8 = (K + 7) / 8;
  g.n_mt = (K + 127) / 128;
  g.Cc8 = (C + 7) / 8;
  g.Cc16 = (g.Cc8 + 15) & ~15;           // 16-byte cores per token row, padded to whole GEMM-2 chunks
  g.nchunk1 = g.Cc16 / 8;
  g.nchunk2 = g.Cc16 / 16;
  g.q_chunk_bytes = (size_t)g.K8 * 1024;
  g.q_bytes = (size_t)g.nchunk1 * g.q_chunk_bytes;
  g.tile_row_bytes = (size_t)g.Cc16 * 128;
  g.img_bytes = (size_t)g.NG * g.tile_row_bytes;
  g.sbo2 = (uint32_t)g.NG * 128 + 16;    // W operand (K-major over p): +16 keeps 8-row groups on different banks
  // A stage = the live Q rows only; the MMA of the last M tile also reads the (n_mt*128 - K) rows behind them, which
  // are whatever follows in shared memory: garbage rows of A only produce accumulator rows >= K, never read.
  g.stage1A = g.q_chunk_bytes;
  g.stage1B = (size_t)g.NG * 1024;
  const size_t stage1 = g.stage1A + g.stage1B;
  const size_t wop = (size_t)32 * g.sbo2;
  int psb = (P + 1) & ~1;
  if (((psb / 2) & 1) == 0) psb += 2;
  g.PSb = psb;
  int ks = (K + 1) & ~1;
  if (((ks / 2) & 1) == 0) ks += 2;
  g.Ks = ks;
  const size_t zb = (size_t)K * psb * 2 > (size_t)P * ks * 2 ? (size_t)K * psb * 2 : (size_t)P * ks * 2;
  const size_t xt = S2 * (size_t)g.NG * 2048;
  // Shared-memory plan (bytes from the base):
  //   [0, S1*stage1)          GEMM-1 stages (phase 0: LayerNorm parameters + token ring) | later [0, wop) W operand
  //   [z_off, z_off + zb)     Z (bf16) / staged SiT logits   (z_off = wop: live together with the W operand; written
  //                           only after the last GEMM-1 MMA has committed, so it may overlap the stages)
  //   [xt_off, xt_off + xt)   GEMM-2 B stages (xt_off = wop)  (overlaps Z, which is dead once W is built)
  const size_t wop_al = (wop + 127) & ~(size_t)127;
  g.z_off = wop_al;
  g.xt_off = wop_al;
  const size_t end0 = g.z_off + zb, end1 = g.xt_off + xt;
  size_t endm = end0 > end1 ? end0 : end1;
  const size_t tail = (size_t)g.n_mt * 16 * 1024 - g.stage1A;      // over-read of the last stage's A operand
  int s1 = (int)((endm > tail ? endm - tail : 0) / stage1);
  s1 = s1 > S1MAX ? S1MAX : s1;
  if (s1 < 2) { s1 = 2; if (2 * stage1 + tail > endm) endm = 2 * stage1 + tail; }
  g.S1 = s1;
  g.vec_off = (endm + 127) & ~(size_t)127;
  g.total = g.vec_off + (size_t)(3 * P + 2 * K) * 4 + 328 + 2 * kRingMax * 8;
  return g;
}

struct Tc2Params {
  const void* x;                  // [B,P,C], images xbs elements apart
  long long xbs;
  const unsigned char* q_packed;  // pack kernel output
  unsigned char* xh;              // [B] x img_bytes scratch tiles
  const float* ln_w;
  const float* ln_b;
  const __nv_bfloat16* logits;
  const float* scale_ptr;
  float scale, log_norm, ln_eps;
  int iters, P, C, K;
  __nv_bfloat16* out;
  float* weights;
  long long* dbg;                 // optional clock64() phase stamps of CTA 0 (tools/phase_times.py)
};

// ------------------------------------------------------------------------------------------ PTX helpers (bulk copy)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(umma::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}

template <typename T> __device__ __forceinline__ void load8(const T* p, bool vec, int valid, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, bool vec, int valid, float (&v)[8]) {
  if (vec && valid >= 8) {
    // one 256-bit load per lane (LDG.256; 32-byte aligned, checked by the caller): 1 KB contiguous per warp
    // instruction.  Two 16-byte halves cost twice the L1 wavefronts, and with L1::no_allocate the second half
    // re-fetched the sector from L2 (phase 0 ran 4x slower that way).
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < valid ? p[i] : 0.f;
  }
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, bool vec, int valid, float (&v)[8]) {
  if (vec && valid >= 8) {
    const int4 raw = ld_stream16(p);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(h[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < valid ? __bfloat162float(p[i]) : 0.f;
  }
}
__device__ __forceinline__ void st_global32(void* p, const int4& a, const int4& b) {      // STG.256, 32-byte aligned
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ int4 pack8(const float (&v)[8]) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return *reinterpret_cast<const int4*>(h);
}

// ------------------------------------------------------------------------------------------ Q pack kernel
// q [K][C] fp32 -> [chunk][row group][core 0..7][row % 8][8 bf16]: exactly the shared-memory image of a K-major stage.
__global__ void __launch_bounds__(256) pack_q_kernel(const float* __restrict__ q, int K, int C, int K8, unsigned char* __restrict__ out) {
  const int chunk = blockIdx.x;
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(q) & 15u) == 0);
  for (int e = threadIdx.x; e < K8 * 64; e += 256) {
    const int row = (e & 7) + ((e >> 6) << 3), core = (e >> 3) & 7;
    const int c = chunk * 64 + core * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (row < K && c < C) {
      const float* p = q + (long long)row * C + c;
      if (vec && c + 8 <= C) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c + i < C) v[i] = p[i];
      }
    }
    *reinterpret_cast<int4*>(out + (size_t)chunk * K8 * 1024 + (size_t)(row >> 3) * 1024 + core * 128 + (row & 7) * 16) = pack8(v);
  }
}

// ------------------------------------------------------------------------------------------ main kernel
template <typename T>
__device__ __forceinline__ void tile_row_load(const T* xb, int p, int P, int C, int lane, bool xvec, float (&v)[4][8]) {
  if (p < P) {
    const T* row = xb + (long long)p * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane * 8 + j * 256;
      if (c < C) load8<T>(row + c, xvec, C - c, v[j]);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
  }
}

// same lane -> channel mapping from a row staged in shared memory (C % 8 == 0 on this path)
template <typename T>
__device__ __forceinline__ void tile_row_load_smem(const T* row, bool live, int C, int lane, float (&v)[4][8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane * 8 + j * 256;
    if (live && c < C) {
      if constexpr (sizeof(T) == 4) {
        const float4 a = *reinterpret_cast<const float4*>(row + c), b = *reinterpret_cast<const float4*>(row + c + 4);
        v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w; v[j][4] = b.x; v[j][5] = b.y; v[j][6] = b.z; v[j][7] = b.w;
      } else {
        const int4 raw = *reinterpret_cast<const int4*>(row + c);
        const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j][i] = __bfloat162float(h[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
    }
  }
}

// Normalise (Sinkhorn: unit L2 norm; PatchMerger: LayerNorm; SiT: identity) one token held by a warp.  The LayerNorm
// parameters sit in shared memory as float4s in (chunk j, half, lane) order so that a warp's read is 512 contiguous
// bytes (the natural channel order is a 32-byte lane stride: 2-way bank conflicts on every read).  A dead row (token
// index >= P) stays all-zero.
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}
__device__ __forceinline__ void ln8(float (&v)[8], float rstd, float shift, const float4& g0, const float4& g1,
                                    const float4& b0, const float4& b1) {      // shift = -mean * rstd
  v[0] = fmaf(g0.x, fmaf(v[0], rstd, shift), b0.x); v[1] = fmaf(g0.y, fmaf(v[1], rstd, shift), b0.y);
  v[2] = fmaf(g0.z, fmaf(v[2], rstd, shift), b0.z); v[3] = fmaf(g0.w, fmaf(v[3], rstd, shift), b0.w);
  v[4] = fmaf(g1.x, fmaf(v[4], rstd, shift), b1.x); v[5] = fmaf(g1.y, fmaf(v[5], rstd, shift), b1.y);
  v[6] = fmaf(g1.z, fmaf(v[6], rstd, shift), b1.z); v[7] = fmaf(g1.w, fmaf(v[7], rstd, shift), b1.w);
}
// FULL8: C % 8 == 0, so a lane's 8-channel chunk is either entirely inside the row or entirely padding (one test per
// chunk instead of one per element: the per-element predicates were a third of the instructions of this phase).
template <int MODE, bool FULL8>
__device__ __forceinline__ void tile_row_normalise(float (&v)[4][8], bool live, int C, int lane, const float4* lng4,
                                                   const float4* lnb4, float ln_eps) {
  if (!live) return;
  if (MODE == MODE_SINKHORN) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) s[j] = fmaf(v[j][i], v[j][i], s[j]);
    const float t = warp_sum((s[0] + s[1]) + (s[2] + s[3]));
    const float inv = 1.0f / fmaxf(sqrtf(t), 1e-12f);       // rounded to bf16 right after: reciprocal form is fine
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) v[j][i] *= inv;
  } else if (MODE == MODE_PATCHMERGER) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) s[j] += v[j][i];
    const float invC = 1.0f / (float)C;
    const float mean = warp_sum((s[0] + s[1]) + (s[2] + s[3])) * invC;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j] = 0.f;
      const int c0 = lane * 8 + j * 256;
      if (FULL8) {
        if (c0 < C) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; s[j] = fmaf(d, d, s[j]); }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + i < C) { const float d = v[j][i] - mean; s[j] = fmaf(d, d, s[j]); }
      }
    }
    const float rstd = rsqrtf(warp_sum((s[0] + s[1]) + (s[2] + s[3])) * invC + ln_eps);
    const float shift = -mean * rstd;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lane * 8 + j * 256 < C)            // channels past C hold gamma = beta = 0 in the padded parameter arrays
        ln8(v[j], rstd, shift, lng4[(j * 2) * 32 + lane], lng4[(j * 2 + 1) * 32 + lane], lnb4[(j * 2) * 32 + lane],
            lnb4[(j * 2 + 1) * 32 + lane]);
  }
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void warp_max2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
  }
}
// Per-lane constants of the weight-row passes (one warp per row of W, lane holds token pairs lane, lane+32, ...):
// hoisted out of the row loops, which were ~350 instructions per row, most of them bounds tests and operand-offset
// arithmetic.  Dead slots (token >= P) re-read token 0 (finite) and carry dead = -inf.
struct RowMap {
  int poff[4];        // token index of the pair (0 for dead slots)
  uint32_t woff[4];   // byte offset of the pair inside a K-major row group of the W operand
  float dead[8];      // 0 for live slots, -inf for dead ones (added to the score)
  bool live0[4], live1[4];
};
__device__ __forceinline__ RowMap make_row_map(int P, int lane) {
  RowMap m;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = 2 * (lane + 32 * i);
    m.live0[i] = p < P; m.live1[i] = p + 1 < P;
    m.poff[i] = m.live0[i] ? p : 0;
    m.woff[i] = (uint32_t)((p >> 3) * 128 + (p & 7) * 2);
    m.dead[2 * i] = m.live0[i] ? 0.f : -CUDART_INF_F;
    m.dead[2 * i + 1] = m.live1[i] ? 0.f : -CUDART_INF_F;
  }
  return m;
}
__device__ __forceinline__ void load_scores(const __nv_bfloat16* zrow, const RowMap& m, float (&x)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 z = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(zrow + m.poff[i]));
    x[2 * i] = z.x; x[2 * i + 1] = z.y;
  }
}
// weights of one row: fp32 to global memory (float2 when the row is 8-byte aligned), bf16 pairs into the K-major A
// operand of GEMM 2 (wgrp = Wop + (k/8)*sbo2 + (k%8)*16); dead slots hold 0 and are not stored
__device__ __forceinline__ void store_weights(const float (&w)[8], float* wrow, unsigned char* wgrp, const RowMap& m) {
  const bool vec2 = (reinterpret_cast<uintptr_t>(wrow) & 7u) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (m.live0[i]) {
      if (m.live1[i] && vec2) *reinterpret_cast<float2*>(wrow + m.poff[i]) = make_float2(w[2 * i], w[2 * i + 1]);
      else {
        wrow[m.poff[i]] = w[2 * i];
        if (m.live1[i]) wrow[m.poff[i] + 1] = w[2 * i + 1];
      }
      *reinterpret_cast<__nv_bfloat162*>(wgrp + m.woff[i]) = __floats2bfloat162_rn(w[2 * i], m.live1[i] ? w[2 * i + 1] : 0.f);
    }
  }
}

// SiT: logits [P, K] (bf16) -> shared memory TRANSPOSED as rows [K][PSb], by `nt` cooperating threads (index t).
// Lanes take consecutive tokens for one group of 8 slots: 16-byte loads (4 kept in flight; the loop is latency bound)
// and conflict-free 2-byte stores.
__device__ __forceinline__ void stage_logits_transposed(__nv_bfloat16* Lt, const __nv_bfloat16* lb, int P, int K, int PSb,
                                                        int t, int nt) {
  if ((K & 7) == 0 && (reinterpret_cast<uintptr_t>(lb) & 15u) == 0) {
    const int K8 = K >> 3, items = P * K8;            // item = (slot group k8, token p), p fastest
    for (int e0 = t; e0 < items; e0 += 4 * nt) {
      int4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nt;
        if (e < items) { const int k8 = e / P, p = e - k8 * P; raw[u] = *reinterpret_cast<const int4*>(lb + (size_t)p * K + k8 * 8); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nt;
        if (e < items) {
          const int k8 = e / P, p = e - k8 * P;
          const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) Lt[(size_t)(k8 * 8 + j) * PSb + p] = h[j];
        }
      }
    }
  } else {
    for (int e = t; e < P * K; e += nt) { const int p = e / K, k = e - p * K; Lt[(size_t)k * PSb + p] = lb[e]; }
  }
  if (P & 1)
    for (int k = t; k < K; k += nt) Lt[(size_t)k * PSb + P] = __float2bfloat16_rn(0.f);      // pad column: finite
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, 1) soft_merge_tc2_kernel(Tc2Params prm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int P = prm.P, C = prm.C, K = prm.K;
  const Geo G = make_geo(P, C, K);
  unsigned char* R0 = smem;
  float* lng = reinterpret_cast<float*>(smem);            // [1024] LayerNorm gamma, lane-major float4 order; lives in the
  float* lnb = lng + 1024;                               // [1024] (idle) GEMM-1 stage area during phase 0 only
  float* s0 = reinterpret_cast<float*>(smem + G.vec_off);      // [P]
  float* s1 = s0 + P;                                    // [P]
  float* uvec = s1 + P;                                  // [K]
  float* vvec = uvec + K;                                // [P]
  float* aux = vvec + P;                                 // [K]   sit: softmax normalisers
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(aux + K) + 7) & ~(uintptr_t)7);
  uint64_t* full1 = bars;            // [S1MAX]
  uint64_t* empty1 = full1 + S1MAX;  // [S1MAX]
  uint64_t* acc1 = empty1 + S1MAX;   // [1]  GEMM 1 finished
  uint64_t* full2 = acc1 + 1;        // [S2]
  uint64_t* empty2 = full2 + S2;     // [S2]
  uint64_t* accfull = empty2 + S2;   // [2]
  uint64_t* accempty = accfull + 2;  // [2]
  uint64_t* pfull = accempty + 2;    // [kRingMax] phase-0 ring: pair of token rows landed
  uint64_t* pempty = pfull + kRingMax;  // [kRingMax] pair consumed
  uint64_t* lgbar = pempty + kRingMax;  // [1] SiT: raw logits landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lgbar + 1);

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = warp < kWWarps;
  const T* xb = reinterpret_cast<const T*>(prm.x) + (long long)b * prm.xbs;
  unsigned char* xh = prm.xh + (size_t)b * G.img_bytes;
  const bool xvec = (C % 8 == 0) && ((reinterpret_cast<uintptr_t>(xb) & (8 * sizeof(T) - 1)) == 0);

  // phase-0 ring: pairs of adjacent token rows (contiguous in x) streamed into the idle GEMM stage area.  SiT has no
  // GEMM 1: its ring lives in the (idle) Z / GEMM-2 stage area instead, and [0, 2PK) receives the raw logits.
  const int row_bytes = C * (int)sizeof(T);
  const size_t ring_off = MODE == MODE_SIT ? G.z_off : (MODE == MODE_PATCHMERGER ? 8192 : 0);   // 8 KB: LayerNorm parameters
  int ring_slots = (int)((G.vec_off - ring_off) / (size_t)(2 * row_bytes));
  ring_slots = ring_slots >= kRingMax ? kRingMax : (ring_slots >= 16 ? 16 : 0);
  const int ring_log = ring_slots == 32 ? 5 : 4;      // 16 or 32 slots
  const bool ring = ring_slots > 0 && xvec && (row_bytes % 16 == 0);
  const uint32_t lg_bytes = (uint32_t)(P * K * 2);
  const bool lg_bulk = MODE == MODE_SIT && ring && (lg_bytes % 16 == 0) &&
                       ((reinterpret_cast<uintptr_t>(prm.logits) & 15u) == 0);
  unsigned char* ring_base = smem + ring_off;

#define STAMP(i) do { if (prm.dbg && blockIdx.x == 0 && tid == 0) prm.dbg[i] = clock64(); } while (0)
  STAMP(0);
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < S1MAX; ++i) { umma::mbar_init(&full1[i], 1); umma::mbar_init(&empty1[i], 1); }
    umma::mbar_init(acc1, 1);
    for (int i = 0; i < S2; ++i) { umma::mbar_init(&full2[i], 1); umma::mbar_init(&empty2[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&accfull[i], 1); umma::mbar_init(&accempty[i], kWWarps); }
    for (int i = 0; i < kRingMax; ++i) { umma::mbar_init(&pfull[i], 1); umma::mbar_init(&pempty[i], 1); }
    umma::mbar_init(lgbar, 1);
    umma::fence_mbar_init();
  }
  if (MODE == MODE_PATCHMERGER)
    for (int c = tid; c < 1024; c += kThreads) {        // (chunk j, half, lane, e) order, zero beyond C
      const int j = c >> 8, ln = (c & 255) >> 3, half = (c & 7) >> 2, e = c & 3;
      const int slot = (((j * 2 + half) * 32 + ln) << 2) + e;
      lng[slot] = c < C ? prm.ln_w[c] : 0.f;
      lnb[slot] = c < C ? prm.ln_b[c] : 0.f;
    }
  __syncthreads();

  // ---- 0. token statistics + bf16 core-matrix tiles in ONE pass over x.  One warp per PAIR of adjacent tokens, rows
  //         in registers; each lane stores the two tokens' 16-byte core rows of a core matrix as one 32-byte STG.256.
  //         The rows arrive through a shared-memory ring filled by the producer warp with cp.async.bulk (one copy per
  //         pair: the two rows are contiguous), so HBM latency is hidden by 16-32 pairs in flight per SM instead of
  //         by the 18 resident warps (direct loads left the phase a load -> reduce -> store latency chain: 58-77k
  //         cycles of which 15k were arithmetic).
  {
    const float4* lng4 = reinterpret_cast<const float4*>(lng);
    const float4* lnb4 = reinterpret_cast<const float4*>(lnb);
    const int npairs = G.Np >> 1;
    if (ring) {
      if (warp == kWWarps && lane == 0) {
        if (lg_bulk) {
          mbar_expect_tx(lgbar, lg_bytes);
          bulk_g2s(smem, prm.logits + (long long)b * P * K, lg_bytes, lgbar);
        }
        const int live_pairs = (P + 1) >> 1;
        for (int m = 0; m < live_pairs; ++m) {
          const int slot = m & (ring_slots - 1);
          if (m >= ring_slots) umma::mbar_wait(&pempty[slot], (uint32_t)(((m >> ring_log) - 1) & 1));
          const uint32_t bytes = (uint32_t)((2 * m + 1 < P ? 2 : 1) * row_bytes);
          mbar_expect_tx(&pfull[slot], bytes);
          bulk_g2s(ring_base + (size_t)slot * 2 * row_bytes, xb + (long long)(2 * m) * C, bytes, &pfull[slot]);
        }
      } else if (worker) {
        for (int m = warp; m < npairs; m += kWWarps) {
          const int p = 2 * m, slot = m & (ring_slots - 1);
          int4 pa[4], pb[4];          // one row live at a time: 18 warps cap the kernel at 96 registers per thread
          if (p < P) {
            umma::mbar_wait(&pfull[slot], (uint32_t)((m >> ring_log) & 1));
            const T* rows = reinterpret_cast<const T*>(ring_base + (size_t)slot * 2 * row_bytes);
            float v[4][8];
            tile_row_load_smem<T>(rows, true, C, lane, v);
            tile_row_normalise<MODE, true>(v, true, C, lane, lng4, lnb4, prm.ln_eps);
#pragma unroll
            for (int j = 0; j < 4; ++j) pa[j] = pack8(v[j]);
            tile_row_load_smem<T>(rows + C, p + 1 < P, C, lane, v);
            // Generic-proxy reads followed by an async-proxy overwrite (the refill of this slot) need a cross-proxy
            // fence: without it the release below overtook loads still queued in the load/store unit and the next
            // occupant's rows showed up in ~0.3 % of images (a few slightly-wrong tokens).
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pempty[slot]);
            tile_row_normalise<MODE, true>(v, p + 1 < P, C, lane, lng4, lnb4, prm.ln_eps);
#pragma unroll
            for (int j = 0; j < 4; ++j) pb[j] = pack8(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { pa[j] = make_int4(0, 0, 0, 0); pb[j] = make_int4(0, 0, 0, 0); }
          }
          unsigned char* dst = xh + (size_t)(p >> 3) * G.tile_row_bytes + (size_t)(p & 7) * 16;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int core = lane + j * 32;
            if (core < G.Cc16) st_global32(dst + (size_t)core * 128, pa[j], pb[j]);
          }
        }
      }
    } else {
      for (int p = 2 * warp; p < G.Np; p += 2 * (kThreads / 32)) {       // rows not 16-byte copyable: direct loads
        int4 pa[4], pb[4];
        float v[4][8];
        tile_row_load<T>(xb, p, P, C, lane, xvec, v);
        tile_row_normalise<MODE, false>(v, p < P, C, lane, lng4, lnb4, prm.ln_eps);
#pragma unroll
        for (int j = 0; j < 4; ++j) pa[j] = pack8(v[j]);
        tile_row_load<T>(xb, p + 1, P, C, lane, xvec, v);
        tile_row_normalise<MODE, false>(v, p + 1 < P, C, lane, lng4, lnb4, prm.ln_eps);
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = pack8(v[j]);
        unsigned char* dst = xh + (size_t)(p >> 3) * G.tile_row_bytes + (size_t)(p & 7) * 16;    // p even: 32-byte aligned
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int core = lane + j * 32;
          if (core < G.Cc16) st_global32(dst + (size_t)core * 128, pa[j], pb[j]);
        }
      }
    }
  }
  __threadfence();          // the tiles are read back through the async proxy (bulk copies) by this CTA
  asm volatile("fence.proxy.async;" ::: "memory");
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  STAMP(1);

  // ---- 1. Z = Q . Xn^T : producer warp + MMA warp, workers wait on acc1
  if (MODE != MODE_SIT) {
    if (warp == kWWarps) {
      // the whole producer warp issues: one lane per 1 KB token-group segment (a single thread issuing the 27 copies
      // of a stage one after the other was the critical path of this GEMM)
      for (int c = 0; c < G.nchunk1; ++c) {
        const int st = c % G.S1;
        if (c >= G.S1) umma::mbar_wait(&empty1[st], (uint32_t)(((c / G.S1) - 1) & 1));
        unsigned char* A = R0 + (size_t)st * (G.stage1A + G.stage1B);
        unsigned char* Bt = A + G.stage1A;
        if (lane == 0) {
          mbar_expect_tx(&full1[st], (uint32_t)(G.q_chunk_bytes + G.stage1B));
          bulk_g2s(A, prm.q_packed + (size_t)c * G.q_chunk_bytes, (uint32_t)G.q_chunk_bytes, &full1[st]);
        }
        __syncwarp();
        for (int pg = lane; pg < G.NG; pg += 32)
          bulk_g2s(Bt + (size_t)pg * 1024, xh + (size_t)pg * G.tile_row_bytes + (size_t)c * 1024, 1024u, &full1[st]);
      }
    } else if (warp == kWWarps + 1 && lane == 0) {
      const uint32_t idesc1 = umma::instr_desc(umma::FMT_BF16, 128, (uint32_t)G.Np);
      for (int c = 0; c < G.nchunk1; ++c) {
        const int st = c % G.S1;
        umma::mbar_wait(&full1[st], (uint32_t)((c / G.S1) & 1));
        umma::tc_fence_after_sync();
        const uint32_t a0 = umma::smem_u32(R0 + (size_t)st * (G.stage1A + G.stage1B)), b0 = a0 + (uint32_t)G.stage1A;
        for (int mt = 0; mt < G.n_mt; ++mt)
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = umma::smem_desc_kmajor(a0 + mt * 16 * 1024 + ks * 256, 128, 1024);
            const uint64_t db = umma::smem_desc_kmajor(b0 + ks * 256, 128, 1024);
            umma::mma_bf16(tmem_base + mt * 256, da, db, idesc1, (c > 0 || ks > 0) ? 1u : 0u);
          }
        umma::mma_commit(&empty1[st]);
      }
      umma::mma_commit(acc1);
    }
    umma::mbar_wait(acc1, 0);
    umma::tc_fence_after_sync();
  }
  STAMP(2);

  // ---- 2. accumulator -> bf16 scores in shared memory
  __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(smem + G.z_off);
  const int PSb = G.PSb;
  if (MODE != MODE_SIT && worker) {
    const int q = warp & 3, nslots = kWWarps >> 2;
    const int parts = nslots / G.n_mt > 0 ? nslots / G.n_mt : 1;
    for (int sl = warp >> 2; sl < G.n_mt * parts; sl += nslots) {
      const int mt = sl / parts, part = sl % parts;
      const int cbeg = ((G.Np / 16) * part / parts) * 16, cend = ((G.Np / 16) * (part + 1) / parts) * 16;
      const int k = mt * 128 + q * 32 + lane;
      for (int c0 = cbeg; c0 < cend; c0 += 16) {
        uint32_t v[16];
        umma::tmem_ld16(umma::tmem_addr(tmem_base, (uint32_t)(q * 32), (uint32_t)(mt * 256 + c0)), v);
        umma::tmem_ld_wait();
        if (k < K) {
#pragma unroll
          for (int j = 0; j < 16; j += 2)
            if (c0 + j < P) {
              const float z0 = bf16_round(bf16_round(__uint_as_float(v[j])) * prm.scale);
              const float z1 = bf16_round(bf16_round(__uint_as_float(v[j + 1])) * prm.scale);
              *reinterpret_cast<__nv_bfloat162*>(Z + (size_t)k * PSb + c0 + j) = __floats2bfloat162_rn(z0, z1);
            }
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  STAMP(3);

  // ---- 3. W from Z: global fp32 + bf16 A operand (K-major over p) in region 0
  unsigned char* Wop = R0;
  float* wout = prm.weights + (long long)b * K * P;
  if (MODE == MODE_SIT) {
    // logits [P, K] -> staged TRANSPOSED as bf16 rows [K][PSb] so that the softmax over tokens is the same
    // register-row pass as PatchMerger's (one warp per slot k, one shared-memory sweep)
    __nv_bfloat16* Lt = reinterpret_cast<__nv_bfloat16*>(smem + G.z_off);
    if (lg_bulk) {
      // the raw [P, K] block was bulk-copied to the bottom of shared memory while the tiles were produced:
      // transpose it shared -> shared (no global latency left in this phase)
      umma::mbar_wait(lgbar, 0);
      const __nv_bfloat16* raw = reinterpret_cast<const __nv_bfloat16*>(smem);
      if ((K & 1) == 0) {
        const int K2 = K >> 1;
        for (int e = tid; e < P * K2; e += kThreads) {
          const int p = e / K2, k = (e - p * K2) * 2;
          const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(raw + (size_t)p * K + k);
          Lt[(size_t)k * PSb + p] = v2.x;
          Lt[(size_t)(k + 1) * PSb + p] = v2.y;
        }
      } else {
        for (int e = tid; e < P * K; e += kThreads) { const int p = e / K, k = e - p * K; Lt[(size_t)k * PSb + p] = raw[e]; }
      }
      if (P & 1)
        for (int k = tid; k < K; k += kThreads) Lt[(size_t)k * PSb + P] = __float2bfloat16_rn(0.f);
      __syncthreads();          // the W operand below overwrites the raw block
    } else {
      stage_logits_transposed(Lt, prm.logits + (long long)b * P * K, P, K, PSb, tid, kThreads);
    }
    for (int e = tid; e < (int)(32 * G.sbo2 / 16); e += kThreads) reinterpret_cast<int4*>(Wop)[e] = make_int4(0, 0, 0, 0);
    const float sc = prm.scale_ptr[0];
    __syncthreads();
    const RowMap rm = make_row_map(P, lane);
    for (int k = warp; k < K; k += kThreads / 32) {
      float x[8];
      load_scores(Lt + (size_t)k * PSb, rm, x);
      float m = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 8; ++i) { x[i] = x[i] * sc + rm.dead[i]; m = fmaxf(m, x[i]); }   // the scale may have either sign
      m = warp_max(m);
      const float nm2 = -m * 1.4426950408889634f;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { x[i] = ex2_ftz(fmaf(x[i], 1.4426950408889634f, nm2)); sum += x[i]; }   // logits are bf16
      const float inv = 1.0f / warp_sum(sum);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] *= inv;
      store_weights(x, wout + (long long)k * P, Wop + (size_t)(k >> 3) * G.sbo2 + (size_t)(k & 7) * 16, rm);
    }
  } else if (MODE == MODE_SINKHORN) {
    // Log-domain Sinkhorn on the bf16 scores.  Row passes keep a row in registers (one shared-memory sweep: max,
    // then exp-sum), column passes split K over several threads per column PAIR with a chunked online log-sum-exp
    // (one sweep, one rescale per 8 rows) and combine the partials through shared memory (the W-operand area, which
    // is zero-filled only afterwards).  13 two-byte sweeps with a separate max pass cost 90k cycles before.
    const float nrm = prm.log_norm;
    constexpr int NW = kThreads / 32;
    constexpr float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const int npp = (P + 1) >> 1;                       // column pairs
    int nsplit = kThreads / npp;
    nsplit = nsplit > 8 ? 8 : nsplit;
    float* part_m = reinterpret_cast<float*>(Wop);       // [nsplit][2*npp]
    float* part_s = part_m + 8 * 256;
    for (int k = tid; k < K; k += kThreads) uvec[k] = 0.f;
    for (int p = tid; p < P; p += kThreads) vvec[p] = 0.f;
    // per-lane constants of the row passes (the phase is instruction-issue bound: 60k warp-instructions per
    // iteration before bounds tests, potential reloads and the 6-instruction __expf were hoisted / replaced by
    // ex2.approx.ftz(fma(x, log2e, -m*log2e)))
    int poff[4];
    bool ok0[4], ok1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pp = 2 * (lane + 32 * i);
      ok0[i] = pp < P; ok1[i] = pp + 1 < P;
      poff[i] = ok0[i] ? pp : 0;                          // dead slots re-read token 0 (finite) and add -inf
    }
    __syncthreads();
    for (int it = 0; it < prm.iters; ++it) {
      float vv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        vv[2 * i] = ok0[i] ? vvec[poff[i]] : -CUDART_INF_F;
        vv[2 * i + 1] = ok1[i] ? vvec[poff[i] + 1] : -CUDART_INF_F;
      }
      for (int k0 = warp; k0 < K; k0 += 2 * NW) {
        const bool two = k0 + NW < K;
        const __nv_bfloat16* za = Z + (size_t)k0 * PSb;
        const __nv_bfloat16* zb = Z + (size_t)(two ? k0 + NW : k0) * PSb;
        float xa[8], xb2[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(za + poff[i]));
          const float2 b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(zb + poff[i]));
          xa[2 * i] = a2.x + vv[2 * i]; xa[2 * i + 1] = a2.y + vv[2 * i + 1];
          xb2[2 * i] = b2.x + vv[2 * i]; xb2[2 * i + 1] = b2.y + vv[2 * i + 1];
        }
        float ma = xa[0], mb = xb2[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) { ma = fmaxf(ma, xa[i]); mb = fmaxf(mb, xb2[i]); }
        warp_max2(ma, mb);
        const float na = -ma * L2E, nb = -mb * L2E;
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { sa += ex2_ftz(fmaf(xa[i], L2E, na)); sb += ex2_ftz(fmaf(xb2[i], L2E, nb)); }
        warp_sum2(sa, sb);
        if (lane == 0) {
          uvec[k0] = nrm - (lg2_fast(sa) * LN2 + ma);
          if (two) uvec[k0 + NW] = nrm - (lg2_fast(sb) * LN2 + mb);
        }
      }
      __syncthreads();
      {
        const int cp = tid % npp, sp = tid / npp;
        if (sp < nsplit) {
          const int kb = K * sp / nsplit, ke = K * (sp + 1) / nsplit;
          const __nv_bfloat16* zc = Z + (size_t)kb * PSb + 2 * cp;
          float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, s0 = 0.f, s1 = 0.f;
          for (int k = kb; k < ke; k += 8, zc += (size_t)8 * PSb) {
            float x0[8], x1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (k + i < ke) {
                const float2 z = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(zc + (size_t)i * PSb));
                const float u = uvec[k + i];
                x0[i] = z.x + u; x1[i] = z.y + u;
              } else { x0[i] = -CUDART_INF_F; x1[i] = -CUDART_INF_F; }
            }
            float n0 = m0, n1 = m1;
#pragma unroll
            for (int i = 0; i < 8; ++i) { n0 = fmaxf(n0, x0[i]); n1 = fmaxf(n1, x1[i]); }
            const float c0 = -n0 * L2E, c1 = -n1 * L2E;
            s0 *= ex2_ftz(fmaf(m0, L2E, c0)); s1 *= ex2_ftz(fmaf(m1, L2E, c1));
#pragma unroll
            for (int i = 0; i < 8; ++i) { s0 += ex2_ftz(fmaf(x0[i], L2E, c0)); s1 += ex2_ftz(fmaf(x1[i], L2E, c1)); }
            m0 = n0; m1 = n1;
          }
          part_m[sp * 256 + 2 * cp] = m0; part_m[sp * 256 + 2 * cp + 1] = m1;
          part_s[sp * 256 + 2 * cp] = s0; part_s[sp * 256 + 2 * cp + 1] = s1;
        }
      }
      __syncthreads();
      for (int p = tid; p < P; p += kThreads) {
        float m = part_m[p];
        for (int sp = 1; sp < nsplit; ++sp) m = fmaxf(m, part_m[sp * 256 + p]);
        float sum = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) sum += part_s[sp * 256 + p] * ex2_ftz((part_m[sp * 256 + p] - m) * L2E);
        vvec[p] = nrm - (lg2_fast(sum) * LN2 + m);
      }
      __syncthreads();
    }
    STAMP(6);
    for (int e = tid; e < (int)(32 * G.sbo2 / 16); e += kThreads) reinterpret_cast<int4*>(Wop)[e] = make_int4(0, 0, 0, 0);
    __syncthreads();
    STAMP(7);
    {
      const RowMap rm = make_row_map(P, lane);
      float vv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        vv[2 * i] = rm.live0[i] ? vvec[rm.poff[i]] : -CUDART_INF_F;
        vv[2 * i + 1] = rm.live1[i] ? vvec[rm.poff[i] + 1] : -CUDART_INF_F;
      }
      for (int k = warp; k < K; k += NW) {
        float x[8];
        load_scores(Z + (size_t)k * PSb, rm, x);
        const float uk = uvec[k];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = ex2_ftz((((x[i] + uk) + vv[i]) - nrm) * L2E);   // rel. error ~1e-6: bf16 scores
        store_weights(x, wout + (long long)k * P, Wop + (size_t)(k >> 3) * G.sbo2 + (size_t)(k & 7) * 16, rm);
      }
    }
  } else {
    for (int e = tid; e < (int)(32 * G.sbo2 / 16); e += kThreads) reinterpret_cast<int4*>(Wop)[e] = make_int4(0, 0, 0, 0);
    __syncthreads();
    STAMP(7);
    const RowMap rm = make_row_map(P, lane);
    for (int k = warp; k < K; k += kThreads / 32) {
      float x[8];
      load_scores(Z + (size_t)k * PSb, rm, x);
      float m = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 8; ++i) { x[i] += rm.dead[i]; m = fmaxf(m, x[i]); }
      m = warp_max(m);
      const float nm2 = -m * 1.4426950408889634f;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { x[i] = ex2_ftz(fmaf(x[i], 1.4426950408889634f, nm2)); sum += x[i]; }   // scores are bf16
      // one reciprocal per row: IEEE division takes its slow path for every zero numerator (the slots past P), which
      // made this loop 11k cycles longer; the product differs from the quotient by at most 1 ulp
      const float inv = 1.0f / warp_sum(sum);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] *= inv;
      store_weights(x, wout + (long long)k * P, Wop + (size_t)(k >> 3) * G.sbo2 + (size_t)(k & 7) * 16, rm);
    }
  }
  umma::fence_proxy_async_smem();      // W operand (generic-proxy writes) -> visible to the tensor core
  __syncthreads();
  STAMP(4);

  // ---- 4. out = W . Xn : B operand = tile segments used MN-major (N = channel, K = token)
  unsigned char* XT = smem + G.xt_off;
  __nv_bfloat16* ob = prm.out + (long long)b * K * C;
  if (warp == kWWarps) {
    for (int cc = 0; cc < G.nchunk2; ++cc) {
      const int st = cc % S2;
      if (cc >= S2) umma::mbar_wait(&empty2[st], (uint32_t)(((cc / S2) - 1) & 1));
      unsigned char* dstb = XT + (size_t)st * G.NG * 2048;
      if (lane == 0) mbar_expect_tx(&full2[st], (uint32_t)(G.NG * 2048));
      __syncwarp();
      for (int pg = lane; pg < G.NG; pg += 32)
        bulk_g2s(dstb + (size_t)pg * 2048, xh + (size_t)pg * G.tile_row_bytes + (size_t)cc * 2048, 2048u, &full2[st]);
    }
  } else if (warp == kWWarps + 1 && lane == 0) {
    const uint32_t idesc2 = umma::instr_desc(umma::FMT_BF16, 128, NC2) | (1u << 16);     // B operand MN-major
    const uint32_t a0 = umma::smem_u32(Wop);
    for (int cc = 0; cc < G.nchunk2; ++cc) {
      const int st = cc % S2, as = cc & 1;
      umma::mbar_wait(&full2[st], (uint32_t)((cc / S2) & 1));
      if (cc >= 2) umma::mbar_wait(&accempty[as], (uint32_t)(((cc >> 1) - 1) & 1));
      umma::tc_fence_after_sync();
      const uint32_t b0 = umma::smem_u32(XT + (size_t)st * G.NG * 2048);
      const uint32_t acc = tmem_base + (uint32_t)(as * 256);
      for (int mt = 0; mt < G.n_mt; ++mt)
        for (int ks = 0; ks < G.Np / 16; ++ks) {
          const uint64_t da = umma::smem_desc_kmajor(a0 + mt * 16 * G.sbo2 + ks * 256, 128, G.sbo2);
          // MN-major, no swizzle: LBO = stride between groups of 8 along K (token groups, 2 KB), SBO = stride between
          // groups of 8 along N (channel cores, 128 B)
          const uint64_t db = umma::smem_desc_kmajor(b0 + ks * 4096, 2048, 128);
          umma::mma_bf16(acc + mt * 128, da, db, idesc2, ks > 0 ? 1u : 0u);
        }
      umma::mma_commit(&empty2[st]);
      umma::mma_commit(&accfull[as]);
    }
  } else if (worker) {
    const int q = warp & 3, nslots = kWWarps >> 2;
    const int parts = nslots / G.n_mt > 0 ? nslots / G.n_mt : 1;
    for (int cc = 0; cc < G.nchunk2; ++cc) {
      const int as = cc & 1;
      umma::mbar_wait(&accfull[as], (uint32_t)((cc >> 1) & 1));
      umma::tc_fence_after_sync();
      for (int sl = warp >> 2; sl < G.n_mt * parts; sl += nslots) {
        const int mt = sl / parts, part = sl % parts;
        const int jbeg = ((NC2 / 16) * part / parts) * 16, jend = ((NC2 / 16) * (part + 1) / parts) * 16;
        const int k = mt * 128 + q * 32 + lane;
        const uint32_t acc = tmem_base + (uint32_t)(as * 256 + mt * 128);
        for (int j0 = jbeg; j0 < jend; j0 += 16) {
          uint32_t v[16];
          umma::tmem_ld16(umma::tmem_addr(acc, (uint32_t)(q * 32), (uint32_t)j0), v);
          umma::tmem_ld_wait();
          const int c = cc * NC2 + j0;
          if (k < K && c < C) {
            float f[8], h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { f[i] = __uint_as_float(v[i]); h[i] = __uint_as_float(v[8 + i]); }
            __nv_bfloat16* dst = ob + (long long)k * C + c;
            if (c + 16 <= C && ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0)) {
              st_global32(dst, pack8(f), pack8(h));          // one STG.256: every lane writes its own output row, so a
                                                             // store instruction costs 32 L1 wavefronts whatever its width
            } else if (c + 16 <= C && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
              *reinterpret_cast<int4*>(dst) = pack8(f);
              *reinterpret_cast<int4*>(dst + 8) = pack8(h);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (c + i < C) dst[i] = __float2bfloat16_rn(f[i]);
                if (c + 8 + i < C) dst[8 + i] = __float2bfloat16_rn(h[i]);
              }
            }
          }
        }
      }
      umma::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accempty[as]);
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  STAMP(5);
#undef STAMP
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

}  // namespace

size_t soft_merge_tc2_workspace_bytes(int B, int P, int C, int K) {
  const Geo g = make_geo(P, C, K);
  return ((g.q_bytes + 255) & ~(size_t)255) + (size_t)B * g.img_bytes;
}

// returns TOKRED_OK after launching, or 1 if this kernel does not cover the shape / workspace (caller falls back)
int launch_soft_merge_tc2(int mode, const void* x, int x_dtype, const float* q, const float* ln_w, const float* ln_b,
                          int B, int P, int C, int K, float scale, float log_norm, float ln_eps, int iters, void* out,
                          float* weights, void* stream, const char* what, const void* logits, const float* scale_ptr,
                          void* workspace, size_t workspace_bytes, long long xbs) {
  if (P > kMaxP || K > kMaxK || P < 8 || K < 1 || C > 1024 || C % 8 != 0) return 1;
  if (!workspace || workspace_bytes < soft_merge_tc2_workspace_bytes(B, P, C, K)) return 1;
  if (reinterpret_cast<uintptr_t>(workspace) & 127u) return 1;
  const Geo g = make_geo(P, C, K);
  if (g.total > 227 * 1024) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* qp = reinterpret_cast<unsigned char*>(workspace);
  unsigned char* xh = qp + ((g.q_bytes + 255) & ~(size_t)255);
  if (mode != MODE_SIT) {
    pack_q_kernel<<<g.nchunk1, 256, 0, st>>>(q, K, C, g.K8, qp);
    if (int e = finish_launch(what)) return e;
  }
  Tc2Params prm{};
  prm.x = x; prm.xbs = xbs; prm.q_packed = qp; prm.xh = xh; prm.ln_w = ln_w; prm.ln_b = ln_b; prm.logits = (const __nv_bfloat16*)logits;
  prm.scale_ptr = scale_ptr; prm.scale = scale; prm.log_norm = log_norm; prm.ln_eps = ln_eps; prm.iters = iters;
  prm.P = P; prm.C = C; prm.K = K; prm.out = (__nv_bfloat16*)out; prm.weights = weights;
  prm.dbg = g_phase_dbg;
#define LAUNCH(T, MODE)                                                                   \
  do {                                                                                    \
    if (int e = allow_smem(soft_merge_tc2_kernel<T, MODE>, g.total, what)) return e;      \
    soft_merge_tc2_kernel<T, MODE><<<B, kThreads, g.total, st>>>(prm);                    \
  } while (0)
  if (mode == MODE_SIT) { if (x_dtype == TOKRED_F32) LAUNCH(float, MODE_SIT); else LAUNCH(__nv_bfloat16, MODE_SIT); }
  else if (mode == MODE_SINKHORN) { if (x_dtype == TOKRED_F32) LAUNCH(float, MODE_SINKHORN); else LAUNCH(__nv_bfloat16, MODE_SINKHORN); }
  else { if (x_dtype == TOKRED_F32) LAUNCH(float, MODE_PATCHMERGER); else LAUNCH(__nv_bfloat16, MODE_PATCHMERGER); }
#undef LAUNCH
  return finish_launch(what);
}

}  // namespace tokred
