// Tensor-core Sinkhorn / PatchMerger / SiT, version 2: bulk-copy fed, warp-specialised.
//
// v1 (softmerge_tc.cu) converted every operand element to bf16 on the CUDA cores inside both GEMM loops; phase
// stamps showed it issue-bound there (GEMM 1 staging 41 %, GEMM 2 staging + store 30 % of the kernel, tensor pipe
// 7 % active).  v2 converts each token ONCE:
//   pack kernel   Q [K,C] fp32 -> bf16 in the canonical K-major stage image, one contiguous block per 64-column chunk.
//   phase 0       token statistics with the row in registers; the normalised row is rounded to bf16 and written to a
//                 per-image scratch of CORE-MATRIX TILES  [p/8][c/8][8 rows x 16 B]  (L2-resident, 2 B/element).
//                 The same 128-byte tile is a K-major core matrix for GEMM 1 (rows = tokens, 16 B along C) and an
//                 MN-major core matrix for GEMM 2 (rows = the contraction index p, 16 B along the N index c).
//   phase 1       warp 16 streams stages with cp.async.bulk (1 KB per token group + one block of Q per chunk) onto
//                 mbarriers (expect_tx); warp 17 issues tcgen05.mma and hands stages back with tcgen05.commit.
//   phase 2/3     as v1: accumulator -> bf16 Z in smem; Sinkhorn iterations / softmax; W -> global fp32 + bf16 A operand.
//   phase 4       same producer / MMA warps: B operand = 2 KB tile segments used MN-major; 16 worker warps drain two
//                 TMEM accumulator sets (tcgen05.ld) and store bf16 rows while the next chunk's MMAs run.
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
constexpr int S1 = 2;      // GEMM 1 stages
constexpr int S2 = 2;      // GEMM 2 B-operand stages
constexpr int NC2 = 128;   // output columns per accumulator set

enum { MODE_SINKHORN = 0, MODE_PATCHMERGER = 1, MODE_SIT = 2 };

struct Geo {
  int Np, NG, K8, n_mt, Cc8, Cc16, nchunk1, nchunk2, PSb, Ks;
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
  g.stage1A = (size_t)g.n_mt * 16 * 1024;
  g.stage1B = (size_t)g.NG * 1024;
  const size_t stages1 = S1 * (g.stage1A + g.stage1B);
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
  //   [0, stages1)            GEMM-1 stages                      | later [0, wop) the W operand of GEMM 2
  //   [z_off, z_off + zb)     Z (bf16) / staged SiT logits       (z_off = stages1: live together with the W operand)
  //   [xt_off, xt_off + xt)   GEMM-2 B stages (xt_off = wop)     (overlaps Z, which is dead once W is built)
  const size_t wop_al = (wop + 127) & ~(size_t)127, st_al = (stages1 + 127) & ~(size_t)127;
  g.z_off = st_al > wop_al ? st_al : wop_al;
  g.xt_off = wop_al;
  const size_t end0 = g.z_off + zb, end1 = g.xt_off + xt;
  g.vec_off = ((end0 > end1 ? end0 : end1) + 127) & ~(size_t)127;
  g.total = g.vec_off + (size_t)(3 * P + 2 * K + 2 * ((C + 3) & ~3)) * 4 + 256;
  return g;
}

struct Tc2Params {
  const void* x;                  // [B,P,C]
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
    // plain (L1-allocating) loads: a lane reads 32 contiguous bytes as two 16-byte halves of the SAME sector; with
    // L1::no_allocate the second half re-fetches the sector from L2 (phase 0 ran 4x slower that way)
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
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

// Normalise (Sinkhorn: unit L2 norm; PatchMerger: LayerNorm; SiT: identity) one token held by a warp and store its
// bf16 core rows into the K-major tile image.  Rows p >= P arrive as zeros and are stored as zeros.
template <int MODE>
__device__ __forceinline__ void tile_row_finish(float (&v)[4][8], int p, int P, int C, int lane, const float* lng,
                                                const float* lnb, float ln_eps, unsigned char* xh,
                                                size_t tile_row_bytes, int Cc16) {
  if (p < P) {
    if (MODE == MODE_SINKHORN) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(v[j][i], v[j][i], s);
      s = warp_sum(s);
      const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);      // rounded to bf16 right after: reciprocal form is fine
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j][i] *= inv;
    } else if (MODE == MODE_PATCHMERGER) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[j][i];
      const float mean = warp_sum(s) / (float)C;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (lane * 8 + j * 256 + i < C) { const float d = v[j][i] - mean; q = fmaf(d, d, q); }
      const float rstd = rsqrtf(warp_sum(q) / (float)C + ln_eps);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = lane * 8 + j * 256;
        if (c + 8 <= C) {
          const float4 g0 = *reinterpret_cast<const float4*>(lng + c), g1 = *reinterpret_cast<const float4*>(lng + c + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(lnb + c), b1 = *reinterpret_cast<const float4*>(lnb + c + 4);
          v[j][0] = g0.x * (rstd * (v[j][0] - mean)) + b0.x; v[j][1] = g0.y * (rstd * (v[j][1] - mean)) + b0.y;
          v[j][2] = g0.z * (rstd * (v[j][2] - mean)) + b0.z; v[j][3] = g0.w * (rstd * (v[j][3] - mean)) + b0.w;
          v[j][4] = g1.x * (rstd * (v[j][4] - mean)) + b1.x; v[j][5] = g1.y * (rstd * (v[j][5] - mean)) + b1.y;
          v[j][6] = g1.z * (rstd * (v[j][6] - mean)) + b1.z; v[j][7] = g1.w * (rstd * (v[j][7] - mean)) + b1.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            v[j][i] = (c + i < C) ? lng[c + i] * (rstd * (v[j][i] - mean)) + lnb[c + i] : 0.f;
        }
      }
    }
  }
  unsigned char* dst = xh + (size_t)(p >> 3) * tile_row_bytes + (size_t)(p & 7) * 16;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int core = lane + j * 32;
    if (core < Cc16) *reinterpret_cast<int4*>(dst + (size_t)core * 128) = pack8(v[j]);
  }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, 1) soft_merge_tc2_kernel(Tc2Params prm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int P = prm.P, C = prm.C, K = prm.K;
  const Geo G = make_geo(P, C, K);
  unsigned char* R0 = smem;
  const int Cpad = (C + 3) & ~3;
  float* lng = reinterpret_cast<float*>(smem + G.vec_off);     // [Cpad]
  float* lnb = lng + Cpad;                               // [Cpad]
  float* s0 = lnb + Cpad;                                // [P]
  float* s1 = s0 + P;                                    // [P]
  float* uvec = s1 + P;                                  // [K]
  float* vvec = uvec + K;                                // [P]
  float* aux = vvec + P;                                 // [K]   sit: softmax normalisers
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(aux + K) + 7) & ~(uintptr_t)7);
  uint64_t* full1 = bars;            // [S1]
  uint64_t* empty1 = full1 + S1;     // [S1]
  uint64_t* acc1 = empty1 + S1;      // [1]  GEMM 1 finished
  uint64_t* full2 = acc1 + 1;        // [S2]
  uint64_t* empty2 = full2 + S2;     // [S2]
  uint64_t* accfull = empty2 + S2;   // [2]
  uint64_t* accempty = accfull + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = warp < kWWarps;
  const T* xb = reinterpret_cast<const T*>(prm.x) + (long long)b * P * C;
  unsigned char* xh = prm.xh + (size_t)b * G.img_bytes;
  const bool xvec = (C % 8 == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15u) == 0);

#define STAMP(i) do { if (prm.dbg && blockIdx.x == 0 && tid == 0) prm.dbg[i] = clock64(); } while (0)
  STAMP(0);
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < S1; ++i) { umma::mbar_init(&full1[i], 1); umma::mbar_init(&empty1[i], 1); }
    umma::mbar_init(acc1, 1);
    for (int i = 0; i < S2; ++i) { umma::mbar_init(&full2[i], 1); umma::mbar_init(&empty2[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&accfull[i], 1); umma::mbar_init(&accempty[i], kWWarps); }
    umma::fence_mbar_init();
  }
  if (MODE == MODE_PATCHMERGER)
    for (int c = tid; c < C; c += kThreads) { lng[c] = prm.ln_w[c]; lnb[c] = prm.ln_b[c]; }
  __syncthreads();

  // ---- 0. token statistics + bf16 core-matrix tiles in ONE pass: one warp per token, row in registers (x read from
  //         HBM once), 16-byte core rows stored straight into the tile.  (Measured alternatives that did not help:
  //         two tokens in flight per warp; assembling tile rows in shared memory and storing them with cp.async.bulk.)
  for (int p = warp; p < G.Np; p += kThreads / 32) {
    float v[4][8];
    tile_row_load<T>(xb, p, P, C, lane, xvec, v);
    tile_row_finish<MODE>(v, p, P, C, lane, lng, lnb, prm.ln_eps, xh, G.tile_row_bytes, G.Cc16);
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
    if (warp == kWWarps && lane == 0) {
      for (int c = 0; c < G.nchunk1; ++c) {
        const int st = c % S1;
        if (c >= S1) umma::mbar_wait(&empty1[st], (uint32_t)(((c / S1) - 1) & 1));
        unsigned char* A = R0 + (size_t)st * (G.stage1A + G.stage1B);
        unsigned char* Bt = A + G.stage1A;
        mbar_expect_tx(&full1[st], (uint32_t)(G.q_chunk_bytes + G.stage1B));
        bulk_g2s(A, prm.q_packed + (size_t)c * G.q_chunk_bytes, (uint32_t)G.q_chunk_bytes, &full1[st]);
        for (int pg = 0; pg < G.NG; ++pg)
          bulk_g2s(Bt + (size_t)pg * 1024, xh + (size_t)pg * G.tile_row_bytes + (size_t)c * 1024, 1024u, &full1[st]);
      }
    } else if (warp == kWWarps + 1 && lane == 0) {
      const uint32_t idesc1 = umma::instr_desc(umma::FMT_BF16, 128, (uint32_t)G.Np);
      for (int c = 0; c < G.nchunk1; ++c) {
        const int st = c % S1;
        umma::mbar_wait(&full1[st], (uint32_t)((c / S1) & 1));
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
  for (int e = tid; e < (int)(32 * G.sbo2 / 16); e += kThreads) reinterpret_cast<int4*>(Wop)[e] = make_int4(0, 0, 0, 0);
  float* wout = prm.weights + (long long)b * K * P;
  if (MODE == MODE_SIT) {
    __nv_bfloat16* Lg = reinterpret_cast<__nv_bfloat16*>(smem + G.z_off);
    const int Ks = G.Ks;
    const __nv_bfloat16* lb = prm.logits + (long long)b * P * K;
    for (int e = tid; e < P * K; e += kThreads) Lg[(e / K) * Ks + e % K] = lb[e];
    const float sc = prm.scale_ptr[0];
    __syncthreads();
    for (int k = tid; k < K; k += kThreads) {
      float m = -CUDART_INF_F;
      for (int p = 0; p < P; ++p) m = fmaxf(m, __bfloat162float(Lg[p * Ks + k]) * sc);
      float sum = 0.f;
      for (int p = 0; p < P; ++p) sum += expf(__bfloat162float(Lg[p * Ks + k]) * sc - m);
      uvec[k] = m;
      aux[k] = sum;
    }
    __syncthreads();
    for (int k = warp; k < K; k += kThreads / 32) {
      const float m = uvec[k], sum = aux[k];
      for (int p = lane; p < P; p += 32) {
        const float w = expf(__bfloat162float(Lg[p * Ks + k]) * sc - m) / sum;
        wout[(long long)k * P + p] = w;
        *reinterpret_cast<__nv_bfloat16*>(Wop + umma::kmajor_offset((uint32_t)k, (uint32_t)p, 2, G.sbo2)) = __float2bfloat16_rn(w);
      }
    }
  } else if (MODE == MODE_SINKHORN) {
    const float nrm = prm.log_norm;
    for (int k = tid; k < K; k += kThreads) uvec[k] = 0.f;
    for (int p = tid; p < P; p += kThreads) vvec[p] = 0.f;
    __syncthreads();
    for (int it = 0; it < prm.iters; ++it) {
      for (int k = warp; k < K; k += kThreads / 32) {
        float m = -CUDART_INF_F;
        for (int p = lane; p < P; p += 32) m = fmaxf(m, __bfloat162float(Z[(size_t)k * PSb + p]) + vvec[p]);
        m = warp_max(m);
        float s = 0.f;
        for (int p = lane; p < P; p += 32) s += __expf(__bfloat162float(Z[(size_t)k * PSb + p]) + vvec[p] - m);
        s = warp_sum(s);
        if (lane == 0) uvec[k] = nrm - (__logf(s) + m);
      }
      __syncthreads();
      // column pass: 2 threads per column (row halves), combined through shared memory
      {
        const int p = tid >> 1, h = tid & 1;
        const int kb = h ? (K + 1) / 2 : 0, ke = h ? K : (K + 1) / 2;
        float m = -CUDART_INF_F, s = 0.f;
        if (p < P) {
#pragma unroll 4
          for (int k = kb; k < ke; ++k) m = fmaxf(m, __bfloat162float(Z[(size_t)k * PSb + p]) + uvec[k]);
        }
        const float mo = __shfl_xor_sync(0xffffffffu, m, 1);
        m = fmaxf(m, mo);
        if (p < P) {
#pragma unroll 4
          for (int k = kb; k < ke; ++k) s += __expf(__bfloat162float(Z[(size_t)k * PSb + p]) + uvec[k] - m);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (p < P && h == 0) vvec[p] = nrm - (__logf(s) + m);
      }
      __syncthreads();
    }
    for (int k = warp; k < K; k += kThreads / 32) {
      const float uk = uvec[k];
      for (int p = lane; p < P; p += 32) {
        const float w = expf(((__bfloat162float(Z[(size_t)k * PSb + p]) + uk) + vvec[p]) - nrm);
        wout[(long long)k * P + p] = w;
        *reinterpret_cast<__nv_bfloat16*>(Wop + umma::kmajor_offset((uint32_t)k, (uint32_t)p, 2, G.sbo2)) = __float2bfloat16_rn(w);
      }
    }
  } else {
    __syncthreads();
    for (int k = warp; k < K; k += kThreads / 32) {
      float m = -CUDART_INF_F;
      for (int p = lane; p < P; p += 32) m = fmaxf(m, __bfloat162float(Z[(size_t)k * PSb + p]));
      m = warp_max(m);
      float s = 0.f;
      for (int p = lane; p < P; p += 32) s += expf(__bfloat162float(Z[(size_t)k * PSb + p]) - m);
      s = warp_sum(s);
      for (int p = lane; p < P; p += 32) {
        const float w = expf(__bfloat162float(Z[(size_t)k * PSb + p]) - m) / s;
        wout[(long long)k * P + p] = w;
        *reinterpret_cast<__nv_bfloat16*>(Wop + umma::kmajor_offset((uint32_t)k, (uint32_t)p, 2, G.sbo2)) = __float2bfloat16_rn(w);
      }
    }
  }
  umma::fence_proxy_async_smem();      // W operand (generic-proxy writes) -> visible to the tensor core
  __syncthreads();
  STAMP(4);

  // ---- 4. out = W . Xn : B operand = tile segments used MN-major (N = channel, K = token)
  unsigned char* XT = smem + G.xt_off;
  __nv_bfloat16* ob = prm.out + (long long)b * K * C;
  if (warp == kWWarps && lane == 0) {
    for (int cc = 0; cc < G.nchunk2; ++cc) {
      const int st = cc % S2;
      if (cc >= S2) umma::mbar_wait(&empty2[st], (uint32_t)(((cc / S2) - 1) & 1));
      unsigned char* dstb = XT + (size_t)st * G.NG * 2048;
      mbar_expect_tx(&full2[st], (uint32_t)(G.NG * 2048));
      for (int pg = 0; pg < G.NG; ++pg)
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
            if (xvec && c + 16 <= C && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
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
                          void* workspace, size_t workspace_bytes) {
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
  prm.x = x; prm.q_packed = qp; prm.xh = xh; prm.ln_w = ln_w; prm.ln_b = ln_b; prm.logits = (const __nv_bfloat16*)logits;
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
