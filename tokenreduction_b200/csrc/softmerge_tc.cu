// Tensor-core (tcgen05 + TMEM) Sinkhorn / PatchMerger for the bf16-autocast path.
// Reference: models/sinkhorn.py:66-86, models/patchmerger.py:35-39 under torch.autocast(bf16): both contractions
// ARE bf16 tensor-core matmuls with fp32 accumulate and bf16-rounded results, everything between them is fp32.
//
// One persistent CTA per image, 256 threads, all 512 TMEM columns:
//   0. token statistics (L2 norm / LayerNorm mean,rstd), warp per row, 16-byte loads            (compulsory x read)
//   1. Z = Q . Xn^T   — operands converted to bf16 and written by the threads straight into the canonical K-major
//      UMMA layout, 64 columns of C per stage, two stages: staging of chunk c+1 overlaps the MMAs of chunk c
//      (tcgen05.commit -> mbarrier hands the stage back).  M = 2 x 128 rows of Q, N = P padded to 16.
//   2. accumulator -> registers (tcgen05.ld, thread = row) -> bf16 (the reference's scores ARE bf16) -> smem Z
//   3. Sinkhorn iterations (row pass: warp per row; column pass: thread per column) / softmax over tokens, fp32;
//      W is written once to global (fp32, the op's second output) and once as bf16 into the A-operand layout
//   4. out = W . Xn   — B operand = Xn^T staged per 128 columns of C by 8x8 register transposes; two TMEM
//      accumulator sets so the epilogue of chunk c overlaps the MMAs of chunk c+1; bf16 rows stored with 16-byte writes.
// Shared memory: [stages | W operand] 116 KB, [Z bf16 | Xn^T chunk] <= 70 KB, vectors 4 KB.
#include <math_constants.h>

#include "common.cuh"
#include "umma.cuh"

namespace tokred {
long long* g_phase_dbg = nullptr;   // set through tokred_debug_phase_buffer()
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int KC1 = 64;           // contraction columns per stage of the first GEMM
constexpr int NC2 = 128;          // output columns per accumulator set of the second GEMM
constexpr int kMaxK = 208, kMaxP = 208;

enum { MODE_SINKHORN = 0, MODE_PATCHMERGER = 1, MODE_SIT = 2 };

struct TcParams {
  const void* x;            // [B,P,C], images xbs elements apart
  long long xbs;
  const float* q;           // [K,C]
  const float* ln_w;
  const float* ln_b;
  const __nv_bfloat16* logits;   // sit: [B,P,K] bf16
  const float* scale_ptr;        // sit: device scalar
  float scale;              // patchmerger: sim * scale ; sinkhorn: 1/eps
  float log_norm;
  float ln_eps;
  int iters;
  int P, C, K;
  __nv_bfloat16* out;       // [B,K,C]
  float* weights;           // [B,K,P]
  long long* dbg;           // optional: clock64() stamps of CTA 0 at the phase boundaries (tools/phase_times.py)
};

struct Layout {
  int Np, Pp, PSb, Ks, n_mt;
  uint32_t sbo1, sbo2;
  size_t stageA, stageB, r0, r1, total;
};

__host__ __device__ inline Layout make_layout(int P, int K, int C) {
  Layout L;
  L.Np = (P + 15) & ~15;
  L.Pp = L.Np;
  L.n_mt = (K + 127) / 128;
  L.sbo1 = (KC1 / 8) * 128;                       // 1024
  L.sbo2 = (uint32_t)(L.Pp / 8) * 128 + 16;       // +16: 8-row groups land on different banks for the 16-B transposed stores
  L.stageA = (size_t)32 * L.sbo1;                 // 256 rows
  L.stageB = (size_t)(L.Np / 8) * L.sbo1;
  const size_t wop = (size_t)32 * L.sbo2;
  const size_t stages = 2 * (L.stageA + L.stageB);
  L.r0 = ((stages > wop ? stages : wop) + 127) & ~(size_t)127;
  int psb = (P + 1) & ~1;
  if (((psb / 2) & 1) == 0) psb += 2;             // odd number of 32-bit words per row -> conflict-free row-per-thread stores
  L.PSb = psb;
  int ks = (K + 1) & ~1;
  if (((ks / 2) & 1) == 0) ks += 2;               // sit: staged logits [P][Ks], odd word count per row
  L.Ks = ks;
  const size_t zb0 = (size_t)K * psb * 2, zb1 = (size_t)P * ks * 2;
  const size_t zbytes = zb0 > zb1 ? zb0 : zb1;
  const size_t xt = (size_t)(NC2 / 8) * L.sbo2;
  L.r1 = ((zbytes > xt ? zbytes : xt) + 127) & ~(size_t)127;
  L.total = L.r0 + L.r1 + (size_t)(3 * P + K + 2 * ((C + 3) & ~3)) * 4 + 64;
  return L;
}

template <typename T> __device__ __forceinline__ void load8(const T* p, bool vec, int valid, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, bool vec, int valid, float (&v)[8]) {
  if (vec && valid >= 8) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < valid ? p[i] : 0.f;
  }
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, bool vec, int valid, float (&v)[8]) {
  if (vec && valid >= 8) {
    const int4 raw = *reinterpret_cast<const int4*>(p);
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

template <int MODE>
struct Xform {
  const float* s0;
  const float* s1;
  const float* g;
  const float* bta;
  __device__ __forceinline__ float operator()(int p, int c, float v) const {
    if (MODE == MODE_SINKHORN) return v / s0[p];
    return g[c] * (s1[p] * (v - s0[p])) + bta[c];
  }
  // 8 consecutive channels c0..c0+7 of token p (c0 % 4 == 0); channels >= C become 0
  __device__ __forceinline__ void apply8(int p, int c0, int C, float (&v)[8]) const {
    if (MODE == MODE_SIT) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (c0 + i < C) ? v[i] : 0.f;
    } else if (MODE == MODE_SINKHORN) {
      // x * (1/||x||): the quotient is rounded to bf16 right after, so the reciprocal form (1 ulp in fp32) is
      // indistinguishable here and avoids 8 IEEE divisions per chunk (15 % of the kernel in the r01 profile)
      const float inv = s1[p];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (c0 + i < C) ? v[i] * inv : 0.f;
    } else {
      const float mean = s0[p], rstd = s1[p];
      if (c0 + 8 <= C) {
        const float4 g0 = *reinterpret_cast<const float4*>(g + c0), g1 = *reinterpret_cast<const float4*>(g + c0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bta + c0), b1 = *reinterpret_cast<const float4*>(bta + c0 + 4);
        v[0] = g0.x * (rstd * (v[0] - mean)) + b0.x; v[1] = g0.y * (rstd * (v[1] - mean)) + b0.y;
        v[2] = g0.z * (rstd * (v[2] - mean)) + b0.z; v[3] = g0.w * (rstd * (v[3] - mean)) + b0.w;
        v[4] = g1.x * (rstd * (v[4] - mean)) + b1.x; v[5] = g1.y * (rstd * (v[5] - mean)) + b1.y;
        v[6] = g1.z * (rstd * (v[6] - mean)) + b1.z; v[7] = g1.w * (rstd * (v[7] - mean)) + b1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (c0 + i < C) ? g[c0 + i] * (rstd * (v[i] - mean)) + bta[c0 + i] : 0.f;
      }
    }
  }
};

template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, 1) soft_merge_tc_kernel(TcParams prm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int P = prm.P, C = prm.C, K = prm.K;
  const Layout L = make_layout(P, K, C);
  unsigned char* R0 = smem;
  unsigned char* R1 = smem + L.r0;
  const int Cpad = (C + 3) & ~3;
  float* lng = reinterpret_cast<float*>(R1 + L.r1);     // [Cpad] LayerNorm weight (patchmerger), 16-byte aligned
  float* lnb = lng + Cpad;                               // [Cpad] LayerNorm bias
  float* s0 = lnb + Cpad;                                // [P]
  float* s1 = s0 + P;                                    // [P]
  float* uvec = s1 + P;                                  // [K]
  float* vvec = uvec + K;                                // [P]
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(vvec + P) + 7) & ~(uintptr_t)7);   // stage / accumulator barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* xb = reinterpret_cast<const T*>(prm.x) + (long long)b * prm.xbs;
  const bool xvec = (C % 8 == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15u) == 0);
  const bool qvec = (C % 8 == 0) && ((reinterpret_cast<uintptr_t>(prm.q) & 15u) == 0);

#define STAMP(i) do { if (prm.dbg && blockIdx.x == 0 && tid == 0) prm.dbg[i] = clock64(); } while (0)
  STAMP(0);
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  if (MODE == MODE_PATCHMERGER)
    for (int c = tid; c < C; c += kThreads) { lng[c] = prm.ln_w[c]; lnb[c] = prm.ln_b[c]; }
  if (tid == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::fence_mbar_init();
  }

  // ---- 0. token statistics: one warp per token, the whole row (<= 1024 channels) held in registers so that x is
  //         read from HBM once and every load of the row is in flight together
  if (MODE == MODE_SIT) {
    // no token statistics: SiT merges the raw tokens
  } else if (C <= 1024) {
    for (int p = warp; p < P; p += kWarps) {
      const T* row = xb + (long long)p * C;
      float v[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = lane * 8 + j * 256;
        if (c < C) load8<T>(row + c, xvec, C - c, v[j]);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
        }
      }
      if (MODE == MODE_SINKHORN) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) s = fmaf(v[j][i], v[j][i], s);
        s = warp_sum(s);
        if (lane == 0) { s0[p] = fmaxf(sqrtf(s), 1e-12f); s1[p] = 1.0f / s0[p]; }
      } else {
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
        const float var = warp_sum(q) / (float)C;
        if (lane == 0) { s0[p] = mean; s1[p] = rsqrtf(var + prm.ln_eps); }
      }
    }
  } else {
    for (int p = warp; p < P; p += kWarps) {
      const T* row = xb + (long long)p * C;
      float s = 0.f, q = 0.f;
      for (int c = lane; c < C; c += 32) { const float v = to_f32(row[c]); s += v; q = fmaf(v, v, q); }
      if (MODE == MODE_SINKHORN) {
        q = warp_sum(q);
        if (lane == 0) { s0[p] = fmaxf(sqrtf(q), 1e-12f); s1[p] = 1.0f / s0[p]; }
      } else {
        const float mean = warp_sum(s) / (float)C;
        float d2 = 0.f;
        for (int c = lane; c < C; c += 32) { const float d = to_f32(row[c]) - mean; d2 = fmaf(d, d, d2); }
        const float var = warp_sum(d2) / (float)C;
        if (lane == 0) { s0[p] = mean; s1[p] = rsqrtf(var + prm.ln_eps); }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const Xform<MODE> xf{s0, s1, lng, lnb};
  STAMP(1);   // stats done

  // ---- 1. Z = Q . Xn^T on tensor cores
  const int nchunk = MODE == MODE_SIT ? 0 : (C + KC1 - 1) / KC1;
  const uint32_t idesc1 = umma::instr_desc(umma::FMT_BF16, 128, (uint32_t)L.Np);
  // Software pipeline: the global loads of chunk c+1 are issued right after the barrier of chunk c and stay in
  // registers while the MMAs of chunk c run and the stage is handed back; convert + store happen one iteration later.
  constexpr int GMAX = (kMaxK / 8 * 64 + kThreads - 1) / kThreads;      // groups of 8 elements per thread and operand
  const int nga = ((K + 7) / 8) * 64, ngb = L.Np * 8;
  float va[GMAX][8], vb[GMAX][8];
  auto load_chunk = [&](int c) {
    const int k0 = c * KC1;
#pragma unroll
    for (int u = 0; u < GMAX; ++u) {
      const int gI = tid + u * kThreads;
      const int row = (gI & 7) + ((gI >> 6) << 3), k = k0 + ((gI >> 3) & 7) * 8;
      if (gI < nga && row < K) load8<float>(prm.q + (long long)row * C + k, qvec, C - k, va[u]);
      if (gI < ngb && row < P) load8<T>(xb + (long long)row * C + k, xvec, C - k, vb[u]);
    }
  };
  if (nchunk > 0) load_chunk(0);
  for (int c = 0; c < nchunk; ++c) {
    const int st = c & 1;
    unsigned char* A = R0 + (size_t)st * (L.stageA + L.stageB);
    unsigned char* Bt = A + L.stageA;
    if (c >= 2) umma::mbar_wait(&bars[st], (uint32_t)(((c - 2) >> 1) & 1));   // MMAs of chunk c-2 released this stage
    const int k0 = c * KC1;
    // 8 consecutive lanes = 8 consecutive rows of one 16-byte K-chunk: 128 contiguous bytes of shared memory
#pragma unroll
    for (int u = 0; u < GMAX; ++u) {
      const int gI = tid + u * kThreads;
      const int row = (gI & 7) + ((gI >> 6) << 3), ch = (gI >> 3) & 7, k = k0 + ch * 8;
      if (gI < nga && row < K)
        *reinterpret_cast<int4*>(A + (row >> 3) * L.sbo1 + (row & 7) * 16 + ch * 128) = pack8(va[u]);
      if (gI < ngb) {
        if (row < P) {
          xf.apply8(row, k, C, vb[u]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) vb[u][i] = 0.f;
        }
        *reinterpret_cast<int4*>(Bt + (row >> 3) * L.sbo1 + (row & 7) * 16 + ch * 128) = pack8(vb[u]);
      }
    }
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after_sync();
      const uint32_t a0 = umma::smem_u32(A), b0 = umma::smem_u32(Bt);
      for (int mt = 0; mt < L.n_mt; ++mt)
        for (int ks = 0; ks < KC1 / 16; ++ks) {
          const uint64_t da = umma::smem_desc_kmajor(a0 + mt * 16 * L.sbo1 + ks * 256, 128, L.sbo1);
          const uint64_t db = umma::smem_desc_kmajor(b0 + ks * 256, 128, L.sbo1);
          umma::mma_bf16(tmem_base + mt * 256, da, db, idesc1, (c > 0 || ks > 0) ? 1u : 0u);
        }
      umma::mma_commit(&bars[st]);
    }
    if (c + 1 < nchunk) load_chunk(c + 1);
  }
  // all MMAs done <=> the last commit completed (tcgen05 ops of one thread complete in order)
  if (nchunk > 0) {
    const int last = nchunk - 1;
    umma::mbar_wait(&bars[last & 1], (uint32_t)((last >> 1) & 1));
    if (nchunk >= 2) umma::mbar_wait(&bars[(last - 1) & 1], (uint32_t)(((last - 1) >> 1) & 1));
  }
  umma::tc_fence_after_sync();
  STAMP(2);   // GEMM 1 done

  // ---- 2. accumulator -> bf16 scores in shared memory (row per thread; padded row stride = odd word count)
  __nv_bfloat16* Z = reinterpret_cast<__nv_bfloat16*>(R1);
  const int PSb = L.PSb;
  if (MODE != MODE_SIT) {
    // warp w owns TMEM lane quarter w % 4; (w / 4) enumerates (M tile, column part) pairs
    const int q = warp & 3, nslots = kWarps >> 2;
    const int parts = nslots / L.n_mt > 0 ? nslots / L.n_mt : 1;
    for (int sl = warp >> 2; sl < L.n_mt * parts; sl += nslots) {
      const int mt = sl / parts, part = sl % parts;
      const int cbeg = ((L.Np / 16) * part / parts) * 16, cend = ((L.Np / 16) * (part + 1) / parts) * 16;
      const int k = mt * 128 + q * 32 + lane;
      for (int c0 = cbeg; c0 < cend; c0 += 16) {
        uint32_t v[16];
        umma::tmem_ld16(umma::tmem_addr(tmem_base, (uint32_t)(q * 32), (uint32_t)(mt * 256 + c0)), v);
        umma::tmem_ld_wait();
        if (k < K) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            if (c0 + j < P) {      // P even-padded row: the pad element is never read
              const float z0 = bf16_round(bf16_round(__uint_as_float(v[j])) * prm.scale);
              const float z1 = bf16_round(bf16_round(__uint_as_float(v[j + 1])) * prm.scale);
              *reinterpret_cast<__nv_bfloat162*>(Z + (size_t)k * PSb + c0 + j) = __floats2bfloat162_rn(z0, z1);
            }
          }
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  STAMP(3);   // Z in smem

  // ---- 3. W from Z (fp32 math), written to global (fp32) and as bf16 into the A-operand layout of GEMM 2
  unsigned char* Wop = R0;
  for (int e = tid; e < (int)(32 * L.sbo2 / 16); e += kThreads) reinterpret_cast<int4*>(Wop)[e] = make_int4(0, 0, 0, 0);
  float* wout = prm.weights + (long long)b * K * P;
  if (MODE == MODE_SIT) {
    // logits [P][K] (bf16, K contiguous) -> smem [P][Ks]; softmax over TOKENS per cluster column (models/sit.py:38)
    __nv_bfloat16* Lg = reinterpret_cast<__nv_bfloat16*>(R1);
    const int Ks = L.Ks;
    const __nv_bfloat16* lb = prm.logits + (long long)b * P * K;
    for (int e = tid; e < P * K; e += kThreads) Lg[(e / K) * Ks + e % K] = lb[e];
    const float sc = prm.scale_ptr[0];
    __syncthreads();
    for (int k = tid; k < K; k += kThreads) {       // per-cluster max and normaliser (lanes on consecutive columns)
      float m = -CUDART_INF_F;
      for (int p = 0; p < P; ++p) m = fmaxf(m, __bfloat162float(Lg[p * Ks + k]) * sc);
      float sum = 0.f;
      for (int p = 0; p < P; ++p) sum += expf(__bfloat162float(Lg[p * Ks + k]) * sc - m);
      uvec[k] = m;
      lng[k] = sum;                                   // lng/lnb are unused by SiT: lng[0..K) holds the normalisers
    }
    __syncthreads();
    for (int k = warp; k < K; k += kWarps) {
      const float m = uvec[k], sum = lng[k];
      for (int p = lane; p < P; p += 32) {
        const float w = expf(__bfloat162float(Lg[p * Ks + k]) * sc - m) / sum;
        wout[(long long)k * P + p] = w;
        *reinterpret_cast<__nv_bfloat16*>(Wop + umma::kmajor_offset((uint32_t)k, (uint32_t)p, 2, L.sbo2)) = __float2bfloat16_rn(w);
      }
    }
  } else if (MODE == MODE_SINKHORN) {
    const float nrm = prm.log_norm;
    for (int k = tid; k < K; k += kThreads) uvec[k] = 0.f;
    for (int p = tid; p < P; p += kThreads) vvec[p] = 0.f;
    __syncthreads();
    for (int it = 0; it < prm.iters; ++it) {
      for (int k = warp; k < K; k += kWarps) {
        float m = -CUDART_INF_F;
        for (int p = lane; p < P; p += 32) m = fmaxf(m, __bfloat162float(Z[(size_t)k * PSb + p]) + vvec[p]);
        m = warp_max(m);
        float s = 0.f;
        for (int p = lane; p < P; p += 32) s += __expf(__bfloat162float(Z[(size_t)k * PSb + p]) + vvec[p] - m);
        s = warp_sum(s);
        if (lane == 0) uvec[k] = nrm - (__logf(s) + m);
      }
      __syncthreads();
      for (int p = tid; p < P; p += kThreads) {
        float m = -CUDART_INF_F;
#pragma unroll 8
        for (int k = 0; k < K; ++k) m = fmaxf(m, __bfloat162float(Z[(size_t)k * PSb + p]) + uvec[k]);
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < K; ++k) s += __expf(__bfloat162float(Z[(size_t)k * PSb + p]) + uvec[k] - m);
        vvec[p] = nrm - (__logf(s) + m);
      }
      __syncthreads();
    }
    for (int k = warp; k < K; k += kWarps) {
      const float uk = uvec[k];
      for (int p = lane; p < P; p += 32) {
        const float w = expf(((__bfloat162float(Z[(size_t)k * PSb + p]) + uk) + vvec[p]) - nrm);
        wout[(long long)k * P + p] = w;
        *reinterpret_cast<__nv_bfloat16*>(Wop + umma::kmajor_offset((uint32_t)k, (uint32_t)p, 2, L.sbo2)) = __float2bfloat16_rn(w);
      }
    }
  } else {
    __syncthreads();
    for (int k = warp; k < K; k += kWarps) {
      float m = -CUDART_INF_F;
      for (int p = lane; p < P; p += 32) m = fmaxf(m, __bfloat162float(Z[(size_t)k * PSb + p]));
      m = warp_max(m);
      float s = 0.f;
      for (int p = lane; p < P; p += 32) s += expf(__bfloat162float(Z[(size_t)k * PSb + p]) - m);
      s = warp_sum(s);
      for (int p = lane; p < P; p += 32) {
        const float w = expf(__bfloat162float(Z[(size_t)k * PSb + p]) - m) / s;
        wout[(long long)k * P + p] = w;
        *reinterpret_cast<__nv_bfloat16*>(Wop + umma::kmajor_offset((uint32_t)k, (uint32_t)p, 2, L.sbo2)) = __float2bfloat16_rn(w);
      }
    }
  }
  __syncthreads();     // Z is dead from here on: its region becomes the Xn^T chunk buffer
  STAMP(4);   // W built

  // ---- 4. out = W . Xn on tensor cores, 128 output columns per accumulator set
  unsigned char* XT = R1;
  const int nc2 = (C + NC2 - 1) / NC2;
  const uint32_t idesc2 = umma::instr_desc(umma::FMT_BF16, 128, NC2);
  const int npg = L.Pp / 8;
  __nv_bfloat16* ob = prm.out + (long long)b * K * C;
  // barrier uses so far: bars[0] ceil(nchunk/2) phases, bars[1] floor(nchunk/2) phases
  uint32_t ph0 = (uint32_t)((nchunk + 1) / 2), ph1 = (uint32_t)(nchunk / 2);
  for (int cc = 0; cc <= nc2; ++cc) {
    if (cc < nc2) {
      // the MMAs of chunk cc-1 read XT: wait for them before overwriting it
      if (cc >= 1) {
        const int pst = (cc - 1) & 1;
        if (pst == 0) { umma::mbar_wait(&bars[0], ph0 & 1); ++ph0; } else { umma::mbar_wait(&bars[1], ph1 & 1); ++ph1; }
        umma::tc_fence_after_sync();
      }
      const int c0 = cc * NC2;
      // 8 tokens x 8 channels per thread-step, transposed in registers: each store is one full 16-byte K-chunk
      for (int blk = tid; blk < (NC2 / 8) * npg; blk += kThreads) {
        const int cg = blk % (NC2 / 8), pg = blk / (NC2 / 8);
        float v[8][8];
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const int p = pg * 8 + pp, c = c0 + cg * 8;
          if (p < P && c < C) {
            load8<T>(xb + (long long)p * C + c, xvec, C - c, v[pp]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[pp][i] = 0.f;
          }
        }
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
          if (pg * 8 + pp < P && c0 + cg * 8 < C) xf.apply8(pg * 8 + pp, c0 + cg * 8, C, v[pp]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float col[8];
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) col[pp] = v[pp][i];
          *reinterpret_cast<int4*>(XT + cg * L.sbo2 + i * 16 + pg * 128) = pack8(col);
        }
      }
      umma::fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        umma::tc_fence_after_sync();
        const uint32_t a0 = umma::smem_u32(Wop), b0 = umma::smem_u32(XT);
        const uint32_t acc = tmem_base + (uint32_t)((cc & 1) * 256);
        for (int mt = 0; mt < L.n_mt; ++mt)
          for (int ks = 0; ks < L.Pp / 16; ++ks) {
            const uint64_t da = umma::smem_desc_kmajor(a0 + mt * 16 * L.sbo2 + ks * 256, 128, L.sbo2);
            const uint64_t db = umma::smem_desc_kmajor(b0 + ks * 256, 128, L.sbo2);
            umma::mma_bf16(acc + mt * 128, da, db, idesc2, ks > 0 ? 1u : 0u);
          }
        umma::mma_commit(&bars[cc & 1]);
      }
    }
    // epilogue of chunk cc-1 (its completion was awaited above, or here for the last chunk)
    if (cc >= 1) {
      const int pc = cc - 1;
      if (cc == nc2) {
        if ((pc & 1) == 0) { umma::mbar_wait(&bars[0], ph0 & 1); ++ph0; } else { umma::mbar_wait(&bars[1], ph1 & 1); ++ph1; }
        umma::tc_fence_after_sync();
      }
      const int q = warp & 3, nslots = kWarps >> 2;
      const int parts = nslots / L.n_mt > 0 ? nslots / L.n_mt : 1;
      for (int sl = warp >> 2; sl < L.n_mt * parts; sl += nslots) {
        const int mt = sl / parts, part = sl % parts;
        const int jbeg = ((NC2 / 16) * part / parts) * 16, jend = ((NC2 / 16) * (part + 1) / parts) * 16;
        const int k = mt * 128 + q * 32 + lane;
        const uint32_t acc = tmem_base + (uint32_t)((pc & 1) * 256 + mt * 128);
        for (int j0 = jbeg; j0 < jend; j0 += 16) {
          uint32_t v[16];
          umma::tmem_ld16(umma::tmem_addr(acc, (uint32_t)(q * 32), (uint32_t)j0), v);
          umma::tmem_ld_wait();
          const int c = pc * NC2 + j0;
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
    }
  }
  __syncthreads();
  STAMP(5);   // GEMM 2 + epilogue done
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
#undef STAMP
}

}  // namespace

// returns TOKRED_OK after launching, or 1 if the shape is outside what this kernel covers (caller falls back to FFMA)
int launch_soft_merge_tc(int mode, const void* x, int x_dtype, const float* q, const float* ln_w, const float* ln_b,
                         int B, int P, int C, int K, float scale, float log_norm, float ln_eps, int iters, void* out,
                         float* weights, void* stream, const char* what, const void* logits, const float* scale_ptr,
                         long long xbs) {
  if (P > kMaxP || K > kMaxK || P < 8 || K < 1) return 1;
  const Layout L = make_layout(P, K, C);
  if (L.total > 227 * 1024) return 1;
  TcParams prm{};
  prm.x = x; prm.xbs = xbs; prm.q = q; prm.ln_w = ln_w; prm.ln_b = ln_b; prm.scale = scale; prm.log_norm = log_norm; prm.ln_eps = ln_eps;
  prm.iters = iters; prm.P = P; prm.C = C; prm.K = K; prm.out = (__nv_bfloat16*)out; prm.weights = weights;
  prm.logits = (const __nv_bfloat16*)logits; prm.scale_ptr = scale_ptr;
  prm.dbg = g_phase_dbg;
  if (mode == MODE_SIT && K > ((C + 3) & ~3)) return 1;      // the normalisers borrow the (unused) LayerNorm slot [C]
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(T, MODE)                                                                  \
  do {                                                                                   \
    if (int e = allow_smem(soft_merge_tc_kernel<T, MODE>, L.total, what)) return e;      \
    soft_merge_tc_kernel<T, MODE><<<B, kThreads, L.total, st>>>(prm);                    \
  } while (0)
  if (mode == MODE_SIT) { if (x_dtype == TOKRED_F32) LAUNCH(float, MODE_SIT); else LAUNCH(__nv_bfloat16, MODE_SIT); }
  else if (mode == MODE_SINKHORN) { if (x_dtype == TOKRED_F32) LAUNCH(float, MODE_SINKHORN); else LAUNCH(__nv_bfloat16, MODE_SINKHORN); }
  else { if (x_dtype == TOKRED_F32) LAUNCH(float, MODE_PATCHMERGER); else LAUNCH(__nv_bfloat16, MODE_PATCHMERGER); }
#undef LAUNCH
  return finish_launch(what);
}

}  // namespace tokred

// debug hook (not part of the product ABI): device buffer of >= 8 int64 receiving clock64() phase stamps of CTA 0
extern "C" __attribute__((visibility("default"))) void tokred_debug_phase_buffer(void* p) { tokred::g_phase_dbg = (long long*)p; }
