// f1 / f2: the producer of the reduction operators' inputs -- multi-head self-attention for the
// token counts of a 224x224 DeiT (N <= 256, head dim 64) that never materialises the [B,H,N,N] probabilities.
//
// Replaces, under bf16 autocast, the ATen sequence of models/topk.py:44-52,59-61 (also evit.py:66-87, tome.py:44-58,
// kmedoids.py:105-112, ats.py:115-127, dyvit.py:53-69 eval branch and the stock timm block):
//     attn = (q @ k^T) * scale            bf16 matmul, bf16 result, bf16 product
//     attn = attn + log(size)             (ToMe proportional attention, tome.py:48-49; fp32 from here on)
//     attn = attn.softmax(-1)             fp32 (autocast runs softmax in fp32)
//     x    = (attn @ v)                   probabilities rounded to bf16, bf16 matmul
// and emits, on request, the only parts of `attn` the reduction operators read: the CLS row of every head
// (Top-K / EViT / ATS scores) -- the [B,H,N,N] tensor is never written.
//
// One CTA per (image, head), 128 threads, two CTAs per SM (<= 79 KB shared memory, <= 256 TMEM columns each), so one
// CTA's loads overlap the other's softmax without any explicit pipeline:
//   load   q, k, v head slices (128-byte rows, row stride 3C) -> shared memory with 16-byte cp.async straight into the
//          canonical no-swizzle UMMA core-matrix layout (8 consecutive lanes = 8 consecutive rows = one contiguous 128-byte
//          core matrix: conflict-free); pad rows are zero-filled by the copy itself (src-size 0)
//   S      = Q_tile K^T      tcgen05.mma kind::f16 (bf16), M = 128, N = ceil16(N), K = 64, accumulator in TMEM cols [0, Np)
//   softmax thread = accumulator row (tcgen05.ld 32x32b): pass A max, pass B exp2 + sum (e written back IN PLACE as
//          fp32), pass C normalise -> bf16 pairs written back IN PLACE to TMEM cols [0, Np/2) (writes trail the reads)
//   O      = P V             tcgen05.mma with the A operand read FROM TMEM (P never touches shared memory), B = the v tile
//          used MN-major (no transpose), accumulator in TMEM cols [o_col, o_col + 64) (dead part of the S region)
//   store  thread = output row: 64 fp32 -> bf16 -> 128 contiguous bytes of out[b, row, h*64 ..]
// Rounding points are the reference's: S to bf16, (S*scale) to bf16, P to bf16, O to bf16; everything else fp32.
#include <cmath>

#include "common.cuh"
#include "umma.cuh"

namespace tokred {
namespace {

constexpr int kThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  const __nv_bfloat16* qkv;     // [B, N, 3, H, 64]
  const float* key_bias;        // [B, N] or null
  __nv_bfloat16* out;           // [B, N, H*64]
  float* cls_row;               // [B, H, N] or null
  int B, N, H;
  float scale;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem] ; A: lane = row, one 32-bit column = two consecutive K elements (bf16)
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two fp32 -> packed bf16x2 (lo = a, hi = b), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// logits of 16 accumulator columns starting at col0, with the reference's roundings: bf16(acc) * scale (rounded again
// unless scale is a power of two, where the product is exact), + bias in fp32
template <bool BIAS, bool ROUND2>
__device__ __forceinline__ void logits16(const uint32_t (&r)[16], float (&x)[16], float scale, const float* bias, int col0) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const uint32_t pk = pack_bf16x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
    float a = __uint_as_float(pk << 16) * scale, b = __uint_as_float(pk & 0xffff0000u) * scale;
    if (ROUND2) {
      const uint32_t p2 = pack_bf16x2(a, b);
      a = __uint_as_float(p2 << 16);
      b = __uint_as_float(p2 & 0xffff0000u);
    }
    x[i] = a;
    x[i + 1] = b;
  }
  if (BIAS) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(bias + col0 + i);
      x[i] += bv.x; x[i + 1] += bv.y; x[i + 2] += bv.z; x[i + 3] += bv.w;
    }
  }
}

template <bool BIAS, bool ROUND2>
__global__ void __launch_bounds__(kThreads) attention_kernel(const AttnParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H, C = H * 64;
  const int Np = (N + 15) & ~15;                 // keys: UMMA N of S, UMMA K of P.V
  const int gq = (N + 7) >> 3, gk = Np >> 3;     // 8-row groups of the Q / K,V tiles
  const int b = blockIdx.x / H, h = blockIdx.x % H;

  unsigned char* Qs = smem;
  unsigned char* Ks = Qs + (size_t)gq * 1024;
  unsigned char* Vs = Ks + (size_t)gk * 1024;
  // the S MMA always reads 128 A rows: the operand area spans at least ntiles*16 row groups past Qs
  const int ggrp = max(gq + 2 * gk, ((N + 127) >> 7) * 16);
  float* bias_s = reinterpret_cast<float*>(smem + (size_t)ggrp * 1024);
  uint64_t* bar = reinterpret_cast<uint64_t*>(bias_s + Np);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const uint32_t ncols = Np <= 128 ? 128u : 256u;
  const uint32_t o_col = Np <= 128 ? 64u : 128u;
  if (warp == 0) umma::tmem_alloc(tmem_slot, ncols);
  if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_mbar_init(); }

  // ---- loads: job = (matrix, 8-row group); lane -> row = lane % 8, 16-byte chunks 2*(lane/8), 2*(lane/8)+1
  {
    const __nv_bfloat16* base = p.qkv + (size_t)b * N * 3 * C + (size_t)h * 64;
    const int r8 = lane & 7, cp = lane >> 3;
    const int njobs = gq + 2 * gk;
    for (int j = warp; j < njobs; j += kThreads / 32) {
      int m, g;
      if (j < gq) { m = 0; g = j; } else if (j < gq + gk) { m = 1; g = j - gq; } else { m = 2; g = j - gq - gk; }
      const int row = g * 8 + r8;
      const bool ok = row < N;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(base + ((size_t)(ok ? row : 0) * 3 + m) * C) + cp * 32;
      unsigned char* tile = (m == 0 ? Qs : m == 1 ? Ks : Vs) + (size_t)g * 1024 + r8 * 16 + cp * 256;
      const uint32_t dst = umma::smem_u32(tile);
      cp_async16(dst, src, ok ? 16u : 0u);
      cp_async16(dst + 128, src + 16, ok ? 16u : 0u);
    }
    if (BIAS)
      for (int j = tid; j < Np; j += kThreads) bias_s[j] = j < N ? p.key_bias[(size_t)b * N + j] : 0.f;
    cp_async_wait_all();
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const uint32_t idesc_s = umma::instr_desc(umma::FMT_BF16, 128, (uint32_t)Np);
  const uint32_t idesc_o = umma::instr_desc(umma::FMT_BF16, 128, 64) | (1u << 16);      // B operand (v) MN-major
  const float scale = p.scale;
  uint32_t phase = 0;
  const int ntiles = (N + 127) >> 7;
  const int nch = Np >> 4;

  for (int t = 0; t < ntiles; ++t) {
    // ---- S = Q_t K^T (rows past the Q tile read the K tile behind it: finite garbage in accumulator rows nobody reads)
    if (tid == 0) {
      const uint32_t a0 = umma::smem_u32(Qs) + (uint32_t)t * 16u * 1024u, b0 = umma::smem_u32(Ks);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma::mma_bf16(tmem, umma::smem_desc_kmajor(a0 + ks * 256, 128, 1024), umma::smem_desc_kmajor(b0 + ks * 256, 128, 1024),
                       idesc_s, ks > 0 ? 1u : 0u);
      umma::mma_commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::tc_fence_after_sync();

    const int row = t * 128 + warp * 32 + lane;
    const bool active = t * 128 + warp * 32 < N;           // warp-uniform
    const uint32_t trow = umma::tmem_addr(tmem, (uint32_t)(warp * 32), 0);
    if (active) {
      uint32_t r[16];
      float x[16];
      // pass A: row maximum
      float mx = -INFINITY;
      for (int c = 0; c < nch; ++c) {
        umma::tmem_ld16(trow + c * 16, r);
        umma::tmem_ld_wait();
        logits16<BIAS, ROUND2>(r, x, scale, bias_s, c * 16);
        if (c * 16 + 16 <= N) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, x[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (c * 16 + i < N) mx = fmaxf(mx, x[i]);
        }
      }
      // pass B: e = exp(x - max) (kept in place as fp32), row sum
      const float mneg = -mx * kLog2e;
      float sum = 0.f;
      for (int c = 0; c < nch; ++c) {
        umma::tmem_ld16(trow + c * 16, r);
        umma::tmem_ld_wait();
        logits16<BIAS, ROUND2>(r, x, scale, bias_s, c * 16);
        const bool full = c * 16 + 16 <= N;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float e = ex2(fmaf(x[i], kLog2e, mneg));
          if (!full && c * 16 + i >= N) e = 0.f;
          sum += e;
          r[i] = __float_as_uint(e);
        }
        tmem_st16(trow + c * 16, r);
      }
      tmem_st_wait();
      // pass C: p = e / sum -> bf16 pairs in place (column j holds keys 2j, 2j+1); CLS row to global in fp32
      const float inv = 1.0f / sum;
      float* cls = (p.cls_row && row == 0) ? p.cls_row + ((size_t)b * H + h) * N : nullptr;
      for (int c = 0; c < nch; ++c) {
        umma::tmem_ld16(trow + c * 16, r);
        umma::tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(r[i]) * inv;
        if (cls) {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (c * 16 + i < N) cls[c * 16 + i] = x[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
        tmem_st8(trow + c * 8, pk);
      }
      tmem_st_wait();
    }
    umma::tc_fence_before_sync();
    __syncthreads();

    // ---- O = P V : A from TMEM, B = v tile MN-major (LBO field = stride between 8-token groups, SBO field = 8-channel cores)
    if (tid == 0) {
      umma::tc_fence_after_sync();
      const uint32_t v0 = umma::smem_u32(Vs);
      for (int ks = 0; ks < nch; ++ks)
        mma_bf16_ts(tmem + o_col, tmem + ks * 8, umma::smem_desc_kmajor(v0 + ks * 2048, 1024, 128), idesc_o, ks > 0 ? 1u : 0u);
      umma::mma_commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::tc_fence_after_sync();
    if (active) {                                          // whole warp: the TMEM loads are .sync.aligned
      __nv_bfloat16* dst = p.out + ((size_t)b * N + (row < N ? row : 0)) * C + h * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[16];
        umma::tmem_ld16(trow + o_col + c * 16, r);
        umma::tmem_ld_wait();
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        if (row < N)
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + c * 16), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                       "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                       : "memory");
      }
    }
    umma::tc_fence_before_sync();
    __syncthreads();
  }
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

size_t attn_smem_bytes(int N) {
  const int Np = (N + 15) & ~15;
  const int groups = (N + 7) / 8 + 2 * (Np / 8), mma_rows = ((N + 127) / 128) * 16;
  return (size_t)(groups > mma_rows ? groups : mma_rows) * 1024 + (size_t)Np * 4 + 64;
}

}  // namespace
}  // namespace tokred

using namespace tokred;

extern "C" int tokred_attention(const void* qkv, int B, int N, int H, int head_dim, float scale, const float* key_bias,
                                void* out, float* cls_row, void* stream) {
  TOKRED_REQUIRE(qkv && out, "attention: null pointer");
  TOKRED_REQUIRE(B >= 1 && H >= 1 && N >= 1, "attention: B=%d N=%d H=%d", B, N, H);
  if (head_dim != 64 || N > 256) {
    set_error("attention: head_dim=%d N=%d outside the fused kernel's range (head_dim 64, N <= 256)", head_dim, N);
    return TOKRED_ERR_UNSUPPORTED;
  }
  TOKRED_REQUIRE((long long)B * H <= 0x7fffffffLL, "attention: B*H too large");
  TOKRED_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0,
                 "attention: qkv must be 16-byte and out 32-byte aligned");
  AttnParams prm{};
  prm.qkv = (const __nv_bfloat16*)qkv; prm.key_bias = key_bias; prm.out = (__nv_bfloat16*)out; prm.cls_row = cls_row;
  prm.B = B; prm.N = N; prm.H = H; prm.scale = scale;
  int ex = 0;
  const bool pow2 = std::frexp(scale, &ex) == 0.5f;
  const size_t smem = attn_smem_bytes(N);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(BIAS, R2)                                                                          \
  do {                                                                                            \
    if (int e = allow_smem(attention_kernel<BIAS, R2>, smem, "attention")) return e;             \
    attention_kernel<BIAS, R2><<<B * H, kThreads, smem, st>>>(prm);                               \
  } while (0)
  if (key_bias) { if (pow2) LAUNCH(true, false); else LAUNCH(true, true); }
  else          { if (pow2) LAUNCH(false, false); else LAUNCH(false, true); }
#undef LAUNCH
  return finish_launch("attention");
}
