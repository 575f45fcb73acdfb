// f1 / f2: the producer of the reduction operators' inputs -- multi-head self-attention for the
// token counts of a 224x224 DeiT (N <= 256, head dim 64) that never materialises the [B,H,N,N] probabilities.
//
// Replaces, under bf16 autocast, the ATen sequence of models/topk.py:44-52,59-61 (also evit.py:66-87, tome.py:44-58,
// kmedoids.py:105-112, ats.py:115-127, dyvit.py:53-69 eval branch and the stock timm block):
//     attn = (q @ k^T) * scale            bf16 matmul, bf16 result, bf16 product
//     attn = attn + log(size)             (ToMe proportional attention, tome.py:48-49; fp32 from here on)
//     attn = masked_fill(~(m_i m_j))      (ATS, ats.py:118-121)
//     attn = attn.softmax(-1)             fp32 (autocast runs softmax in fp32)
//     attn = attn[:, :, ids, :]           (ATS row gather, ats.py:84-87 -- here the QUERY rows are gathered instead)
//     x    = (attn @ v)                   probabilities rounded to bf16, bf16 matmul
// and emits, on request, the only parts of `attn` the reduction operators read: the CLS row of every head (Top-K /
// EViT / ATS scores) and the per-head column sums (K-Medoids token weights) -- the [B,H,N,N] tensor is never written.
//
// Work item = (image, head); persistent CTAs of 256 threads, two per SM (<= 108 KB shared memory, <= 256 TMEM columns
// each): the two CTAs of an SM drift into different phases, so one's MMA / barrier waits are filled by the other's
// softmax, and each CTA prefetches its next item's q, k, v during the last query tile of the current one:
//   load   q, k, v head slices (128-byte rows, row stride 3C) -> shared memory with 16-byte cp.async straight into the
//          canonical no-swizzle UMMA core-matrix layout (8 consecutive lanes = 8 consecutive rows = one contiguous 128-byte
//          core matrix: conflict-free); pad rows are zero-filled by the copy itself (src-size 0)
//   S      = Q_tile K^T      tcgen05.mma kind::f16 (bf16), M = 128, N = ceil16(N), K = 64, accumulator in TMEM cols [0, Np)
//   softmax thread = accumulator row (tcgen05.ld 32x32b); the two warps that share a TMEM lane quarter split the columns
//          for pass A (max) and pass B (exp2 + sum, e written back IN PLACE as fp32) and exchange the partials through
//          shared memory; pass C (normalise -> bf16 pairs written back IN PLACE to TMEM cols [0, Np/2), writes trailing
//          the reads) belongs to one warp of the pair, which therefore takes the smaller share of the columns
//   O      = P V             tcgen05.mma with the A operand read FROM TMEM (P never touches shared memory), B = the v tile
//          used MN-major (no transpose), accumulator in TMEM cols [o_col, o_col + 64) (dead part of the S region)
//   store  thread = output row: 32 fp32 -> bf16 -> 64 contiguous bytes of out[b, row, h*64 ..]
// Rounding points are the reference's: S to bf16, (S*scale) to bf16, P to bf16, O to bf16; everything else fp32.
#include <cmath>
#include <mutex>

#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

#include "common.cuh"
#include "umma.cuh"

namespace tokred {
namespace {

constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  const __nv_bfloat16* qkv;     // [B, N, 3, H, 64]
  const float* key_bias;        // [B, N] or null
  const uint8_t* mask;          // [B, N] or null
  const int64_t* q_ids;         // [B, ids_stride] or null: query row m reads token q_ids[b, m]
  long long ids_stride;
  __nv_bfloat16* out;           // [B, M, H*64] or null (scores only)
  float* cls_row;               // [B, H, N] or null: probabilities of query row 0
  float* colsum;                // [B, H, N] or null: sum over query rows
  int B, N, H, M;               // M query rows (= N without q_ids)
  float scale;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA: one 4-D box of the qkv tensor (d, q/k/v x head, token, image) -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t r4,
                                         uint32_t r5, uint32_t r6, uint32_t r7) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7)
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
__device__ __forceinline__ void pair_sync(int quarter) {      // the two warps of one TMEM lane quarter
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

// the TMEM load is asynchronous: tying the destination registers to the wait keeps every use behind it
__device__ __forceinline__ void ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// f(r, c) over the 16-column chunks [c0, c1) of one accumulator row, two chunks per round: both loads are issued, then
// both chunks are processed -- the two bodies are independent, so the scheduler sees twice the instruction-level
// parallelism of one chunk (a warp shares its scheduler with only three others: its own ILP is what hides latency)
template <typename F>
__device__ __forceinline__ void for_chunks(uint32_t trow, int c0, int c1, F&& f) {
  uint32_t ra[16], rb[16], rc[16];
  int c = c0;
  for (; c + 2 < c1; c += 3) {
    umma::tmem_ld16(trow + c * 16, ra);
    umma::tmem_ld16(trow + (c + 1) * 16, rb);
    umma::tmem_ld16(trow + (c + 2) * 16, rc);
    ld_wait16(ra);
    ld_wait16(rb);
    ld_wait16(rc);
    f(ra, c);
    f(rb, c + 1);
    f(rc, c + 2);
  }
  if (c + 1 < c1) {
    umma::tmem_ld16(trow + c * 16, ra);
    umma::tmem_ld16(trow + (c + 1) * 16, rb);
    ld_wait16(ra);
    ld_wait16(rb);
    f(ra, c);
    f(rb, c + 1);
    c += 2;
  }
  if (c < c1) {
    umma::tmem_ld16(trow + c * 16, ra);
    ld_wait16(ra);
    f(ra, c);
  }
}
__device__ __forceinline__ float bf16r(float v) { return __uint_as_float(pack_bf16x2(v, 0.f) << 16); }

// logits of 16 accumulator columns starting at col0, with the reference's roundings: bf16(acc) * scale (rounded again
// unless scale is a power of two, where the product is exact), + bias in fp32 (-inf at masked keys); a masked query
// row is all-equal in the reference (every entry filled with -max) -> uniform probabilities: logits 0
template <bool BIAS, bool ROUND2, bool MASK>
__device__ __forceinline__ void logits16(const uint32_t (&r)[16], float (&x)[16], float scale, const float* bias, int col0,
                                         bool row_masked) {
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
  if (MASK) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = row_masked ? 0.f : x[i];
  }
}
// exponent arguments (base 2) of 16 columns: without bias and with a power-of-two scale the scaling folds into one FFMA
template <bool BIAS, bool ROUND2, bool MASK>
__device__ __forceinline__ void exp_args16(const uint32_t (&r)[16], float (&x)[16], float scale, const float* bias, int col0,
                                           bool row_masked, float mneg) {
  if (!BIAS && !ROUND2) {
    const float sl2 = scale * kLog2e;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const uint32_t pk = pack_bf16x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      x[i] = fmaf(__uint_as_float(pk << 16), sl2, mneg);
      x[i + 1] = fmaf(__uint_as_float(pk & 0xffff0000u), sl2, mneg);
    }
  } else {
    logits16<BIAS, ROUND2, MASK>(r, x, scale, bias, col0, row_masked);
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], kLog2e, mneg);
  }
}

// column sums of a 32 x 16 block held one row per lane: 16 shuffles; lane l ends with the sum of column l >> 1
__device__ __forceinline__ float colsum16(const float (&v)[16], int lane) {
  float w8[8], w4[4], w2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
    w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? w8[i] : w8[i + 4], keep = b3 ? w8[i + 4] : w8[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? w4[i] : w4[i + 2], keep = b2 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? w2[0] : w2[1], keep = b1 ? w2[1] : w2[0];
  float s = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}

template <bool BIAS, bool ROUND2, bool MASK, bool COLSUM>
__global__ void __launch_bounds__(kThreads, 2) attention_kernel(const __grid_constant__ CUtensorMap tmap, const AttnParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;          // TMEM lane quarter; 0 = "lower" warp of the pair, 1 = "upper"
  const int N = p.N, H = p.H, C = H * 64, M = p.M;
  const int Np = (N + 15) & ~15;                     // keys: UMMA N of S, UMMA K of P.V
  const int ntiles = (M + 127) >> 7;
  const bool want_out = p.out != nullptr;
  const int nitems = p.B * H;

  // Q | K | V[0] | V[1], each [Np rows][128 B] in the 128-byte swizzled UMMA layout (16-byte chunk index XOR row % 8) --
  // exactly what ONE TMA box of (64 elements x Np tokens) with SWIZZLE_128B writes (a first TMA version wrote 8-element
  // boxes into the no-swizzle layout: 4.7 k 16-byte requests per item kept the copy at 4-5 us and slowed the P.V MMA
  // running beside it).  The CTA is persistent: while it works on the last query tile of an item, the next item's q, k
  // (dead once the last S MMA has completed) and v (other buffer) are already on their way.
  const uint32_t mat_bytes = (uint32_t)Np * 128u;
  unsigned char* Qs = smem;
  unsigned char* Ks = Qs + mat_bytes;
  unsigned char* Vs = Ks + mat_bytes;
  // the S MMA always reads 128 A rows: the operand area spans at least ntiles * 16 KB past Qs
  const size_t op_bytes = max((size_t)4 * mat_bytes, (size_t)ntiles * 16384);
  float* bias_s = reinterpret_cast<float*>(smem + op_bytes);                  // [2][Np]
  float* red_max = bias_s + 2 * Np;                                           // [2][128]
  float* red_sum = red_max + 256;                                             // [2][128]
  float* colpart = red_sum + 256;                                             // [4 quarters][Np] (COLSUM)
  uint64_t* bar = reinterpret_cast<uint64_t*>(colpart + (COLSUM ? 4 * Np : 0));
  uint64_t* ldbar = bar + 1;                                                  // [2]: q, k of item it landed (TMA)
  uint64_t* vbar = ldbar + 2;                                                 // [2]: v of item it landed (needed only by P.V)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(vbar + 2);

  const uint32_t ncols = Np <= 128 ? 128u : 256u;
  const uint32_t o_col = Np <= 128 ? 64u : 128u;
  TOKRED_STAMP(tid == 0, 0, 0);
  if (warp == 0) umma::tmem_alloc(tmem_slot, ncols);
  if (tid == 0) { umma::mbar_init(bar, 1); umma::mbar_init(&ldbar[0], 1); umma::mbar_init(&ldbar[1], 1);
                  umma::mbar_init(&vbar[0], 1); umma::mbar_init(&vbar[1], 1); umma::fence_mbar_init(); }
  __syncthreads();

  // ---- loads of one item into (Qs, Ks, V[buf], bias[buf]).  k, v and (without a row gather) q arrive by TMA: one thread
  // issues 8 boxes per matrix and moves on (a 75 KB cp.async prefetch kept its issuing warps blocked on the load/store
  // queue for as long as the copy took); rows past N are zero-filled by the out-of-bounds handling of the tensor map.
  // Gathered query rows (ATS) or a single query row (scores only) come by cp.async into the same layout.
  const bool q_by_tma = p.q_ids == nullptr && M == N;
  const size_t row_bytes = (size_t)3 * C * 2;
  auto issue_tma = [&](int item, int buf) {                       // one thread
    const int b = item / H, h = item - b * H;
    mbar_expect_tx(&ldbar[buf], (q_by_tma ? 2u : 1u) * mat_bytes);
    if (want_out) mbar_expect_tx(&vbar[buf], mat_bytes);
    const uint32_t k0 = umma::smem_u32(Ks), v0 = umma::smem_u32(Vs) + (uint32_t)buf * mat_bytes, q0 = umma::smem_u32(Qs);
    tma_load_4d(k0, &tmap, 0, H + h, 0, b, &ldbar[buf]);
    if (q_by_tma) tma_load_4d(q0, &tmap, 0, h, 0, b, &ldbar[buf]);
    if (want_out) tma_load_4d(v0, &tmap, 0, 2 * H + h, 0, b, &vbar[buf]);
  };
  auto issue_rest = [&](int item, int buf, int wi, int nw) {      // warp wi of nw issuing warps
    const int b = item / H, h = item - b * H;
    if (!q_by_tma) {
      const unsigned char* src0 = reinterpret_cast<const unsigned char*>(p.qkv + (size_t)b * N * 3 * C + (size_t)h * 64);
      const int64_t* ids = p.q_ids ? p.q_ids + (size_t)b * p.ids_stride : nullptr;
      const int rows = (M + 7) & ~7;                              // pad rows of the last group are zero-filled
      for (int e = wi * 32 + lane; e < rows * 8; e += nw * 32) {
        const int row = e >> 3, c = e & 7;
        const bool ok = row < M;
        const int srow = ok ? (ids ? clamp_idx(ids[row], N) : row) : 0;
        cp_async16(umma::smem_u32(Qs) + row * 128 + ((c ^ (row & 7)) << 4), src0 + (size_t)srow * row_bytes + c * 16, ok ? 16u : 0u);
      }
    }
    if (BIAS)
      for (int j = wi * 32 + lane; j < Np; j += nw * 32) {
        float bv = 0.f;
        if (j < N) {
          if (p.key_bias) bv = p.key_bias[(size_t)b * N + j];
          if (MASK && !p.mask[(size_t)b * N + j]) bv = -INFINITY;
        }
        bias_s[buf * Np + j] = bv;
      }
  };
  int item = blockIdx.x;
  if (item < nitems) {
    if (tid == 32) issue_tma(item, 0);
    issue_rest(item, 0, warp, kThreads / 32);
  }
  if (COLSUM)
    for (int j = tid; j < 4 * Np; j += kThreads) colpart[j] = 0.f;
  cp_async_wait_all();
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  TOKRED_STAMP(tid == 0, 0, 1);

  const uint32_t idesc_s = umma::instr_desc(umma::FMT_BF16, 128, (uint32_t)Np);
  const uint32_t idesc_o = umma::instr_desc(umma::FMT_BF16, 128, 64) | (1u << 16);      // B operand (v) MN-major
  const float scale = p.scale;
  uint32_t phase = 0;
  const int nch = Np >> 4;
  // Column split of passes A and B between the two warps of a quarter; pass C (cheap, and its in-place writes must trail
  // its reads) is the lower warp's alone, so the lower warp takes the smaller share of A and B.  Measured at N = 197
  // (13 chunks): 4/9 64.2 us, 5/8 62.9 us, 6/7 63.1 us (B=256), 431 / 429 / 425 us (B=1024).
  // (Splitting pass C too was measured twice: with the upper warp parking its whole packed half in registers until the
  // lower warp has read, 128 registers no longer hold it and the spills made it slower; with at most 32 parked
  // registers and an even A/B split it came out even, and with 2-3 loads in flight per round on both warps it was
  // slower again (68 vs 63 us: the pair barrier in the middle and ~125 registers cost more than the halved chunk count
  // saves).)
  const int nlo = (nch * 15 + 16) >> 5;
  const int c_beg = half ? nlo : 0, c_end = half ? nch : nlo;
  const int c_full = (c_end == nch && c_end > c_beg && (N & 15)) ? c_end - 1 : c_end;      // [c_beg, c_full) full chunks, then the ragged one
  const uint32_t trow = umma::tmem_addr(tmem, (uint32_t)(q * 32), 0);
  const int rl = q * 32 + lane;                            // row inside the tile

  for (int it = 0; item < nitems; ++it, item += gridDim.x) {
    const int buf = it & 1;
    const int b = item / H, h = item - b * H;
    const int64_t* ids = p.q_ids ? p.q_ids + (size_t)b * p.ids_stride : nullptr;
    const float* bias_i = bias_s + buf * Np;
    const uint32_t v_i = umma::smem_u32(Vs) + (uint32_t)buf * mat_bytes;
    if (tid == 0) umma::mbar_wait(&ldbar[buf], (uint32_t)((it >> 1) & 1));      // this item's TMA boxes have landed

    for (int t = 0; t < ntiles; ++t) {
      // ---- S = Q_t K^T (rows past the Q tile read what lies behind it: finite garbage in accumulator rows nobody reads)
      if (tid == 0) {
        const uint32_t a0 = umma::smem_u32(Qs) + (uint32_t)t * 16384u, b0 = umma::smem_u32(Ks);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma::mma_bf16(tmem, umma::smem_desc_sw128(a0 + ks * 32, 16, 1024), umma::smem_desc_sw128(b0 + ks * 32, 16, 1024),
                         idesc_s, ks > 0 ? 1u : 0u);
        umma::mma_commit(bar);
      }
      umma::mbar_wait(bar, phase);
      phase ^= 1u;
      umma::tc_fence_after_sync();
      TOKRED_STAMP(tid == 0 && it == 0, t, 2);
      const int row = t * 128 + rl;                          // query row
      const bool active = t * 128 + q * 32 < M;              // uniform over the warp pair of a quarter
      // The last S MMA of this item has completed: q and k are dead, the other v buffer was last read an item ago ->
      // prefetch the next item (TMA: one thread; the bias vector and cp.async query rows: the warps of quarters that hold
      // no query row of this tile, all warps only when every quarter is busy).
      if (t == ntiles - 1 && item + (int)gridDim.x < nitems) {
        if (tid == 32) issue_tma(item + gridDim.x, buf ^ 1);
        const int aq = min(4, (M - t * 128 + 31) >> 5);      // quarters with query rows in this tile
        if (aq == 4) issue_rest(item + gridDim.x, buf ^ 1, warp, kThreads / 32);
        else if (!active) issue_rest(item + gridDim.x, buf ^ 1, (q - aq) * 2 + half, (4 - aq) * 2);
      }
      if (active) {
        bool row_masked = false;
        if (MASK && row < M) row_masked = !p.mask[(size_t)b * N + (ids ? clamp_idx(ids[row], N) : row)];
        // pass A: row maximum over this warp's columns.  Without a bias the roundings and the (positive) scale are
        // monotone, so the maximum is taken over the raw accumulators and rounded once.
        float m0 = -INFINITY, m1 = -INFINITY;
        auto pass_a = [&](uint32_t (&r)[16], int c, bool ragged) {
          float x[16];
          if (BIAS) logits16<BIAS, ROUND2, MASK>(r, x, scale, bias_i, c * 16, row_masked);
          else {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(r[i]);
          }
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c * 16 + i >= N) x[i] = -INFINITY;
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            m0 = fmaxf(m0, fmaxf(x[i], x[i + 1]));
            m1 = fmaxf(m1, fmaxf(x[i + 2], x[i + 3]));
          }
        };
        for_chunks(trow, c_beg, c_full, [&](uint32_t (&r)[16], int c) { pass_a(r, c, false); });
        if (c_full < c_end) {
          uint32_t r[16];
          umma::tmem_ld16(trow + c_full * 16, r);
          ld_wait16(r);
          pass_a(r, c_full, true);
        }
        float mw = fmaxf(m0, m1);
        if (!BIAS) {
          mw = bf16r(mw) * scale;
          if (ROUND2) mw = bf16r(mw);
        }
        red_max[half * 128 + rl] = mw;
        TOKRED_STAMP(tid == 128 && it == 0, t, 3);
        TOKRED_STAMP(tid == 0 && it == 0, t, 8);
        pair_sync(q);
        const float mx = fmaxf(red_max[rl], red_max[128 + rl]);
        // pass B: e = exp(x - max) (kept in place as fp32), row sum
        const float mneg = -mx * kLog2e;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        auto pass_b = [&](uint32_t (&r)[16], int c, bool ragged) {
          float x[16];
          exp_args16<BIAS, ROUND2, MASK>(r, x, scale, bias_i, c * 16, row_masked, mneg);
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c * 16 + i >= N) x[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4) { s0 += x[i]; s1 += x[i + 1]; s2 += x[i + 2]; s3 += x[i + 3]; }
          uint32_t e[16];                  // a separate array: writing into r would keep the values alive for the next wait
#pragma unroll
          for (int i = 0; i < 16; ++i) e[i] = __float_as_uint(x[i]);
          tmem_st16(trow + c * 16, e);
        };
        for_chunks(trow, c_beg, c_full, [&](uint32_t (&r)[16], int c) { pass_b(r, c, false); });
        if (c_full < c_end) {
          uint32_t r[16];
          umma::tmem_ld16(trow + c_full * 16, r);
          ld_wait16(r);
          pass_b(r, c_full, true);
        }
        red_sum[half * 128 + rl] = (s0 + s1) + (s2 + s3);
        tmem_st_wait();
        TOKRED_STAMP(tid == 128 && it == 0, t, 4);
        TOKRED_STAMP(tid == 0 && it == 0, t, 9);
        pair_sync(q);
        const float inv = 1.0f / (red_sum[rl] + red_sum[128 + rl]);
        if (COLSUM) {
          // column sums of the fp32 probabilities: both warps of the quarter work on their own columns (the butterfly
          // would triple pass C, which only the lower warp can run), before pass C starts overwriting e
          for_chunks(trow, c_beg, c_end, [&](uint32_t (&r)[16], int c) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = row < M ? __uint_as_float(r[i]) * inv : 0.f;
            const float cs = colsum16(v, lane);
            // one slot per quarter: the second query tile adds onto the first (same warp, fixed order: deterministic).
            // Eight slots pushed the CTA past 113 KB of shared memory: one CTA per SM instead of two, 177 us instead of 95.
            if (!(lane & 1)) colpart[(size_t)q * Np + c * 16 + (lane >> 1)] += cs;
          });
          pair_sync(q);
        }
        // pass C (lower warp): p = e / sum -> bf16 pairs in place (column j holds keys 2j, 2j+1; the writes trail the
        // reads); the CLS row goes out in fp32
        if (!half) {
          float* cls = (p.cls_row && row == 0) ? p.cls_row + ((size_t)b * H + h) * N : nullptr;
          for_chunks(trow, 0, nch, [&](uint32_t (&r)[16], int c) {
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(r[i]) * inv;
            if (cls) {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (c * 16 + i < N) cls[c * 16 + i] = x[i];
            }
            tmem_st8(trow + c * 8, pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                     pack_bf16x2(x[6], x[7]), pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]),
                     pack_bf16x2(x[14], x[15]));
          });
          tmem_st_wait();
          TOKRED_STAMP(tid == 0 && it == 0, t, 5);
        }
      }
      umma::tc_fence_before_sync();
      __syncthreads();
      if (!want_out) continue;             // scores only (uniform): nothing reads the accumulator again

      // ---- O = P V : A from TMEM, B = v tile used MN-major (rows = tokens = the K index; 16 tokens per MMA = 2048 bytes)
      if (tid == 0) {
        umma::tc_fence_after_sync();
        if (t == 0) umma::mbar_wait(&vbar[buf], (uint32_t)((it >> 1) & 1));       // v has its own barrier: S never waits for it
        for (int ks = 0; ks < nch; ++ks)
          mma_bf16_ts(tmem + o_col, tmem + ks * 8, umma::smem_desc_sw128(v_i + ks * 2048, 16, 1024), idesc_o, ks > 0 ? 1u : 0u);
        umma::mma_commit(bar);
      }
      umma::mbar_wait(bar, phase);
      phase ^= 1u;
      umma::tc_fence_after_sync();
      TOKRED_STAMP(tid == 0 && it == 0, t, 6);
      if (active) {                                          // whole warp: the TMEM loads are .sync.aligned
        __nv_bfloat16* dst = p.out + ((size_t)b * M + (row < M ? row : 0)) * C + h * 64 + half * 32;
        uint32_t ra[16], rb[16];
        umma::tmem_ld16(trow + o_col + half * 32, ra);
        umma::tmem_ld16(trow + o_col + half * 32 + 16, rb);
        ld_wait16(ra);
        ld_wait16(rb);
        if (row < M) {
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(pack_bf16x2(__uint_as_float(ra[0]), __uint_as_float(ra[1]))),
                       "r"(pack_bf16x2(__uint_as_float(ra[2]), __uint_as_float(ra[3]))), "r"(pack_bf16x2(__uint_as_float(ra[4]), __uint_as_float(ra[5]))),
                       "r"(pack_bf16x2(__uint_as_float(ra[6]), __uint_as_float(ra[7]))), "r"(pack_bf16x2(__uint_as_float(ra[8]), __uint_as_float(ra[9]))),
                       "r"(pack_bf16x2(__uint_as_float(ra[10]), __uint_as_float(ra[11]))), "r"(pack_bf16x2(__uint_as_float(ra[12]), __uint_as_float(ra[13]))),
                       "r"(pack_bf16x2(__uint_as_float(ra[14]), __uint_as_float(ra[15])))
                       : "memory");
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 16), "r"(pack_bf16x2(__uint_as_float(rb[0]), __uint_as_float(rb[1]))),
                       "r"(pack_bf16x2(__uint_as_float(rb[2]), __uint_as_float(rb[3]))), "r"(pack_bf16x2(__uint_as_float(rb[4]), __uint_as_float(rb[5]))),
                       "r"(pack_bf16x2(__uint_as_float(rb[6]), __uint_as_float(rb[7]))), "r"(pack_bf16x2(__uint_as_float(rb[8]), __uint_as_float(rb[9]))),
                       "r"(pack_bf16x2(__uint_as_float(rb[10]), __uint_as_float(rb[11]))), "r"(pack_bf16x2(__uint_as_float(rb[12]), __uint_as_float(rb[13]))),
                       "r"(pack_bf16x2(__uint_as_float(rb[14]), __uint_as_float(rb[15])))
                       : "memory");
        }
      }
      umma::tc_fence_before_sync();
      __syncthreads();
      TOKRED_STAMP(tid == 0 && it == 0, t, 7);
    }
    if (COLSUM) {
      // fixed-order combine of the quarter partials: deterministic
      float* dst = p.colsum + ((size_t)b * H + h) * N;
      for (int j = tid; j < N; j += kThreads) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s += colpart[(size_t)k * Np + j]; colpart[(size_t)k * Np + j] = 0.f; }
        dst[j] = s;
      }
    }
    // the next item's operands (issued during the last tile) have landed
    cp_async_wait_all();
    umma::fence_proxy_async_smem();
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    TOKRED_STAMP(tid == 0 && it < 8, it, 10);
  }
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

size_t attn_smem_bytes(int N, int M, bool colsum) {
  const int Np = (N + 15) & ~15;
  const size_t ops = (size_t)4 * Np * 128, mma = (size_t)((M + 127) / 128) * 16384;
  return (ops > mma ? ops : mma) + (size_t)2 * Np * 4 + 2048 + (colsum ? (size_t)4 * Np * 4 : 0) + 96;
}

}  // namespace
}  // namespace tokred

using namespace tokred;
TOKRED_STAMP_SETTER(attention)

namespace {
// cuTensorMapEncodeTiled through the runtime's driver entry point: the library links cudart only
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}
}  // namespace

extern "C" int tokred_attention(const void* qkv, int B, int N, int H, int head_dim, float scale, const float* key_bias,
                                const uint8_t* mask, const int64_t* q_ids, int64_t ids_stride, int M, void* out,
                                float* cls_row, float* colsum, void* stream) {
  TOKRED_REQUIRE(qkv, "attention: null qkv");
  TOKRED_REQUIRE(out || cls_row || colsum, "attention: no output requested");
  TOKRED_REQUIRE(B >= 1 && H >= 1 && N >= 1, "attention: B=%d N=%d H=%d", B, N, H);
  TOKRED_REQUIRE(scale > 0.f && std::isfinite(scale), "attention: scale %g must be positive", (double)scale);
  if (!q_ids) M = (out || colsum) ? N : 1;        // CLS rows only: query row 0 is all that is needed
  TOKRED_REQUIRE(M >= 1 && (!q_ids || ids_stride >= M), "attention: M=%d ids_stride=%lld", M, (long long)ids_stride);
  if (head_dim != 64 || N > 256 || M > 256) {
    set_error("attention: head_dim=%d N=%d M=%d outside the fused kernel's range (head_dim 64, N <= 256)", head_dim, N, M);
    return TOKRED_ERR_UNSUPPORTED;
  }
  TOKRED_REQUIRE((long long)B * H <= 0x7fffffffLL, "attention: B*H too large");
  TOKRED_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0,
                 "attention: qkv must be 16-byte and out 32-byte aligned");
  AttnParams prm{};
  prm.qkv = (const __nv_bfloat16*)qkv; prm.key_bias = key_bias; prm.mask = mask; prm.q_ids = q_ids; prm.ids_stride = ids_stride;
  prm.out = (__nv_bfloat16*)out; prm.cls_row = cls_row; prm.colsum = colsum;
  prm.B = B; prm.N = N; prm.H = H; prm.M = M; prm.scale = scale;
  int ex = 0;
  const bool r2 = std::frexp(scale, &ex) != 0.5f;          // not a power of two: the scaled logits round to bf16 again
  const bool cs = colsum != nullptr;
  const size_t smem = attn_smem_bytes(N, M, cs);
  // persistent CTAs, two per SM (108 KB of shared memory and <= 256 TMEM columns each)
  const int grid = B * H < 2 * kNumSMs ? B * H : 2 * kNumSMs;
  cudaStream_t st = (cudaStream_t)stream;
  // qkv as a 4-D tensor (d, q/k/v x head, token, image); one box = the 64 elements of Np consecutive tokens of one
  // (q/k/v, head) slot, written 128-byte swizzled.  Tokens past N are out of bounds: the copy writes zeros.
  EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) {
    set_error("attention: cuTensorMapEncodeTiled is not available from this driver");
    return TOKRED_ERR_UNSUPPORTED;
  }
  const int C = H * 64, Np = (N + 15) & ~15;
  CUtensorMap tmap;
  const cuuint64_t dims[4] = {64, (cuuint64_t)(3 * H), (cuuint64_t)N, (cuuint64_t)B};
  const cuuint64_t strides[3] = {128, (cuuint64_t)3 * C * 2, (cuuint64_t)N * 3 * C * 2};
  const cuuint32_t box[4] = {64, 1, (cuuint32_t)Np, 1}, estr[4] = {1, 1, 1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(qkv), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("attention: cuTensorMapEncodeTiled failed (%d) for B=%d N=%d H=%d", (int)cr, B, N, H);
    return TOKRED_ERR_UNSUPPORTED;
  }
#define LAUNCH(BIAS, R2, MASK, CS)                                                                  \
  do {                                                                                              \
    if (int e = allow_smem(attention_kernel<BIAS, R2, MASK, CS>, smem, "attention")) return e;     \
    attention_kernel<BIAS, R2, MASK, CS><<<grid, kThreads, smem, st>>>(tmap, prm);                 \
  } while (0)
#define PICK(BIAS, MASK)                                                                            \
  do {                                                                                              \
    if (r2) { if (cs) LAUNCH(BIAS, true, MASK, true); else LAUNCH(BIAS, true, MASK, false); }       \
    else    { if (cs) LAUNCH(BIAS, false, MASK, true); else LAUNCH(BIAS, false, MASK, false); }     \
  } while (0)
  if (mask) PICK(true, true);
  else if (key_bias) PICK(true, false);
  else PICK(false, false);
#undef PICK
#undef LAUNCH
  return finish_launch("attention");
}
