// Score-ordered token selection kernels: Top-K / DynamicViT keep, EViT select+fuse, row gather, DynamicViT pooling.
//
// All of them are HBM-bound byte movers (0 flop/byte): one CTA-row per image orders <= a few hundred scores in
// shared memory (rank-by-counting, lowest-index ties) and then streams token rows with 16-byte no-allocate
// loads/stores, one warp per row.  grid = (splits, B): the ordering is recomputed by every split (196^2 LDS
// compares, ~1 us) so that small batches still fill 148 SMs without a second launch or a cluster barrier.
#include "common.cuh"

namespace tokred {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct ScoreSrc {
  const void* scores;      // [B,P] strided, or nullptr
  int score_dtype;
  long long stride, batch_stride;
  const void* attn;        // [B,H,N,N], used when scores == nullptr
  int attn_dtype;
  int H;
};

__device__ __forceinline__ float load_any(const void* p, int dtype, long long i) {
  return dtype == TOKRED_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                              : reinterpret_cast<const float*>(p)[i];
}

// s_p for image b.  From attention: head-mean of the CLS row (models/topk.py:60-61); ATen's CUDA mean is
// sum * (1/H) accumulated in fp32 and rounded to the tensor dtype.
__device__ __forceinline__ float fetch_score(const ScoreSrc& s, int b, int p, int N) {
  if (s.scores) return load_any(s.scores, s.score_dtype, (long long)b * s.batch_stride + (long long)p * s.stride);
  float acc = 0.f;
  for (int h = 0; h < s.H; ++h)
    acc += load_any(s.attn, s.attn_dtype, ((long long)(b * s.H + h) * N) * N + 1 + p);
  acc *= (1.0f / (float)s.H);
  return s.attn_dtype == TOKRED_BF16 ? bf16_round(acc) : acc;
}

__device__ __forceinline__ void copy_row(char* dst, const char* src, int row_bytes, int vec16, int elem_size, int lane) {
  if (vec16) {
    warp_copy_row16(dst, src, row_bytes, lane);
  } else if (elem_size == 4) {
    warp_copy_row_elems(reinterpret_cast<uint32_t*>(dst), reinterpret_cast<const uint32_t*>(src), row_bytes / 4, lane);
  } else {
    warp_copy_row_elems(reinterpret_cast<uint16_t*>(dst), reinterpret_cast<const uint16_t*>(src), row_bytes / 2, lane);
  }
}

// x + branch on the fly (fp32 stream, bf16 branch: the block's `x = x + drop_path(attn branch)`, e.g. models/topk.py:87, as
// the same fp32 addition) for the rows that survive only: four rows in flight, one 16-byte + one 8-byte load per row and step
__device__ __forceinline__ int4 add_bf16x4(const int4& a, const uint2& c) {
  int4 r;
  r.x = __float_as_int(__fadd_rn(__int_as_float(a.x), __uint_as_float(c.x << 16)));
  r.y = __float_as_int(__fadd_rn(__int_as_float(a.y), __uint_as_float(c.x & 0xffff0000u)));
  r.z = __float_as_int(__fadd_rn(__int_as_float(a.z), __uint_as_float(c.y << 16)));
  r.w = __float_as_int(__fadd_rn(__int_as_float(a.w), __uint_as_float(c.y & 0xffff0000u)));
  return r;
}
__device__ __forceinline__ void warp_addcopy_rows16x4(char* dbase, const int (&doff)[4], const char* sbase, const char* bbase,
                                                      const int (&soff)[4], int nrows, int bytes, int lane) {
  for (int off = lane * 16; off < bytes; off += 512) {
    int4 a[4];
    uint2 c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nrows) {
        a[u] = ld_stream16(sbase + soff[u] + off);
        c[u] = *reinterpret_cast<const uint2*>(bbase + ((soff[u] + off) >> 1));
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nrows) st_stream16(dbase + doff[u] + off, add_bf16x4(a[u], c[u]));
  }
}

// output row j <- token 0 (j == 0) or kept patch sel[j-1]; the rows of this CTA (blockIdx.x of gridDim.x splits) are
// dealt to its warps four at a time so that every warp has four source rows in flight
template <bool ADD = false>
__device__ __forceinline__ void gather_kept_rows(char* ob, const char* xb, const int* sel, int nrows, int row_bytes, int vec16,
                                                 int elem_size, int warp, int lane, const char* bb = nullptr) {
  const int stride = gridDim.x * kWarps;
  for (int j = blockIdx.x * kWarps + warp; j < nrows; j += 4 * stride) {
    int so[4], dof[4], nr = 0;        // byte offsets inside the image (an image is far below 2 GB)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = j + u * stride;
      const bool live = jj < nrows;
      const int src_row = !live || jj == 0 ? 0 : 1 + sel[jj - 1];
      so[u] = src_row * row_bytes;
      dof[u] = (live ? jj : 0) * row_bytes;
      nr += live;
    }
    if (ADD) {
      warp_addcopy_rows16x4(ob, dof, xb, bb, so, nr, row_bytes, lane);      // host: fp32 stream on the vector path only
    } else if (vec16) {
      warp_copy_rows16x4(ob, dof, xb, so, nr, row_bytes, lane);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nr) copy_row(ob + dof[u], xb + so[u], row_bytes, 0, elem_size, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------ Top-K
template <bool ADD>       // ADD: x + branch on the rows that are kept (own instantiation: the plain gather keeps its 64 registers)
__global__ void __launch_bounds__(kThreads, ADD ? 3 : 4)
topk_gather_kernel(ScoreSrc ss, const char* __restrict__ x, const char* __restrict__ branch, char* __restrict__ x_out,
                   int64_t* __restrict__ idx_out, int N, int k, int row_bytes, int vec16, int elem_size) {
  extern __shared__ float smem[];
  const int P = N - 1, b = blockIdx.y, tid = threadIdx.x;
  float* keys = smem;
  int* sel = reinterpret_cast<int*>(keys + P);

  for (int p = tid; p < P; p += kThreads) keys[p] = fetch_score(ss, b, p, N);
  __syncthreads();
  for (int p = tid; p < P; p += kThreads) {
    int r = rank_desc_window(keys, P, p, p - (tid & 31), p - (tid & 31) + 32);
    if (r < k) {
      sel[r] = p;
      if (blockIdx.x == 0) idx_out[(long long)b * k + r] = p;
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const char* xb = x + (long long)b * N * row_bytes;
  char* ob = x_out + (long long)b * (k + 1) * row_bytes;
  gather_kept_rows<ADD>(ob, xb, sel, k + 1, row_bytes, vec16, elem_size, warp, lane,
                        ADD ? branch + (long long)b * N * (row_bytes >> 1) : nullptr);
}

// ------------------------------------------------------------------------------------------ EViT
// grid.x = number of 1024-byte column slices of a token row (round 1: 512-byte slices -> 768 CTAs at B=128, two waves).
// Split s gathers its share of the kept rows and owns slice s of the fused token: warp w accumulates complement rows
// w, w+8, ... (ascending patch order), the 8 partials are combined in warp order through shared memory -> deterministic.
template <typename T, bool ADD = false>
__global__ void __launch_bounds__(kThreads)
evit_select_fuse_kernel(ScoreSrc ss, const T* __restrict__ x, const __nv_bfloat16* __restrict__ branch, T* __restrict__ x_out,
                        int64_t* __restrict__ idx_out, int64_t* __restrict__ compl_out, int N, int C, int k, int vec16) {
  extern __shared__ float smem[];
  constexpr int VE = 16 / sizeof(T);            // elements per 16-byte lane chunk
  constexpr int SLICE = 2 * 32 * VE;            // elements per 1024-byte slice
  const int P = N - 1, M = P - k, b = blockIdx.y, tid = threadIdx.x;
  float* keys = smem;                                   // [P]
  int* sel = reinterpret_cast<int*>(keys + P);          // [k]
  int* cmp = sel + k;                                   // [M]
  unsigned char* dropped = reinterpret_cast<unsigned char*>(cmp + M);   // [P]
  float* part = reinterpret_cast<float*>(smem) + P + k + M + (P + 3) / 4;   // [kWarps][SLICE]

  for (int p = tid; p < P; p += kThreads) keys[p] = fetch_score(ss, b, p, N);
  __syncthreads();
  for (int p = tid; p < P; p += kThreads) {
    int r = rank_desc_window(keys, P, p, p - (tid & 31), p - (tid & 31) + 32);
    dropped[p] = r >= k;
    if (r < k) {
      sel[r] = p;
      if (blockIdx.x == 0) idx_out[(long long)b * (k + 1) + r] = p;
    }
  }
  if (blockIdx.x == 0 && tid == 0) idx_out[(long long)b * (k + 1) + k] = -1;
  __syncthreads();
  for (int p = tid; p < P; p += kThreads) {
    if (dropped[p]) {
      int pos = 0;
      for (int q = 0; q < p; ++q) pos += dropped[q];
      cmp[pos] = p;
      if (blockIdx.x == 0) compl_out[(long long)b * M + pos] = p;
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int row_bytes = C * (int)sizeof(T);
  const T* xb = x + (long long)b * N * C;
  T* ob = x_out + (long long)b * (k + 2) * C;
  // kept rows
  const __nv_bfloat16* bb = nullptr;                                    // x + branch on every row read (fp32 stream only)
  if constexpr (ADD) bb = branch + (long long)b * N * C;
  gather_kept_rows<ADD>(reinterpret_cast<char*>(ob), reinterpret_cast<const char*>(xb), sel, k + 1, row_bytes, vec16, (int)sizeof(T),
                   warp, lane, reinterpret_cast<const char*>(bb));
  // fused inattentive token, slice blockIdx.x (two 16-byte chunks per lane)
  const int e0 = blockIdx.x * SLICE + lane * VE;
  float acc[2][VE];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[h][i] = 0.f;
  if (vec16) {
    // 4 complement rows x 2 chunks per step: the 16-byte loads are issued together, then consumed in ascending row order
    for (int m0 = warp; m0 < M; m0 += 4 * kWarps) {
      int4 raw[4][2];
      uint2 rb[4][2];
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = m0 + u * kWarps;
        w[u] = 0.f;
        if (m < M) {
          const int p = cmp[m];
          w[u] = keys[p];
          const T* row = xb + (long long)(1 + p) * C + e0;
          if (e0 < C) raw[u][0] = ld_stream16(row);
          if (e0 + 32 * VE < C) raw[u][1] = ld_stream16(row + 32 * VE);
          if constexpr (ADD) {
            {
              const __nv_bfloat16* brow = bb + (long long)(1 + p) * C + e0;
              if (e0 < C) rb[u][0] = *reinterpret_cast<const uint2*>(brow);
              if (e0 + 32 * VE < C) rb[u][1] = *reinterpret_cast<const uint2*>(brow + 32 * VE);
            }
          }
        }
      }
      if constexpr (ADD) {
        {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (m0 + u * kWarps < M) {
              if (e0 < C) raw[u][0] = add_bf16x4(raw[u][0], rb[u][0]);
              if (e0 + 32 * VE < C) raw[u][1] = add_bf16x4(raw[u][1], rb[u][1]);
            }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (m0 + u * kWarps < M) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (e0 + h * 32 * VE < C) {
              const T* v = reinterpret_cast<const T*>(&raw[u][h]);
#pragma unroll
              for (int i = 0; i < VE; ++i) acc[h][i] = fmaf(w[u], to_f32(v[i]), acc[h][i]);
            }
        }
      }
    }
  } else {
    for (int m = warp; m < M; m += kWarps) {
      const int p = cmp[m];
      const float w = keys[p];
      const T* row = xb + (long long)(1 + p) * C + e0;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < VE; ++i)
          if (e0 + h * 32 * VE + i < C) acc[h][i] = fmaf(w, to_f32(row[h * 32 * VE + i]), acc[h][i]);
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < VE; ++i) part[warp * SLICE + h * 32 * VE + lane * VE + i] = acc[h][i];
  __syncthreads();
  for (int e = tid; e < SLICE; e += kThreads) {
    if (blockIdx.x * SLICE + e < C) {
      float s = part[e];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) s += part[w * SLICE + e];
      ob[(long long)(k + 1) * C + blockIdx.x * SLICE + e] = from_f32<T>(s);
    }
  }
}

// ------------------------------------------------------------------------------------------ row gather
// out[b,g,m,:] = src[b,g,ids[b,m],:]; one warp per group of 4 consecutive output rows, grid-stride over groups.
// Attention rows (N = 197 fp32 = 788 B) are only 4-byte aligned, so that path moves 4-byte words — with the loads
// of all 4 rows issued before the first store (28 requests in flight per lane instead of 7).
template <typename W4>
__device__ __forceinline__ void copy_rows4(const char* const (&sp)[4], char* const (&dp)[4], int nrows, int n, int lane) {
  for (int i = lane; i < n; i += 64) {
    W4 a[4], b[4];
    const bool two = i + 32 < n;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nrows) {
        a[u] = reinterpret_cast<const W4*>(sp[u])[i];
        if (two) b[u] = reinterpret_cast<const W4*>(sp[u])[i + 32];
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nrows) {
        reinterpret_cast<W4*>(dp[u])[i] = a[u];
        if (two) reinterpret_cast<W4*>(dp[u])[i + 32] = b[u];
      }
  }
}

template <bool VEC16>       // separate instantiations: the 16-byte path's registers must not cost the word path its occupancy
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(const char* __restrict__ src, const int64_t* __restrict__ ids, long long ids_stride, int G, int N,
                   int M, long long rows, int row_bytes, int vec16, int elem_size, char* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row0 = ((long long)blockIdx.x * kWarps + warp) * 4; row0 < rows; row0 += (long long)gridDim.x * kWarps * 4) {
    const char* sp[4];
    char* dp[4];
    const int nrows = (int)min((long long)4, rows - row0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long row = row0 + (u < nrows ? u : 0);
      const int m = (int)(row % M);
      const long long bg = row / M;
      const int b = (int)(bg / G);
      long long id = ids[(long long)b * ids_stride + m];
      id = id < 0 ? 0 : (id >= N ? N - 1 : id);
      sp[u] = src + (bg * N + id) * row_bytes;
      dp[u] = out + row * row_bytes;
    }
    if (VEC16) {
      // all four rows' loads before the first store (offsets relative to the first row: a few images apart at most)
      int so[4], dof[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { so[u] = (int)(sp[u] - sp[0]); dof[u] = (int)(dp[u] - dp[0]); }
      warp_copy_rows16x4(dp[0], dof, sp[0], so, nrows, row_bytes, lane);
    } else if (elem_size == 4) {
      copy_rows4<uint32_t>(sp, dp, nrows, row_bytes / 4, lane);
    } else {
      copy_rows4<uint16_t>(sp, dp, nrows, row_bytes / 2, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------ DynamicViT pooling
// grid = (128-channel slices of the global half, B); 256 threads = 16 patch-phases x 16 channel-threads, every
// thread owns 8 consecutive channels (one 16-byte load of bf16, two of fp32).
// Phase 1: masked mean over patches of this slice (each phase sums its patches in order, phases combined in fixed
// order -> deterministic).  Phase 2: write both halves of this slice with 16-byte stores.
constexpr int kPoolSlice = 128;
constexpr int kPoolPhases = kThreads / (kPoolSlice / 8);   // 16
template <typename T> __device__ __forceinline__ void ld8(const T* p, bool vec, int valid, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, bool vec, int valid, float (&v)[8]) {
  if (vec && valid >= 8) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < valid ? p[i] : 0.f;
  }
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, bool vec, int valid, float (&v)[8]) {
  if (vec && valid >= 8) {
    const int4 a = ld_stream16(p);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&a);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(h[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < valid ? __bfloat162float(p[i]) : 0.f;
  }
}
template <typename T> __device__ __forceinline__ void st8(T* p, bool vec, int valid, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, bool vec, int valid, const float (&v)[8]) {
  if (vec && valid >= 8) {
    // ONE 32-byte store per thread (STG.256; rows are 32-byte aligned on the vector path): two 16-byte stores at a
    // 32-byte lane stride touch every sector twice
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < valid) p[i] = v[i];
  }
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, bool vec, int valid, const float (&v)[8]) {
  if (vec && valid >= 8) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    st_stream16(p, *reinterpret_cast<const int4*>(h));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < valid) p[i] = __float2bfloat16_rn(v[i]);
  }
}

// raw 8-channel chunk of a row: bf16 stays packed in registers (4 instead of 8) until it is consumed, so that more
// rows can be in flight per thread without pushing the CTA count per SM down
template <typename T> struct Raw8;
template <> struct Raw8<float> { float f[8]; };
template <> struct Raw8<__nv_bfloat16> { int4 q; };
__device__ __forceinline__ void ldraw(const float* p, bool vec, int valid, Raw8<float>& r) { ld8<float>(p, vec, valid, r.f); }
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, bool vec, int valid, Raw8<__nv_bfloat16>& r) {
  if (vec && valid >= 8) {
    r.q = ld_stream16(p);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t lo = 2 * i < valid ? (uint32_t)__bfloat16_as_ushort(p[2 * i]) : 0u;
      const uint32_t hi = 2 * i + 1 < valid ? (uint32_t)__bfloat16_as_ushort(p[2 * i + 1]) : 0u;
      w[i] = lo | (hi << 16);
    }
    r.q = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
  }
}
__device__ __forceinline__ void unraw(const Raw8<float>& r, float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = r.f[i];
}
__device__ __forceinline__ void unraw(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
  const uint32_t w[4] = {(uint32_t)r.q.x, (uint32_t)r.q.y, (uint32_t)r.q.z, (uint32_t)r.q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

template <typename TI, typename TO, bool vec>
__global__ void __launch_bounds__(kThreads, 3)
dyvit_pool_concat_kernel(const TI* __restrict__ h, long long hbs, const float* __restrict__ policy, TO* __restrict__ out, int P,
                         int C, float eps) {
  extern __shared__ float smem[];
  // rows in flight per thread (P = 196: 13 rows per thread -> 2 / 4 rounds); the scalar fallback keeps one
  constexpr int R = !vec ? 1 : (sizeof(TI) == 2 ? 6 : 4);
  float* part = smem;                              // [kPoolPhases][kPoolSlice]
  float* pol = smem + kPoolPhases * kPoolSlice;    // [P]
  const int b = blockIdx.y, tid = threadIdx.x, half = C / 2;
  const int ct = tid % (kPoolSlice / 8), ph = tid / (kPoolSlice / 8);
  const int c0 = blockIdx.x * kPoolSlice + ct * 8;          // first of this thread's 8 channels (within a half)
  const int valid = half - c0;                              // <= 0: thread idle
  const int vld = vec ? 8 : valid;                          // vector path: half % 8 == 0, a live chunk is always whole
  const TI* hb = h + (long long)b * hbs;          // hbs: elements between images (h may be the [:, 1:] view of [B,P+1,C])
  for (int p = tid; p < P; p += kThreads) pol[p] = policy[(long long)b * P + p];
  __syncthreads();
  float psum = 0.f;
  for (int p = 0; p < P; ++p) psum += pol[p];
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (valid > 0) {
    for (int p0 = ph; p0 < P; p0 += R * kPoolPhases) {          // R rows in flight per thread, consumed in ascending order
      Raw8<TI> raw[R];
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (p0 + kPoolPhases * u < P) ldraw(hb + (long long)(p0 + kPoolPhases * u) * C + half + c0, vec, vld, raw[u]);
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (p0 + kPoolPhases * u < P) {
          const float w = pol[p0 + kPoolPhases * u];
          float v[8];
          unraw(raw[u], v);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += v[i] * w;
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[ph * kPoolSlice + ct * 8 + i] = acc[i];
  __syncthreads();
  float g[8];
  {
    // phases combined in ascending order; two 16-byte reads per phase (128 scalar reads scheduled at once held 128 registers)
    const float4* p4 = reinterpret_cast<const float4*>(part + ct * 8);
    float4 s0 = p4[0], s1 = p4[1];
#pragma unroll 4
    for (int q = 1; q < kPoolPhases; ++q) {
      const float4 a = p4[q * (kPoolSlice / 4)], c = p4[q * (kPoolSlice / 4) + 1];
      s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
      s1.x += c.x; s1.y += c.y; s1.z += c.z; s1.w += c.w;
    }
    g[0] = s0.x / psum + eps; g[1] = s0.y / psum + eps; g[2] = s0.z / psum + eps; g[3] = s0.w / psum + eps;
    g[4] = s1.x / psum + eps; g[5] = s1.y / psum + eps; g[6] = s1.z / psum + eps; g[7] = s1.w / psum + eps;
  }
  if (valid > 0) {
    TO* ob = out + (long long)b * P * C;
    for (int p0 = ph; p0 < P; p0 += R * kPoolPhases) {
      Raw8<TI> raw[R];
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (p0 + kPoolPhases * u < P) ldraw(hb + (long long)(p0 + kPoolPhases * u) * C + c0, vec, vld, raw[u]);
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (p0 + kPoolPhases * u < P) {
          const long long p = p0 + kPoolPhases * u;
          float v[8];
          unraw(raw[u], v);
          st8<TO>(ob + p * C + c0, vec, vld, v);
          st8<TO>(ob + p * C + half + c0, vec, vld, g);
        }
    }
  }
}

int check_scores(const void* scores, int score_dtype, const void* attn, int attn_dtype, int H, const char* what) {
  TOKRED_REQUIRE(scores || attn, "%s: neither scores nor attn given", what);
  if (scores) TOKRED_REQUIRE(valid_float_dtype(score_dtype), "%s: bad score dtype %d", what, score_dtype);
  else TOKRED_REQUIRE(valid_float_dtype(attn_dtype) && H > 0, "%s: bad attn dtype %d / H %d", what, attn_dtype, H);
  return TOKRED_OK;
}

}  // namespace
}  // namespace tokred

using namespace tokred;

static int launch_topk_gather(const char* what, const void* x, int x_dtype, const void* branch, const void* scores,
                              int score_dtype, int64_t score_stride, int64_t score_batch_stride, const void* attn, int attn_dtype,
                              int H, int B, int N, int C, int k, void* x_out, int64_t* idx_out, void* stream) {
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && x_out && idx_out, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(x_dtype), "%s: bad x dtype %d", what, x_dtype);
  TOKRED_REQUIRE(B >= 0 && N >= 2 && C >= 1, "%s: bad shape B=%d N=%d C=%d", what, B, N, C);
  TOKRED_REQUIRE(k >= 1 && k <= N - 1, "%s: k=%d outside [1, %d]", what, k, N - 1);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (int e = check_scores(scores, score_dtype, attn, attn_dtype, H, what)) return e;
  if (B == 0) return TOKRED_OK;
  const int P = N - 1, row_bytes = C * dtype_size(x_dtype);
  TOKRED_REQUIRE((long long)N * row_bytes <= 0x7fffffffLL, "%s: an image of %d x %d bytes exceeds 2 GB (32-bit row offsets)", what, N,
                 row_bytes);
  const int vec16 = (row_bytes % 16 == 0) && aligned16(x) && aligned16(x_out);
  const size_t smem = (size_t)(P + k) * 4;
  if (int e = allow_smem(topk_gather_kernel<true>, smem, what)) return e;
  if (int e = allow_smem(topk_gather_kernel<false>, smem, what)) return e;
  // every split re-ranks the scores, and a warp moves four rows at a time: as few splits as give every SM two CTAs
  int splits = ceil_div(4 * kNumSMs, B);
  splits = max(1, min(splits, ceil_div(k + 1, 2 * kWarps)));
  ScoreSrc ss{scores, score_dtype, score_stride, score_batch_stride, attn, attn_dtype, H};
  if (branch && !(x_dtype == TOKRED_F32 && vec16 && (reinterpret_cast<uintptr_t>(branch) & 7u) == 0)) {
    set_error("%s: the branch add needs an fp32 stream with 16-byte rows and an 8-byte aligned bf16 branch", what);
    return TOKRED_ERR_UNSUPPORTED;
  }
  if (branch)
    topk_gather_kernel<true><<<dim3(splits, B), kThreads, smem, (cudaStream_t)stream>>>(
        ss, (const char*)x, (const char*)branch, (char*)x_out, idx_out, N, k, row_bytes, vec16, dtype_size(x_dtype));
  else
    topk_gather_kernel<false><<<dim3(splits, B), kThreads, smem, (cudaStream_t)stream>>>(
        ss, (const char*)x, nullptr, (char*)x_out, idx_out, N, k, row_bytes, vec16, dtype_size(x_dtype));
  return finish_launch(what);
}

extern "C" int tokred_topk_gather(const void* x, int x_dtype, const void* scores, int score_dtype, int64_t score_stride,
                                  int64_t score_batch_stride, const void* attn, int attn_dtype, int H, int B, int N,
                                  int C, int k, void* x_out, int64_t* idx_out, void* stream) {
  return launch_topk_gather("tokred_topk_gather", x, x_dtype, nullptr, scores, score_dtype, score_stride, score_batch_stride, attn,
                            attn_dtype, H, B, N, C, k, x_out, idx_out, stream);
}

extern "C" int tokred_topk_gather_add(const float* x, const void* branch, const void* scores, int score_dtype,
                                      int64_t score_stride, int64_t score_batch_stride, int B, int N, int C, int k, float* x_out,
                                      int64_t* idx_out, void* stream) {
  const char* what = "tokred_topk_gather_add";
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(branch, "%s: null branch", what);
  return launch_topk_gather(what, x, TOKRED_F32, branch, scores, score_dtype, score_stride, score_batch_stride, nullptr, 0, 0, B,
                            N, C, k, x_out, idx_out, stream);
}

static int launch_evit_select_fuse(const char* what, const void* x, int x_dtype, const void* branch, const void* scores,
                                   int score_dtype, const void* attn, int attn_dtype, int H, int B, int N, int C, int k,
                                   void* x_out, int64_t* idx_out, int64_t* compl_out, void* stream) {
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(x && x_out && idx_out && compl_out, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(x_dtype), "%s: bad x dtype %d", what, x_dtype);
  TOKRED_REQUIRE(B >= 0 && N >= 2 && C >= 1, "%s: bad shape B=%d N=%d C=%d", what, B, N, C);
  TOKRED_REQUIRE(k >= 1 && k <= N - 1, "%s: k=%d outside [1, %d]", what, k, N - 1);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (int e = check_scores(scores, score_dtype, attn, attn_dtype, H, what)) return e;
  if (B == 0) return TOKRED_OK;
  const int P = N - 1, esz = dtype_size(x_dtype), row_bytes = C * esz;
  TOKRED_REQUIRE((long long)N * row_bytes <= 0x7fffffffLL, "%s: an image of %d x %d bytes exceeds 2 GB (32-bit row offsets)", what, N,
                 row_bytes);
  const int vec16 = (row_bytes % 16 == 0) && aligned16(x) && aligned16(x_out);
  const int slice = 1024 / esz;
  const int splits = ceil_div(C, slice);
  const size_t smem = (size_t)(P + k + (P - k) + (P + 3) / 4 + kWarps * slice) * 4;
  ScoreSrc ss{scores, score_dtype, 1, P, attn, attn_dtype, H};
  if (branch && !(x_dtype == TOKRED_F32 && vec16 && (reinterpret_cast<uintptr_t>(branch) & 7u) == 0)) {
    set_error("%s: the branch add needs an fp32 stream with 16-byte rows and an 8-byte aligned bf16 branch", what);
    return TOKRED_ERR_UNSUPPORTED;
  }
  if (x_dtype == TOKRED_F32) {
    if (int e = allow_smem(evit_select_fuse_kernel<float>, smem, what)) return e;
    if (branch) {
      if (int e = allow_smem(evit_select_fuse_kernel<float, true>, smem, what)) return e;
      evit_select_fuse_kernel<float, true><<<dim3(splits, B), kThreads, smem, (cudaStream_t)stream>>>(
          ss, (const float*)x, (const __nv_bfloat16*)branch, (float*)x_out, idx_out, compl_out, N, C, k, vec16);
    } else {
      evit_select_fuse_kernel<float><<<dim3(splits, B), kThreads, smem, (cudaStream_t)stream>>>(
          ss, (const float*)x, nullptr, (float*)x_out, idx_out, compl_out, N, C, k, vec16);
    }
  } else {
    if (int e = allow_smem(evit_select_fuse_kernel<__nv_bfloat16>, smem, what)) return e;
    evit_select_fuse_kernel<__nv_bfloat16><<<dim3(splits, B), kThreads, smem, (cudaStream_t)stream>>>(
        ss, (const __nv_bfloat16*)x, nullptr, (__nv_bfloat16*)x_out, idx_out, compl_out, N, C, k, vec16);
  }
  return finish_launch(what);
}

extern "C" int tokred_evit_select_fuse(const void* x, int x_dtype, const void* scores, int score_dtype,
                                       const void* attn, int attn_dtype, int H, int B, int N, int C, int k,
                                       void* x_out, int64_t* idx_out, int64_t* compl_out, void* stream) {
  return launch_evit_select_fuse("tokred_evit_select_fuse", x, x_dtype, nullptr, scores, score_dtype, attn, attn_dtype, H, B, N, C,
                                 k, x_out, idx_out, compl_out, stream);
}

extern "C" int tokred_evit_select_fuse_add(const float* x, const void* branch, const void* scores, int score_dtype, int B, int N,
                                           int C, int k, float* x_out, int64_t* idx_out, int64_t* compl_out, void* stream) {
  const char* what = "tokred_evit_select_fuse_add";
  if (B == 0) return TOKRED_OK;
  TOKRED_REQUIRE(branch, "%s: null branch", what);
  return launch_evit_select_fuse(what, x, TOKRED_F32, branch, scores, score_dtype, nullptr, 0, 0, B, N, C, k, x_out, idx_out,
                                 compl_out, stream);
}

extern "C" int tokred_gather_rows(const void* src, int dtype, const int64_t* ids, int64_t ids_stride, int B, int G,
                                  int N, int W, int M, void* out, void* stream) {
  const char* what = "tokred_gather_rows";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(src && ids && out, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(dtype), "%s: bad dtype %d", what, dtype);
  TOKRED_REQUIRE(B >= 0 && G >= 1 && N >= 1 && W >= 1 && M >= 0 && ids_stride >= M, "%s: bad shape", what);
  const long long rows = (long long)B * G * M;
  if (rows == 0) return TOKRED_OK;
  const int row_bytes = W * dtype_size(dtype);
  // the 16-byte path addresses the four rows of a group by 32-bit offsets from the first (at most four (b, g) slices apart)
  const int vec16 = (row_bytes % 16 == 0) && aligned16(src) && aligned16(out) && 5LL * N * row_bytes <= 0x7fffffffLL &&
                    4LL * row_bytes <= 0x7fffffffLL;
  long long blocks = (rows + 4 * kWarps - 1) / (4 * kWarps);
  if (blocks > 32LL * kNumSMs) blocks = 32LL * kNumSMs;
  if (vec16)
    gather_rows_kernel<true><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
        (const char*)src, ids, ids_stride, G, N, M, rows, row_bytes, vec16, dtype_size(dtype), (char*)out);
  else
    gather_rows_kernel<false><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
        (const char*)src, ids, ids_stride, G, N, M, rows, row_bytes, vec16, dtype_size(dtype), (char*)out);
  return finish_launch(what);
}

extern "C" int tokred_dyvit_pool_concat(const void* h, int h_dtype, int64_t h_batch_stride, const float* policy, int B, int P,
                                        int C, float eps, void* out, int out_dtype, void* stream) {
  const char* what = "tokred_dyvit_pool_concat";
  if (B == 0) return TOKRED_OK;   // empty batch: nothing to enqueue (tensors may be null)
  TOKRED_REQUIRE(h && policy && out, "%s: null tensor", what);
  TOKRED_REQUIRE(valid_float_dtype(h_dtype) && valid_float_dtype(out_dtype), "%s: bad dtype", what);
  TOKRED_REQUIRE(B >= 0 && P >= 1 && C >= 2 && C % 2 == 0, "%s: bad shape B=%d P=%d C=%d", what, B, P, C);
  TOKRED_REQUIRE(B <= 65535, "%s: B=%d > 65535", what, B);
  if (B == 0) return TOKRED_OK;
  const int splits = ceil_div(C / 2, kPoolSlice);
  const size_t smem = (size_t)(kPoolPhases * kPoolSlice + P) * 4;
  // vector path: 8 channels per thread; fp32 sides move 32 bytes per access (LDG.256 / STG.256: 32-byte alignment)
  const bool al_h = h_dtype == TOKRED_F32 ? (reinterpret_cast<uintptr_t>(h) & 31u) == 0 : aligned16(h);
  const bool al_o = out_dtype == TOKRED_F32 ? (reinterpret_cast<uintptr_t>(out) & 31u) == 0 : aligned16(out);
  TOKRED_REQUIRE(h_batch_stride == 0 || h_batch_stride >= (int64_t)P * C, "%s: h_batch_stride=%lld < P*C", what,
                 (long long)h_batch_stride);
  const long long hbs = h_batch_stride ? (long long)h_batch_stride : (long long)P * C;
  const bool al_s = (hbs * dtype_size(h_dtype)) % (h_dtype == TOKRED_F32 ? 32 : 16) == 0;
  const int vec = ((C / 2) % 8 == 0) && al_h && al_o && al_s;
  dim3 grid(splits, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(TI, TO)                                                                                       \
  do {                                                                                                       \
    if (vec) dyvit_pool_concat_kernel<TI, TO, true><<<grid, kThreads, smem, st>>>((const TI*)h, hbs, policy, (TO*)out, P, C, eps);  \
    else dyvit_pool_concat_kernel<TI, TO, false><<<grid, kThreads, smem, st>>>((const TI*)h, hbs, policy, (TO*)out, P, C, eps);    \
  } while (0)
  if (h_dtype == TOKRED_F32 && out_dtype == TOKRED_F32) LAUNCH(float, float);
  else if (h_dtype == TOKRED_BF16 && out_dtype == TOKRED_F32) LAUNCH(__nv_bfloat16, float);
  else if (h_dtype == TOKRED_BF16 && out_dtype == TOKRED_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else LAUNCH(float, __nv_bfloat16);
#undef LAUNCH
  return finish_launch(what);
}
