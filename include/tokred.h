/* libtokred_sm100a.so — C ABI of the B200 token-reduction operators.
 *
 * The reference (JoakimHaurum/TokenReduction) has no FFI layer: its reduction operators are Python
 * functions/modules calling ATen.  Each entry point below replaces the ATen call sequence of the cited
 * reference lines (paths relative to the reference root) with one hand-written sm_100a kernel launch.
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions (all entry points)
 *   - every pointer is a DEVICE pointer to a contiguous row-major tensor allocated by the caller;
 *     the library never allocates, never synchronises, never copies to the host: it only enqueues
 *     kernels on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - floating tensors carry a dtype tag: TOKRED_F32 or TOKRED_BF16.  Index tensors are int64 (torch.long),
 *     masks are uint8 (torch.bool).
 *   - return value: 0 = enqueued; <0 = argument/shape violation detected on the host before any launch
 *     (TOKRED_ERR_*); >0 = cudaError_t reported by the launch.  tokred_last_error() returns a thread-local
 *     human-readable message for the last non-zero return on the calling thread.
 *   - re-entrant and thread-safe; no global mutable state besides per-kernel shared-memory opt-in attributes.
 *   - ordering ties are always broken toward the LOWEST index (ATen max/min/argmin semantics, stable sort).
 *   - tokens 0..N-1 with CLS = 0; P = N-1 patches; "patch index" p = token-1.
 */
#ifndef TOKRED_H_
#define TOKRED_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOKRED_ABI_VERSION 4   /* 2: x_batch_stride on a6-a9, a12, a13; 3: tokred_attention (f1); 4: h_batch_stride on
                                   tokred_dyvit_pool_concat, the fused / deferred entry points (tome_merge_ln, *_add, patchify,
                                   embed_layernorm, residual_add, kmedoids_fit_init) */
#define TOKRED_API __attribute__((visibility("default")))

enum { TOKRED_F32 = 0, TOKRED_BF16 = 1 };
enum { TOKRED_OK = 0, TOKRED_ERR_ARGUMENT = -1, TOKRED_ERR_UNSUPPORTED = -2 };

TOKRED_API int tokred_abi_version(void);
TOKRED_API const char* tokred_last_error(void);
/* Number of kernel launches enqueued by this library in the calling process (for bench.py's gpu_launches). */
TOKRED_API uint64_t tokred_launch_count(void);

/* ---- a1 Top-K / a11 DynamicViT keep -----------------------------------------------------------------
 * models/topk.py:55-65 (cls_attn mean + torch.topk) and :89-93 (gather + cat);
 * models/dyvit.py:231-236 + :340-356 (argsort(score)[:k], batch_index_select).
 *   x        [B,N,C]  x_dtype
 *   scores   [B,P] fp32 or bf16 with element stride `score_stride` between consecutive patches and
 *            `score_batch_stride` between images (DynamicViT passes pred_score[:,:,0], stride 2); or NULL
 *   attn     [B,H,N,N] attn_dtype, used only when scores == NULL: s_p = (sum_h attn[b,h,0,1+p]) * (1/H)
 *   x_out    [B,k+1,C] x_dtype : row 0 = CLS, row 1+j = x[1+idx[j]]
 *   idx_out  [B,k] int64, patch indices in descending score order                                      */
TOKRED_API int tokred_topk_gather(const void* x, int x_dtype, const void* scores, int score_dtype, int64_t score_stride,
                       int64_t score_batch_stride, const void* attn, int attn_dtype, int H, int B, int N, int C,
                       int k, void* x_out, int64_t* idx_out, void* stream);

/* ---- a2 EViT -----------------------------------------------------------------------------------------
 * models/evit.py:77-87 (top-k), :25-46 (complement_idx), :111-123 (gather, inattentive-token fusion, idx -1).
 *   x_out     [B,k+2,C]: CLS, kept tokens (descending score), fused token sum_{p in compl} s_p * x[1+p]
 *   idx_out   [B,k+1] int64 (last column = -1)
 *   compl_out [B,P-k] int64 ascending                                                                  */
TOKRED_API int tokred_evit_select_fuse(const void* x, int x_dtype, const void* scores, int score_dtype, const void* attn,
                            int attn_dtype, int H, int B, int N, int C, int k, void* x_out, int64_t* idx_out,
                            int64_t* compl_out, void* stream);

/* ---- a3 ToMe matching --------------------------------------------------------------------------------
 * models/tome.py:230-277 bipartite_soft_matching (cosine similarity of even vs odd tokens, per-row argmax,
 * descending edge order, CLS protected).  r is clamped to (N - protected)/2 like :252-253; the caller sizes
 * the outputs with tokred_tome_effective_r().
 *   metric [B,N,D] metric_dtype (= k.mean(1), models/tome.py:58); or, heads > 1: the per-head keys themselves,
 *   metric[b,n,h,:] at ((b*N+n)*token_stride + h*D) -- the k slice of the qkv Linear's output (tome.py:33-35) -- and the
 *   head mean of :58 is taken in-kernel (fp32 sum * 1/heads, rounded to bf16 like ATen's mean); bf16, D = 64,
 *   score_lowp = 1 only.  heads <= 1 and token_stride = 0: plain dense metric.
 *   NaN metric rows (zero norm) order like ATen: NaN is the largest key, so all index slots are always written.
 *   score_lowp: 0 = fp32 similarity (FFMA); 1 = operands and similarity rounded to bf16 (what CUDA autocast
 *   does) on tcgen05 tensor cores; 3 = same rounding on the FFMA path (cross-check of the tensor-core path)
 *   class_token: protection flags -- bit 0 = class token (even token 0 never merges away, :263-264), bit 1 =
 *   distillation token (odd token 0 never receives a merge, :265-266); 0/1 keep their plain boolean meaning.
 *   The distilled output row order of :286-287 is a fixed permutation applied by the caller.
 *   unm_idx [B,a-r] (ascending when class_token), src_idx [B,r], dst_idx [B,r] int64; a = ceil(N/2)      */
TOKRED_API int tokred_tome_effective_r(int N, int r, int class_token);
TOKRED_API int tokred_tome_match(const void* metric, int metric_dtype, int heads, int64_t token_stride, int B, int N, int D,
                      int r, int class_token, int score_lowp, int64_t* unm_idx, int64_t* src_idx, int64_t* dst_idx,
                      void* stream);

/* ---- a4/a5 ToMe merge --------------------------------------------------------------------------------
 * models/tome.py:279-289 (merge closure), :309-323 (merge_wavg), :326-337 + Block_ToMe :91-99 (source map).
 *   x [B,N,C], size [B,N] (same dtype as x) or NULL (= ones)
 *   x_out [B,N-r,C] = [unmerged even tokens ; odd tokens + merged sources] size-weighted mean
 *   size_out [B,N-r]; reduced_cluster_idx [B,N-1] fp32 or NULL (output row of every input patch, minus 1)
 *   divide = 1: merge_wavg (weighted mean); divide = 0: the bare merge closure (sums of x*size, no division)
 * Sources are accumulated in src-list order (the CPU scatter_add order): deterministic, no atomics.      */
TOKRED_API int tokred_tome_merge(const void* x, int x_dtype, const void* size, const int64_t* unm_idx,
                      const int64_t* src_idx, const int64_t* dst_idx, int B, int N, int C, int r, void* x_out,
                      void* size_out, float* reduced_cluster_idx, int divide, void* stream);

/* ---- a1 / a2 with the block's residual add fused (bf16-autocast Block_TopK / Block_EVIT) --------------------
 * `x = x + drop_path(attn branch)` (models/topk.py:87, models/evit.py:109) is formed on the fly, as the same fp32 addition,
 * for the rows that are read: x [B,N,C] fp32, branch [B,N,C] bf16 (the attention projection's output).  Outputs as
 * tokred_topk_gather / tokred_evit_select_fuse computed on x + branch; rows that are dropped are never added (Top-K).
 * Rows must be whole 16-byte units (C % 4 == 0), scores given explicitly.                                       */
TOKRED_API int tokred_topk_gather_add(const float* x, const void* branch, const void* scores, int score_dtype,
                           int64_t score_stride, int64_t score_batch_stride, int B, int N, int C, int k, float* x_out,
                           int64_t* idx_out, void* stream);
TOKRED_API int tokred_evit_select_fuse_add(const float* x, const void* branch, const void* scores, int score_dtype, int B, int N,
                                int C, int k, float* x_out, int64_t* idx_out, int64_t* compl_out, void* stream);

/* ---- a4/a5 with the callers either side fused (bf16-autocast Block_ToMe, models/tome.py:88-104) ----------
 * x + drop_path(attn branch)  ->  merge_wavg  ->  norm2  in ONE launch over the fp32 residual stream:
 *   x [B,N,C] fp32; branch [B,N,C] bf16 or NULL (the attention projection's output; added in fp32 on every row read);
 *   size [B,N] fp32 or NULL; index lists as tokred_tome_merge; gamma, beta [C] fp32, eps: norm2's parameters
 *   x_out [B,N-r,C] fp32 = merge_wavg(x + branch), size_out [B,N-r] fp32, reduced_cluster_idx [B,N-1] or NULL,
 *   y [B,N-r,C] bf16 = LayerNorm(x_out) rounded once (what the MLP's first autocast Linear consumes).
 * Bit-identical to add -> tokred_tome_merge -> tokred_add_layernorm.  C must be a multiple of 128 up to 768.      */
TOKRED_API int tokred_tome_merge_ln(const float* x, const void* branch, const float* size, const int64_t* unm_idx,
                         const int64_t* src_idx, const int64_t* dst_idx, int B, int N, int C, int r, const float* gamma,
                         const float* beta, float eps, float* x_out, float* size_out, float* reduced_cluster_idx, void* y,
                         void* stream);

/* ---- pairwise distances (building block of a6 / a8, exported for parity checks) ------------------------
 * torch.cdist(x, x) as called at models/dpcknn.py:59 and models/kmedoids.py:68, times post_scale:
 * matmul expansion sqrt(max(|xi|^2+|xj|^2-2xi.xj, 1e-30)) for P > 25, direct differences otherwise.
 *   x [B,P,C] fp32 -> out [B,P,P] fp32 (bit-symmetric).
 *   exact_fp32 = 0: Gram on tcgen05 tensor cores with 3xTF32 error compensation (fp32-matmul accuracy class);
 *   exact_fp32 = 1: FFMA (true fp32 products).  Same switch on dpcknn_cluster / kmedoids_fit.            */
TOKRED_API int tokred_pairwise_dist(const float* x, int B, int P, int C, float post_scale, int exact_fp32, float* out,
                                    void* stream);

/* x_batch_stride (a6-a9, a12, a13): elements between consecutive images of x; 0 = dense (P*C).  The reference hands
 * these operators x[:, 1:] (class token dropped: models/dpcknn.py:259, kmedoids.py:241, sinkhorn.py:167, patchmerger.py:118, sit.py:118): rows stay
 * C apart but images are (P+1)*C apart, and the stride lets the kernels read that view in place instead of a copy. */

/* ---- a6 DPC-KNN clustering ---------------------------------------------------------------------------
 * models/dpcknn.py:44-100 cluster_dpc_knn (token_mask=None).
 *   x [B,P,C] fp32; noise_u [B,P] fp32 ~ U(0,1) drawn by the caller with the reference's torch.rand call
 *   idx_cluster [B,P] int64, index_down [B,K] int64 (descending centre score)                           */
TOKRED_API int tokred_dpcknn_cluster(const float* x, int64_t x_batch_stride, const float* noise_u, int B, int P, int C, int K, int knn,
                                     int exact_fp32, int64_t* idx_cluster, int64_t* index_down, void* stream);

/* ---- a7 DPC-KNN merge --------------------------------------------------------------------------------
 * models/dpcknn.py:103-140 merge_tokens.  token_weight [B,P] fp32 or NULL (= ones); idx_token [B,T] int64;
 * agg_weight [B,T] fp32 -> x_merged [B,K,C], idx_token_new [B,T], agg_weight_new [B,T].                 */
TOKRED_API int tokred_dpcknn_merge(const float* x, int64_t x_batch_stride, const int64_t* idx_token, const float* agg_weight,
                        const int64_t* idx_cluster, const float* token_weight, int B, int P, int C, int K, int T,
                        float* x_merged, int64_t* idx_token_new, float* agg_weight_new, void* stream);

/* ---- a8 K-Medoids ------------------------------------------------------------------------------------
 * models/kmedoids.py:240 token weights: out[b,p] = sum_h sum_q attn[b,h,q,num_tokens+p]                 */
TOKRED_API int tokred_attn_colsum(const void* attn, int attn_dtype, int B, int H, int N, int num_tokens, float* out,
                       void* stream);
/* models/kmedoids.py:62-85 k_medoids_fit with token weights (topk init, iters x {assign, re-centre}).
 *   x [B,P,C] fp32, token_weight [B,P] fp32 -> centres [B,K,C], cluster_idx [B,K] int64, assignment [B,P] int64 */
TOKRED_API int tokred_kmedoids_fit(const float* x, int64_t x_batch_stride, const float* token_weight, int B, int P, int C, int K, int iters,
                                   int exact_fp32, float* centres, int64_t* cluster_idx, int64_t* assignment, void* stream);

/* The same with caller-supplied initial medoids init_idx [B,K] int64 (models/kmedoids.py:43-61: the equal_weight variant's
 * farthest-point initialisation, built by the caller from tokred_pairwise_dist); token_weight may then be NULL (= ones, :61). */
TOKRED_API int tokred_kmedoids_fit_init(const float* x, int64_t x_batch_stride, const float* token_weight, const int64_t* init_idx,
                                        int B, int P, int C, int K, int iters, int exact_fp32, float* centres,
                                        int64_t* cluster_idx, int64_t* assignment, void* stream);

/* Scratch for the bulk-copy fed tensor-core path of the three soft merges (a9, a12, a13): the caller passes a device
 * buffer of at least this many bytes, 128-byte aligned (bf16 token tiles + packed Q).  With workspace = NULL the
 * entry points use the scratch-free tensor-core kernel instead (slower).  The library never allocates.            */
TOKRED_API size_t tokred_soft_merge_workspace_bytes(int B, int P, int C, int K);
/* Shape limits of a9 / a12 / a13 (one image's score matrix lives in shared memory): P <= 208 and K <= 208 always
 * (TOKRED_ERR_UNSUPPORTED above); the bf16-autocast tensor-core kernels (lowp = 1, bf16 out) cover that whole range
 * for C <= 1024 (C % 8 == 0 with a workspace); the fp32 / FFMA kernel keeps the fp32 K x P matrix and needs
 * 4*(K*ceil4(P) + max(36*(K+P), tile) + 3P + K) bytes <= 227 KB (softmerge.cu soft_smem_bytes) -- it covers every shape of the reference
 * (P <= 196, K <= 176) and answers with an argument error naming the shortfall beyond that (e.g. P = K = 208).   */

/* ---- a9 Sinkhorn -------------------------------------------------------------------------------------
 * models/sinkhorn.py:66-86 with :25-56.  v_hat [K,C] fp32 is the already-normalised parameter.
 *   log_norm = -log(K+P) evaluated by the caller the way the reference does (:44-47; in bf16 under autocast)
 *   lowp = 1: both contractions round operands/result to bf16 (CUDA autocast) and run on tcgen05 tensor cores
 *   when out_dtype is bf16; lowp = 3: same rounding on the FFMA path (cross-check); lowp = 0: exact fp32 (FFMA).
 *   out [B,K,C], weights [B,K,P] fp32                                                                   */
TOKRED_API int tokred_sinkhorn_merge(const void* x, int x_dtype, int64_t x_batch_stride, const float* v_hat, int B, int P, int C, int K, float eps,
                          float log_norm, int iters, int lowp, void* out, int out_dtype, float* weights,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- a12 PatchMerger ---------------------------------------------------------------------------------
 * models/patchmerger.py:35-39: LayerNorm -> queries x^T * scale -> softmax over tokens -> attn x.       */
TOKRED_API int tokred_patchmerger(const void* x, int x_dtype, int64_t x_batch_stride, const float* ln_weight, const float* ln_bias,
                       const float* queries, int B, int P, int C, int K, float scale, float ln_eps, int lowp,
                       void* out, int out_dtype, float* attn, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a13 SiT -----------------------------------------------------------------------------------------
 * models/sit.py:37-40: w = softmax(logits * scale, over tokens)^T ; out = w x.  logits [B,P,K];
 * scale = device pointer to the module's 1-element fp32 parameter (no host read).                       */
TOKRED_API int tokred_sit_merge(const void* x, int x_dtype, int64_t x_batch_stride, const void* logits, int logits_dtype, const float* scale, int B,
                     int P, int C, int K, int lowp, void* out, int out_dtype, float* weights, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- a10 ATS -----------------------------------------------------------------------------------------
 * models/ats.py:52-82: significance score, inverse-CDF sampling, per-image sorted unique ids.
 *   v [B,H,N,Dh] v_dtype with element strides (v_stride_b, v_stride_h, v_stride_n, 1) — the qkv view of
 *   :112-113 is consumed in place; attn fp32: only the CLS rows attn[b,h,0,:] are read, at b*H*attn_head_stride +
 *   h*attn_head_stride (N*N, or 0, for a [B,H,N,N] tensor; N for the [B,H,N] cls_row output of tokred_attention);
 *   mask [B,N] uint8, steps [n_steps] fp32 (= sample_steps, :48)
 *   ids_out  [B,n_steps+1] int64: 0, sorted unique sampled tokens, 0-padding
 *   mask_out [B,n_steps+1] uint8: 1, ids != 0
 *   max_count [1] int32: max_b #unique (caller zeroes it before the call; device atomicMax)             */
TOKRED_API int tokred_ats_sample(const void* v, int v_dtype, int64_t v_stride_b, int64_t v_stride_h, int64_t v_stride_n,
                      const float* attn, int64_t attn_head_stride, const uint8_t* mask, const float* steps, int B, int H,
                      int N, int Dh, int n_steps, float eps, int64_t* ids_out, uint8_t* mask_out, int32_t* max_count,
                      void* stream);

/* models/ats.py:27-41,84-87 (attention rows) and :156-157 (residual tokens): out[b,g,m,:] = src[b,g,ids[b,m],:]
 *   src [B,G,N,W] dtype, ids [B,ids_stride] int64 (first M used) -> out [B,G,M,W]                         */
TOKRED_API int tokred_gather_rows(const void* src, int dtype, const int64_t* ids, int64_t ids_stride, int B, int G, int N,
                       int W, int M, void* out, void* stream);

/* ---- a11 DynamicViT predictor pooling ----------------------------------------------------------------
 * models/dyvit.py:114-118: out = [h[:,:,:C/2] | (sum_p h[:,p,C/2:]*policy[p]) / sum_p policy[p] + eps].
 *   h [B,P,C] h_dtype (bf16 under autocast), images h_batch_stride elements apart (0 = dense P*C; the [:, 1:] view of a
 *   [B,P+1,C] tensor is read in place), policy [B,P] fp32 -> out [B,P,C] out_dtype (fp32: the reference's cat promotes;
 *   bf16: what the autocast Linear consuming it, :119, casts that to -- the same bits)                   */
TOKRED_API int tokred_dyvit_pool_concat(const void* h, int h_dtype, int64_t h_batch_stride, const float* policy, int B, int P,
                             int C, float eps, void* out, int out_dtype, void* stream);

/* ---- f1 / f2: attention that emits only what the reduction operators read ---------------------------
 * models/topk.py:44-52,59-61; evit.py:66-87; tome.py:44-58 (proportional attention :48-49); kmedoids.py:105-112,240;
 * ats.py:115-127,84-87; dyvit.py:53-69 (eval branch); the stock timm block.  bf16-autocast semantics: S = q k^T rounded
 * to bf16, * scale rounded to bf16, (+ key_bias in fp32), masked_fill, softmax in fp32, probabilities rounded to bf16,
 * out = P v rounded to bf16.  The [B,H,N,N] probabilities are never written.
 *   qkv      [B,N,3,H,head_dim] bf16 -- the qkv Linear's output, consumed in place (topk.py:45)
 *   key_bias [B,N] fp32 or NULL: added to every row of the scaled logits (ToMe: log(size), tome.py:48-49)
 *   mask     [B,N] uint8 or NULL (ATS, ats.py:118-121): logits with mask[i]*mask[j] == 0 are filled with -max
 *            (a masked query row therefore comes out uniform, as in the reference)
 *   q_ids    [B,ids_stride] int64 or NULL: query row m of the output is token q_ids[b,m] (first M used) -- the ATS
 *            row gather attn[:, :, ids, :] (ats.py:84-87) applied to the queries, so only M rows are computed
 *   out      [B,M,H*head_dim] bf16 = (attn @ v).transpose(1,2).reshape(B,M,C) (topk.py:51), M = N without q_ids;
 *            NULL = scores only (v is not read, P.v is skipped)
 *   cls_row  [B,H,N] fp32 or NULL: probabilities of query row 0 -- attn[b,h,0,:] (topk.py:60, ats.py:57), fp32
 *            values before the bf16 rounding
 *   colsum   [B,H,N] fp32 or NULL: sum over the query rows of attn[b,h,:,j] (K-Medoids token weights are the sum of
 *            this over heads, kmedoids.py:240); combined in a fixed order: deterministic
 * head_dim must be 64 and N, M <= 256 (TOKRED_ERR_UNSUPPORTED otherwise: the caller keeps its ATen sequence).  */
TOKRED_API int tokred_attention(const void* qkv, int B, int N, int H, int head_dim, float scale, const float* key_bias,
                     const uint8_t* mask, const int64_t* q_ids, int64_t ids_stride, int M, void* out, float* cls_row,
                     float* colsum, void* stream);

/* ---- callers either side of the attention producer: residual add + LayerNorm + bf16 cast in one pass ----------
 * e.g. models/topk.py:87 (x = x + drop_path(tmp)) + :94 (norm2(x)) + the autocast cast in front of the next Linear.
 *   x       [rows,C] fp32 residual stream
 *   branch  [rows,C] branch_dtype (bf16 under autocast) or NULL: x_out = x + branch (fp32); NULL: plain LayerNorm of x
 *   gamma, beta [C] fp32, eps: the LayerNorm's affine parameters
 *   x_out   [rows,C] fp32 (required with branch; may alias x), y [rows,C] bf16 = LayerNorm(x_out) rounded once
 * C must be a multiple of 128 up to 1024 (TOKRED_ERR_UNSUPPORTED otherwise).                               */
TOKRED_API int tokred_add_layernorm(const float* x, const void* branch, int branch_dtype, const float* gamma, const float* beta,
                         float eps, int64_t rows, int C, float* x_out, void* y, void* stream);

/* out[i] = x[i] + float(branch[i]), i < n (n % 4 == 0): the fp32 residual sum alone, for a consumer that is not a LayerNorm
 * (x fp32, branch bf16; the same fp32 addition as ATen's mixed-dtype add, e.g. models/dpcknn.py:258 reading x after :104).  */
TOKRED_API int tokred_residual_add(const float* x, const void* branch, int64_t n, float* out, void* stream);

/* ---- the data formats in front of the first block (bf16 autocast; models/deit_viz.py PatchEmbed + forward_features) ----
 * tokred_patchify: image [B,Cin,H,W] fp32 -> out [B,(H/ph)*(W/pw),Cin*ph*pw] bf16 (round-to-nearest-even), row element
 *   (c*ph+py)*pw+px of patch (gy,gx) = img[b,c,gy*ph+py,gx*pw+px]: the operand of the patch-embedding GEMM (the stride-p
 *   convolution with its weight viewed as [C,Cin*ph*pw]); replaces ATen's cast + permuting copy.  pw % 4 == 0.
 * tokred_embed_layernorm: x_out [B,T+P,C] fp32 = cat(tokens, patches [B,P,C] bf16 | fp32) (+ pos [T+P,C] unless NULL)
 *   and y [B,T+P,C] bf16 = LayerNorm(x_out; gamma, beta, eps) rounded once (the next block's norm1 as its qkv Linear consumes
 *   it).  tokens [T,C] fp32 shared by the batch (tokens_batch_stride = 0: the cls / dist parameters, forward_features) or
 *   per image, tokens_batch_stride elements apart (the class rows x[:, :T] of a [B,N,C] stream read in place: the
 *   re-concatenation after a cluster layer, e.g. models/sinkhorn.py:168, models/dpcknn.py:262).  C a multiple of 128 <= 1024. */
TOKRED_API int tokred_patchify(const float* img, int B, int Cin, int H, int W, int ph, int pw, void* out, void* stream);
TOKRED_API int tokred_embed_layernorm(const void* patches, int patch_dtype, const float* tokens, int64_t tokens_batch_stride,
                           const float* pos, const float* gamma, const float* beta, float eps, int B, int P, int T, int C,
                           float* x_out, void* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOKRED_H_ */
