"""TEST INFRASTRUCTURE ONLY — restatement of the reference's token-reduction operators.

Each function restates, in plain PyTorch tensor arithmetic (device-agnostic, normally run on CPU), what
the cited reference lines compute (paths relative to /root/reference).  It is the checker for the CUDA
kernels in tokenreduction_b200/csrc and the body of bench.py's cpu_baseline / ``--impl reference`` legs.
It is pinned against the UNMODIFIED reference (imported through oracle/timm_shim.py) by
tests/test_oracle_vs_reference.py in the build container and against the committed vectors in
tests/golden/ everywhere else.  The reference has no tests/golden vectors of its own.

Conventions
* tokens 0..N-1, CLS = 0; "patch index" p = token-1.
* every arg-reduction / ordering breaks ties toward the LOWEST index (ATen max/min/argmin semantics and
  stable sort); the reference's topk/argsort tie order is unspecified, tests use tie-free inputs.
* ``lowp`` (None | torch.bfloat16): emulates where CUDA autocast rounds a matmul (operands rounded to
  lowp, fp32 accumulate, result rounded to lowp) — SURVEY.md Appendix D.
* accumulation orders follow the CPU reference (sequential scatter_add / index_add_ order).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------ helpers
def _lowp_mm(a: Tensor, b_t: Tensor, lowp) -> Tensor:
    """a @ b_t^T the way an autocast matmul does it: round operands, fp32 accumulate, round result."""
    if lowp is None:
        return a.float() @ b_t.float().transpose(-1, -2)
    out = a.to(lowp).float() @ b_t.to(lowp).float().transpose(-1, -2)
    return out.to(lowp)


def order_desc(scores: Tensor) -> Tensor:
    """indices that sort each row descending, ties -> lowest index first."""
    return torch.sort(scores, dim=-1, descending=True, stable=True).indices


def order_asc(scores: Tensor) -> Tensor:
    return torch.sort(scores, dim=-1, descending=False, stable=True).indices


def gather_rows(x: Tensor, idx: Tensor) -> Tensor:
    """x [B,N,C], idx [B,M] -> [B,M,C]."""
    return torch.gather(x, 1, idx.unsqueeze(-1).expand(-1, -1, x.shape[-1]))


# ------------------------------------------------------------------------------------------ a1 Top-K
def cls_attention_scores(attn: Tensor) -> Tensor:
    """models/topk.py:60-61 / models/evit.py:82-83 — head-mean of the CLS row over patches. attn [B,H,N,N]."""
    return attn[:, :, 0, 1:].mean(dim=1)


def topk_gather(x: Tensor, scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """models/topk.py:62 + :89-93.  x [B,N,C], scores [B,N-1] -> (x_out [B,k+1,C], idx [B,k] int64 desc)."""
    idx = order_desc(scores)[:, :k]
    out = torch.cat([x[:, :1], gather_rows(x[:, 1:], idx)], dim=1)
    return out, idx


# ------------------------------------------------------------------------------------------ a2 EViT
def complement(idx: Tensor, p: int) -> Tensor:
    """models/evit.py:25-46 — ascending patch ids NOT in idx. idx [B,k] -> [B,p-k]."""
    b, k = idx.shape
    keep = torch.zeros(b, p, dtype=torch.bool, device=idx.device)
    keep.scatter_(1, idx, True)
    ar = torch.arange(p, device=idx.device).expand(b, p)
    return ar[~keep].reshape(b, p - k)


def evit_select_fuse(x: Tensor, scores: Tensor, k: int) -> Tuple[Tensor, Tensor, Tensor]:
    """models/evit.py:84 + :111-123.  -> (x_out [B,k+2,C], idx [B,k+1] (last=-1), compl [B,P-k])."""
    idx = order_desc(scores)[:, :k]
    patches = x[:, 1:]
    compl = complement(idx, patches.shape[1])
    w = torch.gather(scores, 1, compl)
    extra = (gather_rows(patches, compl) * w.unsqueeze(-1)).sum(dim=1, keepdim=True)   # unnormalised (quirk B.6)
    out = torch.cat([x[:, :1], gather_rows(patches, idx), extra.to(x.dtype)], dim=1)
    idx_ret = torch.cat([idx, idx.new_full((idx.shape[0], 1), -1)], dim=1)
    return out, idx_ret, compl


# ------------------------------------------------------------------------------------------ a3-a5 ToMe
def tome_effective_r(n_tokens: int, r: int, class_token: bool = True, distill_token: bool = False) -> int:
    """models/tome.py:244-253."""
    protected = int(class_token) + int(distill_token)
    return max(min(r, (n_tokens - protected) // 2), 0)


def tome_match(metric: Tensor, r: int, class_token: bool = True, lowp=None):
    """models/tome.py:258-277.  metric [B,N,D] -> (unm [B,a-r], src [B,r], dst [B,r]) int64, plus node_max.

    a = ceil(N/2) even tokens, b = floor(N/2) odd tokens.  Under CUDA autocast the metric arrives bf16, is
    normalised in fp32 and the a@b^T matmul runs (and is stored) in bf16 -> pass lowp=torch.bfloat16.
    """
    r = tome_effective_r(metric.shape[1], r, class_token, False)
    m = metric.float() if lowp is not None else metric
    m = m / m.norm(dim=-1, keepdim=True)
    a, b = m[:, ::2], m[:, 1::2]
    scores = _lowp_mm(a, b, lowp)
    if class_token:
        scores[:, 0, :] = -math.inf
    node_max, node_idx = scores.max(dim=-1)
    edge = order_desc(node_max.float())
    src = edge[:, :r]
    unm = edge[:, r:]
    if class_token:
        unm = unm.sort(dim=1).values
    dst = torch.gather(node_idx, 1, src)
    return unm, src, dst, node_max


def tome_merge(x: Tensor, size: Optional[Tensor], unm: Tensor, src: Tensor, dst: Tensor):
    """models/tome.py:279-289 (merge) + :309-323 (merge_wavg) + :326-337 and Block_ToMe :91-99 (source map).

    x [B,N,C]; size [B,N,1] or None -> (x_out [B,N-r,C], size_out [B,N-r,1], reduced_cluster_idx [B,N-1] f32).
    The destination rows accumulate their sources in src-list order (CPU scatter_add order).
    """
    b, n, c = x.shape
    if size is None:
        size = torch.ones_like(x[..., :1])

    def push(t: Tensor) -> Tensor:
        ev, od = t[:, ::2], t[:, 1::2]
        w = t.shape[-1]
        kept = torch.gather(ev, 1, unm.unsqueeze(-1).expand(-1, -1, w))
        moved = torch.gather(ev, 1, src.unsqueeze(-1).expand(-1, -1, w))
        od = od.scatter_add(1, dst.unsqueeze(-1).expand(-1, -1, w), moved)
        return torch.cat([kept, od], dim=1)

    xs = push(x * size)
    size_out = push(size)
    x_out = xs / size_out

    # row of the output every input token lands in (the reference derives this from a [B,N,N] identity)
    n_unm = unm.shape[1]
    row_of = torch.empty(b, n, dtype=torch.long, device=x.device)
    odd_rows = n_unm + torch.arange(n // 2, device=x.device)
    row_of[:, 1::2] = odd_rows
    row_of.scatter_(1, 2 * unm, torch.arange(n_unm, device=x.device).expand(b, -1))
    row_of.scatter_(1, 2 * src, n_unm + dst)
    reduced_cluster_idx = (row_of[:, 1:] - 1).to(torch.float32)
    return x_out, size_out, reduced_cluster_idx


# ------------------------------------------------------------------------------------------ distances
def pairwise_dist(x: Tensor) -> Tensor:
    """torch.cdist(x, x) as the reference calls it (models/dpcknn.py:59, models/kmedoids.py:68).

    ATen picks the matmul expansion when P > 25 (d2 = |xi|^2 + |xj|^2 - 2 xi.xj, clamped at 1e-30 before the
    sqrt, so diag(D) is NOT exactly 0 — SURVEY.md A.7) and the direct difference form otherwise.
    """
    p = x.shape[1]
    x = x.float()
    if p > 25:
        sq = x.pow(2).sum(dim=-1)
        d2 = sq.unsqueeze(2) + sq.unsqueeze(1) - 2.0 * (x @ x.transpose(1, 2))
        return d2.clamp_min(1e-30).sqrt()
    diff = x.unsqueeze(2) - x.unsqueeze(1)
    return diff.pow(2).sum(dim=-1).sqrt()


# ------------------------------------------------------------------------------------------ a6-a7 DPC-KNN
def dpcknn_cluster(x: Tensor, cluster_num: int, k: int, noise_u: Tensor, dist: Optional[Tensor] = None,
                   dist_scaled: bool = False):
    """models/dpcknn.py:56-98 (token_mask=None).  x [B,P,C], noise_u [B,P] ~ U(0,1) (the reference draws it
    inside, :73-74; drawing it outside with the same call keeps the generator in step).
    ``dist`` (optional) replaces cdist(x, x); with ``dist_scaled`` it already carries the 1/sqrt(C) of :59.
    -> (idx_cluster [B,P], index_down [B,K]) int64."""
    b, p, c = x.shape
    d = pairwise_dist(x) if dist is None else dist
    if not (dist is not None and dist_scaled):
        d = d / (c ** 0.5)
    near = torch.topk(d, k=k, dim=-1, largest=False).values
    density = (-(near ** 2).mean(dim=-1)).exp() + noise_u * 1e-6
    denser = density[:, None, :] > density[:, :, None]                     # [b,i,j]: j denser than i
    d_max = d.flatten(1).max(dim=-1).values[:, None, None]
    parent = torch.where(denser, d, d_max.expand_as(d)).min(dim=-1).values
    score = parent * density
    index_down = order_desc(score)[:, :cluster_num]
    d_centres = torch.gather(d, 1, index_down.unsqueeze(-1).expand(-1, -1, p))   # [b,K,p]
    idx_cluster = d_centres.argmin(dim=1)
    idx_cluster.scatter_(1, index_down, torch.arange(cluster_num, device=x.device).expand(b, -1))
    return idx_cluster, index_down


def dpcknn_merge(x: Tensor, idx_token: Tensor, agg_weight: Tensor, idx_cluster: Tensor, cluster_num: int,
                 token_weight: Optional[Tensor]):
    """models/dpcknn.py:117-140.  -> (x_merged [B,K,C], idx_token_new [B,T], agg_weight_new [B,T,1])."""
    b, p, c = x.shape
    if token_weight is None:
        token_weight = x.new_ones(b, p, 1)
    flat = (idx_cluster + torch.arange(b, device=x.device)[:, None] * cluster_num).reshape(-1)
    w_sum = token_weight.new_zeros(b * cluster_num, 1).index_add_(0, flat, token_weight.reshape(-1, 1)) + 1e-6
    w_norm = token_weight / w_sum[flat].reshape(b, p, 1)
    merged = x.new_zeros(b * cluster_num, c).index_add_(0, flat, (x * w_norm).reshape(-1, c).to(x.dtype))
    idx_token_new = torch.gather(idx_cluster, 1, idx_token)
    agg_weight_new = agg_weight * torch.gather(w_norm, 1, idx_token.unsqueeze(-1))
    return merged.reshape(b, cluster_num, c), idx_token_new, agg_weight_new


# ------------------------------------------------------------------------------------------ a8 K-Medoids
def attn_colsum(attn: Tensor, num_tokens: int = 1) -> Tensor:
    """models/kmedoids.py:240 — sum over heads then over query rows; patches only. attn [B,H,N,N] -> [B,P,1]."""
    return attn.sum(dim=1).sum(dim=1)[:, num_tokens:].unsqueeze(2)


def kmedoids_fit(x: Tensor, cluster_num: int, iters: int, token_weight: Tensor, dist: Optional[Tensor] = None):
    """models/kmedoids.py:62-85 (token_weight given; the equal_weight path :43-61 is kmedoids_fit_equal below).

    The reference's K x iters loop of masked clones is restated through S_i = sum_j (D_ij * w_i): rows outside
    cluster k are masked to 1e6 in EVERY column, so they sum to P*1e6 (exact in fp32 for P <= 4096) and an
    empty cluster's argmin is index 0.
    -> (centres [B,K,C], cluster_idx [B,K], assignment [B,P])."""
    b, p, c = x.shape
    centre = order_desc(token_weight.squeeze(2))[:, :cluster_num].clone()
    d = pairwise_dist(x) if dist is None else dist
    s = (d * token_weight).sum(dim=-1)                                   # [b,p]
    big = torch.tensor(1.0e6 * p, dtype=s.dtype, device=s.device)         # P copies of 1e6 sum exactly in fp32
    for _ in range(iters):
        assign = torch.gather(d, 2, centre.unsqueeze(1).expand(-1, p, -1)).argmin(dim=-1)   # [b,p]
        for k in range(cluster_num):
            cand = torch.where(assign == k, s, big.expand_as(s))
            centre[:, k] = cand.argmin(dim=1)
    assign = torch.gather(d, 2, centre.unsqueeze(1).expand(-1, p, -1)).argmin(dim=-1)
    return gather_rows(x, centre), centre, assign


def kmedoids_init_equal(x: Tensor, cluster_num: int, first: int, dist: Optional[Tensor] = None) -> Tensor:
    """models/kmedoids.py:43-59 -- the equal_weight initialisation, loop for loop: medoid 0 = ``first`` (the reference's
    one ``np.random.choice`` draw, shared by the batch); then k = 1..K-1: distances of every token to the medoids so far
    (cdist(x, centers), here columns of the full matrix), rows of chosen medoids zeroed, max over the medoid axis, max
    over tokens.  -> cluster_idx [B,K]."""
    b, n, _ = x.shape
    cluster_idx = torch.ones((b, 1), dtype=torch.long, device=x.device) * int(first)
    for k in range(1, cluster_num):
        if dist is None:
            inter = torch.cdist(x, gather_rows(x, cluster_idx))                                 # [b,n,k], the reference's call (:48-49)
        else:
            inter = torch.gather(dist, 2, cluster_idx.unsqueeze(1).expand(-1, n, -1)).clone()   # columns of a given matrix
        for i in range(b):
            for kt in range(k):
                inter[i, cluster_idx[i, kt]] = 0
        max_dist, _ = torch.max(inter, dim=-1)
        _, new = torch.max(max_dist, dim=-1)
        cluster_idx = torch.cat((cluster_idx, new.reshape(b, -1)), dim=-1)
    return cluster_idx


def kmedoids_fit_equal(x: Tensor, cluster_num: int, iters: int, first: int, dist: Optional[Tensor] = None):
    """models/kmedoids.py:40-85 with token_weight=None: kmedoids_init_equal, unit weights (:61), then the iterations of
    kmedoids_fit.  -> (centres, cluster_idx, assignment)."""
    b, p, c = x.shape
    d = pairwise_dist(x) if dist is None else dist
    centre = kmedoids_init_equal(x, cluster_num, first, dist).clone()
    s = (d * x.new_ones(b, p, 1)).sum(dim=-1)
    big = torch.tensor(1.0e6 * p, dtype=s.dtype, device=s.device)
    for _ in range(iters):
        assign = torch.gather(d, 2, centre.unsqueeze(1).expand(-1, p, -1)).argmin(dim=-1)
        for k in range(cluster_num):
            cand = torch.where(assign == k, s, big.expand_as(s))
            centre[:, k] = cand.argmin(dim=1)
    assign = torch.gather(d, 2, centre.unsqueeze(1).expand(-1, p, -1)).argmin(dim=-1)
    return gather_rows(x, centre), centre, assign


# ------------------------------------------------------------------------------------------ a9 Sinkhorn
def sinkhorn_merge(x: Tensor, v: Tensor, eps: float, iters: int, lowp=None):
    """models/sinkhorn.py:66-86 with :25-56.  x [B,P,C], v [K,C] -> (out [B,K,C], weights [B,K,P], v_hat [K,C]).

    Side effect of the reference (the parameter is overwritten by its normalised value, :73-76) is returned
    as v_hat for the caller to apply.  Merges the NORMALISED tokens (quirk B.9)."""
    xh = F.normalize(x.float() if lowp is not None else x, p=2, dim=-1)
    vh = F.normalize(v, p=2, dim=-1)
    k, p = vh.shape[0], xh.shape[1]
    z = _lowp_mm(vh.unsqueeze(0).expand(x.shape[0], -1, -1), xh, lowp) / eps     # [B,K,P]
    norm = -math.log(k + p)
    z32 = z.float()
    u = torch.zeros(x.shape[0], k, dtype=torch.float32, device=x.device)
    w = torch.zeros(x.shape[0], p, dtype=torch.float32, device=x.device)
    for _ in range(iters):
        u = norm - torch.logsumexp(z32 + w.unsqueeze(1), dim=2)
        w = norm - torch.logsumexp(z32 + u.unsqueeze(2), dim=1)
    weights = (z32 + u.unsqueeze(2) + w.unsqueeze(1) - norm).exp()                 # [B,K,P]
    if lowp is None:
        out = weights @ xh
        return out, weights, vh
    out = (weights.to(lowp).float() @ xh.to(lowp).float()).to(lowp)
    return out, weights, vh


# ------------------------------------------------------------------------------------------ a12 PatchMerger
def patchmerger(x: Tensor, ln_weight: Tensor, ln_bias: Tensor, queries: Tensor, scale: float = 1.0,
                ln_eps: float = 1e-5, lowp=None):
    """models/patchmerger.py:35-39.  -> (out [B,K,C], attn [B,K,P]).  Merges the LayerNorm-ed tokens."""
    xn = F.layer_norm(x.float(), (x.shape[-1],), ln_weight.float(), ln_bias.float(), ln_eps)
    sim = _lowp_mm(queries.unsqueeze(0).expand(x.shape[0], -1, -1), xn, lowp) * scale
    attn = sim.float().softmax(dim=-1)
    if lowp is None:
        return attn @ xn, attn
    out = (attn.to(lowp).float() @ xn.to(lowp).float()).to(lowp)
    return out, attn


# ------------------------------------------------------------------------------------------ a13 SiT
def sit_merge(x: Tensor, logits: Tensor, scale: Tensor, lowp=None):
    """models/sit.py:37-40.  x [B,P,C], logits [B,P,K] (= MLP(x), stays on cuBLAS) -> (out [B,K,C], w [B,K,P])."""
    w = F.softmax(logits.float() * scale.float() if lowp is not None else logits * scale, dim=1).transpose(2, 1)
    if lowp is None:
        return torch.bmm(w, x), w
    out = (w.to(lowp).float() @ x.to(lowp).float()).to(lowp)
    return out, w


# ------------------------------------------------------------------------------------------ a10 ATS
def ats_sample_steps(sample_count: int) -> Tensor:
    """models/ats.py:48 — K-1 fp32 steps (2i+1)/(2K); computed with the identical torch call."""
    k = sample_count
    return torch.arange(1 / (2 * k), (2 * k - 1) / (2 * k), 2 / (2 * k))


def ats_significance(v: Tensor, attn: Tensor, eps: float = 1e-6) -> Tensor:
    """models/ats.py:53-66 — normalised significance score [B,P]."""
    cls_attn = attn[:, :, 0, 1:]
    value_norms = v[:, :, 1:, :].float().norm(dim=-1)
    sig = (cls_attn * value_norms).sum(dim=1)
    return sig / (sig.sum(dim=-1, keepdim=True) + eps)


def _ats_dist(steps: Tensor, cdf: Tensor) -> Tensor:
    """models/ats.py:73 — torch.cdist on 1-d points; ATen uses the matmul expansion when either side has > 25
    points: sqrt(clamp_min(s^2 + c^2 - 2sc, 1e-30)), NOT |s - c|."""
    q, p = steps.shape[0], cdf.shape[1]
    s, c = steps[None, :, None], cdf[:, None, :]
    if q > 25 or p > 25:
        d2 = (-2.0 * s) * c + (s * s) + (c * c)
        return d2.clamp_min(1e-30).sqrt()
    return (s - c).abs()


def ats_sample(v: Tensor, attn: Tensor, mask: Tensor, sample_count: int, eps: float = 1e-6, pad_to: Optional[int] = None):
    """models/ats.py:52-89.  v [B,H,N,Dh], attn [B,H,N,N], mask [B,N] bool
    -> (new_attn [B,H,M+1,N], new_mask [B,M+1], ids [B,M+1] int64), M = max_b #unique (or pad_to-1 if given)."""
    b, h, n = attn.shape[:3]
    cdf = ats_significance(v, attn, eps).cumsum(dim=1)
    cdf = torch.where(mask[:, 1:], cdf, cdf + 0.1)
    steps = ats_sample_steps(sample_count).to(cdf.device)
    ids = _ats_dist(steps, cdf).argmin(dim=-1) + 1                                  # [B,K-1]
    hit = torch.zeros(b, n, dtype=torch.bool, device=attn.device)
    hit.scatter_(1, ids, True)
    count = hit.sum(dim=1)
    m = int(count.max()) if pad_to is None else pad_to - 1
    order = torch.sort(torch.where(hit, torch.arange(n, device=attn.device).expand(b, -1), n), dim=1).values[:, :m]
    uniq = torch.where(order < n, order, 0)
    new_mask = F.pad(uniq != 0, (1, 0), value=True)
    ids_out = F.pad(uniq, (1, 0), value=0)
    new_attn = torch.gather(attn, 2, ids_out[:, None, :, None].expand(-1, h, -1, n))
    return new_attn, new_mask, ids_out


# ------------------------------------------------------------------------------------------ a11 DynamicViT
def dyvit_pool_concat(h: Tensor, policy: Tensor, eps: float = 1e-6) -> Tensor:
    """models/dyvit.py:114-118 — [local half | masked mean of the global half (+eps on the quotient)]."""
    b, p, c = h.shape
    local = h[:, :, : c // 2]
    glob = (h[:, :, c // 2:] * policy).sum(dim=1, keepdim=True) / policy.sum(dim=1, keepdim=True) + eps
    return torch.cat([local, glob.expand(b, p, c // 2)], dim=-1)


def dyvit_keep(x: Tensor, score: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """models/dyvit.py:231-236 + :340-356 — same selection/gather as Top-K with predictor scores."""
    return topk_gather(x, score, k)


# ----------------------------------------------------------------------------------------------- f1: attention producer
def attention_autocast(qkv: Tensor, num_heads: int, scale: float, key_bias: Optional[Tensor] = None,
                       mask: Optional[Tensor] = None, q_ids: Optional[Tensor] = None):
    """The reference's attention under CUDA bf16 autocast with every dtype written out, device
    independent: models/topk.py:44-52 (qkv split, q k^T * scale, softmax, attn @ v, transpose/reshape), tome.py:48-49
    (+ log size: pass key_bias = size.log()), ats.py:118-121 (masked_fill with -finfo.max), ats.py:84-87 (row gather).
    autocast runs the two matmuls in bf16 (fp32 accumulation, bf16 result), ``* scale`` on the bf16 tensor, the softmax
    in fp32, and casts the fp32 probabilities to bf16 for attn @ v.
    qkv [B,N,3*H*Dh] bf16 -> (out [B,M,H*Dh] bf16, attn [B,H,M,N] fp32 probabilities, attn_full_cls [B,H,N] fp32)."""
    b, n, c3 = qkv.shape
    c = c3 // 3
    qkv = qkv.to(torch.bfloat16)
    q, k, v = qkv.reshape(b, n, 3, num_heads, c // num_heads).permute(2, 0, 3, 1, 4)
    dots = (q @ k.transpose(-2, -1)) * scale                            # bf16 matmul, bf16 * python float -> bf16
    if key_bias is not None:
        dots = dots + key_bias.float()[:, None, None, :]                # bf16 + fp32 -> fp32 (tome.py:49)
    if mask is not None:
        m2 = mask.unsqueeze(1).unsqueeze(3) * mask.unsqueeze(1).unsqueeze(2)
        dots = dots.masked_fill(~m2, -torch.finfo(dots.dtype).max)     # ats.py:118-121
    attn = dots.float().softmax(dim=-1)                                 # autocast: softmax in fp32
    cls = attn[:, :, 0, :].clone()
    if q_ids is not None:                                               # ats.py:84-87
        attn = torch.gather(attn, 2, q_ids[:, None, :, None].expand(b, num_heads, q_ids.shape[1], n))
    out = (attn.to(torch.bfloat16) @ v).transpose(1, 2).reshape(b, attn.shape[2], c)
    return out, attn, cls
