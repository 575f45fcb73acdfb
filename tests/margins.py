"""TEST INFRASTRUCTURE — float64 decision margins for the parity protocol of SURVEY.md §8(c).

A reduction operator makes discrete decisions (orderings, arg-reductions) on floating scores.  Two correct fp32
implementations of the same formula differ in the last bits of those scores (summation order, fused vs unfused
multiply-add, exp implementation), so a decision whose margin is within that noise can legitimately differ.  The
functions below re-evaluate an operator's decision inputs in float64, bound the fp32 evaluation error of every
score, and return per image whether ALL of its decisions are separated by more than the bound ("decidable").
The parity tests then require 100 % index equality with the oracle on decidable images and report the excluded
fraction (and assert that it stays small, so the check cannot pass vacuously).

Every function runs on CPU tensors (they move their inputs) and is pure test code: nothing under
tokenreduction_b200/ imports it.
"""
from __future__ import annotations

import math

import torch

U32 = 2.0 ** -24          # fp32 unit roundoff


# ------------------------------------------------------------------------------------------------ orderings
def topk_order_margin(scores: torch.Tensor, k: int, rel: float) -> torch.Tensor:
    """scores [B,P]; decisions = the ORDER of the k largest and the k / k+1 boundary (models/topk.py:62).
    -> [B] bool: every adjacent gap among the k+1 largest exceeds rel * |score|."""
    s = scores.double().cpu().sort(dim=1, descending=True).values[:, : k + 1]
    gap = s[:, :-1] - s[:, 1:]
    return (gap > rel * s[:, :-1].abs().clamp_min(1e-300)).all(dim=1)


def topk_set_margin(scores: torch.Tensor, k: int, rel: float) -> torch.Tensor:
    """only the kept SET matters (k-th vs (k+1)-th largest)."""
    s = scores.double().cpu().sort(dim=1, descending=True).values
    if k >= s.shape[1]:
        return torch.ones(s.shape[0], dtype=torch.bool)
    return (s[:, k - 1] - s[:, k]) > rel * s[:, k - 1].abs().clamp_min(1e-300)


def _per_image(v, b):
    """scalar or [B] bound -> float64 [B,1]."""
    t = torch.as_tensor(v, dtype=torch.float64).cpu().reshape(-1)
    return (t.expand(b) if t.numel() == 1 else t).reshape(b, 1).clone()


# ------------------------------------------------------------------------------------------------ DPC-KNN
def dpcknn_decidable(d_scaled: torch.Tensor, noise_u: torch.Tensor, cluster_num: int, knn: int, eps_d: float = 0.0,
                     eps_diag: float = 0.0):
    """models/dpcknn.py:56-98 re-evaluated in float64 on the fp32 distance matrix ``d_scaled`` (= cdist / sqrt(C)).

    eps_d = bound on the OFF-DIAGONAL |D_impl - d_scaled| when the implementation under test computed its own
    distances (0 when it was handed this very matrix); eps_diag = bound on the self distances of either side (the
    matmul form of cdist leaves up to ~1e-2 of cancellation noise on the diagonal where the exact value is 0, SURVEY
    A.7 -- the self distance is one of the knn neighbours).  Error model of an fp32 evaluation:
      density  rho = exp(-mean(knn smallest d^2)) + 1e-6 U :  |err| <= 4 ulp(1) + 2 rho d_knn eps_d
      parent distance delta = an entry of D                 :  |err| <= eps_d
      score    = delta * rho                                :  |err| <= delta err_rho + rho eps_d + 2 ulp(score)
    Decisions: (1) "j denser than i" for pairs whose flip would change delta_i or delta_j; (2) order of the K largest
    scores and the K / K+1 boundary; (3) nearest centre of every token (exact when eps_d == 0: same D, ties go to
    the lowest k on both sides).
    -> (decidable [B] bool, idx_cluster [B,P], index_down [B,K]) with the float64 decisions."""
    d = d_scaled.double().cpu()
    u = noise_u.double().cpu()
    b, p, _ = d.shape
    has_eps = bool(torch.as_tensor(eps_d).max() > 0)
    eps_d = _per_image(eps_d, b)                                       # [B,1]: a scalar or one bound per image
    eps_diag = _per_image(eps_diag, b)
    near = torch.topk(d, k=knn, dim=-1, largest=False).values
    rho = (-(near ** 2).mean(dim=-1)).exp() + u * float(torch.tensor(1e-6, dtype=torch.float32))
    err_rho = 4 * 2 * U32 + 2.0 * rho * near[..., -1] * eps_d + rho * eps_diag ** 2 / knn
    denser = rho[:, None, :] > rho[:, :, None]                           # [b,i,j]: j denser than i
    d_max = d.flatten(1).max(dim=-1).values
    delta = torch.where(denser, d, d_max[:, None, None].expand_as(d)).min(dim=-1).values
    # (1) ambiguous density pairs that matter
    amb = (rho[:, None, :] - rho[:, :, None]).abs() <= (err_rho[:, None, :] + err_rho[:, :, None])
    amb &= ~torch.eye(p, dtype=torch.bool)[None]
    matters = d <= torch.maximum(delta[:, :, None], delta[:, None, :]) + 2 * eps_d[:, :, None]
    ok = ~(amb & matters).flatten(1).any(dim=1)
    # (2) centre order
    score = delta * rho
    err_s = delta * err_rho + rho * eps_d + 4 * U32 * score
    order = torch.sort(score, dim=-1, descending=True, stable=True).indices
    top = order[:, : cluster_num + 1]
    s_top, e_top = torch.gather(score, 1, top), torch.gather(err_s, 1, top)
    ok &= ((s_top[:, :-1] - s_top[:, 1:]) > (e_top[:, :-1] + e_top[:, 1:])).all(dim=1)
    index_down = order[:, :cluster_num]
    # (3) assignment
    d_c = torch.gather(d, 1, index_down.unsqueeze(-1).expand(-1, -1, p))   # [b,K,p]
    idx_cluster = d_c.argmin(dim=1)
    if has_eps and cluster_num > 1:
        two = torch.topk(d_c, 2, dim=1, largest=False).values
        is_centre = torch.zeros(b, p, dtype=torch.bool).scatter_(1, index_down, True)
        ok &= (((two[:, 1] - two[:, 0]) > 2 * eps_d) | is_centre).all(dim=1)
    idx_cluster.scatter_(1, index_down, torch.arange(cluster_num).expand(b, -1))
    return ok, idx_cluster, index_down


# ------------------------------------------------------------------------------------------------ K-Medoids
def kmedoids_decidable(d: torch.Tensor, token_weight: torch.Tensor, cluster_num: int, iters: int, eps_d: float = 0.0,
                       rel_s: float = 2e-5, rel_w: float = 0.0, eps_diag: float = 0.0):
    """models/kmedoids.py:62-85 re-evaluated in float64 on the fp32 distance matrix ``d`` along the float64 trajectory.

    Scores S_i = w_i sum_j D_ij are sums of P fp32 terms: two summation orders differ by up to ~P u relative
    (rel_s, default 2e-5 > 196 * 2^-24) plus P w_i eps_d.  Decisions: initial centres (top-K of w: exact for a given
    w, rel_w > 0 when w itself was recomputed), per-iteration medoid = argmin of S inside each cluster, nearest-centre
    assignment (exact when eps_d == 0).
    -> (decidable [B] bool, cluster_idx [B,K], assignment [B,P])."""
    d = d.double().cpu()
    w = token_weight.double().cpu().reshape(d.shape[0], -1)
    b, p, _ = d.shape
    has_eps = bool(torch.as_tensor(eps_d).max() > 0)
    eps_d = _per_image(eps_d, b)
    eps_diag = _per_image(eps_diag, b)
    s = w * d.sum(dim=-1)
    err_s = rel_s * s + p * w * eps_d + w * eps_diag
    ok = torch.ones(b, dtype=torch.bool)
    order = torch.sort(w, dim=-1, descending=True, stable=True).indices
    centre = order[:, :cluster_num].clone()
    if rel_w > 0 and cluster_num < p:
        ws = torch.gather(w, 1, order[:, : cluster_num + 1])
        ok &= ((ws[:, :-1] - ws[:, 1:]) > rel_w * ws[:, :-1]).all(dim=1)
    inf = torch.tensor(float("inf"), dtype=torch.float64)

    def assign_of(centre):
        dc = torch.gather(d, 2, centre.unsqueeze(1).expand(-1, p, -1))     # [b,p,K]  D[i, c_k]
        a = dc.argmin(dim=-1)
        good = torch.ones(b, dtype=torch.bool)
        if has_eps and cluster_num > 1:
            two = torch.topk(dc, 2, dim=-1, largest=False).values
            good = ((two[..., 1] - two[..., 0]) > 2 * eps_d).all(dim=1)
        return a, good

    for _ in range(iters):
        assign, good = assign_of(centre)
        ok &= good
        for k in range(cluster_num):
            member = assign == k
            cand = torch.where(member, s, inf)
            lo = cand - torch.where(member, err_s, torch.zeros_like(s))
            best = cand.argmin(dim=1)
            has = member.any(dim=1)
            # the winner's upper bound must stay below every other member's lower bound
            sb, eb = torch.gather(s, 1, best[:, None])[:, 0], torch.gather(err_s, 1, best[:, None])[:, 0]
            lo_others = lo.scatter(1, best[:, None], float("inf")).min(dim=1).values
            ok &= (~has) | (sb + eb < lo_others)
            centre[:, k] = torch.where(has, best, torch.zeros_like(best))      # empty cluster -> token 0 (A.8)
    assign, good = assign_of(centre)
    ok &= good
    return ok, centre, assign


# ------------------------------------------------------------------------------------------------ ToMe
def _bf16_boundary_distance(s: torch.Tensor) -> torch.Tensor:
    """distance of every (float64) value to the nearest bf16 round-to-nearest decision boundary."""
    r = s.float().to(torch.bfloat16).double()                  # nearest bf16
    expo = torch.floor(torch.log2(r.abs().clamp_min(2.0 ** -126)))
    ulp = torch.pow(torch.tensor(2.0, dtype=torch.float64), expo - 7)
    return ulp / 2 - (s - r).abs()


def tome_bf16_decidable(metric: torch.Tensor, r: int, class_token: bool = True, delta: float = 3e-7, rel_op: float = 2.5e-7):
    """bf16-autocast matching (models/tome.py:258-277 with a bf16 matmul).  Two correct implementations can disagree
    on the bf16 value of a similarity for exactly two reasons:
      (a) the fp32 accumulation of the exact bf16 x bf16 products runs in another order (cuBLAS / CPU / tcgen05, whose
          fp32 accumulator truncates): a few fp32 ulps (delta = 3e-7 ~ 5 ulps at |s| <= 0.5) BEFORE the one rounding to
          bf16 -- the entry flips only if its exact value is within delta of a bf16 rounding boundary;
      (b) the normalised operand m / |m| is rounded to bf16 after an fp32 norm whose summation order differs (ATen CPU
          vs ATen CUDA vs the kernel's 8-lane partial sums): the quotient differs by ~2 fp32 ulps (rel_op), which flips
          the bf16 rounding of an operand element sitting on a boundary -> that token's similarities move by up to
          ulp_bf16(element) * |partner element| (measured: 8e-5 on a 0.31 score, image 15 of the B=256 test).
    All exact bf16 ties are resolved identically by both sides (lowest index), so an image is decidable iff no entry
    that can influence a row maximum (bf16 value within one bf16 ulp of its row's maximum) can change its bf16 value
    under (a) + (b).
    -> (decidable [B] bool, scores_bf16 [B,a,b] as float64 with the CLS row at -inf)."""
    m = metric.float().cpu().double()
    x = m / m.norm(dim=-1, keepdim=True)                        # exact (float64) normalised operands
    xb = x.float().to(torch.bfloat16).double()
    two = torch.tensor(2.0, dtype=torch.float64)
    ulp_op = torch.pow(two, torch.floor(torch.log2(xb.abs().clamp_min(2.0 ** -126))) - 7)
    risky_op = (_bf16_boundary_distance(x) < rel_op * x.abs()).double() * ulp_op       # possible operand shift
    a_, b_ = xb[:, ::2], xb[:, 1::2]
    s = a_ @ b_.transpose(1, 2)                                 # exact products, float64 sum = exact similarity
    shift = risky_op[:, ::2] @ b_.abs().transpose(1, 2) + a_.abs() @ risky_op[:, 1::2].transpose(1, 2)
    sb = s.float().to(torch.bfloat16).double()
    unstable = _bf16_boundary_distance(s) < delta + shift
    row_max = sb.max(dim=-1, keepdim=True).values
    ulp = torch.pow(two, torch.floor(torch.log2(row_max.abs().clamp_min(2.0 ** -126))) - 7)
    relevant = sb >= row_max - ulp
    if class_token:
        relevant[:, 0] = False
        sb[:, 0] = -math.inf
    ok = ~(relevant & unstable).flatten(1).any(dim=1)
    return ok, sb


def tome_fp32_margin(metric: torch.Tensor) -> torch.Tensor:
    """fp32 matching: min over (best vs second-best match of every row, adjacent sorted node_max) in float64."""
    m = metric.double().cpu()
    m = m / m.norm(dim=-1, keepdim=True)
    s = m[:, ::2] @ m[:, 1::2].transpose(1, 2)
    top2 = s[:, 1:].topk(2, dim=-1).values
    row_margin = (top2[..., 0] - top2[..., 1]).min(dim=-1).values
    nm = s.max(dim=-1).values[:, 1:].sort(dim=-1).values
    order_margin = (nm[:, 1:] - nm[:, :-1]).min(dim=-1).values
    return torch.minimum(row_margin, order_margin)


# ------------------------------------------------------------------------------------------------ ATS
def ats_decidable(cdf: torch.Tensor, steps: torch.Tensor, tol: float = 2e-6) -> torch.Tensor:
    """models/ats.py:71-75: every step's nearest CDF entry must beat the runner-up by more than tol in d^2
    (cdist's matmul expansion carries ~1e-7 of cancellation noise on d^2).  cdf [B,P] (masked slots already +0.1)."""
    c = cdf.double().cpu()
    d2 = (steps.double().cpu()[None, :, None] - c[:, None, :]) ** 2
    two = torch.topk(d2, 2, dim=-1, largest=False).values
    return ((two[..., 1] - two[..., 0]) > tol).all(dim=1)
