"""``torch.ops.tokred.*`` — the reduction operators as torch.library custom ops over the C-ABI library.

Every op: checks arguments on the host, allocates its outputs with torch (caching allocator, current device),
and enqueues exactly one hand-written sm_100a kernel on the current CUDA stream through ctypes.  No op has a
CPU implementation: calling one with CPU tensors raises.  ``register_fake`` gives shape/dtype inference so
CPU-only CI can trace graphs with FakeTensors.  All ops are inference-only (the reference runs its
matching / clustering under no_grad: models/tome.py:258, models/dpcknn.py:56).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import BF16, F32, TokredError

__all__ = [
    "topk_gather", "topk_gather_attn", "evit_select_fuse", "evit_select_fuse_attn", "tome_effective_r", "tome_match",
    "tome_match_qkv", "tome_merge", "pairwise_dist", "dpcknn_cluster", "dpcknn_merge", "attn_colsum", "kmedoids_fit", "sinkhorn_merge", "patchmerger",
    "sit_merge", "ats_sample", "gather_rows", "dyvit_pool_concat", "attention", "attention_supported", "add_layernorm",
]


# ----------------------------------------------------------------------------------------------- plumbing
def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TokredError(f"unsupported dtype {t.dtype}: tokred ops take float32 or bfloat16")


def _tdt(code: int) -> torch.dtype:
    return torch.float32 if code == F32 else torch.bfloat16


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(name: str, *tensors: Optional[Tensor]) -> None:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise TokredError(f"{name}: tensor on {t.device}; tokred ops run on CUDA only (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise TokredError(f"{name}: tensors on different devices ({dev} vs {t.device})")
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        raise TokredError(f"{name}: tensors on {dev} but current device is cuda:{torch.cuda.current_device()}")


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _up(*ts):
    """fp16 boundary (the reference's eval default is fp16 autocast, validate.py:52-54): half tensors enter the kernels as
    fp32 -- an exact conversion -- everything else passes through."""
    return tuple(t.float() if (t is not None and t.dtype == torch.float16) else t for t in ts)


def _down(out: Tensor, like: Tensor) -> Tensor:
    """floating outputs go back to fp16 when the tensor they derive from was fp16."""
    return out.half() if (like.dtype == torch.float16 and out.is_floating_point()) else out


def _rows(x: Tensor) -> Tuple[Tensor, int]:
    """[B,P,C] token tensor for the entry points that take x_batch_stride: a view whose rows are dense and C apart
    (x[:, 1:] of a contiguous [B,N,C] tensor -- what the cluster / soft-merge layers receive) is passed in place with
    its batch stride in elements; anything else is made contiguous (stride 0 = dense)."""
    b, p, c = x.shape
    if x.is_contiguous():
        return x, 0
    if b > 0 and p > 0 and x.stride(2) == 1 and x.stride(1) == c and (b == 1 or x.stride(0) >= p * c):
        return x, int(x.stride(0)) if b > 1 else 0
    return x.contiguous(), 0


# ----------------------------------------------------------------------------------------------- Top-K / DyViT
@torch.library.custom_op("tokred::topk_gather", mutates_args=(), device_types="cuda")
def _topk_gather(x: Tensor, scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    _need_cuda("topk_gather", x, scores)
    b, n, c = x.shape
    if scores.dim() != 2 or scores.shape[0] != b or scores.shape[1] != n - 1:
        raise TokredError(f"topk_gather: scores {tuple(scores.shape)} does not match x {tuple(x.shape)}")
    x = _c(x)
    out = torch.empty((b, k + 1, c), dtype=x.dtype, device=x.device)
    idx = torch.empty((b, k), dtype=torch.int64, device=x.device)
    _lib.call("tokred_topk_gather", _ptr(x), _dt(x), _ptr(scores), _dt(scores), scores.stride(1), scores.stride(0),
              None, 0, 0, b, n, c, k, _ptr(out), _ptr(idx), _stream())
    return out, idx


@_topk_gather.register_fake
def _(x, scores, k):
    b, n, c = x.shape
    return x.new_empty((b, k + 1, c)), x.new_empty((b, k), dtype=torch.int64)


@torch.library.custom_op("tokred::topk_gather_attn", mutates_args=(), device_types="cuda")
def _topk_gather_attn(x: Tensor, attn: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    _need_cuda("topk_gather_attn", x, attn)
    b, n, c = x.shape
    if attn.dim() != 4 or attn.shape[0] != b or attn.shape[2] != n or attn.shape[3] != n:
        raise TokredError(f"topk_gather_attn: attn {tuple(attn.shape)} does not match x {tuple(x.shape)}")
    x, attn = _c(x), _c(attn)
    out = torch.empty((b, k + 1, c), dtype=x.dtype, device=x.device)
    idx = torch.empty((b, k), dtype=torch.int64, device=x.device)
    _lib.call("tokred_topk_gather", _ptr(x), _dt(x), None, 0, 0, 0, _ptr(attn), _dt(attn), attn.shape[1], b, n, c, k,
              _ptr(out), _ptr(idx), _stream())
    return out, idx


@_topk_gather_attn.register_fake
def _(x, attn, k):
    b, n, c = x.shape
    return x.new_empty((b, k + 1, c)), x.new_empty((b, k), dtype=torch.int64)


def topk_gather(x: Tensor, scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """models/topk.py:62 + :89-93 (also DynamicViT keep, models/dyvit.py:231-236): (x_out [B,k+1,C], idx [B,k])."""
    xf, sf = _up(x, scores)
    out, idx = torch.ops.tokred.topk_gather(xf, sf, k)
    return _down(out, x), idx


def topk_gather_attn(x: Tensor, attn: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """Same, with the head-mean of the CLS attention row computed in-kernel (no [B,P] round trip)."""
    xf, af = _up(x, attn)
    out, idx = torch.ops.tokred.topk_gather_attn(xf, af, k)
    return _down(out, x), idx


# ----------------------------------------------------------------------------------------------- EViT
def _evit_call(x, scores, attn, k):
    b, n, c = x.shape
    x = _c(x)
    out = torch.empty((b, k + 2, c), dtype=x.dtype, device=x.device)
    idx = torch.empty((b, k + 1), dtype=torch.int64, device=x.device)
    compl = torch.empty((b, n - 1 - k), dtype=torch.int64, device=x.device)
    if scores is not None:
        scores = _c(scores)
        _lib.call("tokred_evit_select_fuse", _ptr(x), _dt(x), _ptr(scores), _dt(scores), None, 0, 0, b, n, c, k,
                  _ptr(out), _ptr(idx), _ptr(compl), _stream())
    else:
        attn = _c(attn)
        _lib.call("tokred_evit_select_fuse", _ptr(x), _dt(x), None, 0, _ptr(attn), _dt(attn), attn.shape[1], b, n, c, k,
                  _ptr(out), _ptr(idx), _ptr(compl), _stream())
    return out, idx, compl


@torch.library.custom_op("tokred::evit_select_fuse", mutates_args=(), device_types="cuda")
def _evit_select_fuse(x: Tensor, scores: Tensor, k: int) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("evit_select_fuse", x, scores)
    if scores.dim() != 2 or scores.shape[0] != x.shape[0] or scores.shape[1] != x.shape[1] - 1:
        raise TokredError(f"evit_select_fuse: scores {tuple(scores.shape)} does not match x {tuple(x.shape)}")
    return _evit_call(x, scores, None, k)


@_evit_select_fuse.register_fake
def _(x, scores, k):
    b, n, c = x.shape
    return (x.new_empty((b, k + 2, c)), x.new_empty((b, k + 1), dtype=torch.int64),
            x.new_empty((b, n - 1 - k), dtype=torch.int64))


@torch.library.custom_op("tokred::evit_select_fuse_attn", mutates_args=(), device_types="cuda")
def _evit_select_fuse_attn(x: Tensor, attn: Tensor, k: int) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("evit_select_fuse_attn", x, attn)
    if attn.dim() != 4 or attn.shape[0] != x.shape[0] or attn.shape[2] != x.shape[1]:
        raise TokredError(f"evit_select_fuse_attn: attn {tuple(attn.shape)} does not match x {tuple(x.shape)}")
    return _evit_call(x, None, attn, k)


@_evit_select_fuse_attn.register_fake
def _(x, attn, k):
    b, n, c = x.shape
    return (x.new_empty((b, k + 2, c)), x.new_empty((b, k + 1), dtype=torch.int64),
            x.new_empty((b, n - 1 - k), dtype=torch.int64))


def evit_select_fuse(x: Tensor, scores: Tensor, k: int):
    """models/evit.py:84 + :111-123: (x_out [B,k+2,C], idx [B,k+1] with trailing -1, compl [B,P-k])."""
    xf, sf = _up(x, scores)
    out, idx, compl = torch.ops.tokred.evit_select_fuse(xf, sf, k)
    return _down(out, x), idx, compl


def evit_select_fuse_attn(x: Tensor, attn: Tensor, k: int):
    xf, af = _up(x, attn)
    out, idx, compl = torch.ops.tokred.evit_select_fuse_attn(xf, af, k)
    return _down(out, x), idx, compl


# ----------------------------------------------------------------------------------------------- ToMe
def tome_effective_r(n_tokens: int, r: int, class_token: bool = True, distill_token: bool = False) -> int:
    """models/tome.py:244-253."""
    return max(min(r, (n_tokens - int(class_token) - int(distill_token)) // 2), 0)


@torch.library.custom_op("tokred::tome_match", mutates_args=(), device_types="cuda")
def _tome_match(metric: Tensor, r: int, class_token: bool, lowp: bool, tensor_cores: bool,
                distill_token: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("tome_match", metric)
    b, n, d = metric.shape
    re = tome_effective_r(n, r, class_token, distill_token)
    if re <= 0:
        raise TokredError(f"tome_match: effective r = {re}; nothing to merge (caller must skip, models/tome.py:255)")
    metric = _c(metric)
    na = (n + 1) // 2
    unm = torch.empty((b, na - re), dtype=torch.int64, device=metric.device)
    src = torch.empty((b, re), dtype=torch.int64, device=metric.device)
    dst = torch.empty((b, re), dtype=torch.int64, device=metric.device)
    mode = (1 if tensor_cores else 3) if lowp else 0
    _lib.call("tokred_tome_match", _ptr(metric), _dt(metric), 1, 0, b, n, d, r, int(class_token) | (2 if distill_token else 0),
              mode, _ptr(unm), _ptr(src), _ptr(dst), _stream())
    return unm, src, dst


@torch.library.custom_op("tokred::tome_match_qkv", mutates_args=(), device_types="cuda")
def _tome_match_qkv(qkv: Tensor, num_heads: int, r: int, class_token: bool, distill_token: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("tome_match_qkv", qkv)
    if qkv.dim() != 3 or qkv.dtype != torch.bfloat16 or qkv.shape[2] != 3 * 64 * num_heads:
        raise TokredError(f"tome_match_qkv: qkv {tuple(qkv.shape)} {qkv.dtype}; expected bf16 [B,N,3*H*64]")
    b, n, c3 = qkv.shape
    c = c3 // 3
    re = tome_effective_r(n, r, class_token, distill_token)
    if re <= 0:
        raise TokredError(f"tome_match_qkv: effective r = {re}; nothing to merge (caller must skip, models/tome.py:255)")
    qkv = _c(qkv)
    na = (n + 1) // 2
    unm = torch.empty((b, na - re), dtype=torch.int64, device=qkv.device)
    src = torch.empty((b, re), dtype=torch.int64, device=qkv.device)
    dst = torch.empty((b, re), dtype=torch.int64, device=qkv.device)
    # the k slice of token 0 starts C elements into the row; consecutive tokens are 3C apart
    _lib.call("tokred_tome_match", qkv.data_ptr() + 2 * c, BF16, num_heads, c3, b, n, 64, r,
              int(class_token) | (2 if distill_token else 0), 1, _ptr(unm), _ptr(src), _ptr(dst), _stream())
    return unm, src, dst


@_tome_match_qkv.register_fake
def _(qkv, num_heads, r, class_token, distill_token=False):
    b, n, _ = qkv.shape
    re = tome_effective_r(n, r, class_token, distill_token)
    na = (n + 1) // 2
    mk = lambda m: qkv.new_empty((b, m), dtype=torch.int64)
    return mk(na - re), mk(re), mk(re)


@_tome_match.register_fake
def _(metric, r, class_token, lowp, tensor_cores, distill_token=False):
    b, n, d = metric.shape
    re = tome_effective_r(n, r, class_token, distill_token)
    na = (n + 1) // 2
    mk = lambda m: metric.new_empty((b, m), dtype=torch.int64)
    return mk(na - re), mk(re), mk(re)


@torch.library.custom_op("tokred::tome_merge", mutates_args=(), device_types="cuda")
def _tome_merge(x: Tensor, size: Optional[Tensor], unm: Tensor, src: Tensor, dst: Tensor, want_map: bool,
                divide: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("tome_merge", x, size, unm, src, dst)
    b, n, c = x.shape
    r = src.shape[1]
    if unm.shape != (b, (n + 1) // 2 - r) or dst.shape != (b, r) or src.shape[0] != b:
        raise TokredError("tome_merge: index tensors do not match x")
    x, unm, src, dst = _c(x), _c(unm), _c(src), _c(dst)
    if size is not None:
        if size.numel() != b * n:
            raise TokredError(f"tome_merge: size {tuple(size.shape)} does not match x {tuple(x.shape)}")
        size = _c(size.to(x.dtype))
    out = torch.empty((b, n - r, c), dtype=x.dtype, device=x.device)
    size_out = torch.empty((b, n - r, 1), dtype=x.dtype, device=x.device)
    rci = torch.empty((b, n - 1) if want_map else (0,), dtype=torch.float32, device=x.device)
    _lib.call("tokred_tome_merge", _ptr(x), _dt(x), _ptr(size), _ptr(unm), _ptr(src), _ptr(dst), b, n, c, r, _ptr(out),
              _ptr(size_out), _ptr(rci) if want_map else None, int(divide), _stream())
    return out, size_out, rci


@_tome_merge.register_fake
def _(x, size, unm, src, dst, want_map, divide):
    b, n, c = x.shape
    r = src.shape[1]
    return (x.new_empty((b, n - r, c)), x.new_empty((b, n - r, 1)),
            x.new_empty((b, n - 1) if want_map else (0,), dtype=torch.float32))


def tome_match(metric: Tensor, r: int, class_token: bool = True, lowp: bool = False, tensor_cores: bool = True,
               distill_token: bool = False):
    """models/tome.py:258-277: (unm_idx [B,a-r], src_idx [B,r], dst_idx [B,r]) int64.
    lowp=True reproduces the bf16 autocast matmul (on tcgen05 tensor cores; tensor_cores=False keeps the same
    rounding on the FFMA path, used as a cross-check).  distill_token protects odd token 0 as a destination (:265-266)."""
    return torch.ops.tokred.tome_match(_up(metric)[0], r, class_token, lowp, tensor_cores, distill_token)


def tome_match_qkv(qkv: Tensor, num_heads: int, r: int, class_token: bool = True, distill_token: bool = False):
    """bipartite matching on ``metric = k.mean(1)`` (models/tome.py:58, :258-277) taken straight from the qkv Linear's
    bf16 output [B,N,3C]: the head mean is computed in-kernel (fp32 sum * 1/H, rounded to bf16 like ATen's mean), so the
    metric tensor is never materialised.  bf16-autocast semantics (= tome_match(..., lowp=True))."""
    return torch.ops.tokred.tome_match_qkv(qkv, num_heads, r, class_token, distill_token)


def tome_merge(x: Tensor, size: Optional[Tensor], unm: Tensor, src: Tensor, dst: Tensor, want_map: bool = True,
               divide: bool = True):
    """models/tome.py:279-289,309-323 + Block_ToMe :91-99: (x_out [B,N-r,C], size_out [B,N-r,1], map [B,N-1] f32).
    divide=False returns the bare merge closure's sums instead of the size-weighted mean."""
    xf, sf = _up(x, size)
    out, size_out, rci = torch.ops.tokred.tome_merge(xf, sf, unm, src, dst, want_map, divide)
    return _down(out, x), _down(size_out, x), rci


# ----------------------------------------------------------------------------------------------- distances
@torch.library.custom_op("tokred::pairwise_dist", mutates_args=(), device_types="cuda")
def _pairwise_dist(x: Tensor, post_scale: float, exact_fp32: bool) -> Tensor:
    _need_cuda("pairwise_dist", x)
    b, p, c = x.shape
    x = _c(x.float())
    out = torch.empty((b, p, p), dtype=torch.float32, device=x.device)
    _lib.call("tokred_pairwise_dist", _ptr(x), b, p, c, float(post_scale), int(exact_fp32), _ptr(out), _stream())
    return out


@_pairwise_dist.register_fake
def _(x, post_scale, exact_fp32):
    b, p, _ = x.shape
    return x.new_empty((b, p, p), dtype=torch.float32)


def pairwise_dist(x: Tensor, post_scale: float = 1.0, exact_fp32: bool = False) -> Tensor:
    """torch.cdist(x, x) * post_scale with ATen's formula selection (models/dpcknn.py:59, models/kmedoids.py:68).
    Default: Gram on tcgen05 with 3xTF32 compensation; exact_fp32=True: FFMA."""
    return torch.ops.tokred.pairwise_dist(_up(x)[0], post_scale, exact_fp32)


# ----------------------------------------------------------------------------------------------- DPC-KNN
@torch.library.custom_op("tokred::dpcknn_cluster", mutates_args=(), device_types="cuda")
def _dpcknn_cluster(x: Tensor, noise_u: Tensor, cluster_num: int, knn: int, exact_fp32: bool) -> Tuple[Tensor, Tensor]:
    _need_cuda("dpcknn_cluster", x, noise_u)
    b, p, c = x.shape
    if x.dtype != torch.float32:
        x = x.float()     # cdist runs in fp32 under autocast
    (x, xbs), noise_u = _rows(x), _c(noise_u.float())
    if noise_u.shape != (b, p):
        raise TokredError(f"dpcknn_cluster: noise {tuple(noise_u.shape)} != {(b, p)}")
    idx_cluster = torch.empty((b, p), dtype=torch.int64, device=x.device)
    index_down = torch.empty((b, cluster_num), dtype=torch.int64, device=x.device)
    _lib.call("tokred_dpcknn_cluster", _ptr(x), xbs, _ptr(noise_u), b, p, c, cluster_num, knn, int(exact_fp32), _ptr(idx_cluster),
              _ptr(index_down), _stream())
    return idx_cluster, index_down


@_dpcknn_cluster.register_fake
def _(x, noise_u, cluster_num, knn, exact_fp32):
    b, p, _ = x.shape
    return x.new_empty((b, p), dtype=torch.int64), x.new_empty((b, cluster_num), dtype=torch.int64)


@torch.library.custom_op("tokred::dpcknn_merge", mutates_args=(), device_types="cuda")
def _dpcknn_merge(x: Tensor, idx_token: Tensor, agg_weight: Tensor, idx_cluster: Tensor,
                  token_weight: Optional[Tensor], cluster_num: int) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("dpcknn_merge", x, idx_token, agg_weight, idx_cluster, token_weight)
    b, p, c = x.shape
    t = idx_token.shape[1]
    if x.dtype != torch.float32:
        raise TokredError("dpcknn_merge: x must be float32 (the residual stream is fp32 under autocast)")
    (x, xbs), idx_token, idx_cluster = _rows(x), _c(idx_token), _c(idx_cluster)
    agg = _c(agg_weight.float())
    tw = None if token_weight is None else _c(token_weight.float())
    merged = torch.empty((b, cluster_num, c), dtype=torch.float32, device=x.device)
    idx_token_new = torch.empty((b, t), dtype=torch.int64, device=x.device)
    agg_new = torch.empty((b, t, 1), dtype=torch.float32, device=x.device)
    _lib.call("tokred_dpcknn_merge", _ptr(x), xbs, _ptr(idx_token), _ptr(agg), _ptr(idx_cluster), _ptr(tw), b, p, c,
              cluster_num, t, _ptr(merged), _ptr(idx_token_new), _ptr(agg_new), _stream())
    return merged, idx_token_new, agg_new


@_dpcknn_merge.register_fake
def _(x, idx_token, agg_weight, idx_cluster, token_weight, cluster_num):
    b, p, c = x.shape
    t = idx_token.shape[1]
    return (x.new_empty((b, cluster_num, c)), x.new_empty((b, t), dtype=torch.int64),
            x.new_empty((b, t, 1), dtype=torch.float32))


def dpcknn_cluster(x: Tensor, noise_u: Tensor, cluster_num: int, knn: int = 5, exact_fp32: bool = False):
    """models/dpcknn.py:44-100: (idx_cluster [B,P], index_down [B,K]) int64; noise_u = torch.rand(B,P)."""
    return torch.ops.tokred.dpcknn_cluster(_up(x)[0], noise_u, cluster_num, knn, exact_fp32)


def dpcknn_merge(x, idx_token, agg_weight, idx_cluster, token_weight, cluster_num):
    """models/dpcknn.py:103-140: (x_merged [B,K,C], idx_token_new [B,T], agg_weight_new [B,T,1])."""
    xf, aw, tw = _up(x, agg_weight, token_weight)
    merged, idx_new, agg_new = torch.ops.tokred.dpcknn_merge(xf, idx_token, aw, idx_cluster, tw, cluster_num)
    return _down(merged, x), idx_new, _down(agg_new, agg_weight)


# ----------------------------------------------------------------------------------------------- K-Medoids
@torch.library.custom_op("tokred::attn_colsum", mutates_args=(), device_types="cuda")
def _attn_colsum(attn: Tensor, num_tokens: int) -> Tensor:
    _need_cuda("attn_colsum", attn)
    b, h, n, n2 = attn.shape
    if n != n2:
        raise TokredError("attn_colsum: attention must be [B,H,N,N]")
    attn = _c(attn)
    out = torch.empty((b, n - num_tokens, 1), dtype=torch.float32, device=attn.device)
    _lib.call("tokred_attn_colsum", _ptr(attn), _dt(attn), b, h, n, num_tokens, _ptr(out), _stream())
    return out


@_attn_colsum.register_fake
def _(attn, num_tokens):
    b, _, n, _ = attn.shape
    return attn.new_empty((b, n - num_tokens, 1), dtype=torch.float32)


@torch.library.custom_op("tokred::kmedoids_fit", mutates_args=(), device_types="cuda")
def _kmedoids_fit(x: Tensor, token_weight: Tensor, cluster_num: int, iters: int, exact_fp32: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("kmedoids_fit", x, token_weight)
    b, p, c = x.shape
    if x.dtype != torch.float32:
        raise TokredError("kmedoids_fit: x must be float32 (cdist runs in fp32)")
    if token_weight.numel() != b * p:
        raise TokredError(f"kmedoids_fit: token_weight {tuple(token_weight.shape)} does not match x")
    (x, xbs), tw = _rows(x), _c(token_weight.float())
    centres = torch.empty((b, cluster_num, c), dtype=torch.float32, device=x.device)
    cidx = torch.empty((b, cluster_num), dtype=torch.int64, device=x.device)
    assign = torch.empty((b, p), dtype=torch.int64, device=x.device)
    _lib.call("tokred_kmedoids_fit", _ptr(x), xbs, _ptr(tw), b, p, c, cluster_num, iters, int(exact_fp32), _ptr(centres), _ptr(cidx),
              _ptr(assign), _stream())
    return centres, cidx, assign


@_kmedoids_fit.register_fake
def _(x, token_weight, cluster_num, iters, exact_fp32):
    b, p, c = x.shape
    return (x.new_empty((b, cluster_num, c)), x.new_empty((b, cluster_num), dtype=torch.int64),
            x.new_empty((b, p), dtype=torch.int64))


@torch.library.custom_op("tokred::kmedoids_fit_init", mutates_args=(), device_types="cuda")
def _kmedoids_fit_init(x: Tensor, init_idx: Tensor, iters: int, exact_fp32: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("kmedoids_fit_init", x, init_idx)
    b, p, c = x.shape
    if x.dtype != torch.float32:
        raise TokredError("kmedoids_fit_init: x must be float32 (cdist runs in fp32)")
    if init_idx.dim() != 2 or init_idx.shape[0] != b or init_idx.dtype != torch.int64 or not 1 <= init_idx.shape[1] <= p:
        raise TokredError(f"kmedoids_fit_init: init_idx {tuple(init_idx.shape)} {init_idx.dtype}; expected int64 [B,K<=P]")
    k = init_idx.shape[1]
    (x, xbs), init_idx = _rows(x), _c(init_idx)
    centres = torch.empty((b, k, c), dtype=torch.float32, device=x.device)
    cidx = torch.empty((b, k), dtype=torch.int64, device=x.device)
    assign = torch.empty((b, p), dtype=torch.int64, device=x.device)
    _lib.call("tokred_kmedoids_fit_init", _ptr(x), xbs, None, _ptr(init_idx), b, p, c, k, iters, int(exact_fp32), _ptr(centres),
              _ptr(cidx), _ptr(assign), _stream())
    return centres, cidx, assign


@_kmedoids_fit_init.register_fake
def _(x, init_idx, iters, exact_fp32):
    b, p, c = x.shape
    k = init_idx.shape[1]
    return (x.new_empty((b, k, c)), x.new_empty((b, k), dtype=torch.int64), x.new_empty((b, p), dtype=torch.int64))


def kmedoids_init_farthest(dist: Tensor, cluster_num: int, first: int) -> Tensor:
    """models/kmedoids.py:43-59, the equal_weight initialisation, on a full distance matrix [B,P,P]: start from token
    ``first`` (the reference's one numpy draw, shared by the batch); k-th medoid = the token whose LARGEST distance to
    the medoids chosen so far is largest (``torch.max`` over the medoid axis, then over tokens: first index on ties), the
    rows of chosen medoids zeroed (:53-55).  A running maximum replaces the reference's growing [B,P,k] cdist.
    -> init_idx [B,K] int64."""
    b, p, _ = dist.shape
    idx = torch.full((b, 1), int(first), dtype=torch.int64, device=dist.device)
    chosen = torch.zeros((b, p), dtype=torch.bool, device=dist.device)
    chosen[:, int(first)] = True
    far = dist[:, :, int(first)].clone()
    for _ in range(1, cluster_num):
        new = far.masked_fill(chosen, 0.0).argmax(dim=-1, keepdim=True)                  # [B,1]
        idx = torch.cat((idx, new), dim=-1)
        chosen.scatter_(1, new, True)
        far = torch.maximum(far, torch.gather(dist, 2, new.unsqueeze(1).expand(-1, p, -1)).squeeze(-1))
    return idx


def kmedoids_fit_equal(x: Tensor, cluster_num: int, iters: int, first: int, exact_fp32: bool = False):
    """models/kmedoids.py:40-85 with token_weight=None (``--equal_weight``): farthest-point initialisation from token
    ``first`` on the product's distance matrix, then the medoid iterations with unit weights in the K-Medoids kernel."""
    xf = _up(x)[0]
    dist = torch.ops.tokred.pairwise_dist(xf, 1.0, exact_fp32)
    init = kmedoids_init_farthest(dist, cluster_num, first)
    centres, cidx, assign = torch.ops.tokred.kmedoids_fit_init(xf, init, iters, exact_fp32)
    return _down(centres, x), cidx, assign


def attn_colsum(attn: Tensor, num_tokens: int = 1) -> Tensor:
    """models/kmedoids.py:240: token weights [B,P,1] = sum over heads and query rows of attention columns."""
    return torch.ops.tokred.attn_colsum(_up(attn)[0], num_tokens)


def kmedoids_fit(x: Tensor, token_weight: Tensor, cluster_num: int, iters: int, exact_fp32: bool = False):
    """models/kmedoids.py:62-85: (centres [B,K,C], cluster_idx [B,K], assignment [B,P])."""
    xf, tw = _up(x, token_weight)
    centres, cidx, assign = torch.ops.tokred.kmedoids_fit(xf, tw, cluster_num, iters, exact_fp32)
    return _down(centres, x), cidx, assign


# ----------------------------------------------------------------------------------------------- soft merges
def sinkhorn_log_norm(k: int, p: int, score_dtype: torch.dtype) -> float:
    """-log(K+P) the way models/sinkhorn.py:43-47 evaluates it: (m*one + n*one) in the dtype of the score matrix
    (bf16 under autocast: sums above 256 are rounded to even multiples of 2), then an fp32 log."""
    one = torch.ones((), dtype=score_dtype)
    return float(-((k * one).to(score_dtype) + (p * one).to(score_dtype)).float().log())


# False: never hand scratch to the soft merges -> they use the scratch-free tensor-core kernel (softmerge_tc.cu)
SOFT_MERGE_SCRATCH = True


def _soft_workspace(x: Tensor, b: int, p: int, c: int, k: int, lowp: bool, tensor_cores: bool):
    """scratch for the bulk-copy fed tensor-core kernels (bf16 token tiles + packed Q); (None, 0) when not used."""
    if not (lowp and tensor_cores and SOFT_MERGE_SCRATCH):
        return None, 0
    n = int(_lib.load().tokred_soft_merge_workspace_bytes(b, p, c, k))
    if n == 0:
        return None, 0
    ws = torch.empty(n + 128, dtype=torch.uint8, device=x.device)
    off = (-ws.data_ptr()) % 128
    return ws[off:off + n], n


def _lowp_mode(lowp: bool, tensor_cores: bool) -> int:
    """C-ABI lowp code: 0 exact fp32, 1 bf16-autocast rounding on tcgen05, 3 same rounding on the FFMA path."""
    return (1 if tensor_cores else 3) if lowp else 0


def _soft_out_dtype(x: Tensor, lowp: bool) -> torch.dtype:
    return torch.bfloat16 if (lowp or x.dtype == torch.bfloat16) else torch.float32


@torch.library.custom_op("tokred::sinkhorn_merge", mutates_args=(), device_types="cuda")
def _sinkhorn_merge(x: Tensor, v_hat: Tensor, eps: float, iters: int, lowp: bool, tensor_cores: bool) -> Tuple[Tensor, Tensor]:
    _need_cuda("sinkhorn_merge", x, v_hat)
    b, p, c = x.shape
    k = v_hat.shape[0]
    if v_hat.shape != (k, c):
        raise TokredError("sinkhorn_merge: v_hat must be [K,C]")
    (x, xbs), v_hat = _rows(x), _c(v_hat.float())
    odt = _soft_out_dtype(x, lowp)
    out = torch.empty((b, k, c), dtype=odt, device=x.device)
    weights = torch.empty((b, k, p), dtype=torch.float32, device=x.device)
    ws, ws_bytes = _soft_workspace(x, b, p, c, k, lowp, tensor_cores)
    _lib.call("tokred_sinkhorn_merge", _ptr(x), _dt(x), xbs, _ptr(v_hat), b, p, c, k, float(eps),
              sinkhorn_log_norm(k, p, torch.bfloat16 if lowp else x.dtype), iters, _lowp_mode(lowp, tensor_cores), _ptr(out), _dt(out),
              _ptr(weights), _ptr(ws), ws_bytes, _stream())
    return out, weights


@_sinkhorn_merge.register_fake
def _(x, v_hat, eps, iters, lowp, tensor_cores):
    b, p, c = x.shape
    k = v_hat.shape[0]
    return x.new_empty((b, k, c), dtype=_soft_out_dtype(x, lowp)), x.new_empty((b, k, p), dtype=torch.float32)


@torch.library.custom_op("tokred::patchmerger", mutates_args=(), device_types="cuda")
def _patchmerger(x: Tensor, ln_weight: Tensor, ln_bias: Tensor, queries: Tensor, scale: float, ln_eps: float,
                 lowp: bool, tensor_cores: bool) -> Tuple[Tensor, Tensor]:
    _need_cuda("patchmerger", x, ln_weight, ln_bias, queries)
    b, p, c = x.shape
    k = queries.shape[0]
    if queries.shape != (k, c) or ln_weight.numel() != c or ln_bias.numel() != c:
        raise TokredError("patchmerger: parameter shapes do not match x")
    (x, xbs), queries = _rows(x), _c(queries.float())
    lw, lb = _c(ln_weight.float()), _c(ln_bias.float())
    odt = _soft_out_dtype(x, lowp)
    out = torch.empty((b, k, c), dtype=odt, device=x.device)
    attn = torch.empty((b, k, p), dtype=torch.float32, device=x.device)
    ws, ws_bytes = _soft_workspace(x, b, p, c, k, lowp, tensor_cores)
    _lib.call("tokred_patchmerger", _ptr(x), _dt(x), xbs, _ptr(lw), _ptr(lb), _ptr(queries), b, p, c, k, float(scale),
              float(ln_eps), _lowp_mode(lowp, tensor_cores), _ptr(out), _dt(out), _ptr(attn), _ptr(ws), ws_bytes, _stream())
    return out, attn


@_patchmerger.register_fake
def _(x, ln_weight, ln_bias, queries, scale, ln_eps, lowp, tensor_cores):
    b, p, c = x.shape
    k = queries.shape[0]
    return x.new_empty((b, k, c), dtype=_soft_out_dtype(x, lowp)), x.new_empty((b, k, p), dtype=torch.float32)


@torch.library.custom_op("tokred::sit_merge", mutates_args=(), device_types="cuda")
def _sit_merge(x: Tensor, logits: Tensor, scale: Tensor, lowp: bool, tensor_cores: bool) -> Tuple[Tensor, Tensor]:
    _need_cuda("sit_merge", x, logits, scale)
    b, p, c = x.shape
    k = logits.shape[2]
    if logits.shape != (b, p, k):
        raise TokredError("sit_merge: logits must be [B,P,K]")
    (x, xbs), logits = _rows(x), _c(logits)
    if scale.numel() != 1:
        raise TokredError("sit_merge: scale must hold one element")
    scale = _c(scale.detach().float().reshape(1))
    odt = _soft_out_dtype(x, lowp)
    out = torch.empty((b, k, c), dtype=odt, device=x.device)
    w = torch.empty((b, k, p), dtype=torch.float32, device=x.device)
    ws, ws_bytes = _soft_workspace(x, b, p, c, k, lowp, tensor_cores)
    _lib.call("tokred_sit_merge", _ptr(x), _dt(x), xbs, _ptr(logits), _dt(logits), _ptr(scale), b, p, c, k, _lowp_mode(lowp, tensor_cores),
              _ptr(out), _dt(out), _ptr(w), _ptr(ws), ws_bytes, _stream())
    return out, w


@_sit_merge.register_fake
def _(x, logits, scale, lowp, tensor_cores):
    b, p, c = x.shape
    k = logits.shape[2]
    return x.new_empty((b, k, c), dtype=_soft_out_dtype(x, lowp)), x.new_empty((b, k, p), dtype=torch.float32)


def sinkhorn_merge(x: Tensor, v_hat: Tensor, eps: float, iters: int, lowp: bool = False, tensor_cores: bool = True):
    """models/sinkhorn.py:66-86: (out [B,K,C], weights [B,K,P]); v_hat = F.normalize(v)."""
    return torch.ops.tokred.sinkhorn_merge(_up(x)[0], _up(v_hat)[0], eps, iters, lowp, tensor_cores)


def patchmerger(x, ln_weight, ln_bias, queries, scale: float = 1.0, ln_eps: float = 1e-5, lowp: bool = False,
                tensor_cores: bool = True):
    """models/patchmerger.py:35-39: (out [B,K,C], attn [B,K,P])."""
    return torch.ops.tokred.patchmerger(_up(x)[0], ln_weight, ln_bias, queries, scale, ln_eps, lowp, tensor_cores)


def sit_merge(x: Tensor, logits: Tensor, scale: Tensor, lowp: bool = False, tensor_cores: bool = True):
    """models/sit.py:37-40: (out [B,K,C], weight [B,K,P])."""
    xf, lf = _up(x, logits)
    return torch.ops.tokred.sit_merge(xf, lf, scale, lowp, tensor_cores)


# ----------------------------------------------------------------------------------------------- ATS
@torch.library.custom_op("tokred::ats_sample", mutates_args=(), device_types="cuda")
def _ats_sample(v: Tensor, attn: Tensor, mask: Tensor, steps: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("ats_sample", v, attn, mask, steps)
    b, h, n, dh = v.shape
    if attn.shape not in ((b, h, n, n), (b, h, n)) or mask.shape != (b, n):          # [B,H,N] = CLS rows only
        raise TokredError("ats_sample: attn/mask shapes do not match v")
    if v.stride(3) != 1:
        v = v.contiguous()
    attn = _c(attn.float())
    mask8 = _c(mask.to(torch.uint8)) if mask.dtype != torch.bool else _c(mask).view(torch.uint8)
    steps = _c(steps.float())
    ns = steps.numel()
    ids = torch.empty((b, ns + 1), dtype=torch.int64, device=v.device)
    mask_out = torch.empty((b, ns + 1), dtype=torch.bool, device=v.device)
    max_count = torch.zeros((1,), dtype=torch.int32, device=v.device)
    _lib.call("tokred_ats_sample", _ptr(v), _dt(v), v.stride(0), v.stride(1), v.stride(2), _ptr(attn),
              n * n if attn.dim() == 4 else n, _ptr(mask8), _ptr(steps), b, h, n, dh, ns, float(eps),
              _ptr(ids), mask_out.data_ptr(), _ptr(max_count), _stream())
    return ids, mask_out, max_count


@_ats_sample.register_fake
def _(v, attn, mask, steps, eps):
    b = v.shape[0]
    ns = steps.numel()
    return (v.new_empty((b, ns + 1), dtype=torch.int64), v.new_empty((b, ns + 1), dtype=torch.bool),
            v.new_empty((1,), dtype=torch.int32))


@torch.library.custom_op("tokred::gather_rows", mutates_args=(), device_types="cuda")
def _gather_rows(src: Tensor, ids: Tensor, m: int) -> Tensor:
    _need_cuda("gather_rows", src, ids)
    if src.dim() == 3:
        b, n, w = src.shape
        g = 1
    elif src.dim() == 4:
        b, g, n, w = src.shape
    else:
        raise TokredError("gather_rows: src must be [B,N,W] or [B,G,N,W]")
    if ids.dim() != 2 or ids.shape[0] != b or ids.shape[1] < m or ids.stride(1) != 1:
        raise TokredError("gather_rows: ids must be [B,>=M] with unit inner stride")
    src = _c(src)
    shape = (b, m, w) if src.dim() == 3 else (b, g, m, w)
    out = torch.empty(shape, dtype=src.dtype, device=src.device)
    _lib.call("tokred_gather_rows", _ptr(src), _dt(src), _ptr(ids), ids.stride(0), b, g, n, w, m, _ptr(out), _stream())
    return out


@_gather_rows.register_fake
def _(src, ids, m):
    shape = list(src.shape)
    shape[-2] = m
    return src.new_empty(shape)


def ats_sample(v: Tensor, attn: Tensor, mask: Tensor, steps: Tensor, eps: float = 1e-6):
    """models/ats.py:52-82: (ids [B,n_steps+1] zero-padded sorted unique, mask [B,n_steps+1], max_count int32[1])."""
    vf, af = _up(v, attn)
    return torch.ops.tokred.ats_sample(vf, af, mask, steps, eps)


def gather_rows(src: Tensor, ids: Tensor, m: Optional[int] = None) -> Tensor:
    """out[b,(g,)j,:] = src[b,(g,)ids[b,j],:] for j < m (models/ats.py:84-87, :156-157)."""
    return _down(torch.ops.tokred.gather_rows(_up(src)[0], ids, ids.shape[1] if m is None else m), src)


# ----------------------------------------------------------------------------------------------- DynamicViT
@torch.library.custom_op("tokred::dyvit_pool_concat", mutates_args=(), device_types="cuda")
def _dyvit_pool_concat(h: Tensor, policy: Tensor, eps: float) -> Tensor:
    _need_cuda("dyvit_pool_concat", h, policy)
    b, p, c = h.shape
    if policy.numel() != b * p:
        raise TokredError("dyvit_pool_concat: policy must be [B,P,1]")
    h, hbs = _rows(h)
    pol = _c(policy.float())
    odt = torch.promote_types(h.dtype, policy.dtype)      # the reference's torch.cat promotes (models/dyvit.py:118)
    out = torch.empty((b, p, c), dtype=odt, device=h.device)
    _lib.call("tokred_dyvit_pool_concat", _ptr(h), _dt(h), hbs, _ptr(pol), b, p, c, float(eps), _ptr(out), _dt(out), _stream())
    return out


@_dyvit_pool_concat.register_fake
def _(h, policy, eps):
    return h.new_empty(h.shape, dtype=torch.promote_types(h.dtype, policy.dtype))


@torch.library.custom_op("tokred::dyvit_pool_concat_lowp", mutates_args=(), device_types="cuda")
def _dyvit_pool_concat_lowp(h: Tensor, policy: Tensor, eps: float) -> Tensor:
    _need_cuda("dyvit_pool_concat_lowp", h, policy)
    b, p, c = h.shape
    if policy.numel() != b * p or h.dtype != torch.bfloat16:
        raise TokredError("dyvit_pool_concat_lowp: h must be bf16 [B,P,C], policy [B,P,1]")
    (h, hbs), pol = _rows(h), _c(policy.float())
    out = torch.empty((b, p, c), dtype=torch.bfloat16, device=h.device)
    _lib.call("tokred_dyvit_pool_concat", _ptr(h), _dt(h), hbs, _ptr(pol), b, p, c, float(eps), _ptr(out), _dt(out), _stream())
    return out


@_dyvit_pool_concat_lowp.register_fake
def _(h, policy, eps):
    return h.new_empty(h.shape)


def dyvit_pool_concat(h: Tensor, policy: Tensor, eps: float = 1e-6, lowp_out: bool = False) -> Tensor:
    """models/dyvit.py:114-118: [local half | masked mean of the global half + eps].
    lowp_out (bf16 h, inference under bf16 autocast): emit the concatenation in bf16 -- what the autocast Linear that
    consumes it (out_conv, :119) would cast the reference's fp32 ``torch.cat`` result to: the local half is bf16 already,
    the pooled half is rounded once either way, so the Linear sees the same bits and the fp32 tensor + its cast disappear."""
    if lowp_out and h.dtype == torch.bfloat16 and not (torch.is_grad_enabled() and (h.requires_grad or policy.requires_grad)):
        return torch.ops.tokred.dyvit_pool_concat_lowp(h, policy, eps)
    hf, pf = _up(h, policy)
    out = torch.ops.tokred.dyvit_pool_concat(hf, pf, eps)
    return out.to(torch.promote_types(h.dtype, policy.dtype))      # torch.cat's promotion (models/dyvit.py:118)


# ----------------------------------------------------------------------------------------------- f1: attention producer
def attention_supported(qkv: Tensor, num_heads: int) -> bool:
    """shapes the fused attention kernel covers: bf16 qkv [B,N,3C] on CUDA, head dim 64, N <= 256."""
    return (qkv.is_cuda and qkv.dtype == torch.bfloat16 and qkv.dim() == 3 and qkv.shape[1] <= 256
            and qkv.shape[2] == 3 * 64 * num_heads)


@torch.library.custom_op("tokred::attention", mutates_args=(), device_types="cuda")
def _attention(qkv: Tensor, num_heads: int, scale: float, key_bias: Optional[Tensor], mask: Optional[Tensor],
               q_ids: Optional[Tensor], want_out: bool, want_cls: bool, want_colsum: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("attention", qkv, key_bias, mask, q_ids)
    if qkv.dim() != 3 or qkv.dtype != torch.bfloat16 or qkv.shape[2] % (3 * num_heads) != 0:
        raise TokredError(f"attention: qkv {tuple(qkv.shape)} {qkv.dtype}; expected bf16 [B,N,3*H*Dh]")
    b, n, c3 = qkv.shape
    c = c3 // 3
    qkv = _c(qkv)
    if key_bias is not None:
        if key_bias.numel() != b * n:
            raise TokredError(f"attention: key_bias {tuple(key_bias.shape)} does not match qkv {tuple(qkv.shape)}")
        key_bias = _c(key_bias.to(torch.float32))
    if mask is not None:
        if mask.shape != (b, n):
            raise TokredError(f"attention: mask {tuple(mask.shape)} does not match qkv {tuple(qkv.shape)}")
        mask = _c(mask.to(torch.bool)).view(torch.uint8)
    m, ids_stride = n, 0
    if q_ids is not None:
        if q_ids.dim() != 2 or q_ids.shape[0] != b or q_ids.dtype != torch.int64:
            raise TokredError(f"attention: q_ids {tuple(q_ids.shape)} {q_ids.dtype}; expected int64 [B,M]")
        q_ids = _c(q_ids)
        m, ids_stride = q_ids.shape[1], q_ids.shape[1]
    dev = qkv.device
    out = torch.empty((b, m, c) if want_out else (0,), dtype=torch.bfloat16, device=dev)
    cls = torch.empty((b, num_heads, n) if want_cls else (0,), dtype=torch.float32, device=dev)
    colsum = torch.empty((b, num_heads, n) if want_colsum else (0,), dtype=torch.float32, device=dev)
    _lib.call("tokred_attention", _ptr(qkv), b, n, num_heads, c // num_heads, float(scale), _ptr(key_bias), _ptr(mask),
              _ptr(q_ids), ids_stride, m, _ptr(out) if want_out else None, _ptr(cls) if want_cls else None,
              _ptr(colsum) if want_colsum else None, _stream())
    return out, cls, colsum


@_attention.register_fake
def _(qkv, num_heads, scale, key_bias, mask, q_ids, want_out, want_cls, want_colsum):
    b, n, c3 = qkv.shape
    m = n if q_ids is None else q_ids.shape[1]
    f32 = lambda on: qkv.new_empty((b, num_heads, n) if on else (0,), dtype=torch.float32)
    return qkv.new_empty((b, m, c3 // 3) if want_out else (0,)), f32(want_cls), f32(want_colsum)


def attention(qkv: Tensor, num_heads: int, scale: float, key_bias: Optional[Tensor] = None, mask: Optional[Tensor] = None,
              q_ids: Optional[Tensor] = None, want_out: bool = True, want_cls: bool = False, want_colsum: bool = False):
    """softmax(q k^T * scale [+ key_bias] [masked]) v with bf16-autocast roundings (models/topk.py:44-52; tome.py:44-58;
    ats.py:115-127) from the qkv Linear's output [B,N,3C] -> (x [B,M,C] bf16 ready for proj | None, cls_row [B,H,N]
    fp32 | None, colsum [B,H,N] fp32 | None).  The [B,H,N,N] probabilities are never materialised: cls_row is
    attn[:, :, 0, :] (models/topk.py:60), colsum is attn.sum(2) (models/kmedoids.py:240 sums it over heads), q_ids
    [B,M] selects the query rows to compute (the ATS row gather, models/ats.py:84-87)."""
    out, cls, colsum = torch.ops.tokred.attention(qkv, num_heads, scale, key_bias, mask, q_ids, want_out, want_cls, want_colsum)
    return (out if want_out else None), (cls if want_cls else None), (colsum if want_colsum else None)


# ----------------------------------------------------------------------------------------------- residual + LayerNorm
@torch.library.custom_op("tokred::add_layernorm", mutates_args=(), device_types="cuda")
def _add_layernorm(x: Tensor, branch: Optional[Tensor], weight: Tensor, bias: Tensor, eps: float) -> Tuple[Tensor, Tensor]:
    _need_cuda("add_layernorm", x, branch, weight, bias)
    if x.dtype != torch.float32 or x.shape[-1] != weight.numel():
        raise TokredError(f"add_layernorm: x {tuple(x.shape)} {x.dtype}; expected fp32 [..., {weight.numel()}]")
    x = _c(x)
    c = x.shape[-1]
    rows = x.numel() // c
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if branch is not None:
        if branch.shape != x.shape:
            raise TokredError("add_layernorm: branch shape differs from x")
        branch = _c(branch if branch.dtype in (torch.float32, torch.bfloat16) else branch.float())
        x_out = torch.empty_like(x)
    else:
        x_out = x.new_empty((0,))
    _lib.call("tokred_add_layernorm", _ptr(x), _ptr(branch), _dt(branch) if branch is not None else 0, _ptr(_c(weight.float())),
              _ptr(_c(bias.float())), float(eps), rows, c, _ptr(x_out) if branch is not None else None, _ptr(y), _stream())
    return x_out, y


@_add_layernorm.register_fake
def _(x, branch, weight, bias, eps):
    return (torch.empty_like(x) if branch is not None else x.new_empty((0,))), x.new_empty(x.shape, dtype=torch.bfloat16)


def add_layernorm(x: Tensor, branch: Optional[Tensor], weight: Tensor, bias: Tensor, eps: float):
    """x_new = x + branch (fp32 residual stream); y = LayerNorm(x_new) rounded to bf16 (what the next autocast Linear
    would cast it to) -- one pass instead of add / layer_norm / cast (e.g. models/topk.py:87,94).  branch=None: plain
    LayerNorm of x.  Returns (x_new | x, y)."""
    x_out, y = torch.ops.tokred.add_layernorm(x, branch, weight, bias, eps)
    return (x_out if branch is not None else x), y


@torch.library.custom_op("tokred::tome_merge_ln", mutates_args=(), device_types="cuda")
def _tome_merge_ln(x: Tensor, branch: Optional[Tensor], size: Optional[Tensor], unm: Tensor, src: Tensor, dst: Tensor,
                   weight: Tensor, bias: Tensor, eps: float, want_map: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    _need_cuda("tome_merge_ln", x, branch, size, unm, src, dst, weight, bias)
    b, n, c = x.shape
    r = src.shape[1]
    if x.dtype != torch.float32 or c != weight.numel():
        raise TokredError(f"tome_merge_ln: x {tuple(x.shape)} {x.dtype}; expected fp32 [B,N,{weight.numel()}]")
    if unm.shape != (b, (n + 1) // 2 - r) or dst.shape != (b, r) or src.shape[0] != b:
        raise TokredError("tome_merge_ln: index tensors do not match x")
    if branch is not None and (branch.shape != x.shape or branch.dtype != torch.bfloat16):
        raise TokredError(f"tome_merge_ln: branch {tuple(branch.shape)} {branch.dtype}; expected bf16 {tuple(x.shape)}")
    x, unm, src, dst = _c(x), _c(unm), _c(src), _c(dst)
    if branch is not None:
        branch = _c(branch)
    if size is not None:
        if size.numel() != b * n:
            raise TokredError(f"tome_merge_ln: size {tuple(size.shape)} does not match x {tuple(x.shape)}")
        size = _c(size.float())
    out = torch.empty((b, n - r, c), dtype=torch.float32, device=x.device)
    size_out = torch.empty((b, n - r, 1), dtype=torch.float32, device=x.device)
    rci = torch.empty((b, n - 1) if want_map else (0,), dtype=torch.float32, device=x.device)
    y = torch.empty((b, n - r, c), dtype=torch.bfloat16, device=x.device)
    _lib.call("tokred_tome_merge_ln", _ptr(x), _ptr(branch), _ptr(size), _ptr(unm), _ptr(src), _ptr(dst), b, n, c, r,
              _ptr(_c(weight.float())), _ptr(_c(bias.float())), float(eps), _ptr(out), _ptr(size_out),
              _ptr(rci) if want_map else None, _ptr(y), _stream())
    return out, size_out, rci, y


@_tome_merge_ln.register_fake
def _(x, branch, size, unm, src, dst, weight, bias, eps, want_map):
    b, n, c = x.shape
    r = src.shape[1]
    return (x.new_empty((b, n - r, c)), x.new_empty((b, n - r, 1)),
            x.new_empty((b, n - 1) if want_map else (0,), dtype=torch.float32), x.new_empty((b, n - r, c), dtype=torch.bfloat16))


def tome_merge_ln_supported(x: Tensor, branch: Optional[Tensor]) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[2] % 128 == 0 and x.shape[2] <= 768
            and (branch is None or (branch.dtype == torch.bfloat16 and branch.shape == x.shape)))


def tome_merge_ln(x: Tensor, branch: Optional[Tensor], size: Optional[Tensor], unm: Tensor, src: Tensor, dst: Tensor,
                  weight: Tensor, bias: Tensor, eps: float, want_map: bool = True):
    """models/tome.py:88 (x + attn branch), :100-101 (merge_wavg), :104 (norm2, as the MLP's autocast Linear consumes it)
    in ONE launch on the fp32 residual stream: (x_out [B,N-r,C] fp32, size_out [B,N-r,1], map [B,N-1], y bf16).
    Bit-identical to ``add -> tome_merge -> add_layernorm``."""
    return torch.ops.tokred.tome_merge_ln(x, branch, size, unm, src, dst, weight, bias, eps, want_map)


# ----------------------------------------------------------------------------------------------- select + residual add
def select_add_supported(x: Tensor, branch: Tensor, scores: Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[2] % 4 == 0 and branch.dtype == torch.bfloat16
            and branch.shape == x.shape and scores.dtype in (torch.float32, torch.bfloat16)
            and not (torch.is_grad_enabled() and (x.requires_grad or branch.requires_grad)))


@torch.library.custom_op("tokred::topk_gather_add", mutates_args=(), device_types="cuda")
def _topk_gather_add(x: Tensor, branch: Tensor, scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    _need_cuda("topk_gather_add", x, branch, scores)
    b, n, c = x.shape
    if x.dtype != torch.float32 or branch.dtype != torch.bfloat16 or branch.shape != x.shape:
        raise TokredError(f"topk_gather_add: x {tuple(x.shape)} {x.dtype}, branch {tuple(branch.shape)} {branch.dtype}")
    if scores.dim() != 2 or scores.shape[0] != b or scores.shape[1] != n - 1:
        raise TokredError(f"topk_gather_add: scores {tuple(scores.shape)} does not match x {tuple(x.shape)}")
    x, branch = _c(x), _c(branch)
    out = torch.empty((b, k + 1, c), dtype=torch.float32, device=x.device)
    idx = torch.empty((b, k), dtype=torch.int64, device=x.device)
    _lib.call("tokred_topk_gather_add", _ptr(x), _ptr(branch), _ptr(scores), _dt(scores), scores.stride(1), scores.stride(0),
              b, n, c, k, _ptr(out), _ptr(idx), _stream())
    return out, idx


@_topk_gather_add.register_fake
def _(x, branch, scores, k):
    b, n, c = x.shape
    return x.new_empty((b, k + 1, c)), x.new_empty((b, k), dtype=torch.int64)


def topk_gather_add(x: Tensor, branch: Tensor, scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """``topk_gather(x + branch, scores, k)`` without materialising the sum (models/topk.py:87 + :62, :89-93): the fp32 add
    happens on the kept rows as they are gathered.  Inference path of the bf16-autocast blocks."""
    return torch.ops.tokred.topk_gather_add(x, branch, scores, k)


@torch.library.custom_op("tokred::evit_select_fuse_add", mutates_args=(), device_types="cuda")
def _evit_select_fuse_add(x: Tensor, branch: Tensor, scores: Tensor, k: int) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda("evit_select_fuse_add", x, branch, scores)
    b, n, c = x.shape
    if x.dtype != torch.float32 or branch.dtype != torch.bfloat16 or branch.shape != x.shape:
        raise TokredError(f"evit_select_fuse_add: x {tuple(x.shape)} {x.dtype}, branch {tuple(branch.shape)} {branch.dtype}")
    if scores.dim() != 2 or scores.shape[0] != b or scores.shape[1] != n - 1:
        raise TokredError(f"evit_select_fuse_add: scores {tuple(scores.shape)} does not match x {tuple(x.shape)}")
    x, branch, scores = _c(x), _c(branch), _c(scores)
    out = torch.empty((b, k + 2, c), dtype=torch.float32, device=x.device)
    idx = torch.empty((b, k + 1), dtype=torch.int64, device=x.device)
    compl = torch.empty((b, n - 1 - k), dtype=torch.int64, device=x.device)
    _lib.call("tokred_evit_select_fuse_add", _ptr(x), _ptr(branch), _ptr(scores), _dt(scores), b, n, c, k, _ptr(out), _ptr(idx),
              _ptr(compl), _stream())
    return out, idx, compl


@_evit_select_fuse_add.register_fake
def _(x, branch, scores, k):
    b, n, c = x.shape
    return (x.new_empty((b, k + 2, c)), x.new_empty((b, k + 1), dtype=torch.int64),
            x.new_empty((b, n - 1 - k), dtype=torch.int64))


def evit_select_fuse_add(x: Tensor, branch: Tensor, scores: Tensor, k: int):
    """``evit_select_fuse(x + branch, scores, k)`` without materialising the sum (models/evit.py:109 + :84, :111-123)."""
    return torch.ops.tokred.evit_select_fuse_add(x, branch, scores, k)


# ----------------------------------------------------------------------------------------------- in front of block 0
@torch.library.custom_op("tokred::patchify", mutates_args=(), device_types="cuda")
def _patchify(img: Tensor, ph: int, pw: int) -> Tensor:
    _need_cuda("patchify", img)
    if img.dim() != 4 or img.dtype != torch.float32:
        raise TokredError(f"patchify: img {tuple(img.shape)} {img.dtype}; expected fp32 [B,C,H,W]")
    img = _c(img)
    b, c, h, w = img.shape
    out = torch.empty((b, (h // ph) * (w // pw), c * ph * pw), dtype=torch.bfloat16, device=img.device)
    _lib.call("tokred_patchify", _ptr(img), b, c, h, w, ph, pw, _ptr(out), _stream())
    return out


@_patchify.register_fake
def _(img, ph, pw):
    b, c, h, w = img.shape
    return img.new_empty((b, (h // ph) * (w // pw), c * ph * pw), dtype=torch.bfloat16)


def patchify(img: Tensor, ph: int, pw: int) -> Tensor:
    """[B,C,H,W] fp32 -> [B, patches, C*ph*pw] bf16: ``img.to(bf16).view(b,c,gh,ph,gw,pw).permute(0,2,4,1,3,5).reshape``
    (the patch-embedding GEMM's operand) in one pass."""
    return torch.ops.tokred.patchify(img, ph, pw)


@torch.library.custom_op("tokred::embed_layernorm", mutates_args=(), device_types="cuda")
def _embed_layernorm(patches: Tensor, tokens: Tensor, pos: Optional[Tensor], weight: Tensor, bias: Tensor,
                     eps: float) -> Tuple[Tensor, Tensor]:
    _need_cuda("embed_layernorm", patches, tokens, pos, weight, bias)
    b, p, c = patches.shape
    if patches.dtype not in (torch.bfloat16, torch.float32) or weight.numel() != c or tokens.dtype != torch.float32:
        raise TokredError(f"embed_layernorm: patches {tuple(patches.shape)} {patches.dtype}, tokens {tokens.dtype}")
    if tokens.dim() == 2:                                  # [T,C]: shared by the batch
        t, tok, tok_bs = tokens.shape[0], _c(tokens), 0
    else:                                                  # [B,T,C]: per image, rows dense (e.g. the x[:, :T] view)
        t = tokens.shape[1]
        if tokens.shape[0] != b or tokens.shape[2] != c:
            raise TokredError(f"embed_layernorm: tokens {tuple(tokens.shape)} do not match patches {tuple(patches.shape)}")
        if not (tokens.stride(2) == 1 and (t == 1 or tokens.stride(1) == c) and tokens.stride(0) % 4 == 0 and tokens.stride(0) >= t * c):
            tokens = tokens.contiguous()
        tok, tok_bs = tokens, (int(tokens.stride(0)) if b > 1 else t * c)
    if tok.shape[-1] != c or (pos is not None and tuple(pos.shape) != (t + p, c)):
        raise TokredError(f"embed_layernorm: tokens {tuple(tokens.shape)} / pos do not match patches {tuple(patches.shape)}")
    patches = _c(patches)
    x_out = torch.empty((b, t + p, c), dtype=torch.float32, device=patches.device)
    y = torch.empty((b, t + p, c), dtype=torch.bfloat16, device=patches.device)
    _lib.call("tokred_embed_layernorm", _ptr(patches), _dt(patches), _ptr(tok), tok_bs, _ptr(_c(pos.float())) if pos is not None else None,
              _ptr(_c(weight.float())), _ptr(_c(bias.float())), float(eps), b, p, t, c, _ptr(x_out), _ptr(y), _stream())
    return x_out, y


@_embed_layernorm.register_fake
def _(patches, tokens, pos, weight, bias, eps):
    b, p, c = patches.shape
    t = tokens.shape[-2]
    return (patches.new_empty((b, t + p, c), dtype=torch.float32), patches.new_empty((b, t + p, c), dtype=torch.bfloat16))


def embed_layernorm(patches: Tensor, tokens: Tensor, pos: Optional[Tensor], weight: Tensor, bias: Tensor, eps: float):
    """(x, y): x = cat(tokens, patches) (+ pos) in fp32, y = LayerNorm(x) rounded to bf16 -- the cat, the positional add, the
    next norm1 and its autocast cast in one pass.  tokens [T,C] (the cls / dist parameters: models/deit_viz.py
    forward_features) or [B,T,C] (the class rows kept aside around a cluster layer, e.g. models/sinkhorn.py:166-168; the
    x[:, :T] view is read in place); patches bf16 or fp32; pos [T+P,C] or None."""
    return torch.ops.tokred.embed_layernorm(patches, tokens, pos, weight, bias, eps)


@torch.library.custom_op("tokred::residual_add", mutates_args=(), device_types="cuda")
def _residual_add(x: Tensor, branch: Tensor) -> Tensor:
    _need_cuda("residual_add", x, branch)
    if x.dtype != torch.float32 or branch.dtype != torch.bfloat16 or branch.shape != x.shape or x.numel() % 4 != 0:
        raise TokredError(f"residual_add: x {tuple(x.shape)} {x.dtype}, branch {tuple(branch.shape)} {branch.dtype}")
    x, branch = _c(x), _c(branch)
    out = torch.empty_like(x)
    _lib.call("tokred_residual_add", _ptr(x), _ptr(branch), x.numel(), _ptr(out), _stream())
    return out


@_residual_add.register_fake
def _(x, branch):
    return torch.empty_like(x)


def residual_add(x: Tensor, branch: Tensor) -> Tensor:
    """``x + branch`` for an fp32 stream and a bf16 branch (the same fp32 addition as ATen's mixed-dtype add, with 16-byte
    accesses); anything else falls back to ATen."""
    if (x.is_cuda and x.dtype == torch.float32 and branch.dtype == torch.bfloat16 and branch.shape == x.shape
            and x.numel() % 4 == 0 and x.numel() > 0 and not (torch.is_grad_enabled() and (x.requires_grad or branch.requires_grad))):
        return torch.ops.tokred.residual_add(x, branch)
    return x + branch


# ----------------------------------------------------------------------------------------------- f4: autograd formulas
# SURVEY §8f row 4.  The reference fine-tunes its reduced models with the reduction operators inside the autograd graph
# (train.py; the discrete selections themselves carry no gradient: topk / argsort / argmax indices).  The forward of
# every op below stays the tokred kernel; the backward is the adjoint of the gather / merge written with ATen ops
# (training throughput is not a metric of this path).  Index outputs are non-differentiable.
def _bwd_topk_gather(ctx, g_out, g_idx):
    (idx,) = ctx.saved_tensors
    b, n, c = ctx.x_shape
    gx = g_out.new_zeros((b, n, c))
    gx[:, 0] = g_out[:, 0]
    gx.scatter_(1, (idx + 1).unsqueeze(-1).expand(-1, -1, c), g_out[:, 1:])        # kept indices are distinct
    return gx, None, None


def _setup_topk_gather(ctx, inputs, output):
    x, scores, k = inputs
    ctx.x_shape = tuple(x.shape)
    ctx.save_for_backward(output[1])


_topk_gather.register_autograd(_bwd_topk_gather, setup_context=_setup_topk_gather)


def _bwd_evit(ctx, g_out, g_idx, g_compl):
    x, scores, idx, compl = ctx.saved_tensors
    b, n, c = x.shape
    k = idx.shape[1] - 1
    kept = idx[:, :k]
    gx = g_out.new_zeros((b, n, c))
    gx[:, 0] = g_out[:, 0]
    gx.scatter_(1, (kept + 1).unsqueeze(-1).expand(-1, -1, c), g_out[:, 1:k + 1])
    g_extra = g_out[:, k + 1:k + 2]                                                # [B,1,C]: fused token = sum s_p x_p
    s_c = torch.gather(scores.to(g_out.dtype), 1, compl).unsqueeze(-1)             # [B,P-k,1]
    gx.scatter_(1, (compl + 1).unsqueeze(-1).expand(-1, -1, c), s_c * g_extra)
    x_c = torch.gather(x, 1, (compl + 1).unsqueeze(-1).expand(-1, -1, c))
    gs = torch.zeros_like(scores, dtype=g_out.dtype).scatter_(1, compl, (x_c * g_extra).sum(-1))
    return gx, gs.to(scores.dtype), None


def _setup_evit(ctx, inputs, output):
    x, scores, k = inputs
    ctx.save_for_backward(x, scores, output[1], output[2])


_evit_select_fuse.register_autograd(_bwd_evit, setup_context=_setup_evit)


def _bwd_gather_rows(ctx, g_out):
    (ids,) = ctx.saved_tensors
    shape, m = ctx.src_shape, ctx.m
    gs = g_out.new_zeros(shape)
    ix = ids[:, :m]
    if len(shape) == 3:
        gs.scatter_add_(1, ix.unsqueeze(-1).expand(-1, -1, shape[2]), g_out)     # ids may repeat (ATS 0-padding)
    else:
        gs.scatter_add_(2, ix[:, None, :, None].expand(-1, shape[1], -1, shape[3]), g_out)
    return gs, None, None


def _setup_gather_rows(ctx, inputs, output):
    src, ids, m = inputs
    ctx.src_shape, ctx.m = tuple(src.shape), m
    ctx.save_for_backward(ids)


_gather_rows.register_autograd(_bwd_gather_rows, setup_context=_setup_gather_rows)


def _tome_rows(unm, src, dst, n):
    """[B,n] output row of every input token (models/tome.py:279-289 order: unmerged even tokens, then all odd tokens)."""
    b, u = unm.shape
    rows = torch.empty(b, n, dtype=torch.long, device=unm.device)
    rows[:, 1::2] = u + torch.arange(n // 2, device=unm.device)
    rows.scatter_(1, 2 * unm, torch.arange(u, device=unm.device).expand(b, -1))
    rows.scatter_(1, 2 * src, u + dst)
    return rows


def _bwd_tome_merge(ctx, g_out, g_size, g_map):
    size, size_out, unm, src, dst = ctx.saved_tensors
    b, n, c = ctx.x_shape
    rows = _tome_rows(unm, src, dst, n)
    g = g_out / size_out if ctx.divide else g_out                                 # out = merge(x * size) / size_out
    gx = torch.gather(g, 1, rows.unsqueeze(-1).expand(-1, -1, c))
    if ctx.has_size:
        gx = gx * size.reshape(b, n, 1)
    # sizes are integer token counts (sums of ones): no parameter reaches them, so they carry no gradient
    return gx, None, None, None, None, None, None


def _setup_tome_merge(ctx, inputs, output):
    x, size, unm, src, dst, want_map, divide = inputs
    ctx.x_shape, ctx.divide, ctx.has_size = tuple(x.shape), bool(divide), size is not None
    ctx.save_for_backward(size if size is not None else x.new_empty(0), output[1], unm, src, dst)


_tome_merge.register_autograd(_bwd_tome_merge, setup_context=_setup_tome_merge)


def _bwd_dyvit_pool(ctx, g_out):
    h, policy = ctx.saved_tensors
    b, p, c = h.shape
    half = c // 2
    pol = policy.reshape(b, p, 1).to(g_out.dtype)
    s = pol.sum(dim=1, keepdim=True)                                               # [B,1,1]
    gsum = g_out[:, :, half:].sum(dim=1, keepdim=True)                             # every row received the same pooled vector
    hf = h[:, :, half:].to(g_out.dtype)
    gh = torch.cat([g_out[:, :, :half], (pol / s) * gsum], dim=-1).to(h.dtype)
    gbar = (hf * pol).sum(dim=1, keepdim=True) / s
    gp = ((hf - gbar) * gsum).sum(dim=-1) / s[:, :, 0]                             # [B,P]
    return gh, gp.reshape(policy.shape).to(policy.dtype), None


def _setup_dyvit_pool(ctx, inputs, output):
    h, policy, eps = inputs
    ctx.save_for_backward(h, policy)


_dyvit_pool_concat.register_autograd(_bwd_dyvit_pool, setup_context=_setup_dyvit_pool)
