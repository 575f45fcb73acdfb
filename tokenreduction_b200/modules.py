"""Drop-in reduction modules and functions: same names, constructor arguments, parameter names and forward
signatures as the reference's (cited per class), with the ATen call sequences replaced by the tokred CUDA ops.

Only the attention sub-modules differ internally from the reference: they hand the *scores* the block needs
(CLS-attention, key mean, attention probabilities) to the block, which then issues ONE fused select+gather /
match+merge launch instead of topk -> expand -> gather -> cat.  Block-level signatures are unchanged.

Precision: ``_lowp()`` is True under ``torch.autocast('cuda', dtype=torch.bfloat16)`` — the kernels then round
exactly where the reference's autocast matmuls round.  fp16 autocast (the reference's
validate.py default) converts at the op boundary, see ``_lowp``.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .vit import DropPath, Mlp


def _lowp() -> bool:
    """True under CUDA autocast.  bf16 autocast (BASELINE) is reproduced rounding for rounding.  fp16 autocast (the
    reference's validate.py:52-54 default) runs the same reduced-precision kernels: half tensors are converted at the op
    boundary (fp16 -> fp32 is exact) and the matmul roundings are bf16's, a documented deviation."""
    if not torch.is_autocast_enabled("cuda"):
        return False
    dt = torch.get_autocast_dtype("cuda")
    if dt not in (torch.bfloat16, torch.float16):
        raise RuntimeError(f"tokred kernels support float32, bfloat16 and float16 autocast, not {dt}")
    return True


def _f16_to_bf16(t: Tensor) -> Tensor:
    """fp16 tensors enter the kernels as fp32 (exact); fp32 / bf16 pass through untouched."""
    return t.float() if t.dtype == torch.float16 else t


def _amp_out(t: Tensor) -> Tensor:
    """a reduced-precision op output in the dtype the reference's autocast matmul would have produced (fp16 under fp16
    autocast; the kernels emit bf16)."""
    if t.dtype == torch.bfloat16 and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.float16:
        return t.half()
    return t


def _train_guard(mod: nn.Module, differentiable: bool = False) -> None:
    """Top-K / EViT / ToMe / DynamicViT ops carry autograd formulas (ops.py) and may sit in a training
    graph; the cluster layers, soft merges and ATS are inference-only and refuse."""
    if mod.training and torch.is_grad_enabled() and not differentiable:
        raise NotImplementedError(
            f"{type(mod).__name__}: this tokred reduction kernel is inference-only (no autograd formula); "
            "call .eval() and run under torch.no_grad()")


FUSED_ATTENTION = os.environ.get("TOKRED_FUSED_ATTENTION", "1") != "0"


def fused_attention_ok(mod: nn.Module, x: Tensor, num_heads: int) -> bool:
    """True where the fused attention producer (ops.attention) reproduces the module's ATen sequence:
    bf16 autocast (or bf16 tensors) in inference, head dim 64, N <= 256.  fp32 models, fp16 autocast, training with
    attention dropout and other shapes keep the reference's own sequence."""
    if not (FUSED_ATTENTION and x.is_cuda and x.dim() == 3 and x.shape[1] <= 256 and x.shape[2] == 64 * num_heads):
        return False
    if torch.is_autocast_enabled("cuda"):
        if torch.get_autocast_dtype("cuda") != torch.bfloat16:
            return False
    elif x.dtype != torch.bfloat16:
        return False
    drop = getattr(mod, "attn_drop", None)
    if mod.training and (torch.is_grad_enabled() or (drop is not None and drop.p > 0)):
        return False
    return True


def _norm_fusable(norm: nn.Module, x: Tensor) -> bool:
    return (FUSED_ATTENTION and isinstance(norm, nn.LayerNorm) and norm.elementwise_affine and norm.bias is not None
            and x.is_cuda and x.dtype == torch.float32 and x.shape[-1] % 128 == 0 and x.shape[-1] <= 1024
            and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
            and not (norm.training and torch.is_grad_enabled()))


def norm_lowp(norm: nn.Module, x: Tensor) -> Tensor:
    """``norm(x)`` as the next autocast Linear consumes it: under bf16 autocast one kernel computes the fp32 LayerNorm
    and rounds it to bf16 (the reference: layer_norm in fp32, then a separate cast)."""
    if _norm_fusable(norm, x):
        return ops.add_layernorm(x, None, norm.weight, norm.bias, norm.eps)[1]
    return norm(x)


def add_norm(x: Tensor, branch: Tensor, norm: nn.Module) -> Tuple[Tensor, Tensor]:
    """(x + branch, norm(x + branch)) -- e.g. models/topk.py:87 + :94 -- in one pass under bf16 autocast."""
    if _norm_fusable(norm, x) and branch.shape == x.shape and branch.is_cuda:
        return ops.add_layernorm(x, branch, norm.weight, norm.bias, norm.eps)
    x = x + branch
    return x, norm(x)


DEFER_RESIDUAL = os.environ.get("TOKRED_DEFER_RESIDUAL", "1") != "0"


class Residual:
    """``x + branch`` that has not been added yet.  Every block ends with ``x + drop_path(mlp(...))`` (e.g.
    models/tome.py:104) and the next block starts with ``norm1(x)`` (:87): under bf16 autocast the block returns the two
    terms and the consumer's add_layernorm forms the sum, the new residual stream and the normalised bf16 activations
    in ONE pass (the eager add was 15 % of a ToMe step).  The sum is the same fp32 addition: results are bit-identical.
    Anything else that reads a block's output calls :func:`value` first."""
    __slots__ = ("x", "branch")

    def __init__(self, x: Tensor, branch: Tensor):
        self.x, self.branch = x, branch

    @property
    def shape(self):
        return self.x.shape

    def value(self) -> Tensor:
        return ops.residual_add(self.x, self.branch)


class Embedded:
    """``cat(cls [, dist], patches) + pos_embed`` that has not been formed yet (models/deit_viz.py forward_features): under
    bf16 autocast the first block's ``enter_norm`` builds the fp32 stream and norm1's bf16 output from the patch GEMM's
    output in one pass (ops.embed_layernorm) instead of cat -> add -> layer_norm -> cast."""
    __slots__ = ("patches", "tokens", "pos")

    def __init__(self, patches: Tensor, tokens: Tensor, pos: Optional[Tensor]):
        # [B,P,C] bf16 | fp32; [T,C] fp32 shared or [B,T,C] fp32 per image (the class rows kept aside around a cluster
        # layer, e.g. models/sinkhorn.py:166-168); [1,T+P,C] fp32 or None
        self.patches, self.tokens, self.pos = patches, tokens, pos

    @property
    def shape(self):
        b, p, c = self.patches.shape
        return torch.Size((b, p + self.tokens.shape[-2], c))

    @property
    def dtype(self):
        return self.tokens.dtype

    @property
    def device(self):
        return self.patches.device

    def value(self) -> Tensor:
        b = self.patches.shape[0]
        tok = self.tokens.unsqueeze(0).expand(b, -1, -1) if self.tokens.dim() == 2 else self.tokens
        x = torch.cat((tok, self.patches.to(tok.dtype)), dim=1)
        return x if self.pos is None else x + self.pos


def value(x):
    """the tensor a block input / output stands for (materialises a deferred residual sum or embedding)."""
    return x.value() if isinstance(x, (Residual, Embedded)) else x


def embed_tokens(patches: Tensor, tokens: Tensor, pos: Tensor, norm_probe: nn.Module):
    """``cat(tokens, patches) + pos`` -- deferred to the first block's norm1 where the fused kernel applies."""
    if (DEFER_RESIDUAL and patches.is_cuda and patches.dtype == torch.bfloat16 and patches.dim() == 3
            and _norm_fusable(norm_probe, pos) and pos.shape[1] == patches.shape[1] + tokens.shape[0]):
        return Embedded(patches, tokens, pos)
    b = patches.shape[0]
    return torch.cat((tokens.unsqueeze(0).expand(b, -1, -1), patches), dim=1) + pos


def cls_value(x) -> Tensor:
    """``value(x)[:, 0]`` without forming the other rows (elementwise add: the same bits)."""
    if isinstance(x, Residual):
        return x.x[:, 0] + x.branch[:, 0]
    return value(x)[:, 0]


def defer_add(x: Tensor, branch: Tensor, norm: nn.Module, owner: Optional[nn.Module] = None):
    """``x + branch`` -- deferred to the next block's norm1 where that block can fuse it (``norm``: a LayerNorm of the
    stream's width, used as the probe for the fused kernel's conditions).  Only a block whose ``defer_out`` flag was set by
    the model that owns the block loop (VisionTransformer.embed_tokens) defers: a block called on its own returns a plain
    tensor, like the reference's."""
    if (DEFER_RESIDUAL and getattr(owner, "defer_out", False) and _norm_fusable(norm, x) and branch.shape == x.shape
            and branch.is_cuda):
        return Residual(x, branch)
    return x + branch


def concat_tokens(tokens: Tensor, x: Tensor, norm_probe: nn.Module):
    """``torch.cat((global_tokens, x.to(global_tokens.dtype)), dim=1)`` after a cluster layer (models/sinkhorn.py:168,
    patchmerger.py:119, sit.py:119, dpcknn.py:262, kmedoids.py:249) -- deferred to the next block's norm1 where the fused kernel
    applies: ATen's cat of the fp32 stream ran at 1.4 TB/s (100 us per stage at B=128 DeiT-B) after a separate bf16 -> fp32
    conversion of the merged tokens."""
    if (DEFER_RESIDUAL and x.is_cuda and x.dim() == 3 and x.dtype in (torch.bfloat16, torch.float32)
            and tokens.dtype == torch.float32 and tokens.dim() == 3 and tokens.shape[0] == x.shape[0]
            and tokens.shape[2] == x.shape[2] and _norm_fusable(norm_probe, tokens)):
        return Embedded(x, tokens, None)
    return torch.cat((tokens, x.to(tokens.dtype)), dim=1)


def enter_norm(norm: nn.Module, x) -> Tuple[Tensor, Tensor]:
    """(x, norm(x) as the next autocast Linear consumes it) for a block input that may be a deferred residual sum."""
    if isinstance(x, Residual):
        if _norm_fusable(norm, x.x):
            return ops.add_layernorm(x.x, x.branch, norm.weight, norm.bias, norm.eps)
        x = x.value()
    elif isinstance(x, Embedded):
        if _norm_fusable(norm, x.tokens):
            return ops.embed_layernorm(x.patches, x.tokens, None if x.pos is None else x.pos[0], norm.weight, norm.bias, norm.eps)
        x = x.value()
    return x, norm_lowp(norm, x)


class _AttentionBase(nn.Module):
    """qkv / proj layout shared by every reference attention variant (e.g. models/topk.py:27-52)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def _qkv(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        return qkv[0], qkv[1], qkv[2]

    def _out(self, attn, v):
        b, _, n, _ = attn.shape
        x = (attn @ v).transpose(1, 2).reshape(b, n, -1)
        return self.proj_drop(self.proj(x))

    def _fused(self, x) -> bool:
        return fused_attention_ok(self, x, self.num_heads)

    def _attend_fused(self, x, key_bias=None, mask=None, q_ids=None, want_cls=False, want_colsum=False):
        """qkv Linear -> ONE attention launch (no [B,H,N,N] tensor) -> proj.  Returns (x, cls_row, colsum, qkv)."""
        qkv = self.qkv(x)
        if qkv.dtype != torch.bfloat16:
            qkv = qkv.to(torch.bfloat16)
        out, cls, colsum = ops.attention(qkv, self.num_heads, self.scale, key_bias, mask, q_ids, True, want_cls, want_colsum)
        return self.proj_drop(self.proj(out)), cls, colsum, qkv


# =============================================================================================== Top-K
class Attention_TopK(_AttentionBase):
    """models/topk.py:27-67.  forward(x) -> (x, cls_attn [B,N-1] | None, left_tokens | None)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0, keep_rate=1.0):
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        assert 0 < keep_rate <= 1, "keep_rate must > 0 and <= 1, got {0}".format(keep_rate)
        self.keep_rate = keep_rate
        self.init_n = 14 * 14

    def forward(self, x):
        n = x.shape[1]
        if self._fused(x):
            left_tokens = int(self.keep_rate * self.init_n)
            reduce = self.keep_rate < 1 and left_tokens != n - 1
            x, cls_row, _, _ = self._attend_fused(x, want_cls=reduce)
            if not reduce:
                return x, None, None
            assert left_tokens >= 1
            return x, cls_row[:, :, 1:].mean(dim=1), left_tokens      # = attn[:, :, 0, 1:].mean(1), models/topk.py:60
        q, k, v = self._qkv(x)
        attn = self.attn_drop(((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1))
        x = self._out(attn, v)
        if self.keep_rate < 1:
            left_tokens = int(self.keep_rate * self.init_n)
            if left_tokens == n - 1:
                return x, None, None
            assert left_tokens >= 1
            cls_attn = attn[:, :, 0, 1:].mean(dim=1)           # identical ATen call -> bit-identical decision input
            return x, cls_attn, left_tokens
        return x, None, None


class Block_TopK(nn.Module):
    """models/topk.py:70-99.  forward(x) -> (x, n_tokens, idx | None)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, keep_rate=0.0):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention_TopK(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop,
                                   keep_rate=keep_rate)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        x, y = enter_norm(self.norm1, x)
        tmp, cls_attn, left_tokens = self.attn(y)
        idx = None
        if cls_attn is not None:
            _train_guard(self, True)
            branch = self.drop_path(tmp)
            if DEFER_RESIDUAL and ops.select_add_supported(x, branch, cls_attn):
                x, idx = ops.topk_gather_add(x, branch, cls_attn, left_tokens)    # + the residual add, on the kept rows only
            else:
                x = x + branch
                x, idx = ops.topk_gather(x, cls_attn, left_tokens)      # select + gather + cat in one launch
            y = norm_lowp(self.norm2, x)
        else:
            x, y = add_norm(x, self.drop_path(tmp), self.norm2)
        x = defer_add(x, self.drop_path(self.mlp(y)), self.norm1, self)
        return x, x.shape[1] - 1, idx


# =============================================================================================== EViT
def complement_idx(idx: Tensor, dim: int) -> Tensor:
    """models/evit.py:25-46 — ascending indices of range(dim) not in idx (trailing dimension)."""
    keep = torch.zeros(idx.shape[:-1] + (dim,), dtype=torch.bool, device=idx.device)
    keep.scatter_(-1, idx, True)
    ar = torch.arange(dim, device=idx.device).expand(keep.shape)
    return ar[~keep].reshape(idx.shape[:-1] + (dim - idx.shape[-1],))


class Attention_EVIT(Attention_TopK):
    """models/evit.py:49-91 — same scores as Top-K."""


class Block_EVIT(nn.Module):
    """models/evit.py:94-129.  forward(x) -> (x, n_tokens, idx (trailing -1) | None, compl | None)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, keep_rate=0.0):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention_EVIT(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop,
                                   keep_rate=keep_rate)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        x, y = enter_norm(self.norm1, x)
        tmp, cls_attn, left_tokens = self.attn(y)
        idx = compl = None
        if cls_attn is not None:
            _train_guard(self, True)
            branch = self.drop_path(tmp)
            if DEFER_RESIDUAL and ops.select_add_supported(x, branch, cls_attn):
                x, idx, compl = ops.evit_select_fuse_add(x, branch, cls_attn, left_tokens)
            else:
                x = x + branch
                x, idx, compl = ops.evit_select_fuse(x, cls_attn, left_tokens)
            y = norm_lowp(self.norm2, x)
        else:
            x, y = add_norm(x, self.drop_path(tmp), self.norm2)
        x = defer_add(x, self.drop_path(self.mlp(y)), self.norm1, self)
        return x, x.shape[1] - 1, idx, compl


# =============================================================================================== ToMe
class _ToMeMerge:
    """The ``merge`` closure of models/tome.py:279-289, carrying the three index lists on the device."""

    def __init__(self, unm_idx: Tensor, src_idx: Tensor, dst_idx: Tensor, n_tokens: int, class_token: bool = True,
                 distill_token: bool = False):
        self.unm_idx, self.src_idx, self.dst_idx, self.n_tokens = unm_idx, src_idx, dst_idx, n_tokens
        self.class_token, self.distill_token = class_token, distill_token

    @property
    def r(self) -> int:
        return self.src_idx.shape[1]

    def reorder(self, out: Tensor) -> Tensor:
        """distilled models keep [cls, dist] in front: cat([unm[:1], dst[:1], unm[1:], dst[1:]]) (models/tome.py:286-287)."""
        if not self.distill_token:
            return out
        u = self.unm_idx.shape[1]
        return torch.cat([out[:, :1], out[:, u:u + 1], out[:, 1:u], out[:, u + 1:]], dim=1)

    def __call__(self, x: Tensor, mode: str = "mean") -> Tensor:        # mode is ignored by the reference too
        out, _, _ = ops.tome_merge(x, None, self.unm_idx, self.src_idx, self.dst_idx, False, False)
        return self.reorder(out)


def do_nothing(x, mode=None):
    return x


def bipartite_soft_matching(metric: Tensor, r: int, class_token: bool = False,
                            distill_token: bool = False) -> Tuple[Callable, Callable]:
    """models/tome.py:230-306.  Returns (merge, unmerge)."""
    t = metric.shape[1]
    r = ops.tome_effective_r(t, r, class_token, distill_token)
    if r <= 0:
        return do_nothing, do_nothing
    lowp = _lowp() or metric.dtype in (torch.bfloat16, torch.float16)
    with torch.no_grad():                                                 # models/tome.py:258
        unm_idx, src_idx, dst_idx = ops.tome_match(_f16_to_bf16(metric), r, class_token, lowp, True, distill_token)
    merge = _ToMeMerge(unm_idx, src_idx, dst_idx, t, class_token, distill_token)

    def unmerge(x: Tensor) -> Tensor:
        """models/tome.py:291-304 — one gather launch: input token t reads the merged row it went to."""
        return ops.gather_rows(x, tome_token_rows(merge, x))

    return merge, unmerge


def tome_token_rows(merge: "_ToMeMerge", like: Tensor) -> Tensor:
    """[B, t] int64: the merged-output row every input token landed in, computed on the device from the three index
    lists (the reference derives the same map by pushing a [B,t,t] identity through the merge, :326-337)."""
    unm, src, dst = merge.unm_idx, merge.src_idx, merge.dst_idx
    b, t = unm.shape[0], merge.n_tokens
    u = unm.shape[1]
    rows = torch.empty(b, t, dtype=torch.long, device=unm.device)
    rows[:, 1::2] = u + torch.arange(t // 2, device=unm.device)
    rows.scatter_(1, 2 * unm, torch.arange(u, device=unm.device).expand(b, -1))
    rows.scatter_(1, 2 * src, u + dst)
    if merge.distill_token:          # row permutation of models/tome.py:286-287
        perm = torch.cat([torch.tensor([0], device=unm.device), torch.tensor([u], device=unm.device),
                          torch.arange(1, u, device=unm.device), torch.arange(u + 1, t - merge.r, device=unm.device)])
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.numel(), device=unm.device)
        rows = inv[rows]
    return rows


def merge_wavg(merge: Callable, x: Tensor, size: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """models/tome.py:309-323 — one fused launch when ``merge`` came from bipartite_soft_matching above."""
    if isinstance(merge, _ToMeMerge):
        out, size_out, _ = ops.tome_merge(x, size, merge.unm_idx, merge.src_idx, merge.dst_idx, False, True)
        return merge.reorder(out), merge.reorder(size_out)
    if size is None:
        size = torch.ones_like(x[..., 0, None])
    x = merge(x * size, mode="sum")
    size = merge(size, mode="sum")
    return x / size, size


def merge_source(merge: Callable, x: Tensor, source: Optional[Tensor] = None) -> Tensor:
    """models/tome.py:326-337 — adjacency [B, N-r, t] between input tokens and merged groups."""
    if source is None:
        n, t, _ = x.shape
        if isinstance(merge, _ToMeMerge):
            rows = tome_token_rows(merge, x)                                  # valid with or without CLS / dist token
            out = torch.zeros(n, t - merge.r, t, device=x.device)
            return out.scatter_(1, rows.unsqueeze(1), 1.0)
        source = torch.eye(t, device=x.device)[None, ...].expand(n, t, t)
    return merge(source, mode="amax")


def _reduced_cluster_idx(source: Tensor, cls_token: bool) -> Tensor:
    """Block_ToMe's map from the adjacency (models/tome.py:92-99): float32, group id of every input token."""
    ar = torch.arange(1, source.shape[1] + 1, device=source.device, dtype=source.dtype)
    rci = torch.amax(source * ar[None, :, None], dim=-2)
    return (rci - 2)[:, 1:] if cls_token else rci - 1


class KeyMean:
    """``k.mean(1)`` (models/tome.py:58) not yet taken: the qkv Linear's output and the head count.  ops.tome_match_qkv
    consumes it in place (head mean in-kernel); ``tensor()`` materialises the reference's [B,N,64] metric."""

    def __init__(self, qkv: Tensor, num_heads: int):
        self.qkv, self.num_heads = qkv, num_heads

    def tensor(self) -> Tensor:
        b, n, c3 = self.qkv.shape
        return self.qkv.view(b, n, 3, self.num_heads, c3 // (3 * self.num_heads))[:, :, 1].mean(2)


_LOG_SIZE = [None, None]        # (size tensor, its log) of the last call


def _log_size(size: Tensor) -> Tensor:
    """``size.log()[..., 0]`` (models/tome.py:48).  The sizes only change where tokens merge, so the blocks between two
    merges are handed the SAME tensor object and reuse its logarithm (three launches per forward instead of eight)."""
    if _LOG_SIZE[0] is not size:
        _LOG_SIZE[0], _LOG_SIZE[1] = size, size.log()[..., 0]
    return _LOG_SIZE[1]


class Attention_ToMe(_AttentionBase):
    """models/tome.py:29-59.  forward(x, size) -> (x, metric = k.mean(1))."""

    need_metric = True      # Block_ToMe switches the key mean off in blocks that do not merge (r == 0)
    lazy_metric = False     # Block_ToMe: hand over the keys themselves (KeyMean) -- the match kernel takes the head mean

    def forward(self, x, size=None):
        if self._fused(x):
            bias = None if size is None else _log_size(size)                  # proportional attention, :48-49
            x, _, _, qkv = self._attend_fused(x, key_bias=bias)
            if not self.need_metric:
                return x, None
            km = KeyMean(qkv, self.num_heads)
            return x, (km if self.lazy_metric else km.tensor())               # = k.mean(1), :58
        q, k, v = self._qkv(x)
        attn = (q @ k.transpose(-2, -1)) * self.scale
        if size is not None:
            attn = attn + size.log()[:, None, None, :, 0]        # proportional attention
        attn = self.attn_drop(attn.softmax(dim=-1))
        return self._out(attn, v), k.mean(1)


class Block_ToMe(nn.Module):
    """models/tome.py:61-104.  forward(x, attn_size) -> (x, attn_size, reduced_cluster_idx | None)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, r=0, cls_token=True, dist_token=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention_ToMe(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.r = r
        self.cls_token = cls_token
        self.dist_token = dist_token
        self.attn.need_metric = r > 0
        self.attn.lazy_metric = True

    def forward(self, x, attn_size=None):
        x, y = enter_norm(self.norm1, x)
        x_attn, metric = self.attn(y, attn_size)
        reduced_cluster_idx = None
        if self.r <= 0:
            x, y = add_norm(x, self.drop_path(x_attn), self.norm2)
            return defer_add(x, self.drop_path(self.mlp(y)), self.norm1, self), attn_size, reduced_cluster_idx
        if self.r > 0:
            _train_guard(self, True)
            re = ops.tome_effective_r(x.shape[1], self.r, self.cls_token, self.dist_token)
            if re > 0 and self.cls_token and not self.dist_token and isinstance(metric, KeyMean):
                # keys straight from the qkv output: head mean + matching in ONE launch, no [B,N,64] metric tensor
                unm, src, dst = ops.tome_match_qkv(metric.qkv, metric.num_heads, self.r, True)
                branch = self.drop_path(x_attn)
                if _norm_fusable(self.norm2, x) and ops.tome_merge_ln_supported(x, branch):
                    # x + branch, merge_wavg and norm2 in ONE launch (:88, :100-101, :104): the residual add and the
                    # LayerNorm ride on the row the merge already holds in registers
                    x, attn_size, reduced_cluster_idx, y = ops.tome_merge_ln(
                        x, branch, attn_size, unm, src, dst, self.norm2.weight, self.norm2.bias, self.norm2.eps, True)
                    return defer_add(x, self.drop_path(self.mlp(y)), self.norm1, self), attn_size, reduced_cluster_idx
                x, attn_size, reduced_cluster_idx = ops.tome_merge(x + branch, attn_size, unm, src, dst, True, True)
                metric = None
                x_attn = None
        if x_attn is not None:
            x = x + self.drop_path(x_attn)
        if self.r > 0:
            if isinstance(metric, KeyMean):
                metric = metric.tensor()
            if metric is None:
                pass
            elif re > 0 and self.cls_token and not self.dist_token:
                lowp = _lowp() or metric.dtype in (torch.bfloat16, torch.float16)
                with torch.no_grad():                                     # models/tome.py:258
                    unm, src, dst = ops.tome_match(_f16_to_bf16(metric), self.r, True, lowp, True)
                # merged tokens, new sizes and the source map from ONE launch (the reference pushes a [B,t,t]
                # identity through the merge to get the map, models/tome.py:91-99)
                x, attn_size, reduced_cluster_idx = ops.tome_merge(x, attn_size, unm, src, dst, True, True)
            else:
                # nothing left to merge (identity map, like the reference's do_nothing), no class token (different
                # offset) or a distillation token (row permutation): the reference's own sequence over the same ops
                merge, _ = bipartite_soft_matching(metric, self.r, self.cls_token, self.dist_token)
                reduced_cluster_idx = _reduced_cluster_idx(merge_source(merge, x, None), self.cls_token)
                x, attn_size = merge_wavg(merge, x, attn_size)
        x = defer_add(x, self.drop_path(self.mlp(norm_lowp(self.norm2, x))), self.norm1, self)
        return x, attn_size, reduced_cluster_idx


# =============================================================================================== DPC-KNN
def index_points(points: Tensor, idx: Tensor) -> Tensor:
    """models/dpcknn.py:25-41 — points [B,N,C], idx [B,S] -> [B,S,C]."""
    b = points.shape[0]
    batch = torch.arange(b, dtype=torch.long, device=points.device).view(b, *([1] * (idx.dim() - 1))).expand_as(idx)
    return points[batch, idx, :]


def cluster_dpc_knn(x: Tensor, cluster_num: int, k: int = 5, token_mask: Optional[Tensor] = None):
    """models/dpcknn.py:44-100 -> (idx_cluster [B,N], index_down [B,cluster_num])."""
    if token_mask is not None:
        raise NotImplementedError("cluster_dpc_knn: token_mask is never passed by the reference models")
    b, n, _ = x.shape
    # the reference draws its tie-breaking noise inside the op (:73-74); same call, same generator position
    noise = torch.rand((b, n), device=x.device, dtype=torch.float32)
    return ops.dpcknn_cluster(x, noise, cluster_num, k, False)


def merge_tokens(x, idx_token, agg_weight, idx_cluster, cluster_num, token_weight=None):
    """models/dpcknn.py:103-140 -> (x_merged, idx_token_new, agg_weight_new)."""
    return ops.dpcknn_merge(x, idx_token, agg_weight, idx_cluster, token_weight, cluster_num)


class CTM(nn.Module):
    """models/dpcknn.py:143-172.  forward(x, idx_token, agg_weight, viz_mode) -> 6-tuple."""

    def __init__(self, embed_dim, cluster_num, k=5, equal_weight=False):
        super().__init__()
        self.cluster_num = cluster_num
        self.equal_weight = equal_weight
        self.k = k
        if not self.equal_weight:
            self.score = nn.Linear(embed_dim, 1)

    def forward(self, x, idx_token, agg_weight, viz_mode=False):
        _train_guard(self)
        token_weight = None if self.equal_weight else self.score(x).exp()
        idx_cluster, idx_centers = cluster_dpc_knn(x, self.cluster_num, self.k)
        cluster_centers = index_points(x, idx_centers) if viz_mode else None
        x, idx_token, agg_weight = merge_tokens(x, idx_token, agg_weight, idx_cluster, self.cluster_num, token_weight)
        if viz_mode:
            return x, idx_token, agg_weight, idx_centers, idx_cluster, cluster_centers
        return x, idx_token, agg_weight, None, None, None


# =============================================================================================== K-Medoids
def k_medoids_fit(x: Tensor, cluster_num: int, iterations: int = 5, token_weight: Optional[Tensor] = None):
    """models/kmedoids.py:40-85 -> (centres, cluster_idx, assignment)."""
    if token_weight is None:
        # --equal_weight (:43-61): one numpy draw picks the first medoid for the whole batch -- the same call, so the same
        # generator position as the reference -- then the farthest-point initialisation and unit weights
        import numpy as np
        first = int(np.random.choice(np.arange(x.shape[1]), 1)[0])
        return ops.kmedoids_fit_equal(x, cluster_num, iterations, first, False)
    return ops.kmedoids_fit(x, token_weight, cluster_num, iterations, False)


class KMedoids(nn.Module):
    """models/kmedoids.py:135-148.  forward(x, token_weights) -> (centres, idx_center, idx_cluster)."""

    def __init__(self, num_clusters, iters, equal_weights=False):
        super().__init__()
        self.cluster_count = num_clusters
        self.iters = iters
        self.equal_weights = equal_weights

    def forward(self, x, token_weights):
        _train_guard(self)
        if self.equal_weights:
            token_weights = None
        return k_medoids_fit(x, self.cluster_count, self.iters, token_weights)


class AttentionWithProbs(_AttentionBase):
    """models/kmedoids.py:88-112.  forward(x) -> (x, attn)."""

    # "full": the reference's materialised probabilities (default; the block-level signature of the reference).
    # "colsum" / "none": set by KMedoidsVisionTransformer for the blocks whose probabilities feed the next cluster
    # layer / nobody: forward returns (x, column sums [B,H,N] | None) and [B,H,N,N] is never written.
    probs_mode = "full"

    def forward(self, x):
        if self.probs_mode != "full" and self._fused(x):
            x, _, colsum, _ = self._attend_fused(x, want_colsum=self.probs_mode == "colsum")
            return x, colsum
        q, k, v = self._qkv(x)
        attn = self.attn_drop(((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1))
        return self._out(attn, v), attn


class BlockWithProbs(nn.Module):
    """models/kmedoids.py:115-132.  forward(x) -> (x, attn)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = AttentionWithProbs(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        x, y = enter_norm(self.norm1, x)
        x_attn, attn = self.attn(y)
        x, y = add_norm(x, self.drop_path(x_attn), self.norm2)
        x = defer_add(x, self.drop_path(self.mlp(y)), self.norm1, self)
        return x, attn


# =============================================================================================== Sinkhorn
class Sinkhorn(nn.Module):
    """models/sinkhorn.py:59-86.  forward(x) -> (x [B,K,C], weights [B,K,P])."""

    def __init__(self, embed_dim, cluster_centers, eps, iters):
        super().__init__()
        self.v = nn.Parameter(torch.randn(cluster_centers, embed_dim))
        self.eps = eps
        self.iters = iters

    def forward(self, x):
        _train_guard(self)
        with torch.no_grad():                       # the reference overwrites its parameter (:73-76); idempotent
            self.v.copy_(F.normalize(self.v.clone(), p=2, dim=-1))
        out, w = ops.sinkhorn_merge(x, self.v.detach(), self.eps, self.iters, _lowp(), True)
        return _amp_out(out), w


# =============================================================================================== PatchMerger
class PatchMerger(nn.Module):
    """models/patchmerger.py:24-39.  forward(x) -> (x [B,K,C], attn [B,K,P])."""

    def __init__(self, embed_dim, cluster_centers, scaled_attention=False):
        super().__init__()
        self.scale = embed_dim ** -0.5 if scaled_attention else 1.0
        self.norm = nn.LayerNorm(embed_dim)
        self.queries = nn.Parameter(torch.randn(cluster_centers, embed_dim))

    def forward(self, x):
        _train_guard(self)
        out, attn = ops.patchmerger(x, self.norm.weight.detach(), self.norm.bias.detach(), self.queries.detach(), self.scale,
                                    self.norm.eps, _lowp(), True)
        return _amp_out(out), attn


# =============================================================================================== SiT
class TokenSlimmingModule(nn.Module):
    """models/sit.py:25-40.  forward(x) -> (x [B,K,C], weight [B,K,P])."""

    def __init__(self, embed_dim, cluster_centers, ratio=0.5):
        super().__init__()
        hidden_dim = int(embed_dim * ratio)
        self.weight = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, hidden_dim), nn.GELU(),
                                    nn.Linear(hidden_dim, cluster_centers))
        self.scale = nn.Parameter(torch.ones(1, 1, 1))

    def forward(self, x):
        _train_guard(self)
        out, w = ops.sit_merge(x, self.weight(x), self.scale.detach(), _lowp(), True)
        return _amp_out(out), w


# =============================================================================================== ATS
def batched_index_select(values: Tensor, indices: Tensor, dim: int = 1) -> Tensor:
    """models/ats.py:27-41 for the two shapes the reference uses (tokens [B,N,C] with ids [B,M];
    attention [B,H,N,N] with ids [B,H,M] identical across heads)."""
    if values.dim() == 3 and indices.dim() == 2 and dim == 1:
        return ops.gather_rows(values, indices)
    if values.dim() == 4 and indices.dim() == 3 and dim == 2:
        return ops.gather_rows(values, indices[:, 0].contiguous())
    raise NotImplementedError("batched_index_select: only the token / attention-row gathers of models/ats.py")


class AdaptiveTokenSampling(nn.Module):
    """models/ats.py:44-89.  forward(v, attn, mask) -> (new_attn [B,H,M+1,N], new_mask [B,M+1], ids [B,M+1]).

    ``static_width=True`` pads to sample_count instead of the batch maximum of unique ids: no host read at all
    (the padded rows are copies of CLS that new_mask switches off, so logits are unchanged); the default keeps
    the reference's data-dependent width at the cost of ONE scalar read (the reference syncs once per image).
    """

    def __init__(self, sample_count, eps=1e-6, static_width=False):
        super().__init__()
        self.sample_count = sample_count
        self.sample_steps = torch.arange(1 / (2 * sample_count), (2 * sample_count - 1) / (2 * sample_count),
                                         2 / (2 * sample_count))
        self.eps = eps
        self.static_width = static_width
        self._steps_dev = None

    def sample(self, x, attn, mask):
        """ids / new mask only.  ``attn`` may be the [B,H,N,N] probabilities or just their CLS rows [B,H,N]."""
        _train_guard(self)
        if self._steps_dev is None or self._steps_dev.device != attn.device:
            self._steps_dev = self.sample_steps.to(attn.device)
        ids, new_mask, max_count = ops.ats_sample(x, attn, mask, self._steps_dev, self.eps)
        if not self.static_width:
            m = int(max_count.item()) + 1
            ids, new_mask = ids[:, :m], new_mask[:, :m]
        return ids, new_mask

    def forward(self, x, attn, mask):
        ids, new_mask = self.sample(x, attn, mask)
        new_attn = ops.gather_rows(attn, ids, ids.shape[1])
        return new_attn, new_mask, ids


class ATSAttention(_AttentionBase):
    """models/ats.py:92-134.  forward(x, mask) -> (x, mask, sample_ids | None)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0, ats_sample_count=0):
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        self.ats_sample_count = ats_sample_count
        if self.ats_sample_count:
            self.ats = AdaptiveTokenSampling(ats_sample_count)
        # set by ATSVisionTransformer for the blocks up to its first sampling stage: their mask is the all-true tensor the
        # model creates, so the fused attention runs its unmasked instantiation (same numbers, 63 instead of 81 us)
        self.mask_all_true = False

    def forward(self, x, mask):
        if self._fused(x):
            # scores first (CLS rows only: v is not read), then the attention of the SAMPLED query rows only -- the
            # row gather attn[:, :, ids, :] of models/ats.py:84-87 commutes with the row-wise softmax
            b, n, c = x.shape
            qkv = self.qkv(x)
            qkv = qkv if qkv.dtype == torch.bfloat16 else qkv.to(torch.bfloat16)
            sample_ids = None
            old_mask = None if self.mask_all_true else mask
            if self.ats_sample_count:
                _, cls_row, _ = ops.attention(qkv, self.num_heads, self.scale, None, old_mask, None, False, True, False)
                v = qkv.view(b, n, 3, self.num_heads, c // self.num_heads)[:, :, 2].permute(0, 2, 1, 3)
                full = mask if mask is not None else torch.ones(b, n, dtype=torch.bool, device=x.device)
                sample_ids, mask = self.ats.sample(v, cls_row, full)
            out, _, _ = ops.attention(qkv, self.num_heads, self.scale, None, old_mask, sample_ids, True, False, False)
            return self.proj_drop(self.proj(out)), mask, sample_ids
        q, k, v = self._qkv(x)
        dots = (q @ k.transpose(-2, -1)) * self.scale
        if mask is not None:
            dots_mask = mask.unsqueeze(1).unsqueeze(3) * mask.unsqueeze(1).unsqueeze(2)
            dots = dots.masked_fill(~dots_mask, -torch.finfo(dots.dtype).max)
        attn = self.attn_drop(dots.softmax(dim=-1))
        sample_ids = None
        if self.ats_sample_count:
            attn, mask, sample_ids = self.ats(v, attn, mask)
        return self._out(attn, v), mask, sample_ids


class ATSBlock(nn.Module):
    """models/ats.py:137-162.  forward(x, mask) -> (x, mask, sample_ids | None)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, init_values=None,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, ats_sample_count=0):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = ATSAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop,
                                 ats_sample_count=ats_sample_count)
        self.drop_path1 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.drop_path2 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x, mask):
        x, y = enter_norm(self.norm1, x)
        x_tmp, mask, sample_ids = self.attn(y, mask)
        if sample_ids is not None:
            x = ops.gather_rows(x, sample_ids)
        x, y = add_norm(x, self.drop_path1(x_tmp), self.norm2)
        x = defer_add(x, self.drop_path2(self.mlp(y)), self.norm1, self)
        return x, mask, sample_ids


# =============================================================================================== DynamicViT
def batch_index_select(x: Tensor, idx: Tensor) -> Tensor:
    """models/dyvit.py:340-356."""
    if x.dim() == 3:
        return ops.gather_rows(x, idx) if x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) else \
            torch.gather(x, 1, idx.unsqueeze(-1).expand(-1, -1, x.shape[-1]))
    if x.dim() == 2:
        return torch.gather(x, 1, idx)
    raise NotImplementedError


def _slice_base(x: Tensor) -> Optional[Tensor]:
    """the contiguous [B,N,C] tensor of which ``x`` is the ``[:, 1:]`` view, or None."""
    base = x._base
    if (base is None or x.dim() != 3 or base.dim() != 3 or not base.is_contiguous() or base.dtype != x.dtype
            or base.shape[0] != x.shape[0] or base.shape[1] != x.shape[1] + 1 or base.shape[2] != x.shape[2]
            or x.stride() != base.stride() or x.data_ptr() != base.data_ptr() + base.shape[2] * base.element_size()):
        return None
    return base


class PredictorLG(nn.Module):
    """models/dyvit.py:91-119.  forward(x, policy) -> log-softmax keep/drop scores [B,P,2]."""

    def __init__(self, embed_dim=384, eps=1e-6):
        super().__init__()
        self.in_conv = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, embed_dim), nn.GELU())
        self.out_conv = nn.Sequential(nn.Linear(embed_dim, embed_dim // 2), nn.GELU(),
                                      nn.Linear(embed_dim // 2, embed_dim // 4), nn.GELU(),
                                      nn.Linear(embed_dim // 4, 2), nn.LogSoftmax(dim=-1))
        self.eps = eps

    def forward(self, x, policy):
        full = _slice_base(x)
        if full is not None and _norm_fusable(self.in_conv[0], full) and not self.training:
            # x is the [:, 1:] view of the contiguous token tensor (models/dyvit.py:232): ATen's layer_norm would copy the
            # view (fp32), normalise it and cast it for the Linear -- three passes.  LayerNorm, Linear and GELU are per
            # token, so they run on ALL tokens from one add_layernorm pass (the fused LayerNorm every block uses under
            # bf16 autocast: one rounding to bf16) and the class row is skipped by the pooling kernel's batch stride.
            h = self.in_conv[2](self.in_conv[1](norm_lowp(self.in_conv[0], full)))[:, 1:]
        else:
            h = self.in_conv(x)
        # under bf16 autocast out_conv's first Linear casts its input to bf16: ask the kernel for that tensor directly
        lowp = (h.is_cuda and h.dtype == torch.bfloat16 and torch.is_autocast_enabled("cuda")
                and torch.get_autocast_dtype("cuda") == torch.bfloat16 and isinstance(self.out_conv[0], nn.Linear))
        return self.out_conv(ops.dyvit_pool_concat(h, policy, self.eps, lowp_out=lowp))


class Policy_Attention(_AttentionBase):
    """models/dyvit.py:23-69 (eval branch: policy is None -> plain softmax)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        if qk_scale is not None:
            self.scale = qk_scale

    @staticmethod
    def softmax_with_policy(attn, policy, eps=1e-6):
        """models/dyvit.py:39-51 (training): dropped keys are masked out of every row but their own diagonal entry."""
        b, n, _ = policy.size()
        attn_policy = policy.reshape(b, 1, 1, n)
        eye = torch.eye(n, dtype=attn_policy.dtype, device=attn_policy.device).view(1, 1, n, n)
        attn_policy = attn_policy + (1.0 - attn_policy) * eye
        max_att = torch.max(attn, dim=-1, keepdim=True)[0]
        attn = attn - max_att
        attn = attn.to(torch.float32).exp_() * attn_policy.to(torch.float32)
        attn = (attn + eps / n) / (attn.sum(dim=-1, keepdim=True) + eps)
        return attn.type_as(max_att)

    def forward(self, x, policy=None):
        if policy is None and self._fused(x):
            return self._attend_fused(x)[0]
        q, k, v = self._qkv(x)
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = attn.softmax(dim=-1) if policy is None else self.softmax_with_policy(attn, policy)
        return self._out(attn, v)


class Block_DyVIT(nn.Module):
    """models/dyvit.py:72-88."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Policy_Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                     proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, policy=None):
        x, y = enter_norm(self.norm1, x)
        x, y = add_norm(x, self.drop_path(self.attn(y, policy=policy)), self.norm2)
        return defer_add(x, self.drop_path(self.mlp(y)), self.norm1, self)
