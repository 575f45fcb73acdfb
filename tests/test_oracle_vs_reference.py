"""Pins the oracle restatement (oracle/ops.py) against the UNMODIFIED reference functions, imported from
/root/reference through the timm shim.  Runs only where the reference tree exists (the build container);
elsewhere tests/test_oracle_golden.py checks the same oracle against vectors the reference produced here."""
import math

import pytest
import torch

from oracle import ops as O
from oracle import timm_shim

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    timm_shim.import_reference()
    import importlib
    mods = {n: importlib.import_module(f"models.{n}") for n in
            ["topk", "evit", "tome", "dpcknn", "kmedoids", "sinkhorn", "ats", "dyvit", "patchmerger", "sit"]}
    return mods


def g(seed):
    return torch.Generator().manual_seed(seed)


def rand_attn(b, h, n, seed):
    return torch.softmax(4 * torch.randn(b, h, n, n, generator=g(seed)), dim=-1)


@pytest.mark.parametrize("n,k", [(197, 137), (138, 96), (97, 67), (197, 98)])
def test_topk(ref, n, k):
    b, h, c = 3, 6, 48
    attn, x = rand_attn(b, h, n, 1), torch.randn(b, n, c, generator=g(2))
    scores = attn[:, :, 0, 1:].mean(dim=1)
    _, idx_ref = torch.topk(scores, k, dim=1, largest=True, sorted=True)
    x_ref = torch.cat([x[:, :1], torch.gather(x[:, 1:], 1, idx_ref.unsqueeze(-1).expand(-1, -1, c))], 1)
    assert torch.equal(O.cls_attention_scores(attn), scores)
    out, idx = O.topk_gather(x, scores, k)
    assert torch.equal(idx, idx_ref) and torch.equal(out, x_ref)


@pytest.mark.parametrize("n,k", [(197, 98), (100, 49), (51, 24)])
def test_evit(ref, n, k):
    b, h, c = 3, 4, 40
    attn, x = rand_attn(b, h, n, 3), torch.randn(b, n, c, generator=g(4))
    scores = attn[:, :, 0, 1:].mean(dim=1)
    _, idx_ref = torch.topk(scores, k, dim=1)
    compl_ref = ref["evit"].complement_idx(idx_ref, n - 1)
    non_cls = x[:, 1:]
    extra = torch.sum(torch.gather(non_cls, 1, compl_ref.unsqueeze(-1).expand(-1, -1, c))
                      * torch.gather(scores, 1, compl_ref).unsqueeze(-1), dim=1, keepdim=True)
    x_ref = torch.cat([x[:, :1], torch.gather(non_cls, 1, idx_ref.unsqueeze(-1).expand(-1, -1, c)), extra], 1)
    out, idx, compl = O.evit_select_fuse(x, scores, k)
    assert torch.equal(compl, compl_ref)
    assert torch.equal(idx[:, :-1], idx_ref) and bool((idx[:, -1] == -1).all())
    assert torch.equal(out, x_ref)


@pytest.mark.parametrize("n,r", [(197, 59), (138, 41), (97, 29), (197, 98), (50, 30)])
@pytest.mark.parametrize("with_size", [False, True])
def test_tome(ref, n, r, with_size):
    T = ref["tome"]
    b, c, d = 3, 48, 64
    metric, x = torch.randn(b, n, d, generator=g(5)), torch.randn(b, n, c, generator=g(6))
    size = torch.randint(1, 4, (b, n, 1), generator=g(7)).float() if with_size else None
    merge, _ = T.bipartite_soft_matching(metric, r, True, False)
    x_ref, size_ref = T.merge_wavg(merge, x, size)
    source = T.merge_source(merge, x, None)
    rci = source * ((torch.ones(source.shape).permute(0, 2, 1)) * torch.arange(1, source.shape[1] + 1)).permute(0, 2, 1)
    rci = (torch.amax(rci, dim=-2) - 2)[:, 1:]
    unm, src, dst, _ = O.tome_match(metric, r, True)
    out, size_out, rci_o = O.tome_merge(x, size, unm, src, dst)
    assert out.shape[1] == n - O.tome_effective_r(n, r)
    assert torch.equal(out, x_ref) and torch.equal(size_out, size_ref) and torch.equal(rci_o, rci)


@pytest.mark.parametrize("p", [196, 49, 12])
def test_pairwise_dist(ref, p):
    x = torch.randn(2, p, 96, generator=g(8))
    d_ref = torch.cdist(x, x)
    d = O.pairwise_dist(x)
    assert torch.allclose(d, d_ref, rtol=1e-4, atol=2e-2 if p > 25 else 1e-6)
    off = ~torch.eye(p, dtype=torch.bool)
    assert torch.allclose(d[:, off], d_ref[:, off], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("p,k", [(196, 49), (49, 12), (12, 3)])
def test_dpcknn(ref, p, k):
    D = ref["dpcknn"]
    b, c = 3, 64
    x = torch.randn(b, p, c, generator=g(9))
    dist = torch.cdist(x, x)
    torch.manual_seed(11)
    idx_cluster_ref, index_down_ref = D.cluster_dpc_knn(x, k, 5)
    torch.manual_seed(11)
    noise = torch.rand(b, p)
    idx_cluster, index_down = O.dpcknn_cluster(x, k, 5, noise, dist=dist)
    assert torch.equal(index_down, index_down_ref) and torch.equal(idx_cluster, idx_cluster_ref)
    tw = torch.randn(b, p, 1, generator=g(10)).exp()
    idx_token = torch.randint(0, p, (b, 196), generator=g(12))
    agg = torch.rand(b, 196, 1, generator=g(13))
    xm_ref, it_ref, aw_ref = D.merge_tokens(x, idx_token, agg, idx_cluster_ref, k, tw)
    xm, it, aw = O.dpcknn_merge(x, idx_token, agg, idx_cluster, k, tw)
    assert torch.equal(xm, xm_ref) and torch.equal(it, it_ref) and torch.equal(aw, aw_ref)


@pytest.mark.parametrize("p,k", [(196, 49), (49, 12), (12, 3)])
def test_kmedoids(ref, p, k):
    K = ref["kmedoids"]
    b, c = 3, 64
    x = torch.randn(b, p, c, generator=g(14))
    attn = rand_attn(b, 3, p + 1, 15)
    tw_ref = torch.sum(torch.sum(attn, dim=1), dim=1)[:, 1:].unsqueeze(2)
    assert torch.equal(O.attn_colsum(attn), tw_ref)
    c_ref, ci_ref, as_ref = K.k_medoids_fit(x, k, 3, tw_ref)
    cen, ci, asg = O.kmedoids_fit(x, k, 3, tw_ref, dist=torch.cdist(x, x))
    assert torch.equal(ci, ci_ref) and torch.equal(asg, as_ref) and torch.equal(cen, c_ref)


@pytest.mark.parametrize("p,k", [(196, 49), (49, 12), (12, 3)])
def test_kmedoids_equal_weight(ref, p, k):
    """--equal_weight (models/kmedoids.py:43-61): token_weight=None -> one numpy draw, farthest-point initialisation, unit
    weights.  The oracle takes the drawn index as an argument; same numpy seed on both sides."""
    import numpy as np
    K = ref["kmedoids"]
    b, c = 3, 64
    x = torch.randn(b, p, c, generator=g(114))
    np.random.seed(5)
    c_ref, ci_ref, as_ref = K.k_medoids_fit(x, k, 3, None)
    np.random.seed(5)
    first = int(np.random.choice(np.arange(p), 1)[0])
    cen, ci, asg = O.kmedoids_fit_equal(x, k, 3, first, dist=None)
    assert torch.equal(ci, ci_ref) and torch.equal(asg, as_ref) and torch.equal(cen, c_ref)


@pytest.mark.parametrize("p,k", [(196, 176), (176, 158), (60, 20)])
def test_sinkhorn(ref, p, k):
    S = ref["sinkhorn"]
    x = torch.randn(2, p, 64, generator=g(16))
    mod = S.Sinkhorn(64, k, 1.0, 3)
    v0 = mod.v.detach().clone()
    with torch.no_grad():
        out_ref, w_ref = mod(x)
    out, w, vh = O.sinkhorn_merge(x, v0, 1.0, 3)
    assert torch.allclose(vh, mod.v.detach(), rtol=0, atol=0)
    assert torch.allclose(w, w_ref, rtol=1e-5, atol=1e-7) and torch.allclose(out, out_ref, rtol=1e-5, atol=1e-6)


def test_patchmerger(ref):
    P = ref["patchmerger"]
    x = torch.randn(2, 196, 64, generator=g(17))
    mod = P.PatchMerger(64, 176)
    with torch.no_grad():
        mod.norm.weight.copy_(torch.rand(64, generator=g(18)) + 0.5)
        mod.norm.bias.copy_(torch.randn(64, generator=g(19)) * 0.1)
        out_ref, attn_ref = mod(x)
    out, attn = O.patchmerger(x, mod.norm.weight.detach(), mod.norm.bias.detach(), mod.queries.detach())
    assert torch.allclose(attn, attn_ref, rtol=1e-5, atol=1e-7) and torch.allclose(out, out_ref, rtol=1e-5, atol=1e-6)


def test_sit(ref):
    S = ref["sit"]
    x = torch.randn(2, 196, 64, generator=g(20))
    mod = S.TokenSlimmingModule(64, 176)
    with torch.no_grad():
        mod.scale.fill_(1.7)
        out_ref, w_ref = mod(x)
        logits = mod.weight(x)
    out, w = O.sit_merge(x, logits, mod.scale.detach())
    assert torch.allclose(w, w_ref, rtol=1e-6, atol=1e-8) and torch.allclose(out, out_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,count", [(197, 177), (177, 159), (197, 60)])
def test_ats(ref, n, count):
    A = ref["ats"]
    b, h = 3, 4
    attn = rand_attn(b, h, n, 21)
    v = torch.randn(b, h, n, 16, generator=g(22))
    mask = torch.ones(b, n, dtype=torch.bool)
    mask[1, n - 20:] = False
    mod = A.AdaptiveTokenSampling(count)
    na_ref, nm_ref, ids_ref = mod(v, attn, mask)
    assert torch.equal(O.ats_sample_steps(count), mod.sample_steps)
    na, nm, ids = O.ats_sample(v, attn, mask, count)
    assert torch.equal(ids, ids_ref) and torch.equal(nm, nm_ref) and torch.equal(na, na_ref)


def test_dyvit(ref):
    Dy = ref["dyvit"]
    b, p, c = 3, 196, 64
    mod = Dy.PredictorLG(c).eval()
    x = torch.randn(b, p, c, generator=g(23))
    policy = (torch.rand(b, p, 1, generator=g(24)) > 0.3).float()
    with torch.no_grad():
        h = mod.in_conv(x)
        ref_out = mod(x, policy)
        mine = mod.out_conv(O.dyvit_pool_concat(h, policy))
    assert torch.equal(mine, ref_out)
    score = torch.randn(b, p, generator=g(25))
    keep = torch.argsort(score, dim=1, descending=True)[:, :98]
    xx = torch.randn(b, p + 1, c, generator=g(26))
    now = torch.cat([torch.zeros(b, 1, dtype=keep.dtype), keep + 1], dim=1)
    x_ref = Dy.batch_index_select(xx, now)
    out, idx = O.dyvit_keep(xx, score, 98)
    assert torch.equal(idx, keep) and torch.equal(out, x_ref)


# ---------------------------------------------------------------------------------------------- model level
import argparse
import contextlib
import io

from oracle import model as OM

_KR = {"topk": 0.7, "evit": 0.5, "tome": 0.7, "dyvit": 0.5, "dpcknn": 0.25, "kmedoids": 0.25, "sinkhorn": 0.9,
       "patchmerger": 0.9, "ats": 0.9, "sit": 0.9}


@pytest.mark.parametrize("name", sorted(_KR))
def test_model_tiny_vs_reference(ref, name):
    """oracle.model.forward == the reference model (DeiT-tiny, reduction_loc 3 6 9) on the same weights and input."""
    from timm.models import create_model
    args = argparse.Namespace(keep_rate=[_KR[name]], reduction_loc=[3, 6, 9], distillation_type="none", k_neighbors=5,
                              cluster_iters=3, sinkhorn_eps=1.0, equal_weight=False, dyvit_distill=False)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(f"{name}_tiny_patch16_224", pretrained=False, num_classes=16, drop_rate=0.0,
                             drop_path_rate=0.0, drop_block_rate=None, img_size=224, args=args).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x = torch.randn(2, 3, 224, 224, generator=g(77))
    torch.manual_seed(5)
    with torch.no_grad():
        y_ref = model(x)
    torch.manual_seed(5)
    y = OM.forward(name, sd, x, OM.cfg_for("tiny", keep_rate=[_KR[name]]))
    assert torch.allclose(y, y_ref, rtol=1e-4, atol=1e-5), float((y - y_ref).abs().max())


def test_dyvit_training_path_vs_reference(ref):
    """f4 pin: oracle.model.dyvit_train_forward == the UNMODIFIED reference model in train() mode (gumbel keep decisions,
    softmax_with_policy, models/dyvit.py:205-229) -- logits, per-stage hard decisions and parameter gradients, with the
    same generator state on both sides so that F.gumbel_softmax draws the same noise."""
    from timm.models import create_model
    args = argparse.Namespace(keep_rate=[0.5], reduction_loc=[3, 6, 9], distillation_type="none", k_neighbors=5,
                              cluster_iters=3, sinkhorn_eps=1.0, equal_weight=False, dyvit_distill=False)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model("dyvit_tiny_patch16_224", pretrained=False, num_classes=16, drop_rate=0.0,
                             drop_path_rate=0.0, drop_block_rate=None, img_size=224, args=args).train()
    with torch.no_grad():          # spread the predictor so that keep / drop decisions are not all ties
        for n_, p_ in model.named_parameters():
            if n_.startswith("score_predictor") and p_.dim() >= 2:
                p_.mul_(4.0)
    x = torch.randn(2, 3, 224, 224, generator=g(78))
    names = ["score_predictor.0.in_conv.1.weight", "score_predictor.2.out_conv.4.weight", "blocks.4.attn.qkv.weight",
             "blocks.0.mlp.fc1.weight", "patch_embed.proj.weight"]
    torch.manual_seed(6)
    y_ref, dec_ref = model(x)
    (y_ref.square().sum() + sum(d.sum() for d in dec_ref)).backward()
    g_ref = {n: dict(model.named_parameters())[n].grad.clone() for n in names}
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    torch.manual_seed(6)
    y, dec = OM.dyvit_train_forward(sd, x, OM.cfg_for("tiny", keep_rate=[0.5]))
    (y.square().sum() + sum(d.sum() for d in dec)).backward()
    assert torch.allclose(y, y_ref, rtol=1e-4, atol=1e-5), float((y - y_ref).abs().max())
    for a, b_ in zip(dec, dec_ref):
        assert torch.equal(a, b_)
    for n in names:
        assert torch.allclose(sd[n].grad, g_ref[n], rtol=1e-3, atol=1e-6), (n, float((sd[n].grad - g_ref[n]).abs().max()))
