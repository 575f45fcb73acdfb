"""CPU: the oracle restatement (oracle/ops.py, oracle/model.py) against the committed golden vectors that the
UNMODIFIED reference produced in the build container (tests/golden/make_golden.py).  Runs anywhere."""
import os

import pytest
import torch

from oracle import model as OM
from oracle import ops as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gops():
    return torch.load(os.path.join(GOLD, "ops.pt"))


@pytest.fixture(scope="module")
def gmodels():
    return torch.load(os.path.join(GOLD, "models.pt"))


def test_select(gops):
    d = gops["select"]
    assert torch.equal(O.cls_attention_scores(d["attn"]), d["cls_attn"])
    out, idx = O.topk_gather(d["x"], d["cls_attn"], d["k"])
    assert torch.equal(idx, d["idx"]) and torch.equal(out, d["x_topk"])
    out, idx, compl = O.evit_select_fuse(d["x"], d["cls_attn"], d["k"])
    assert torch.equal(compl, d["compl"]) and torch.equal(idx[:, :-1], d["idx"]) and torch.equal(out, d["x_evit"])


def test_tome(gops):
    d = gops["tome"]
    unm, src, dst, _ = O.tome_match(d["metric"], d["r"], True)
    out, size, rci = O.tome_merge(d["x"], d["size"], unm, src, dst)
    assert torch.equal(out, d["x_out"]) and torch.equal(size, d["size_out"]) and torch.equal(rci, d["rci"])


def test_dpcknn(gops):
    d = gops["dpcknn"]
    ic, idn = O.dpcknn_cluster(d["x"], d["K"], d["knn"], d["noise"], dist=d["dist"])
    assert torch.equal(ic, d["idx_cluster"]) and torch.equal(idn, d["index_down"])
    xm, it, aw = O.dpcknn_merge(d["x"], d["idx_token"], d["agg_weight"], ic, d["K"], d["token_weight"])
    assert torch.equal(xm, d["x_merged"]) and torch.equal(it, d["idx_token_new"]) and torch.equal(aw, d["agg_weight_new"])
    assert torch.allclose(O.pairwise_dist(d["x"]), d["dist"], rtol=1e-4, atol=2e-2)


def test_kmedoids(gops):
    d = gops["kmedoids"]
    assert torch.equal(O.attn_colsum(d["attn"]), d["token_weight"])
    cen, ci, asg = O.kmedoids_fit(d["x"], d["K"], d["iters"], d["token_weight"], dist=d["dist"])
    assert torch.equal(ci, d["cluster_idx"]) and torch.equal(asg, d["assignment"]) and torch.equal(cen, d["centres"])


def test_soft_merges(gops):
    d = gops["sinkhorn"]
    out, w, vh = O.sinkhorn_merge(d["x"], d["v"], d["eps"], d["iters"])
    assert torch.equal(vh, d["v_after"])
    assert torch.allclose(w, d["weights"], rtol=1e-5, atol=1e-7) and torch.allclose(out, d["out"], rtol=1e-5, atol=1e-6)
    d = gops["patchmerger"]
    out, attn = O.patchmerger(d["x"], d["ln_w"], d["ln_b"], d["queries"])
    assert torch.allclose(attn, d["attn"], rtol=1e-5, atol=1e-7) and torch.allclose(out, d["out"], rtol=1e-5, atol=1e-6)
    d = gops["sit"]
    out, w = O.sit_merge(d["x"], d["logits"], d["scale"])
    assert torch.allclose(w, d["weights"], rtol=1e-6, atol=1e-8) and torch.allclose(out, d["out"], rtol=1e-5, atol=1e-6)


def test_ats(gops):
    d = gops["ats"]
    na, nm, ids = O.ats_sample(d["v"], d["attn"], d["mask"], d["sample_count"])
    assert torch.equal(ids, d["ids"]) and torch.equal(nm, d["new_mask"]) and torch.equal(na, d["new_attn"])


def test_dyvit(gops):
    d = gops["dyvit"]
    assert torch.equal(O.dyvit_pool_concat(d["h"], d["policy"]), d["feat"])
    out, keep = O.dyvit_keep(d["x"], d["score"], d["k"])
    assert torch.equal(keep, d["keep"]) and torch.equal(out, d["x_out"])


def micro_cfg(gm, name):
    m = gm["micro"]
    return OM.Cfg(embed_dim=m["embed_dim"], num_heads=m["num_heads"], depth=m["depth"],
                  keep_rate=[gm["methods"][name]["keep_rate"]], reduction_loc=gm["reduction_loc"])


@pytest.mark.parametrize("name", ["topk", "evit", "tome", "dyvit", "dpcknn", "kmedoids", "sinkhorn", "patchmerger", "ats", "sit"])
def test_model_micro(gmodels, name):
    ent = gmodels["methods"][name]
    sd = {k: v.clone() for k, v in ent["state_dict"].items()}
    rec = {}
    torch.manual_seed(300)
    logits = OM.forward(name, sd, gmodels["images_fp16"].float(), micro_cfg(gmodels, name), record=rec)
    assert torch.allclose(logits, ent["logits"], rtol=1e-4, atol=1e-5), float((logits - ent["logits"]).abs().max())
    dec = ent["decisions"]
    for i in gmodels["reduction_loc"]:
        if name in ("topk", "dyvit"):
            assert torch.equal(rec[i], dec["Kept_Tokens"][i])
        elif name == "evit":
            assert torch.equal(rec[i], dec["Kept_Tokens"][i])
        elif name == "tome":
            assert torch.equal(rec[i], dec["Assignment_Maps"][i])
        elif name in ("dpcknn", "kmedoids"):
            assert torch.equal(rec[i][0], dec["Kept_Tokens"][i]) and torch.equal(rec[i][1], dec["Assignment_Maps"][i])
        elif name == "ats":
            assert torch.equal(rec[i][:, 1:] - 1, dec["Kept_Tokens"][i])
        else:
            assert (torch.argmax(rec[i], dim=-2) == dec["Assignment_Maps"][i]).float().mean() > 0.99
