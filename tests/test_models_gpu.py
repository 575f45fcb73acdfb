"""GPU model-level parity: the drop-in models (tokred kernels) against (1) the golden logits/decisions the
unmodified reference produced on CPU for the micro models and (2) the oracle port run on the SAME device and
weights (identical cuBLAS backbone, only the reduction operators differ), fp32 and bf16 autocast."""
import contextlib
import io
import os
from argparse import Namespace

import pytest
import torch

from oracle import model as OM
from oracle import ops as OP

import margins as MG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
METHODS = ["topk", "evit", "tome", "dyvit", "dpcknn", "kmedoids", "sinkhorn", "patchmerger", "ats", "sit"]
KR = {"topk": 0.7, "evit": 0.5, "tome": 0.7, "dyvit": 0.5, "dpcknn": 0.25, "kmedoids": 0.25, "sinkhorn": 0.9,
      "patchmerger": 0.9, "ats": 0.9, "sit": 0.9}


def margs(kr, loc=(3, 6, 9), **kw):
    return Namespace(keep_rate=[kr], reduction_loc=list(loc), distillation_type="none", k_neighbors=5, cluster_iters=3,
                     sinkhorn_eps=1.0, equal_weight=False, dyvit_distill=False, **kw)


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "models.pt"))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ------------------------------------------------------------------------------------------------ SURVEY §8c.3
TAU = 1e-5          # relative fp32 decision margin below which an image counts as tied


def _product_dist_eps(x, scale):
    """measured discrepancy between the product's distance kernel (3xTF32 Gram on tcgen05) and the reference's cdist on
    the very tokens of this stage: (off-diagonal max |diff|, max self distance) -- the inputs of margins.*_decidable."""
    from tokenreduction_b200 import ops as T
    d_own = T.pairwise_dist(x.contiguous(), scale, False)
    d_ref = OP.pairwise_dist(x) * scale
    p = d_own.shape[-1]
    eye = torch.eye(p, dtype=torch.bool, device=x.device)
    eps = ((d_own - d_ref).abs().masked_fill(eye, 0.0).flatten(1).max(dim=1).values * 1.01 + 1e-9).cpu()
    diag = torch.maximum(d_own.diagonal(dim1=1, dim2=2).abs().max(dim=1).values,
                         d_ref.diagonal(dim1=1, dim2=2).abs().max(dim=1).values).cpu()
    return d_ref, eps, diag


def stage_tie_free(name, inp, amp, tau=TAU, stage=1, same_device=True):
    """[B] bool: every decision of this stage of the ORACLE run has a float64 margin above tau (margins.py).  Soft
    merges make no discrete decision.  eps_d: the product kernel computes its own distances (3xTF32 Gram), measured
    at ~5e-7 (scaled by 1/sqrt(C)) / ~1e-5 (unscaled) from the reference's cdist (test_pairwise_dist).
    Where the product's decision inputs are bit-identical to the oracle's by construction (Top-K scores
    come from the identical ATen call and its gather is a verbatim copy, so at EVERY stage; EViT at the first stage)
    no margin is needed: ties break toward the lowest index on both sides."""
    if same_device and (name == "topk" or (name == "evit" and stage == 0)):
        return torch.ones(inp["scores"].shape[0], dtype=torch.bool)
    if name in ("topk", "evit", "dyvit"):
        return MG.topk_order_margin(inp["scores"], inp["k"], tau)
    if name == "tome":
        if amp:
            return MG.tome_bf16_decidable(inp["metric"], inp["r"])[0]
        return MG.tome_fp32_margin(inp["metric"]) > tau
    if name == "dpcknn":
        x = inp["x"]
        d, eps, diag = _product_dist_eps(x, float(torch.tensor(1.0) / torch.tensor(float(x.shape[-1]) ** 0.5)))
        if not same_device:
            eps = eps.clamp_min(tau)
        return MG.dpcknn_decidable(d, inp["noise"], inp["K"], inp["knn"], eps_d=eps, eps_diag=diag)[0]
    if name == "kmedoids":
        d, eps, diag = _product_dist_eps(inp["x"], 1.0)
        if not same_device:
            eps = eps.clamp_min(tau * float(d.max()))
        return MG.kmedoids_decidable(d, inp["tw"], inp["K"], inp["iters"], eps_d=eps, eps_diag=diag,
                                     rel_w=0.0 if (stage == 0 and same_device) else tau)[0]
    if name == "ats":
        cdf = OP.ats_significance(inp["v"], inp["attn"]).cumsum(dim=1)
        cdf = torch.where(inp["mask"][:, 1:], cdf, cdf + 0.1)
        return MG.ats_decidable(cdf, OP.ats_sample_steps(inp["count"]))
    raise KeyError(name)


def product_decisions(name, viz, i):
    """the product model's stage-i decisions from its viz dict, in the layout of the oracle's record."""
    if name in ("topk", "evit", "dyvit", "ats"):
        return [torch.as_tensor(viz["Kept_Tokens"][i])]
    if name == "tome":
        return [torch.as_tensor(viz["Assignment_Maps"][i])]
    return [torch.as_tensor(viz["Kept_Tokens"][i]), torch.as_tensor(viz["Assignment_Maps"][i])]


def oracle_decisions(name, rec, i):
    r = rec[i]
    if name == "ats":
        return [(r[:, 1:] - 1).cpu()]
    if isinstance(r, tuple):
        return [t.cpu() for t in r]
    return [r.cpu()]


def same_rows(a, b):
    if a.shape != b.shape:      # ATS: padded widths may differ when a tied image changes the batch maximum
        w = min(a.shape[1], b.shape[1])
        rest_a, rest_b = a[:, w:], b[:, w:]
        pad_ok = (rest_a <= 0).all(dim=1) if rest_a.numel() else torch.ones(a.shape[0], dtype=torch.bool)
        pad_ok &= (rest_b <= 0).all(dim=1) if rest_b.numel() else torch.ones(a.shape[0], dtype=torch.bool)
        return (a[:, :w] == b[:, :w]).all(dim=1) & pad_ok
    return (a == b).flatten(1).all(dim=1)


def _canon(name, tensors):
    """order-insensitive form of one stage's decisions (the reference's own consumers compare kept indices as SETS and
    assignment maps as partitions): kept ids sorted; cluster labels replaced by their centre's token id."""
    if name in ("topk", "evit", "dyvit", "ats"):
        t = tensors[0].long()
        return [torch.where(t < 0, torch.full_like(t, 1 << 30), t).sort(dim=1).values]
    if name in ("dpcknn", "kmedoids"):
        centres, assign = tensors[0].long(), tensors[1].long()
        return [centres.sort(dim=1).values, torch.gather(centres, 1, assign.clamp(0, centres.shape[1] - 1))]
    return [t.long() for t in tensors]


def compare_with_oracle(name, model, sd, x, cfg, amp, tol, seed=7):
    """runs product and oracle on the same device / weights / generator state; returns a dict of per-image flags."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        y, viz = model(x)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    rec = {}
    y_ref = OM.forward(name, sd, x, cfg, amp=amp, record=rec).float()
    y = y.float()
    b = x.shape[0]
    stages = sorted(k for k in rec if isinstance(k, int))
    has_decisions = name not in ("sinkhorn", "patchmerger", "sit")
    tie_free = torch.ones(b, dtype=torch.bool)
    first_tie_free = None
    agree = torch.ones(b, dtype=torch.bool)
    set_agree = torch.ones(b, dtype=torch.bool)
    first_agree = None
    for si, i in enumerate(stages):
        if not has_decisions:
            break
        tf = stage_tie_free(name, rec[("in", i)], amp, stage=si)
        eq = torch.ones(b, dtype=torch.bool)
        for got, want in zip(product_decisions(name, viz, i), oracle_decisions(name, rec, i)):
            eq &= same_rows(got.long() if got.dtype != want.dtype else got, want.long() if got.dtype != want.dtype else want)
        for got, want in zip(_canon(name, product_decisions(name, viz, i)), _canon(name, oracle_decisions(name, rec, i))):
            set_agree &= same_rows(got, want)
        if si == 0:
            first_tie_free, first_agree = tf.clone(), eq.clone()
        tie_free &= tf
        agree &= eq
    rel = ((y - y_ref).norm(dim=1) / y_ref.norm(dim=1)).cpu()
    return {"tie_free": tie_free, "agree": agree, "set_agree": set_agree, "first_tie_free": first_tie_free,
            "first_agree": first_agree, "rel": rel, "finite": bool(torch.isfinite(y).all()), "y": y, "y_ref": y_ref}


@pytest.mark.parametrize("name", METHODS)
def test_micro_model_vs_reference_golden(gold, name, monkeypatch):
    """the drop-in model with the UNMODIFIED reference's state_dict on the golden images: decisions equal the reference's
    recorded decisions on every image the oracle run marks tie-free, logits within tolerance on those images.
    DPC-KNN: the golden run drew its tie-break noise from the CPU generator (seed 300) -- torch.rand is redirected to the
    CPU generator here so that both sides see the same noise."""
    from tokenreduction_b200 import factory
    ent = gold["methods"][name]
    cls = factory._METHODS[name]
    extra = {"dyvit_distillation": False} if name == "dyvit" else {}
    model = quiet(cls, args=margs(ent["keep_rate"], gold["reduction_loc"], viz_mode=True), **gold["micro"], **extra).eval()
    model.load_state_dict(ent["state_dict"])
    model = model.cuda()
    images = gold["images_fp16"].float().cuda()
    orig_rand = torch.rand

    def cpu_rand(*size, device=None, **kw):
        return orig_rand(*size, **kw).to(device) if device is not None else orig_rand(*size, **kw)
    monkeypatch.setattr(torch, "rand", cpu_rand)
    torch.manual_seed(300)
    with torch.no_grad():
        logits, viz = model(images)
    # tie-free images according to the oracle run on this device with the same weights
    torch.manual_seed(300)
    rec = {}
    micro = gold["micro"]
    cfg = OM.Cfg(embed_dim=micro["embed_dim"], num_heads=micro["num_heads"], depth=micro["depth"],
                 keep_rate=[ent["keep_rate"]], reduction_loc=list(gold["reduction_loc"]))
    sd = {k: v.detach().clone().cuda() for k, v in ent["state_dict"].items()}
    OM.forward(name, sd, images, cfg, record=rec)
    b = images.shape[0]
    tie_free = torch.ones(b, dtype=torch.bool)
    if name not in ("sinkhorn", "patchmerger", "sit"):
        for i in sorted(k for k in rec if isinstance(k, int)):
            tie_free &= stage_tie_free(name, rec[("in", i)], False, tau=1e-4, stage=99, same_device=False)  # CPU vs cuBLAS backbone
    same = torch.ones(b, dtype=torch.bool)
    soft = name in ("sinkhorn", "patchmerger", "sit")      # hard assignment = argmax of the soft one: no margin model
    for key, stages in ent["decisions"].items():
        for i, ref in stages.items():
            got, ref = torch.as_tensor(viz[key][i]).long(), torch.as_tensor(ref).long()
            same &= same_rows(got, ref)
            if got.shape == ref.shape:          # every method: near-complete agreement even on tied images
                frac = float((got == ref).float().mean())
                assert frac > (0.9 if soft or name == "dpcknn" else 0.97), f"{name} stage {i} {key}: only {frac:.3f} of decisions match"
    print(f"{name}: golden tie-free {tie_free.tolist()} decisions identical {same.tolist()}")
    if not soft:
        assert bool(same[tie_free].all()), f"{name}: a tie-free image differs from the reference's golden decisions"
    ok = tie_free & same
    if bool(ok.any()):
        err = ((logits.cpu() - ent["logits"]).norm(dim=1) / ent["logits"].norm(dim=1))[ok].max()
        assert float(err) < 2e-3, f"{name}: logits differ from the reference golden by {float(err):.2e} (relative)"


# north_star bars: 1e-5 (fp32) holds at the LOGITS too (measured worst case over the ten families: 2.5e-6).  bf16: the
# 1e-2 bar is the op-level one (identical bf16 inputs, tests/test_ops_gpu.py); at the logits a one-ulp bf16 difference
# in one element (4e-3) passes through up to 9 more blocks -- most images are bit-identical, the rest reach 5e-3 .. 1.2e-2,
# inside the model's own bf16-vs-fp32 noise of 1-3e-2: model-level bf16 bar 2e-2.
TOL_FP32 = 1e-5
TOL_BF16 = 2e-2


@pytest.mark.parametrize("amp", [False, True])
@pytest.mark.parametrize("name", METHODS)
def test_small_model_vs_oracle_same_device(name, amp, monkeypatch):
    """SURVEY §8c.3.  Product model vs oracle port on the SAME device, weights and generator state (identical cuBLAS
    backbone; only the reduction operators differ), DeiT-S, 16 images:
      * fp32: every image whose decisions are ALL above the margin (tie-free) must make identical decisions at every
        stage and agree in the logits; the excluded fraction is printed and bounded.
      * bf16 autocast: stage-1 decisions (bit-identical inputs on both sides) must be identical on stage-1 tie-free
        images; later stages see inputs that differ in the last bf16 bit, where bf16 scores are full of exact ties
        (SURVEY A.10), so the logits bar applies to images whose decisions all agree (conditioned parity)."""
    from tokenreduction_b200 import create_model, modules
    # the reduction operators alone: the backbone's attention stays the reference's ATen sequence on both sides (the
    # fused attention producer has its own model-level test below, test_small_model_fused_attention)
    monkeypatch.setattr(modules, "FUSED_ATTENTION", False)
    b = 16
    size = "small"
    torch.manual_seed(0)
    model = quiet(create_model, f"{name}_{size}_patch16_224", num_classes=100, args=margs(KR[name], viz_mode=True)).eval().cuda()
    with torch.no_grad():      # spread the reduction parameters: random init leaves every decision tied
        for n_, p_ in model.named_parameters():
            if n_.startswith("cluster_layers") and p_.dim() >= 2 and "queries" not in n_ and not n_.endswith(".v"):
                p_.mul_(20.0)
            if n_.startswith("score_predictor") and p_.dim() >= 2:
                p_.mul_(4.0)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x = torch.randn(b, 3, 224, 224, generator=torch.Generator().manual_seed(1)).cuda()
    r = compare_with_oracle(name, model, sd, x, OM.cfg_for(size, keep_rate=[KR[name]]), amp, None)
    assert r["finite"]
    tol = TOL_BF16 if amp else TOL_FP32
    tf, ag, rel = r["tie_free"], r["agree"], r["rel"]
    print(f"{name} amp={amp}: tie-free {int(tf.sum())}/{b}, decisions identical {int(ag.sum())}/{b}, "
          f"worst rel err on identical-decision images {float(rel[ag].max()) if bool(ag.any()) else float('nan'):.2e}, "
          f"all {['%.1e' % v for v in rel.tolist()]}")
    if not amp:
        assert bool(ag[tf].all()), f"{name}: {int((~ag[tf]).sum())} tie-free images made different decisions"
        if bool(tf.any()):
            assert float(rel[tf].max()) <= tol, f"{name}: tie-free image logits differ by {float(rel[tf].max()):.2e} > {tol}"
        # the margin models are worst-case bounds (and random-init ATS / DynamicViT scores are tied on every image,
        # A.10): an empirical floor keeps the check from passing vacuously when few images can be certified
        assert float(ag.float().mean()) >= 0.75, f"{name}: only {int(ag.sum())}/{b} images made identical decisions"
        assert float(rel[ag].max()) <= tol, f"{name}: logits differ by {float(rel[ag].max()):.2e} > {tol} with identical decisions"
    else:
        if r["first_tie_free"] is not None:
            f_tf, f_ag = r["first_tie_free"], r["first_agree"]
            assert bool(f_ag[f_tf].all()), f"{name}: stage-1 decisions differ on {int((~f_ag[f_tf]).sum())} tie-free images"
        if bool(ag.any()):
            assert float(rel[ag].max()) <= tol, f"{name}: logits differ by {float(rel[ag].max()):.2e} > {tol} with identical decisions"
        assert float(ag.float().mean()) >= 0.25, f"{name} amp: only {int(ag.sum())}/{b} images with identical decisions"


@pytest.mark.parametrize("name", METHODS)
def test_small_model_fused_attention(name):
    """SURVEY §8f row 1 at model level: bf16 autocast with the fused attention producer (no [B,H,N,N] tensor) against
    the oracle port (materialised attention) on the same device and weights.  The producer's outputs differ from the
    ATen sequence's in the last bf16 bit of a few elements (op-level: tests/test_ops_gpu.py::test_attention_*), so from
    the first block on both sides see different inputs and bit-identical DECISIONS are no longer defined (bf16 scores
    are full of exact ties).  What must hold:
      * the logits stay inside the model's own bf16 noise, measured on the very same images as the oracle's
        autocast-vs-fp32 discrepancy n_i = || y_oracle_bf16 - y_oracle_fp32 ||: over the batch, max and median of
        d_i = || y - y_oracle_bf16 || are <= 2x those of n_i (a flipped decision moves an image by 0.1-0.3 of its
        logit norm on either side, SURVEY A.3, so the comparison is between distributions);
      * conditioned parity: an image whose decisions agree with the oracle's as sets / partitions (always, for the
        soft merges) meets the bf16 bar or d_i <= 2 n_i."""
    from tokenreduction_b200 import create_model, modules
    assert modules.FUSED_ATTENTION
    b, size = 16, "small"
    torch.manual_seed(0)
    model = quiet(create_model, f"{name}_{size}_patch16_224", num_classes=100, args=margs(KR[name], viz_mode=True)).eval().cuda()
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if n_.startswith("cluster_layers") and p_.dim() >= 2 and "queries" not in n_ and not n_.endswith(".v"):
                p_.mul_(20.0)
            if n_.startswith("score_predictor") and p_.dim() >= 2:
                p_.mul_(4.0)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x = torch.randn(b, 3, 224, 224, generator=torch.Generator().manual_seed(1)).cuda()
    cfg = OM.cfg_for(size, keep_rate=[KR[name]])
    from tokenreduction_b200 import _lib
    n0 = _lib.launch_count()
    r = compare_with_oracle(name, model, sd, x, cfg, True, None)
    assert _lib.launch_count() - n0 >= 12, "the fused attention kernel did not run in every block"
    assert r["finite"]
    torch.manual_seed(7)
    torch.cuda.manual_seed(7)
    y32 = OM.forward(name, sd, x, cfg, amp=False).float()
    noise = (r["y_ref"] - y32).norm(dim=1).cpu()
    diff = (r["y"] - r["y_ref"]).norm(dim=1).cpu()
    sa, rel = r["set_agree"], r["rel"]
    print(f"{name} fused: decisions agree as sets {int(sa.sum())}/{b}; |y-y_ref| / bf16 noise of the oracle: "
          f"{['%.2f' % v for v in (diff / noise).tolist()]}; rel {['%.1e' % v for v in rel.tolist()]}")
    assert float(diff.max()) <= 2.0 * float(noise.max()) and float(diff.median()) <= 2.0 * float(noise.median()), \
        f"{name}: logits outside the model's own bf16 noise (max {float(diff.max()):.3f} vs {float(noise.max()):.3f}, " \
        f"median {float(diff.median()):.3f} vs {float(noise.median()):.3f})"
    ok = (rel <= TOL_BF16) | (diff <= 2.0 * noise)
    assert bool(ok[sa].all()), f"{name}: image {int((~ok & sa).nonzero()[0])} differs with set-identical decisions"
    if name in ("sinkhorn", "patchmerger", "sit"):
        assert bool(sa.all())


def test_batch_shard_equivalence():
    """(e) multi-GPU = pure batch sharding: per-image independence means shard outputs concatenate exactly."""
    from tokenreduction_b200 import create_model
    torch.manual_seed(0)
    model = quiet(create_model, "tome_small_patch16_224", num_classes=50, args=margs(0.7)).eval().cuda()
    x = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.no_grad():
        full = model(x)
        parts = torch.cat([model(x[:4]), model(x[4:])])
    assert torch.allclose(full, parts, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["tome", "topk", "kmedoids", "ats", "sinkhorn"])
def test_cuda_graph_replay_equals_eager(name):
    """tokenreduction_b200.graph.GraphedForward: the whole bf16 forward (every tokred entry point only enqueues on the
    current stream) captured into one CUDA graph reproduces the eager forward bit for bit, also on a second input."""
    from tokenreduction_b200 import create_model
    from tokenreduction_b200.graph import GraphedForward
    torch.manual_seed(0)
    model = quiet(create_model, f"{name}_small_patch16_224", num_classes=100, args=margs(KR[name])).eval().cuda()
    xs = [torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(s)).cuda() for s in (3, 4)]
    run = GraphedForward(model, xs[0], torch.bfloat16)
    for x in xs:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            want = model(x).clone()
        got = run(x)
        assert torch.equal(got, want)


@pytest.mark.parametrize("name", METHODS)
def test_fp16_autocast_reference_default(name):
    """validate.py:52-54 evaluates under torch.cuda.amp.autocast() = fp16.  The drop-in runs it: half tensors are converted
    at the op boundary (fp16 -> fp32 is exact), the fused bf16 producers stay off, and the logits stay within the
    half-precision noise of the fp32 forward of the same model."""
    from tokenreduction_b200 import create_model
    torch.manual_seed(0)
    model = quiet(create_model, f"{name}_tiny_patch16_224", num_classes=50, args=margs(KR[name])).eval().cuda()
    x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
    torch.manual_seed(9)
    torch.cuda.manual_seed(9)
    with torch.no_grad():
        y32 = model(x)
    torch.manual_seed(9)
    torch.cuda.manual_seed(9)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        y16 = model(x)
    assert torch.isfinite(y16).all()
    rel = ((y16.float() - y32).norm(dim=1) / y32.norm(dim=1)).max()
    assert float(rel) < 0.35, f"{name}: fp16-autocast logits {float(rel):.3f} away from fp32"


@pytest.mark.parametrize("name", METHODS)
@pytest.mark.parametrize("viz", [False, True])
def test_deferred_sums_are_bit_identical(name, viz, monkeypatch):
    """modules.DEFER_RESIDUAL (residual sums formed by the next block's add_layernorm, the embedding formed by the first
    norm1, the ToMe merge and the Top-K / EViT selects carrying the block's residual add) only moves WHERE the same fp32
    additions and LayerNorms happen: logits -- and in viz mode every recorded decision and feature map -- are bit-identical
    to the undeferred bf16 forward, for all ten families."""
    from tokenreduction_b200 import create_model, modules
    torch.manual_seed(0)
    model = quiet(create_model, f"{name}_small_patch16_224", num_classes=100, args=margs(KR[name], viz_mode=viz)).eval().cuda()
    x = torch.randn(6, 3, 224, 224, generator=torch.Generator().manual_seed(11)).cuda()

    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}

    def run(defer):
        model.load_state_dict(sd)                # Sinkhorn re-normalises (overwrites) its parameter on every forward (:73-76)
        monkeypatch.setattr(modules, "DEFER_RESIDUAL", defer)
        torch.manual_seed(5)
        torch.cuda.manual_seed(5)                # DPC-KNN draws its tie-breaking noise from the CUDA generator
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(x)

    made = []

    class CountingResidual(modules.Residual):
        def __init__(self, *a_, **k_):
            made.append(1)
            super().__init__(*a_, **k_)

    monkeypatch.setattr(modules, "Residual", CountingResidual)
    a = run(True)
    assert len(made) >= 8, "the residual sums were not deferred"
    made.clear()
    b = run(False)
    assert not made
    if viz:
        (ya, va), (yb, vb) = a, b
        assert torch.equal(ya, yb)
        assert va.keys() == vb.keys()
        for key in va:
            assert va[key].keys() == vb[key].keys(), key
            for i in va[key]:
                ta, tb = torch.as_tensor(va[key][i]), torch.as_tensor(vb[key][i])
                assert torch.equal(ta, tb), f"{name}: viz output {key}[{i}] differs"
    else:
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["kmedoids", "dpcknn"])
@pytest.mark.parametrize("amp", [False, True])
def test_equal_weight_models_run(name, amp):
    """`--equal_weight` (train.py flag; models/kmedoids.py:43-61, models/dpcknn.py:151-158): the drop-in models run it -- round 1
    raised NotImplementedError for K-Medoids.  Same numpy / torch seeds -> same logits; a different first medoid -> the
    K-Medoids logits move."""
    import numpy as np
    from tokenreduction_b200 import create_model
    torch.manual_seed(0)
    args = margs(KR[name])
    args.equal_weight = True
    model = quiet(create_model, f"{name}_small_patch16_224", num_classes=100, args=args).eval().cuda()
    x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(21)).cuda()

    def run(seed):
        np.random.seed(seed)
        torch.manual_seed(9)
        torch.cuda.manual_seed(9)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            return model(x).float()

    a, b, c = run(1), run(1), run(2)
    assert bool(torch.isfinite(a).all()) and a.shape == (4, 100)
    assert torch.equal(a, b)
    if name == "kmedoids":
        assert not torch.equal(a, c)


def test_standalone_block_returns_tensors():
    """the drop-in blocks keep the reference's return types when called on their own: residual sums are deferred only inside the
    block loops of this package's models (which set ``defer_out`` on their blocks and materialise on every other read)."""
    from tokenreduction_b200 import modules as M
    torch.manual_seed(0)
    x = torch.randn(4, 197, 384, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for blk in (M.Block_ToMe(384, 6, qkv_bias=True, r=59), M.Block_TopK(384, 6, qkv_bias=True, keep_rate=0.7),
                    M.Block_EVIT(384, 6, qkv_bias=True, keep_rate=0.7), M.BlockWithProbs(384, 6, qkv_bias=True)):
            blk = blk.eval().cuda()
            out = blk(x)
            first = out[0] if isinstance(out, tuple) else out
            assert isinstance(first, torch.Tensor) and first.dtype == torch.float32 and bool(torch.isfinite(first).all())
            blk.defer_out = True                      # what a model of this package sets on the blocks of its loop
            out2 = blk(x)
            first2 = out2[0] if isinstance(out2, tuple) else out2
            assert isinstance(first2, M.Residual) and torch.equal(first2.value(), first)
