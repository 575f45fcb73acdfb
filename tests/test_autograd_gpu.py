"""autograd through the tokred gather / merge ops and the DynamicViT training path.
The forward of every op is the tokred kernel; its backward formula (ops.py) is checked against PyTorch autograd through
the oracle restatement of the same op on the same device and inputs."""
import contextlib
import io
from argparse import Namespace

import pytest
import torch

from oracle import model as OM
from oracle import ops as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def T():
    import tokenreduction_b200.ops as ops
    from tokenreduction_b200 import _lib
    _lib.load()
    return ops


def g(seed):
    return torch.Generator().manual_seed(seed)


def _leaf(t):
    return t.to(DEV).requires_grad_(True)


def _close(a, b, what, rtol=1e-5):
    a, b = a.detach(), b.detach()
    err = float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    assert err <= rtol, f"{what}: relative error {err:.3e}"


def _scores(b, p, seed):
    base = torch.linspace(0.05, 1.0, p)
    return torch.stack([base[torch.randperm(p, generator=g(seed + i))] for i in range(b)])


def test_topk_gather_backward(T):
    b, n, c, k = 5, 197, 384, 137
    x0, s0, w = torch.randn(b, n, c, generator=g(1)), _scores(b, n - 1, 2), torch.randn(b, k + 1, c, generator=g(3)).to(DEV)
    x, s = _leaf(x0), s0.to(DEV)
    out, idx = T.topk_gather(x, s, k)
    (out * w).sum().backward()
    xr = _leaf(x0)
    out_r, idx_r = O.topk_gather(xr, s, k)
    (out_r * w).sum().backward()
    assert torch.equal(idx, idx_r) and torch.equal(out, out_r)
    assert torch.equal(x.grad, xr.grad)


def test_evit_select_fuse_backward(T):
    b, n, c, k = 4, 197, 768, 98
    x0, s0 = torch.randn(b, n, c, generator=g(4)), _scores(b, n - 1, 5)
    w = torch.randn(b, k + 2, c, generator=g(6)).to(DEV)
    x, s = _leaf(x0), _leaf(s0)
    out, idx, compl = T.evit_select_fuse(x, s, k)
    (out * w).sum().backward()
    xr, sr = _leaf(x0), _leaf(s0)
    out_r, idx_r, compl_r = O.evit_select_fuse(xr, sr, k)
    (out_r * w).sum().backward()
    assert torch.equal(idx, idx_r) and torch.equal(compl, compl_r)
    _close(x.grad, xr.grad, "evit dx")
    _close(s.grad, sr.grad, "evit dscores")


def test_gather_rows_backward_with_repeated_ids(T):
    b, n, c, m = 3, 177, 768, 143
    x0 = torch.randn(b, n, c, generator=g(7))
    ids = torch.sort(torch.randint(0, n, (b, m), generator=g(8)), dim=1).values
    ids[:, -9:] = 0                                  # ATS 0-padding: repeated index
    ids = ids.to(DEV)
    w = torch.randn(b, m, c, generator=g(9)).to(DEV)
    x = _leaf(x0)
    (T.gather_rows(x, ids) * w).sum().backward()
    xr = _leaf(x0)
    (O.gather_rows(xr, ids) * w).sum().backward()
    _close(x.grad, xr.grad, "gather_rows dx")


@pytest.mark.parametrize("with_size", [False, True])
def test_tome_merge_backward(T, with_size):
    b, n, c, r = 4, 197, 384, 59
    x0 = torch.randn(b, n, c, generator=g(10))
    metric = torch.randn(b, n, 64, generator=g(11))
    size0 = torch.randint(1, 5, (b, n, 1), generator=g(12)).float() if with_size else None
    unm, src, dst, _ = O.tome_match(metric, r, True)
    unm, src, dst = unm.to(DEV), src.to(DEV), dst.to(DEV)
    size = None if size0 is None else size0.to(DEV)
    w = torch.randn(b, n - r, c, generator=g(13)).to(DEV)
    x = _leaf(x0)
    out, size_out, _ = T.tome_merge(x, size, unm, src, dst, True, True)
    (out * w).sum().backward()
    xr = _leaf(x0)
    out_r, size_r, _ = O.tome_merge(xr, size, unm, src, dst)
    (out_r * w).sum().backward()
    _close(out, out_r, "tome_merge out")      # CUDA scatter_add order is not defined: bit-exactness is checked against the CPU oracle
    assert torch.equal(size_out, size_r)
    _close(x.grad, xr.grad, "tome_merge dx")


def test_dyvit_pool_concat_backward(T):
    b, p, c = 4, 196, 768
    h0 = torch.randn(b, p, c, generator=g(14))
    pol0 = (torch.rand(b, p, 1, generator=g(15)) > 0.4).float()
    w = torch.randn(b, p, c, generator=g(16)).to(DEV)
    h, pol = _leaf(h0), _leaf(pol0)
    (T.dyvit_pool_concat(h, pol, 1e-6) * w).sum().backward()
    hr, polr = _leaf(h0), _leaf(pol0)
    (O.dyvit_pool_concat(hr, polr, 1e-6) * w).sum().backward()
    _close(h.grad, hr.grad, "pool dh")
    _close(pol.grad, polr.grad, "pool dpolicy", 1e-4)


def _margs(kr):
    return Namespace(keep_rate=[kr], reduction_loc=[3, 6, 9], distillation_type="none", k_neighbors=5, cluster_iters=3,
                     sinkhorn_eps=1.0, equal_weight=False, dyvit_distill=False)


def test_dyvit_training_forward_backward_vs_oracle():
    """the drop-in DynamicViT in train() mode (gumbel keep decisions, softmax_with_policy, models/dyvit.py:205-229; the
    predictor pooling is the tokred kernel with its autograd formula) against the oracle restatement, which
    tests/test_oracle_vs_reference.py pins to the unmodified reference: logits, hard decisions, parameter gradients."""
    from tokenreduction_b200 import create_model
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model("dyvit_tiny_patch16_224", num_classes=16, args=_margs(0.5)).cuda().train()
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if n_.startswith("score_predictor") and p_.dim() >= 2:
                p_.mul_(4.0)
    x = torch.randn(4, 3, 224, 224, generator=g(20)).cuda()
    names = ["score_predictor.0.in_conv.1.weight", "score_predictor.2.out_conv.4.weight", "blocks.4.attn.qkv.weight",
             "blocks.0.mlp.fc1.weight", "patch_embed.proj.weight"]
    torch.manual_seed(8)
    torch.cuda.manual_seed(8)
    y, dec = model(x)
    (y.square().sum() + sum(d.sum() for d in dec)).backward()
    params = dict(model.named_parameters())
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    torch.manual_seed(8)
    torch.cuda.manual_seed(8)
    y_r, dec_r = OM.dyvit_train_forward(sd, x, OM.cfg_for("tiny", keep_rate=[0.5]))
    (y_r.square().sum() + sum(d.sum() for d in dec_r)).backward()
    for a, b_ in zip(dec, dec_r):
        assert torch.equal(a, b_), "hard keep decisions differ"
    _close(y, y_r, "logits", 1e-4)
    for n in names:
        _close(params[n].grad, sd[n].grad, n, 2e-3)


@pytest.mark.parametrize("name,kr", [("topk", 0.7), ("evit", 0.5), ("tome", 0.7)])
def test_reduced_models_train_through_the_kernels(name, kr):
    """fine-tuning (train.py): forward + backward through the tokred select / merge kernels gives finite gradients for
    every parameter, and the kernels did run (launch counter)."""
    from tokenreduction_b200 import _lib, create_model
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(f"{name}_tiny_patch16_224", num_classes=16, args=_margs(kr)).cuda().train()
    x = torch.randn(4, 3, 224, 224, generator=g(21)).cuda()
    n0 = _lib.launch_count()
    y = model(x)
    y.square().sum().backward()
    assert _lib.launch_count() - n0 >= 3
    for n_, p_ in model.named_parameters():
        assert p_.grad is not None and torch.isfinite(p_.grad).all(), n_
    assert float(model.blocks[0].attn.qkv.weight.grad.abs().sum()) > 0
