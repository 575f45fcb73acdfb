"""CPU property tests of the oracle restatement: invariants that must hold for ANY input, used
by the GPU tests at full BASELINE sizes where element-wise comparison against a CPU run is too slow."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import ops as O


def g(seed):
    return torch.Generator().manual_seed(seed)


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 60), st.integers(1, 59), st.integers(0, 10 ** 6))
def test_topk_is_descending_subset(p, k, seed):
    k = min(k, p)
    x = torch.randn(2, p + 1, 8, generator=g(seed))
    s = torch.randn(2, p, generator=g(seed + 1))
    out, idx = O.topk_gather(x, s, k)
    vals = torch.gather(s, 1, idx)
    assert bool((vals[:, 1:] <= vals[:, :-1]).all())                      # descending score order (quirk B.2)
    assert all(len(set(r.tolist())) == k for r in idx)                     # a subset, no repeats
    assert torch.equal(out[:, 0], x[:, 0])                                 # CLS untouched
    assert bool((torch.gather(s, 1, idx).min(dim=1).values >= s.kthvalue(p - k + 1, dim=1).values).all())


@settings(max_examples=25, deadline=None)
@given(st.integers(3, 60), st.integers(1, 58), st.integers(0, 10 ** 6))
def test_evit_fused_token_and_partition(p, k, seed):
    k = min(k, p - 1)
    x = torch.randn(2, p + 1, 6, generator=g(seed))
    s = torch.rand(2, p, generator=g(seed + 1))
    out, idx, compl = O.evit_select_fuse(x, s, k)
    assert out.shape == (2, k + 2, 6) and bool((idx[:, -1] == -1).all())
    both = torch.cat([idx[:, :-1], compl], dim=1).sort(dim=1).values
    assert torch.equal(both, torch.arange(p).expand(2, -1))                # kept + complement partition the patches
    assert bool((compl[:, 1:] > compl[:, :-1]).all())                      # complement ascending
    ref = (torch.gather(s, 1, compl).unsqueeze(-1) * torch.gather(x[:, 1:], 1, compl.unsqueeze(-1).expand(-1, -1, 6))).sum(1)
    assert torch.allclose(out[:, -1], ref, rtol=1e-5, atol=1e-6)           # unnormalised weighted sum (quirk B.6)


@settings(max_examples=25, deadline=None)
@given(st.integers(4, 80), st.integers(1, 60), st.integers(0, 10 ** 6))
def test_tome_mass_and_sizes(n, r, seed):
    metric = torch.randn(2, n, 8, generator=g(seed))
    x = torch.randn(2, n, 5, generator=g(seed + 1))
    size = torch.randint(1, 4, (2, n, 1), generator=g(seed + 2)).float()
    re = O.tome_effective_r(n, r)
    if re == 0:
        return
    unm, src, dst, _ = O.tome_match(metric, r, True)
    out, size_out, rci = O.tome_merge(x, size, unm, src, dst)
    assert out.shape[1] == n - re
    assert torch.allclose(size_out.sum(1), size.sum(1))                    # sizes are conserved
    assert torch.allclose((out * size_out).sum(1), (x * size).sum(1), rtol=1e-4, atol=1e-4)   # mass is conserved
    # CLS first and never merged; (x*size)/size is the reference's arithmetic, so only equal up to one rounding
    assert bool((unm[:, 0] == 0).all()) and torch.allclose(out[:, 0], x[:, 0], rtol=1e-6, atol=1e-7)
    assert torch.equal(size_out[:, 0], size[:, 0])
    assert int(rci.min()) >= 0 and int(rci.max()) <= n - re - 2
    # every output row except CLS receives at least one input patch
    for b in range(2):
        assert set(rci[b].long().tolist()) == set(range(n - re - 1))


@settings(max_examples=15, deadline=None)
@given(st.integers(6, 40), st.integers(1, 6), st.integers(0, 10 ** 6))
def test_dpcknn_centres_and_weights(p, k, seed):
    k = min(k, p)
    x = torch.randn(2, p, 7, generator=g(seed))
    noise = torch.rand(2, p, generator=g(seed + 1))
    idx_cluster, index_down = O.dpcknn_cluster(x, k, min(5, p), noise)
    assert torch.equal(torch.gather(idx_cluster, 1, index_down), torch.arange(k).expand(2, -1))   # centres own their cluster
    assert int(idx_cluster.min()) >= 0 and int(idx_cluster.max()) < k
    tw = torch.rand(2, p, 1, generator=g(seed + 2)) + 0.1
    idx_token = torch.arange(p).expand(2, -1).contiguous()
    xm, it, aw = O.dpcknn_merge(x, idx_token, torch.ones(2, p, 1), idx_cluster, k, tw)
    # normalised weights of a cluster sum to ~1 (up to the +1e-6 in the denominator)
    sums = torch.zeros(2, k).scatter_add_(1, idx_cluster, aw.squeeze(-1))
    assert torch.allclose(sums, torch.ones(2, k), atol=1e-4)
    assert torch.equal(it, idx_cluster)


@settings(max_examples=15, deadline=None)
@given(st.integers(6, 40), st.integers(1, 6), st.integers(0, 3), st.integers(0, 10 ** 6))
def test_kmedoids_outputs_are_medoids(p, k, iters, seed):
    k = min(k, p)
    x = torch.randn(2, p, 5, generator=g(seed))
    tw = torch.rand(2, p, 1, generator=g(seed + 1)) + 0.5
    centres, cidx, assign = O.kmedoids_fit(x, k, iters, tw)
    assert torch.equal(centres, O.gather_rows(x, cidx))                    # medoid tokens verbatim (quirk B.8)
    d = O.pairwise_dist(x)
    best = torch.gather(d, 2, cidx.unsqueeze(1).expand(-1, p, -1)).argmin(-1)
    assert torch.equal(assign, best)                                       # final assignment = nearest centre


@settings(max_examples=15, deadline=None)
@given(st.integers(4, 40), st.integers(2, 30), st.integers(1, 5), st.integers(0, 10 ** 6))
def test_sinkhorn_marginals(p, k, iters, seed):
    x = torch.randn(2, p, 9, generator=g(seed))
    v = torch.randn(k, 9, generator=g(seed + 1))
    out, w, vh = O.sinkhorn_merge(x, v, 1.0, iters)
    assert torch.allclose(vh.norm(dim=-1), torch.ones(k), atol=1e-5)
    # after the last column update every token carries unit mass (transport plan scaled by K+P)
    assert torch.allclose(w.sum(dim=1), torch.ones(2, p), rtol=1e-4, atol=1e-5)
    assert bool((w >= 0).all()) and out.shape == (2, k, 9)


@settings(max_examples=15, deadline=None)
@given(st.integers(8, 60), st.integers(2, 40), st.integers(0, 10 ** 6))
def test_ats_ids_sorted_unique_padded(n, count, seed):
    count = min(count, n - 1)
    if count < 2:
        return
    attn = torch.softmax(3 * torch.randn(2, 2, n, n, generator=g(seed)), dim=-1)
    v = torch.randn(2, 2, n, 4, generator=g(seed + 1))
    mask = torch.ones(2, n, dtype=torch.bool)
    na, nm, ids = O.ats_sample(v, attn, mask, count)
    assert bool((ids[:, 0] == 0).all()) and bool(nm[:, 0].all())
    body = ids[:, 1:]
    nz = body != 0
    assert torch.equal(nm[:, 1:], nz)
    srt = torch.where(nz, body, torch.full_like(body, 10 ** 6))
    assert bool((srt[:, 1:] >= srt[:, :-1]).all())                          # ascending, zero padding at the end
    # width <= #steps + 1 (torch.arange with a float step yields K-1 or K steps depending on rounding)
    assert ids.shape[1] <= O.ats_sample_steps(count).numel() + 1 and na.shape == (2, 2, ids.shape[1], n)
