"""GPU parity: piece 4 (tcgen05 catalog sweep: dense scores, target scores, fused top-K + rank),
piece 5 merge kernel and H1 metrics, against the oracle.  Bit-exact on grid-valued data (every
bf16 x bf16 -> fp32 partial sum is exactly representable, so the score does not depend on the
tensor core's summation order); on Gaussian data the fused kernel is checked bit-exactly against
the oracle's rank/top-K applied to the kernel's own dense scores, and the scores against fp64
within the stated tolerance."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import ref_path as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_lib):
    from recguru_b200 import ops as _ops
    return _ops


def grid(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-32, 33, shape, generator=g).float() / 16.0


def bf16(x):
    return x.to(torch.bfloat16).cuda()


@pytest.mark.parametrize("d", [128, 64])
@pytest.mark.parametrize("U,N", [(1, 1), (128, 128), (200, 1000), (129, 257), (300, 5000)])
def test_dense_scores_exact_on_grid(ops, d, U, N):
    users, items = grid((U, d), 1), grid((N, d), 2)
    ref = O.catalog_scores(users, items)
    out = ops.raw_catalog_scores_dense(bf16(users), bf16(items)).cpu()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("d", [128, 64])
def test_dense_scores_gaussian_tolerance(ops, d):
    torch.manual_seed(0)
    users, items = torch.randn(260, d), torch.randn(3000, d)
    ref = O.catalog_scores(users, items)          # fp64 accumulate, rounded once
    out = ops.raw_catalog_scores_dense(bf16(users), bf16(items)).cpu()
    bound = (users.to(torch.bfloat16).double().abs() @ items.to(torch.bfloat16).double().abs().t()).float()
    # stated tolerance: |err| <= 2^-8 * sum|u_i w_i|; fp32 accumulation is far inside it
    assert bool(((out - ref).abs() <= bound * 2.0 ** -8).all())
    assert float(((out - ref).abs() / bound).max()) < 1e-5


@pytest.mark.parametrize("d", [128, 64])
def test_target_scores_same_data_path(ops, d):
    """target score == the dense score of that column, bit for bit, also on Gaussian data."""
    torch.manual_seed(1)
    U, N, id_lo = 333, 2000, 1
    users, items = torch.randn(U, d), torch.randn(N, d)
    target = torch.randint(id_lo, id_lo + N, (U,))
    dense = ops.raw_catalog_scores_dense(bf16(users), bf16(items))
    ts = ops.raw_catalog_target_scores(bf16(users), bf16(items), target.cuda(), id_lo)
    assert torch.equal(ts.cpu(), dense.cpu()[torch.arange(U), target - id_lo])
    # targets outside the slice contribute exactly 0 (the all-reduce over shards adds them up)
    target2 = target.clone(); target2[::3] = id_lo + N + 5; target2[1::3] = 0
    ts2 = ops.raw_catalog_target_scores(bf16(users), bf16(items), target2.cuda(), id_lo).cpu()
    assert float(ts2[::3].abs().max()) == 0.0 and float(ts2[1::3].abs().max()) == 0.0
    assert torch.equal(ts2[2::3], ts.cpu()[2::3])


def _check_fused(ops, users, items, target, k, id_lo, scores):
    """fused kernel == oracle rank/top-k evaluated on `scores` (exact reference values)."""
    U, N = scores.shape
    tcol = target - id_lo
    ts = ops.raw_catalog_target_scores(bf16(users), bf16(items), target.cuda(), id_lo)
    val, idx, cnt = ops.raw_catalog_topk_rank(bf16(users), bf16(items), target.cuda(), ts, k, id_lo)
    ref_rank = O.catalog_rank(scores, tcol)
    rv, ri = O.catalog_topk(scores, min(k, N))
    assert torch.equal(cnt.cpu(), ref_rank)
    assert torch.equal(idx.cpu()[:, :min(k, N)], ri + id_lo)
    assert torch.equal(val.cpu()[:, :min(k, N)], rv)
    if N < k:
        assert bool((idx.cpu()[:, N:] == -1).all())
    # rank-only mode (k = 0: no maxima, no lists) must give the same counts
    _, _, cnt0 = ops.raw_catalog_topk_rank(bf16(users), bf16(items), target.cuda(), ts, 0, id_lo)
    assert torch.equal(cnt0, cnt)
    return cnt


@pytest.mark.parametrize("d", [128, 64])
@pytest.mark.parametrize("U,N,k", [(1, 1, 1), (100, 130, 10), (128, 1000, 10), (333, 4097, 30), (50, 20000, 5),
                                   (700, 2500, 32), (40, 7, 10)])
def test_fused_topk_rank_bit_exact_grid(ops, d, U, N, k):
    users, items = grid((U, d), 10 + U), grid((N, d), 20 + N)
    if N > 300:
        items[N // 2:N // 2 + 40] = items[3]                 # exact duplicate rows: tie-breaks by id
    id_lo = 1
    g = torch.Generator().manual_seed(U + N)
    target = torch.randint(id_lo, id_lo + N, (U,), generator=g)
    if N > 300:
        target[:20] = id_lo + N // 2 + torch.arange(20)      # targets inside the duplicate block
    scores = O.catalog_scores(users, items)
    _check_fused(ops, users, items, target, k, id_lo, scores)


@pytest.mark.parametrize("d", [128])
def test_fused_matches_own_dense_scores_gaussian(ops, d):
    torch.manual_seed(7)
    U, N, k, id_lo = 500, 6000, 10, 1
    users, items = torch.randn(U, d), torch.randn(N, d)
    target = torch.randint(id_lo, id_lo + N, (U,))
    dense = ops.raw_catalog_scores_dense(bf16(users), bf16(items)).cpu()
    _check_fused(ops, users, items, target, k, id_lo, dense)
    # ... and against the fp64 oracle, ranks agree except where a near-tie (gap below the fp32
    # accumulation noise) decides
    ref = O.catalog_scores(users, items)
    r_ref = O.catalog_rank(ref, target - id_lo)
    r_gpu = O.catalog_rank(dense, target - id_lo)
    assert float((r_ref != r_gpu).float().mean()) < 0.01


def test_fused_target_outside_slice_and_id_offset(ops):
    """Shard semantics: candidates are ids [id_lo, id_lo+N); the target may live on another shard."""
    d, U, N, k, id_lo = 128, 300, 3000, 10, 5001
    users, items = grid((U, d), 3), grid((N, d), 4)
    tscore = grid((U,), 5) * 8                       # as if reduced from the owner shard
    for target in (torch.full((U,), 17), torch.full((U,), id_lo + N + 100)):
        val, idx, cnt = ops.raw_catalog_topk_rank(bf16(users), bf16(items), target.cuda(), tscore.cuda(), k, id_lo)
        s = O.catalog_scores(users, items)
        lower = bool(target[0] < id_lo)
        # target has a LOWER id than every candidate -> only strictly greater scores outrank it;
        # target has a HIGHER id -> ties outrank it too
        ref = (s > tscore[:, None]).sum(1) if lower else (s >= tscore[:, None]).sum(1)
        assert torch.equal(cnt.cpu(), ref)
        rv, ri = O.catalog_topk(s, k)
        assert torch.equal(idx.cpu(), ri + id_lo) and torch.equal(val.cpu(), rv)


def test_fused_small_user_batch_uses_n_splits(ops):
    """Few users -> the catalog is split across CTAs and merged (same kernel as the shard merge)."""
    d, U, N, k, id_lo = 128, 130, 40000, 10, 1
    users, items = grid((U, d), 8), grid((N, d), 9)
    target = torch.randint(id_lo, id_lo + N, (U,), generator=torch.Generator().manual_seed(1))
    _check_fused(ops, users, items, target, k, id_lo, O.catalog_scores(users, items))


def test_fused_large_catalog_property(ops):
    """1M-item catalog (cfg4's N), 2048 users, grid data: exact agreement with a brute-force
    fp32 matmul on the device (exact arithmetic => order independent)."""
    d, U, N, k, id_lo = 128, 2048, 1_000_000, 10, 1
    users = grid((U, d), 31).cuda(); items = grid((N, d), 32).cuda()
    target = torch.randint(id_lo, id_lo + N, (U,), generator=torch.Generator().manual_seed(2)).cuda()
    ub, ib = users.to(torch.bfloat16), items.to(torch.bfloat16)
    ts = ops.raw_catalog_target_scores(ub, ib, target, id_lo)
    val, idx, cnt = ops.raw_catalog_topk_rank(ub, ib, target, ts, k, id_lo)
    torch.backends.cuda.matmul.allow_tf32 = False
    rank_ref = torch.empty(U, dtype=torch.int64, device="cuda")
    for u0 in range(0, U, 256):
        s = users[u0:u0 + 256] @ items.t()                                  # exact in fp32
        tc = (target[u0:u0 + 256] - id_lo)
        st = s.gather(1, tc[:, None])
        cols = torch.arange(N, device="cuda")[None, :]
        rank_ref[u0:u0 + 256] = (s > st).sum(1) + ((s == st) & (cols < tc[:, None])).sum(1)
        order = torch.sort(-s, dim=1, stable=True).indices[:, :k]
        assert torch.equal(idx[u0:u0 + 256], order + id_lo)
        assert torch.equal(val[u0:u0 + 256], s.gather(1, order))
        assert torch.equal(ts[u0:u0 + 256], st[:, 0])
    assert torch.equal(cnt, rank_ref)


# ---------------------------------------------------------------- K5 / H1 --------------------
def test_topk_merge_kernel(ops):
    torch.manual_seed(4)
    G, U, k = 4, 257, 10
    s = torch.randint(-5, 6, (U, G * 40)).float()                 # ties across shards
    vals, ids, cnts = [], [], []
    for g in range(G):
        v, i = O.catalog_topk(s[:, g * 40:(g + 1) * 40], k)
        vals.append(v); ids.append(i + g * 40); cnts.append(torch.full((U,), g + 1))
    v, i, r = ops.raw_topk_merge(torch.stack(vals).cuda(), torch.stack(ids).cuda(), torch.stack(cnts).cuda())
    rv, ri = O.catalog_topk(s, k)
    assert torch.equal(v.cpu(), rv) and torch.equal(i.cpu(), ri)
    assert torch.equal(r.cpu(), torch.full((U,), 1 + 2 + 3 + 4))
    # lists padded with -1 ids (shard smaller than k)
    vals[0][:, 5:] = float("-inf"); ids[0][:, 5:] = -1
    v, i, _ = ops.raw_topk_merge(torch.stack(vals).cuda(), torch.stack(ids).cuda(), torch.stack(cnts).cuda())
    assert bool((i.cpu() >= 0).all())


def test_rank_histogram_and_metrics(ops):
    from recguru_b200 import catalog
    g = load_golden("eval_rank_metrics.npz")
    rank = t(g["rank"]).cuda()
    ks = g["ks"].tolist()
    m = catalog.rank_metrics(rank, ks)
    for k in ks:
        assert abs(m[k]["hr"] - float(g[f"hr_{k}"])) < 1e-12
        assert abs(m[k]["ndcg"] - float(g[f"ndcg_{k}"])) < 1e-12
        assert abs(m[k]["mrr"] - float(g[f"mrr_{k}"])) < 1e-12
    assert abs(catalog.hit_at_k_batch([0, 3, 12], 10) - float(g["kat_hr_10"])) < 1e-12
    assert abs(catalog.NDCG_at_k_batch([0, 3, 12], 10) - float(g["kat_ndcg_10"])) < 1e-12
    assert abs(catalog.mrr_at_k_batch([0, 3, 12], 10) - float(g["kat_mrr_10"])) < 1e-12
    torch.manual_seed(0)
    big = torch.randint(0, 1000, (1_000_003,))
    h = ops.raw_rank_histogram(big.cuda(), 30).cpu()
    ref = torch.bincount(big.clamp(max=30), minlength=31)
    assert torch.equal(h, ref)
    mm = catalog.metrics_from_histogram(h, big.numel(), [10])
    assert abs(mm[10]["hr"] - O.hit_at_k_batch(big.numpy(), 10)) < 1e-12
    rb = big.numpy().astype(np.float64)
    assert abs(mm[10]["ndcg"] - float(np.where(rb < 10, 1.0 / np.log2(rb + 2), 0.0).mean())) < 1e-12
    assert abs(mm[10]["mrr"] - float(np.where(rb < 10, 1.0 / (rb + 1), 0.0).mean())) < 1e-12


def test_catalog_evaluator_single_gpu(ops):
    from recguru_b200 import catalog
    d, n_items, U, k = 128, 5000, 400, 10
    W = grid((n_items + 2, d), 77); W[0] = 0
    users = grid((U, d), 78)
    target = torch.randint(1, n_items + 1, (U,), generator=torch.Generator().manual_seed(3))
    ev = catalog.CatalogEvaluator(W.cuda(), k=k)
    val, idx, rank = ev.rank_topk(users.cuda(), target.cuda())
    s = O.catalog_scores(users, W[1:n_items + 1])
    assert torch.equal(rank.cpu(), O.catalog_rank(s, target - 1))
    rv, ri = O.catalog_topk(s, k)
    assert torch.equal(idx.cpu(), ri + 1) and torch.equal(val.cpu(), rv)
    rank2, met = ev.evaluate(users.cuda(), target.cuda(), ks=(5, 10))
    r = rank.cpu().tolist()
    assert abs(met[10]["hr"] - O.hit_at_k_batch(r, 10)) < 1e-12
    assert abs(met[10]["ndcg"] - O.NDCG_at_k_batch(r, 10)) < 1e-9
    assert abs(met[5]["mrr"] - O.mrr_at_k_batch(r, 5)) < 1e-9
    # the torch.library op
    v2, i2, r2 = torch.ops.recguru.catalog_topk_rank(users.cuda(), ev.items, target.cuda(), k, 1)
    assert torch.equal(r2, rank) and torch.equal(i2, idx)


def test_tie_modes_and_bracket(ops):
    """explicit tie rules for the target's rank.  Degenerate scores (few distinct values, an
    all-equal row) on grid data: target_first = #{s > t}, target_last = #{s >= t} - 1, id = the stable
    sort; the bracket's width is the number of items tied with the target."""
    from recguru_b200 import catalog
    d, n_items, U = 64, 3000, 260
    g = torch.Generator().manual_seed(5)
    W = torch.randint(-1, 2, (n_items + 2, d), generator=g).float()      # scores collide massively
    W[0] = 0
    users = torch.randint(-1, 2, (U, d), generator=g).float()
    users[7] = 0                                                         # an all-equal score row (all zeros)
    target = torch.randint(1, n_items + 1, (U,), generator=g)
    s = O.catalog_scores(users, W[1:n_items + 1])
    t = s.gather(1, (target - 1)[:, None])
    best_ref = (s > t).sum(1)
    worst_ref = (s >= t).sum(1) - 1
    id_ref = O.catalog_rank(s, target - 1)
    ev = catalog.CatalogEvaluator(W.cuda(), k=10)
    assert torch.equal(ev.target_rank(users.cuda(), target.cuda()).cpu(), id_ref)
    best, worst = ev.tie_bracket(users.cuda(), target.cuda())
    assert torch.equal(best.cpu(), best_ref) and torch.equal(worst.cpu(), worst_ref)
    assert int(best[7]) == 0 and int(worst[7]) == n_items - 1            # the degenerate row: rank 0 .. N-1
    assert bool(((best.cpu() <= id_ref) & (id_ref <= worst.cpu())).all())
    for tie, ref in (("target_first", best_ref), ("target_last", worst_ref)):
        e2 = catalog.CatalogEvaluator(W.cuda(), k=10, tie=tie)
        val, idx, rk = e2.rank_topk(users.cuda(), target.cuda())
        assert torch.equal(rk.cpu(), ref)
        rv, ri = O.catalog_topk(s, 10)
        assert torch.equal(idx.cpu(), ri + 1) and torch.equal(val.cpu(), rv)   # lists do not depend on the rule
