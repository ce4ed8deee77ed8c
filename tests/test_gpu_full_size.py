"""The training-step hot path at BASELINE.json's full sizes.

cfg2 (configs[1]: 100k items, L=200, d=128, B=1024, K=5, BPRLoss_sas) on the bench's own
Amazon-shaped synthetic batch, against the oracle run through torch autograd on the CPU in fp32
(the reference's arithmetic) and in fp64 (the yardstick): K1 and the mask bit-exact, loss rtol 1e-5,
dh rtol 1e-4, and the table gradient at least as close to the fp64 result as the reference's own
fp32 arithmetic is (hot Zipf items sum thousands of terms per row, so a fixed rtol against the fp32
oracle would only measure ITS rounding).  Plus the size-independent properties: run-to-run and
graph-vs-eager bit identity, untouched rows exactly zero, column checksums.

cfg3-shaped (configs[2]: two 500k-item domains, L=100, d=128, K=30, sampled cross-entropy, tables
without padding_idx) at B=256 per domain against the same oracle."""
import pytest
import torch

from oracle import ref_path as O
from test_gpu_hotpath import oracle_step

pytestmark = pytest.mark.gpu


def _cfg2_inputs():
    from recguru_b200 import synth
    n_items, L, d, B, K = 100_000, 200, 128, 1024, 5
    W = synth.item_table(n_items, d)
    pe = O.pe_table(256, d)
    tb = synth.train_batch(B, L, n_items, K)
    h = synth.hidden_states(B, L, d, seed=50)
    g = synth.hidden_states(B, L, d, seed=90) * 0.01
    return W, pe, tb, h, g, (n_items, L, d, B, K)


def test_cfg2_step_vs_oracle_fp32_and_fp64(built_lib):
    from recguru_b200.hotpath import TrainHotPath
    W, pe, tb, h, g, (n_items, L, d, B, K) = _cfg2_inputs()
    x0_ref, mask_ref, loss_ref, dh_ref, dW_ref = oracle_step(W, pe, tb.dec_in, tb.dec_out, tb.n_items, h, g, K, "bpr_sas")
    *_, dW_64 = oracle_step(W.double(), pe.double(), tb.dec_in, tb.dec_out, tb.n_items, h.double(), g.double(), K,
                            "bpr_sas")
    hp = TrainHotPath(W.cuda(), pe.cuda(), B, L, K, "bpr_sas")
    c = lambda t: t.cuda()
    loss, dh, dW = hp.step(c(tb.dec_in), c(tb.dec_out), c(tb.n_items), c(h), c(g))
    torch.cuda.synchronize()
    assert torch.equal(hp.x0.cpu(), x0_ref), "K1 not bit-exact at cfg2"
    assert torch.equal(hp.mask.cpu(), mask_ref)
    assert torch.allclose(loss.cpu(), loss_ref, rtol=1e-5, atol=0), (float(loss), float(loss_ref))
    assert torch.allclose(dh.cpu(), dh_ref, rtol=1e-4, atol=1e-9)
    dWc = dW.cpu()
    err_ours = (dWc.double() - dW_64).abs().max().item()
    err_ref32 = (dW_ref.double() - dW_64).abs().max().item()        # ~4e-5 here: sequential fp32 sums of hot rows
    assert err_ours <= 2.0 * err_ref32 + 1e-10, (err_ours, err_ref32)
    assert torch.allclose(dWc, dW_ref, rtol=1e-3, atol=4 * err_ref32 + 1e-6)
    # rows no live lookup touches are exactly zero; so is the pad row
    touched = torch.zeros(n_items + 2, dtype=torch.bool)
    live = tb.dec_in.view(-1) != 0
    touched[tb.dec_in.view(-1)[live]] = True
    touched[tb.dec_out.view(-1)[live]] = True
    touched[tb.n_items.view(B * L, K)[live].reshape(-1)] = True
    touched[0] = False
    assert not bool(dWc[~touched].any())
    # column checksum (sum over all rows, accumulated in fp64) against the reference's fp32 arithmetic.
    # Not against the fp64 run: its COEFFICIENTS differ (in fp32 the loss's sigmoids saturate and
    # 1 - sigma + 1e-24 clamps, lossfunctions.py:90-93), which moves the checksums by ~5e-4 for
    # both fp32 computations alike.
    assert torch.allclose(dWc.double().sum(0), dW_ref.double().sum(0), rtol=1e-5, atol=2e-5)


def test_cfg2_bit_identity_across_runs_schedules_and_graph(built_lib):
    from recguru_b200.hotpath import TrainHotPath
    W, pe, tb, h, g, (n_items, L, d, B, K) = _cfg2_inputs()
    ins = [t.cuda() for t in (tb.dec_in, tb.dec_out, tb.n_items, h, g)]
    hp = TrainHotPath(W.cuda(), pe.cuda(), B, L, K, "bpr_sas")
    loss, dh, dW = hp.step(*ins)
    first = (loss.clone(), dh.clone(), dW.clone())
    hp.dW.fill_(3.0)                                      # stale contents must not leak into the next step
    again = hp.step(*ins)
    assert all(torch.equal(a, b) for a, b in zip(first, again))
    serial = TrainHotPath(W.cuda(), pe.cuda(), B, L, K, "bpr_sas", overlap_plan=False).step(*ins)
    assert all(torch.equal(a, b) for a, b in zip(first, serial))
    graph = hp.capture_step(*ins)
    hp.dW.fill_(5.0)
    hp.dh.fill_(5.0)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(hp.sbuf["aux"][0], first[0]) and torch.equal(hp.dh, first[1]) and torch.equal(hp.dW, first[2])
    # linearity in the upstream gradient of the input site: doubling g_x0 (a power of two: exact in
    # fp32) changes dW by exactly the input site's share
    ins2 = ins[:4] + [ins[4] * 2]
    _, _, dW2 = hp.step(*ins2)
    dW2 = dW2.clone()
    zero = ins[:4] + [torch.zeros_like(ins[4])]
    _, _, dW0 = hp.step(*zero)
    assert torch.allclose(dW2 - dW0, 2 * (first[2] - dW0), rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("domain_seed", [0, 1])
def test_cfg3_shaped_domain_step_vs_oracle(built_lib, domain_seed):
    from recguru_b200 import synth
    from recguru_b200.hotpath import TrainHotPath
    n_items, L, d, B, K = 500_000, 100, 128, 256, 30
    W = synth.item_table(n_items, d, seed=synth.SEED_WEIGHTS + domain_seed)
    W[0].normal_()                                       # cross-domain tables have no padding_idx: row 0 is live data
    pe = O.pe_table(128, d)
    tb = synth.train_batch(B, L, n_items, K, seed=synth.SEED_DATA + 7 * domain_seed + 1)
    h = synth.hidden_states(B, L, d, seed=60 + domain_seed)
    g = synth.hidden_states(B, L, d, seed=70 + domain_seed) * 0.01
    x0_ref, mask_ref, loss_ref, dh_ref, _ = oracle_step(W, pe, tb.dec_in, tb.dec_out, tb.n_items, h, g, K, "sce")
    hp = TrainHotPath(W.cuda(), pe.cuda(), B, L, K, "sce", padding_idx=None)
    c = lambda t: t.cuda()
    loss, dh, dW = hp.step(c(tb.dec_in), c(tb.dec_out), c(tb.n_items), c(h), c(g))
    torch.cuda.synchronize()
    assert torch.equal(hp.x0.cpu(), x0_ref) and torch.equal(hp.mask.cpu(), mask_ref)
    assert torch.allclose(loss.cpu(), loss_ref, rtol=1e-5, atol=0), (float(loss), float(loss_ref))
    assert torch.allclose(dh.cpu(), dh_ref, rtol=1e-4, atol=1e-9)
    # the table gradient through ATen's own backward, without the pad-row zeroing oracle_step applies
    Wp = W.clone().requires_grad_(True)
    x0, mask = O.gather_pe_mask(Wp, tb.dec_in, pe, 0)
    p, n = O.sampled_scores(h, Wp, tb.dec_out, tb.n_items, L, K)
    (O.LOSSES["sce"](p, n, mask.view(-1)) + (x0 * g).sum()).backward()
    assert torch.allclose(dW.cpu(), Wp.grad, rtol=1e-3, atol=5e-5)


# ------------------------------------------------------------------------------------------------
# catalog scoring at the shapes bench.py TIMES: cfg4's 75 776 users x 1M items
# takes the single-split path (7 813 tiles per CTA) that no smaller test reaches, and nothing
# touched a 10M-item catalog.  Grid-valued data (k/16): every bf16 x bf16 -> fp32 score is exact in
# any summation order, so ranks / top-k ids / values must be bit-identical to fp32 torch arithmetic.
# ------------------------------------------------------------------------------------------------
def _grid_dev(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(-32, 33, shape, generator=g, device="cuda").float() / 16.0


def _reference_rank(users, items, target_local, u_chunk, n_chunk):
    """count-greater rank with ties -> lower id, chunked fp32 matmuls (exact on grid data)."""
    U, N = users.shape[0], items.shape[0]
    rank = torch.zeros(U, dtype=torch.int64, device="cuda")
    tsc = (users * items[target_local]).sum(1)                       # exact: same products, any order
    for u0 in range(0, U, u_chunk):
        u1 = min(U, u0 + u_chunk)
        t_ = tsc[u0:u1, None]
        tl = target_local[u0:u1, None]
        for n0 in range(0, N, n_chunk):
            n1 = min(N, n0 + n_chunk)
            s = users[u0:u1] @ items[n0:n1].t()
            cols = torch.arange(n0, n1, device="cuda")[None, :]
            rank[u0:u1] += (s > t_).sum(1) + ((s == t_) & (cols < tl)).sum(1)
    return rank, tsc


def _check_topk_exact(users, items, val, idx, id_lo, k, n_chunk):
    """top-k (score desc, id asc) without sorting a [U, N] matrix: the k best VALUES are unique
    as a multiset (merge of per-chunk torch.topk values); every returned id must carry its
    returned value; ties must be in ascending id order; and at the boundary value the returned ids
    must be the smallest ids that carry it."""
    U, N = users.shape[0], items.shape[0]
    best = torch.full((U, k), float("-inf"), device="cuda")
    for n0 in range(0, N, n_chunk):
        s = users @ items[n0:min(N, n0 + n_chunk)].t()
        best = torch.cat([best, s.topk(min(k, s.shape[1]), dim=1).values], 1).topk(k, dim=1).values
    assert torch.equal(val, best), "top-k values differ"
    loc = idx - id_lo
    assert bool(((loc >= 0) & (loc < N)).all())
    got = (users[:, None, :] * items[loc]).sum(2)                      # [U, k] exact
    assert torch.equal(got, val), "a returned id does not carry its returned value"
    same = val[:, 1:] == val[:, :-1]
    assert bool((~same | (idx[:, 1:] > idx[:, :-1])).all()), "ties not in ascending id order"
    vk = val[:, -1:]                                                   # boundary value
    at_b = val == vk
    m = at_b.sum(1)
    max_id = torch.where(at_b, loc, torch.full_like(loc, -1)).max(1).values
    cnt = torch.zeros(U, dtype=torch.int64, device="cuda")
    for n0 in range(0, N, n_chunk):
        n1 = min(N, n0 + n_chunk)
        s = users @ items[n0:n1].t()
        cols = torch.arange(n0, n1, device="cuda")[None, :]
        cnt += ((s == vk) & (cols <= max_id[:, None])).sum(1)
    assert torch.equal(cnt, m), "boundary ties are not the lowest ids"


def test_catalog_cfg4_timed_shape_single_split(built_lib):
    """75 776 users (= bench.USERS_PER_STEP) x 1M items, d=128, top-10: n_splits == 1."""
    from recguru_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    U, N, d, k, id_lo = 148 * 4 * 128, 1_000_000, 128, 10, 1
    users, items = _grid_dev((U, d), 41), _grid_dev((N, d), 42)
    items[123_456:123_556] = items[7]                                  # 100 exact duplicates of item 8
    target = torch.randint(id_lo, id_lo + N, (U,), generator=torch.Generator(device="cuda").manual_seed(3), device="cuda")
    target[:64] = id_lo + 123_500                                      # targets inside the tie block
    target[64:128] = id_lo + N - 1                                     # ... and in the ragged last tile
    ub, ib = ops.raw_cast_bf16(users), ops.raw_cast_bf16(items)
    ts = ops.raw_catalog_target_scores(ub, ib, target, id_lo)
    val, idx, cnt = ops.raw_catalog_topk_rank(ub, ib, target, ts, k, id_lo)
    rank_only = ops.raw_catalog_topk_rank(ub, ib, target, ts, 0, id_lo)[2]
    ref_rank, ref_ts = _reference_rank(users, items, target - id_lo, 4096, 250_000)
    assert torch.equal(ts, ref_ts)
    assert torch.equal(cnt, ref_rank), f"{int((cnt != ref_rank).sum())} ranks differ"
    assert torch.equal(rank_only, ref_rank)
    sub = torch.cat([torch.arange(0, 256), torch.arange(U - 256, U), torch.arange(30_000, 30_512)]).cuda()
    _check_topk_exact(users[sub], items, val[sub], idx[sub], id_lo, k, 250_000)
    # size-independent checksum of the whole result (what bench.py prints per N)
    assert int(idx.min()) >= id_lo and int(idx.max()) < id_lo + N


def test_catalog_10m_items(built_lib):
    """2 048 users x 10M items (cfg5's catalog on one GPU; multi-split + merge at this user count)
    and 12 800 users x 10M in rank-only mode."""
    from recguru_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    U, N, d, k, id_lo = 2048, 10_000_000, 128, 10, 1
    users, items = _grid_dev((U, d), 51), _grid_dev((N, d), 52)
    items[9_999_900:] = items[5]                                       # ties in the last tiles
    target = torch.randint(id_lo, id_lo + N, (U,), generator=torch.Generator(device="cuda").manual_seed(4), device="cuda")
    target[:32] = id_lo + N - 3
    ub, ib = ops.raw_cast_bf16(users), ops.raw_cast_bf16(items)
    ts = ops.raw_catalog_target_scores(ub, ib, target, id_lo)
    val, idx, cnt = ops.raw_catalog_topk_rank(ub, ib, target, ts, k, id_lo)
    ref_rank, ref_ts = _reference_rank(users, items, target - id_lo, 2048, 500_000)
    assert torch.equal(ts, ref_ts) and torch.equal(cnt, ref_rank)
    _check_topk_exact(users[:512], items, val[:512], idx[:512], id_lo, k, 500_000)
    U2 = 12_800
    users2 = _grid_dev((U2, d), 53)
    target2 = torch.randint(id_lo, id_lo + N, (U2,), generator=torch.Generator(device="cuda").manual_seed(5), device="cuda")
    ub2 = ops.raw_cast_bf16(users2)
    ts2 = ops.raw_catalog_target_scores(ub2, ib, target2, id_lo)
    r2 = ops.raw_catalog_topk_rank(ub2, ib, target2, ts2, 0, id_lo)[2]
    ref2, _ = _reference_rank(users2, items, target2 - id_lo, 3200, 500_000)
    assert torch.equal(r2, ref2)
