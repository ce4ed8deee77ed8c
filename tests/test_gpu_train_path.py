"""GPU parity: pieces 1-3 (gather+PE+mask, deterministic table backward, sampled scoring)
against the oracle and the reference-generated golden fixtures.  All calls go through the
C ABI (recguru_b200.ops raw layer -> ctypes -> librecguru_b200.so)."""
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


def dev(x):
    return x.cuda()


# ---------------------------------------------------------------- piece 1 -------------------
@pytest.mark.parametrize("d", [32, 64, 128, 256, 48])
@pytest.mark.parametrize("B,L", [(1, 1), (3, 7), (64, 50), (17, 200)])
def test_gather_pe_mask_bit_exact(ops, d, B, L):
    torch.manual_seed(d * 1000 + B * 10 + L)
    V1 = 503
    W = torch.randn(V1, d); W[0] = 0
    ids = torch.randint(0, V1, (B, L))
    ids[torch.rand(B, L) < 0.4] = 0
    pe = O.pe_table(max(L, 8), d)
    ref, ref_mask = O.gather_pe_mask(W, ids, pe, 0)
    out, mask = ops.raw_gather_pe_mask(dev(W), dev(ids), dev(pe), 0)
    assert torch.equal(out.cpu(), ref)                      # values
    assert torch.equal(out.cpu().view(torch.int32), ref.view(torch.int32))   # bits (incl. -0.0)
    assert torch.equal(mask.cpu(), ref_mask)


def test_gather_golden_all_models(ops):
    g = load_golden("myrec_sas.npz")
    out, mask = ops.raw_gather_pe_mask(dev(t(g["weight"])), dev(t(g["dec_in"])), dev(t(g["pe"])), 0)
    assert torch.equal(out.cpu(), t(g["x0"])) and torch.equal(mask.cpu().view(-1), t(g["mask"]))
    r = load_golden("myrec_recon.npz")
    for ids, x in (("enc_in", "x_enc"), ("dec_in", "x_dec")):
        out, _ = ops.raw_gather_pe_mask(dev(t(r["weight"])), dev(t(r[ids])), dev(t(r["pe"])), 0)
        assert torch.equal(out.cpu(), t(r[x]))
    c = load_golden("cross_a.npz")
    L = int(c["L"])
    out, _ = ops.raw_gather_pe_mask(dev(t(c["weight"])), dev(t(c["enc_in"])), dev(t(c["pe"])), -1,
                                    mask_in=dev(t(c["enc_mask"]).view(-1, L)))
    assert torch.equal(out.cpu(), t(c["x_enc"]))


def test_plain_gather_and_pe_mask_ops(ops):
    torch.manual_seed(1)
    W = torch.randn(1000, 64)
    ids = torch.randint(0, 1000, (9, 33))
    out, _ = ops.raw_gather_pe_mask(dev(W), dev(ids), None, -1, want_mask=False)
    assert torch.equal(out.cpu(), W[ids])
    x = torch.randn(9, 33, 64)
    mask = (torch.rand(9, 33) > 0.3).float()
    pe = O.pe_table(40, 64)
    ref = (x + pe[:, :33]) * mask.unsqueeze(2)
    assert torch.equal(ops.raw_pe_mask(dev(x), dev(pe), dev(mask)).cpu(), ref)


def test_gather_empty_and_errors(ops):
    W = torch.randn(10, 32).cuda()
    out, _ = ops.raw_gather_pe_mask(W, torch.zeros(0, 5, dtype=torch.int64).cuda(), None, -1, want_mask=False)
    assert out.shape == (0, 5, 32)
    with pytest.raises(RuntimeError):
        ops.raw_gather_pe_mask(W.cpu(), torch.zeros(2, 2, dtype=torch.int64), None, -1)      # CPU tensor
    with pytest.raises(RuntimeError):
        ops.raw_gather_pe_mask(W, torch.zeros(2, 2, dtype=torch.int32).cuda(), None, -1)     # wrong dtype
    from recguru_b200._lib import RecGuruError
    with pytest.raises(RecGuruError):
        ops.raw_gather_pe_mask(torch.randn(10, 30).cuda(), torch.zeros(2, 2, dtype=torch.int64).cuda(), None, -1)


# ---------------------------------------------------------------- piece 2 -------------------
@pytest.mark.parametrize("n,V1", [(1, 2), (31, 5), (4096, 70), (4097, 100003), (100000, 511),
                                  (300001, 1 << 20), (50000, 10_000_002)])
def test_sort_ids_bit_exact(ops, n, V1):
    torch.manual_seed(n)
    ids = torch.randint(0, V1, (n,))
    if n > 100:
        ids[torch.rand(n) < 0.3] = int(ids[0])      # heavy duplicates: stability matters
    s_ref, p_ref = O.stable_sort_ids(ids)
    s, p = ops.raw_sort_ids(dev(ids), V1)
    assert torch.equal(s.cpu(), s_ref)
    assert torch.equal(p.cpu(), p_ref)


def _zipf(n, V1, seed):
    from recguru_b200 import synth
    return synth.zipf_ids(n, V1 - 2, seed=seed)


@pytest.mark.parametrize("d", [32, 64, 128, 256])
@pytest.mark.parametrize("n", [5, 64, 65, 1000, 30000])
def test_embedding_bwd_single_site(ops, d, n):
    torch.manual_seed(n + d)
    V1 = 997
    ids = _zipf(n, V1, seed=n)
    ids[torch.rand(n) < 0.3] = 0
    g = torch.randn(n, d)
    ref = O.embedding_backward(ids, g, V1, padding_idx=0)
    out = ops.raw_embedding_bwd([ops.Site(dev(ids), dev(g))], V1, d, padding_idx=0)
    # fp32 summation-order tolerance: a few ulps of the row's sum of magnitudes (the Zipf head
    # item owns thousands of occurrences); plus an aggregate relative-norm bound
    mag = O.embedding_backward(ids, g.abs(), V1, padding_idx=0)
    assert bool(((out.cpu() - ref).abs() <= 2e-6 * mag + 1e-6).all())
    assert float((out.cpu() - ref).norm() / ref.norm()) < 1e-6
    assert float(out[0].abs().max()) == 0.0                       # padding row skipped
    untouched = torch.ones(V1, dtype=torch.bool); untouched[ids] = False
    assert float(out.cpu()[untouched].abs().max()) == 0.0         # dense zero fill
    # no padding_idx: row 0 accumulates like any other
    ref2 = O.embedding_backward(ids, g, V1, padding_idx=None)
    out2 = ops.raw_embedding_bwd([ops.Site(dev(ids), dev(g))], V1, d, padding_idx=-1)
    mag2 = O.embedding_backward(ids, g.abs(), V1, padding_idx=None)
    assert bool(((out2.cpu() - ref2).abs() <= 2e-6 * mag2 + 1e-6).all())


def test_embedding_bwd_hot_item_and_determinism(ops):
    """One id owns 60 % of the occurrences (runs spanning hundreds of chunks); two runs must
    agree bit for bit (no floating-point atomics)."""
    torch.manual_seed(5)
    n, V1, d = 200_000, 5003, 128
    ids = torch.randint(1, V1, (n,))
    ids[torch.rand(n) < 0.6] = 77
    ids[:100] = 1          # run at the very start
    ids[-300:] = V1 - 1    # run at the very end
    g = torch.randn(n, d)
    coef = torch.rand(n)
    a = ops.raw_embedding_bwd([ops.Site(dev(ids), dev(g), dev(coef))], V1, d, padding_idx=0)
    b = ops.raw_embedding_bwd([ops.Site(dev(ids), dev(g), dev(coef))], V1, d, padding_idx=0)
    assert torch.equal(a, b)
    ref = O.embedding_backward_sites([(ids, g, coef, 1)], V1, d, 0).float()
    assert torch.allclose(a.cpu(), ref, rtol=2e-5, atol=2e-3)
    rel = (a.cpu() - ref).norm() / ref.norm()
    assert rel < 1e-6


def test_embedding_bwd_multi_site_golden(ops):
    """All lookup sites of a reference training step in ONE launch == the reference's W.grad."""
    g = load_golden("myrec_sas.npz")
    W = t(g["weight"]); V1, d = W.shape; K = int(g["K"])
    h = dev(t(g["h"]).reshape(-1, d))
    sites = [ops.Site(dev(t(g["dec_in"])), dev(t(g["g_in"]).reshape(-1, d))),
             ops.Site(dev(t(g["dec_out"])), h, dev(t(g["dp"]).reshape(-1))),
             ops.Site(dev(t(g["n_items"])), h, dev(t(g["dn"]).reshape(-1)), group=K)]
    dW = ops.raw_embedding_bwd(sites, V1, d, padding_idx=0)
    assert torch.allclose(dW.cpu(), t(g["dW"]), rtol=1e-5, atol=1e-6)

    c = load_golden("cross_a.npz")                       # no padding_idx
    W = t(c["weight"]); V1, d = W.shape; K = int(c["K"])
    h = t(c["h"]).reshape(-1, d)
    p = t(c["p_logits"]).clone().requires_grad_(True); n = t(c["n_logits"]).clone().requires_grad_(True)
    O.bpr_loss(p, n, t(c["dec_mask"])).backward()
    sites = [ops.Site(dev(t(c["enc_in"])), dev(t(c["g_enc"]).reshape(-1, d))),
             ops.Site(dev(t(c["dec_in"])), dev(t(c["g_dec"]).reshape(-1, d))),
             ops.Site(dev(t(c["dec_out"])), dev(h), dev(p.grad.reshape(-1))),
             ops.Site(dev(t(c["n_items"])), dev(h), dev(n.grad.reshape(-1)), group=K)]
    dW = ops.raw_embedding_bwd(sites, V1, d, padding_idx=-1)
    assert torch.allclose(dW.cpu(), t(c["dW"]), rtol=1e-5, atol=1e-6)


def test_embedding_bwd_accumulate_mode(ops):
    torch.manual_seed(9)
    V1, d, n = 300, 64, 5000
    ids = torch.randint(0, V1, (n,)); g = torch.randn(n, d)
    base = torch.randn(V1, d)
    out = dev(base.clone())
    ops.raw_embedding_bwd([ops.Site(dev(ids), dev(g))], V1, d, padding_idx=-1, out=out, zero_fill=False)
    ref = base + O.embedding_backward(ids, g, V1, None)
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-4)


# ---------------------------------------------------------------- piece 3 -------------------
@pytest.mark.parametrize("kind", ["bpr", "bpr_sas", "sce"])
@pytest.mark.parametrize("d,K", [(32, 5), (64, 5), (128, 5), (128, 30), (64, 30), (256, 3), (32, 1), (128, 31)])
def test_sampled_score_loss_vs_oracle(ops, kind, d, K):
    torch.manual_seed(d + K)
    B, L, V1 = 7, 13, 401
    W = torch.randn(V1, d) * 0.3; W[0] = 0
    h = torch.randn(B, L, d) * 0.3
    pos = torch.randint(0, V1, (B, L)); neg = torch.randint(1, V1, (B, L * K))
    mask = (pos != 0).float().view(-1)
    p_ref, n_ref = O.sampled_scores(h, W, pos, neg, L, K)
    pr = p_ref.clone().requires_grad_(True); nr = n_ref.clone().requires_grad_(True)
    loss_ref = O.LOSSES[kind](pr, nr, mask)
    loss_ref.backward()
    p, n, aux, dp, dn = ops.raw_sampled_score_loss(dev(h), dev(W), dev(pos), dev(neg), dev(mask), kind)
    assert torch.allclose(p.cpu(), p_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(n.cpu(), n_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(aux[0].cpu(), loss_ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(aux[1].cpu(), 1.0 / mask.sum(), rtol=1e-6)
    assert torch.allclose((dp * aux[1]).cpu(), pr.grad.view(-1), rtol=1e-4, atol=1e-7)
    assert torch.allclose((dn * aux[1]).cpu(), nr.grad.view(-1, K), rtol=1e-4, atol=1e-7)
    # loss-only entry point on the oracle's logits
    aux2, dp2, dn2 = ops.raw_sampled_loss(dev(p_ref), dev(n_ref), dev(mask), kind)
    assert torch.allclose(aux2[0].cpu(), loss_ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose((dp2 * aux2[1]).cpu(), pr.grad.view(-1), rtol=1e-4, atol=1e-7)
    assert torch.allclose((dn2 * aux2[1]).cpu(), nr.grad.view(-1, K), rtol=1e-4, atol=1e-7)
    # unmasked mean
    _, _, aux3, _, _ = ops.raw_sampled_score_loss(dev(h), dev(W), dev(pos), dev(neg), None, kind)
    assert torch.allclose(aux3[0].cpu(), O.LOSSES[kind](p_ref, n_ref, None), rtol=1e-5, atol=1e-6)
    # padding-aware variant (RECGURU_LOSS_SKIP_MASKED): derivatives bit-identical, loss to 1e-6,
    # logits identical where mask != 0 and reported as 0 under the mask
    p4, n4, aux4, dp4, dn4 = ops.raw_sampled_score_loss(dev(h), dev(W), dev(pos), dev(neg), dev(mask), kind,
                                                        skip_masked=True)
    live = dev(mask) != 0
    # (the padding-aware kernel reduces the loss over a different CTA partition: same terms,
    # different fp32 summation order)
    assert torch.allclose(aux4, aux, rtol=1e-6, atol=0) and torch.equal(dp4, dp) and torch.equal(dn4, dn)
    assert torch.equal(p4.view(-1)[live], p.view(-1)[live]) and torch.equal(n4.view(-1, K)[live], n.view(-1, K)[live])
    assert float(p4.view(-1)[~live].abs().sum()) == 0.0 and float(n4.view(-1, K)[~live].abs().sum()) == 0.0


def test_sampled_golden_losses(ops):
    for name, kind, mkey in (("myrec_sas.npz", "bpr_sas", "mask"), ("cross_a.npz", "bpr", "dec_mask")):
        g = load_golden(name)
        p, n, aux, dp, dn = ops.raw_sampled_score_loss(dev(t(g["h"])), dev(t(g["weight"])), dev(t(g["dec_out"])),
                                                       dev(t(g["n_items"])), dev(t(g[mkey])), kind)
        assert torch.allclose(p.cpu(), t(g["p_logits"]), rtol=1e-5, atol=1e-5)
        assert torch.allclose(n.cpu(), t(g["n_logits"]), rtol=1e-5, atol=1e-5)
        assert torch.allclose(aux[0].cpu(), t(g["loss"]), rtol=1e-5)
    r = load_golden("myrec_recon.npz")
    K = int(r["K"])
    p, n, aux, dp, dn = ops.raw_sampled_score_loss(dev(t(r["h"])), dev(t(r["weight"])), dev(t(r["dec_out"])),
                                                   dev(t(r["n_items"])), dev(t(r["mask"])), "sce")
    assert torch.allclose(torch.cat([p, n], 2).cpu(), t(r["logits"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(aux[0].cpu(), t(r["loss"]), rtol=1e-5)
    dl = t(r["dlogits"]).reshape(-1, K + 1)
    assert torch.allclose((dp * aux[1]).cpu(), dl[:, 0], rtol=1e-4, atol=1e-7)
    assert torch.allclose((dn * aux[1]).cpu(), dl[:, 1:], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("d,K", [(32, 5), (64, 30), (128, 5), (256, 2)])
def test_sampled_score_bwd_dh(ops, d, K):
    torch.manual_seed(3 * d + K)
    B, L, V1 = 5, 11, 211
    W = torch.randn(V1, d); h = torch.randn(B, L, d, requires_grad=True)
    pos = torch.randint(0, V1, (B, L)); neg = torch.randint(0, V1, (B, L * K))
    p, n = O.sampled_scores(h, W, pos, neg, L, K)
    gp = torch.randn_like(p); gn = torch.randn_like(n)
    (p * gp).sum().add((n * gn).sum()).backward()
    scale = torch.tensor([0.5]); scale2 = torch.tensor([4.0])
    dh = ops.raw_sampled_score_bwd_dh(dev(gp.reshape(-1)), dev(gn.reshape(-1, K)), dev(W), dev(pos), dev(neg),
                                      dev(scale), dev(scale2))
    assert torch.allclose(dh.cpu(), 2.0 * h.grad, rtol=1e-5, atol=1e-5)


def test_pad_mask_count(ops):
    g = torch.Generator().manual_seed(2)
    count3 = torch.zeros(3, dtype=torch.int32, device="cuda")
    for n, frac in [(1, 0.0), (7, 0.5), (4096, 0.9), (204800, 0.926), (1_000_003, 0.1)]:
        ids = torch.randint(1, 1000, (n,), generator=g)
        ids[torch.rand(n, generator=g) < frac] = 0
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        for _ in range(2):                                 # self-clearing scratch: a second call gives the same count
            ops.raw_pad_mask_count(dev(ids), 0, out, count3)
            assert count3.tolist() == [0, 0, int((ids != 0).sum())]
        assert torch.equal(out.cpu(), (ids != 0).float())


@pytest.mark.parametrize("kind", ["bpr", "bpr_sas", "sce"])
@pytest.mark.parametrize("d,K", [(32, 5), (64, 30), (128, 5), (128, 30), (256, 3), (64, 1)])
def test_fused_forward_dh_bit_identical_to_two_kernels(ops, kind, d, K):
    """recguru_sampled_score_loss_fwd_dh writes the same dh bits as the forward followed by
    recguru_sampled_score_bwd_dh, and leaves loss / logits / derivatives untouched; several chunks
    per CTA, dead positions in between, a non-unit upstream scale."""
    torch.manual_seed(5 * d + K)
    B, L, V1 = 37, 61, 977
    W = dev(torch.randn(V1, d) * 0.3)
    h = dev(torch.randn(B, L, d) * 0.3)
    pos = dev(torch.randint(1, V1, (B, L)))
    neg = dev(torch.randint(1, V1, (B, L * K)))
    ids = torch.randint(1, V1, (B, L))
    ids[torch.rand(B, L) < 0.7] = 0
    mask = torch.empty(B, L, device="cuda")
    count3 = torch.zeros(3, dtype=torch.int32, device="cuda")
    ops.raw_pad_mask_count(dev(ids), 0, mask, count3)
    go = dev(torch.tensor([0.75]))
    m = mask.view(-1)
    p1, n1, aux1, dp1, dn1 = ops.raw_sampled_score_loss(h, W, pos, neg, m, kind, skip_masked=True)
    dh_ref = ops.raw_sampled_score_bwd_dh(dp1, dn1, W, pos, neg, go, aux1[1:2])
    dh = torch.full((B * L, d), float("nan"), device="cuda")
    p2, n2, aux2, dp2, dn2 = ops.raw_sampled_score_loss(h, W, pos, neg, m, kind, skip_masked=True, dh_out=dh,
                                                        live_count=count3, grad_scale=go)
    assert torch.equal(aux1, aux2) and torch.equal(p1, p2) and torch.equal(n1, n2)
    assert torch.equal(dp1, dp2) and torch.equal(dn1, dn2)
    assert torch.equal(dh.view_as(dh_ref), dh_ref)
    assert not bool(dh.view(B * L, d)[m == 0].any())
    # no upstream scale (NULL) == scale 1
    dh1 = torch.empty(B * L, d, device="cuda")
    ops.raw_sampled_score_loss(h, W, pos, neg, m, kind, skip_masked=True, dh_out=dh1, live_count=count3)
    one = dev(torch.ones(1))
    assert torch.equal(dh1.view_as(dh_ref), ops.raw_sampled_score_bwd_dh(dp1, dn1, W, pos, neg, one, aux1[1:2]))
    # the fused form needs the padding-aware kernel
    with pytest.raises(RuntimeError):
        ops.raw_sampled_score_loss(h, W, pos, neg, m, kind, skip_masked=False, dh_out=dh1, live_count=count3)


@pytest.mark.parametrize("kind", ["bpr", "bpr_sas", "sce"])
@pytest.mark.parametrize("d,K,frac", [(32, 5, 0.7), (64, 5, 0.93), (64, 30, 0.5), (128, 8, 0.2), (256, 3, 0.99),
                                      (64, 1, 0.0), (128, 31, 1.0)])
def test_score_over_live_row_list_bit_identical(ops, kind, d, K, frac):
    """recguru_sampled_score_loss_fwd_dh_rows (one warp per live position, the list coming from the row-mode
    compaction of the table-backward plan) == the compacting kernel: logits, derivatives and dh bit for bit,
    the loss up to the order of its fp32 partial sums.  frac = share of padded positions (0 and 1 included)."""
    torch.manual_seed(7 * d + K)
    B, L, V1 = 41, 53, 1201
    n = B * L
    W = dev(torch.randn(V1, d) * 0.3)
    h = dev(torch.randn(B, L, d) * 0.3)
    pos = dev(torch.randint(1, V1, (B, L)))
    neg = dev(torch.randint(1, V1, (B, L * K)))
    ids = torch.randint(1, V1, (B, L))
    ids[torch.rand(B, L) < frac] = 0
    if frac >= 1.0:
        ids[:] = 0
        ids[3, 5] = 9                                     # one live position (a loss over none is 0/0)
    idc = dev(ids)
    mask = torch.empty(B, L, device="cuda")
    count3 = torch.zeros(3, dtype=torch.int32, device="cuda")
    ops.raw_pad_mask_count(idc, 0, mask, count3)
    m = mask.view(-1)
    go = dev(torch.tensor([1.25]))
    # the live-row list: compaction stage of the plan over (in, pos, neg) sharing the one mask
    sites = [ops.Site(idc, None, live=m), ops.Site(pos, None, live=m), ops.Site(neg, None, group=K, live=m)]
    ws = torch.zeros(ops.embedding_bwd_workspace_bytes(n * (2 + K), V1, d), dtype=torch.uint8, device="cuda")
    ops.raw_embedding_bwd_plan(sites, V1, 0, ws, stage=ops.PLAN_COMPACT)
    rows = ops.raw_embedding_bwd_plan_live_rows(sites, V1, ws)
    # the list itself: ascending live rows
    base = ws.data_ptr()
    n_live = int(ws[rows[1] - base: rows[1] - base + 4].view(torch.int32).item())
    got_rows = ws[rows[0] - base: rows[0] - base + 4 * n_live].view(torch.int32).cpu()
    assert torch.equal(got_rows.long(), torch.nonzero(ids.view(-1) != 0).view(-1))
    # the same list without a plan (recguru_live_rows)
    rows2, n2 = ops.raw_live_rows(m)
    assert int(n2.item()) == n_live and torch.equal(rows2[:n_live].cpu(), got_rows)
    dh1 = torch.full((n, d), float("nan"), device="cuda")
    p1, n1, aux1, dp1, dn1 = ops.raw_sampled_score_loss(h, W, pos, neg, m, kind, skip_masked=True, dh_out=dh1,
                                                        live_count=count3, grad_scale=go)
    dh2 = torch.full((n, d), float("nan"), device="cuda")
    bufs = {k: torch.full_like(v, float("nan")) for k, v in (("p", p1), ("n", n1), ("dp", dp1), ("dn", dn1))}
    p2, n2, aux2, dp2, dn2 = ops.raw_sampled_score_loss(h, W, pos, neg, m, kind, skip_masked=True, dh_out=dh2,
                                                        live_count=count3, grad_scale=go, live_rows=rows, bufs=bufs)
    assert torch.equal(p1, p2) and torch.equal(n1, n2) and torch.equal(dp1, dp2) and torch.equal(dn1, dn2)
    assert torch.equal(dh1, dh2)
    assert torch.equal(aux1[1], aux2[1])
    torch.testing.assert_close(aux2[0], aux1[0], rtol=2e-6, atol=0)
    # bit-reproducible run to run
    p3, n3, aux3, dp3, dn3 = ops.raw_sampled_score_loss(h, W, pos, neg, m, kind, skip_masked=True, dh_out=dh2,
                                                        live_count=count3, grad_scale=go, live_rows=rows)
    assert torch.equal(aux3, aux2)
    # the sort stage after the compaction == the one-call plan
    ops.raw_embedding_bwd_plan(sites, V1, 0, ws, stage=ops.PLAN_SORT)
    X = dev(torch.randn(n, d))
    full = [ops.Site(idc, X, coef=m, live=m), ops.Site(pos, X, coef=dp1, live=m),
            ops.Site(neg, X, coef=dn1.view(-1), group=K, live=m)]
    dWa = torch.empty(V1, d, device="cuda")
    ops.raw_embedding_bwd_apply(full, V1, d, dWa, ws, mode=ops.DW_ZERO_FILL)
    ws2 = torch.zeros_like(ws)
    ops.raw_embedding_bwd_plan(sites, V1, 0, ws2)
    dWb = torch.empty(V1, d, device="cuda")
    ops.raw_embedding_bwd_apply(full, V1, d, dWb, ws2, mode=ops.DW_ZERO_FILL)
    assert torch.equal(dWa, dWb)


# ---------------------------------------------------------------- op layer / autograd --------
def test_modules_autograd_end_to_end_golden(ops):
    """FusedItemEmbedding + FusedPositionalEncoding + fused loss reproduce the reference's loss
    and W.grad for a SASRec BPR step, given the transformer's output h and its input gradient."""
    from recguru_b200 import modules as M
    g = load_golden("myrec_sas.npz")
    emb = M.FusedItemEmbedding(*g["weight"].shape, padding_idx=0, _weight=t(g["weight"]).clone()).cuda()
    pos_emb = M.FusedPositionalEncoding(32, dropout=0.5, max_len=64).cuda().eval()
    assert torch.equal(pos_emb.pe.cpu(), t(g["pe"]))
    x0, mask = emb.lookup_pe_mask(dev(t(g["dec_in"])), pos_emb, pad_index=0)
    assert torch.equal(x0.detach().cpu(), t(g["x0"]))
    # decomposed drop-in path gives the same bits
    x0b = pos_emb(emb(dev(t(g["dec_in"]))), mask)
    assert torch.equal(x0b.detach().cpu(), t(g["x0"]))
    h = dev(t(g["h"])).requires_grad_(True)
    loss, p, n = M.sampled_loss_fused(h, emb, dev(t(g["dec_out"])), dev(t(g["n_items"])), dev(t(g["mask"])), "bpr_sas")
    assert torch.allclose(loss.detach().cpu(), t(g["loss"]), rtol=1e-5)
    # the transformer sits between x0 and h in the reference; feed its recorded input gradient
    (loss + (x0 * dev(t(g["g_x0"]))).sum()).backward()
    assert torch.allclose(h.grad.cpu(), t(g["dh"]), rtol=1e-4, atol=1e-6)
    assert torch.allclose(emb.weight.grad.cpu(), t(g["dW"]), rtol=1e-5, atol=1e-6)
    assert float(emb.weight.grad[0].abs().max()) == 0.0

    # decomposed: scores op + loss module, as MyRec.forward + loss_bpr (train_auto.py:57-69)
    emb.weight.grad = None
    h2 = dev(t(g["h"])).requires_grad_(True)
    p2, n2 = M.sampled_logits(h2, emb, dev(t(g["dec_out"])), dev(t(g["n_items"])))
    assert p2.shape == (6, 12, 1) and n2.shape == (6, 12, 5)
    l2 = M.FusedBPRLoss_sas()(p2, n2, mask=dev(t(g["mask"])))
    x0c = pos_emb(emb(dev(t(g["dec_in"]))), mask)
    (l2 + (x0c * dev(t(g["g_x0"]))).sum()).backward()
    assert torch.allclose(l2.detach().cpu(), t(g["loss"]), rtol=1e-5)
    assert torch.allclose(h2.grad.cpu(), t(g["dh"]), rtol=1e-4, atol=1e-6)
    assert torch.allclose(emb.weight.grad.cpu(), t(g["dW"]), rtol=1e-5, atol=1e-6)


def test_sce_module_and_cat(ops):
    from recguru_b200 import modules as M
    r = load_golden("myrec_recon.npz")
    K = int(r["K"])
    emb = M.FusedItemEmbedding(*r["weight"].shape, padding_idx=0, _weight=t(r["weight"]).clone()).cuda()
    logits = M.sampled_logits_cat(dev(t(r["h"])), emb, dev(t(r["dec_out"])), dev(t(r["n_items"])))
    assert torch.allclose(logits.cpu(), t(r["logits"]), rtol=1e-5, atol=1e-5)
    loss = M.FusedSampledCrossEntropyLoss()(logits, None, K + 1, mask=dev(t(r["mask"])))
    assert torch.allclose(loss.cpu(), t(r["loss"]), rtol=1e-5)


# ---------------------------------------------------------------- fused plan (round 2) ---------
@pytest.mark.parametrize("V1,n_rows,K,frac", [(100_002, 204_800, 5, 0.926), (500_002, 25_600, 30, 0.5),
                                              (5_002, 3_000, 5, 0.0), (70, 1_000, 3, 0.3), (300_000, 40_000, 5, 1.0)])
def test_fused_plan_equals_round1_chain(ops, monkeypatch, V1, n_rows, K, frac):
    """The 2 + passes launch plan (plan_count / plan_scatter / radix_pass) leaves the SAME workspace
    contents -- live count, sorted keys, sorted occurrence indices -- as the default ten-launch chain,
    in row mode and in by-index mode, so apply's sums are bit-identical.  Checked through apply: dW from
    both plans must be bit-equal.  (The switch RECGURU_BWD_PLAN is read once per process: the fused
    plan runs in a child process.)"""
    import os
    import subprocess
    import sys
    g = torch.Generator().manual_seed(V1 + n_rows)
    d = 32
    mask = (torch.rand(n_rows, generator=g) >= frac).float()
    ids_in = torch.randint(0, V1, (n_rows,), generator=g)
    ids_pos = torch.randint(0, V1, (n_rows,), generator=g)
    ids_neg = torch.randint(0, V1, (n_rows * K,), generator=g)
    ids_in[torch.rand(n_rows, generator=g) < 0.05] = 0           # padding ids at live rows too
    X = torch.randn(n_rows, d, generator=g)
    cp = torch.randn(n_rows, generator=g)
    cn = torch.randn(n_rows * K, generator=g)
    dev_ = lambda x: x.cuda()  # noqa: E731
    m, Xd = dev_(mask), dev_(X)

    def run(row_mode):
        live2 = m if row_mode else m.clone()                      # a different pointer => by-index compaction
        sites = [ops.Site(dev_(ids_in), Xd, coef=m, live=m), ops.Site(dev_(ids_pos), Xd, coef=dev_(cp), live=m),
                 ops.Site(dev_(ids_neg), Xd, coef=dev_(cn), group=K, live=live2)]
        return ops.raw_embedding_bwd(sites, V1, d, 0)

    # same process, both plans: the switch is read once per process, so the v1 chain runs in a child
    script = f"""
import sys, torch, pickle
sys.path.insert(0, {str(__import__('pathlib').Path(__file__).resolve().parent.parent)!r})
from recguru_b200 import ops
g = torch.Generator().manual_seed({V1 + n_rows})
V1, n_rows, K, frac, d = {V1}, {n_rows}, {K}, {frac}, 32
mask = (torch.rand(n_rows, generator=g) >= frac).float()
ids_in = torch.randint(0, V1, (n_rows,), generator=g)
ids_pos = torch.randint(0, V1, (n_rows,), generator=g)
ids_neg = torch.randint(0, V1, (n_rows * K,), generator=g)
ids_in[torch.rand(n_rows, generator=g) < 0.05] = 0
X = torch.randn(n_rows, d, generator=g); cp = torch.randn(n_rows, generator=g); cn = torch.randn(n_rows * K, generator=g)
m, Xd = mask.cuda(), X.cuda()
res = []
for row_mode in (True, False):
    live2 = m if row_mode else m.clone()
    sites = [ops.Site(ids_in.cuda(), Xd, coef=m, live=m), ops.Site(ids_pos.cuda(), Xd, coef=cp.cuda(), live=m),
             ops.Site(ids_neg.cuda(), Xd, coef=cn.cuda(), group=K, live=live2)]
    res.append(ops.raw_embedding_bwd(sites, V1, d, 0).cpu())
torch.save(res, sys.argv[1])
"""
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        r = subprocess.run([sys.executable, "-c", script, td + "/v1.pt"], env=dict(os.environ, RECGURU_BWD_PLAN="2"),
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        v1 = torch.load(td + "/v1.pt")
    for i, row_mode in enumerate((True, False)):
        dW = run(row_mode).cpu()
        assert torch.equal(dW, v1[i]), f"row_mode={row_mode}: fused plan differs from the default chain"
    assert torch.equal(v1[0], v1[1])                              # by-row == by-index (same order of sums)
    # and against plain autograd
    Wr = torch.zeros(V1, d, requires_grad=True)
    emb = torch.nn.functional.embedding
    tot = (emb(ids_in, Wr) * X * mask[:, None]).sum() + (emb(ids_pos, Wr) * X * (cp * mask)[:, None]).sum() \
        + (emb(ids_neg, Wr).view(n_rows, K, d) * X[:, None, :] * (cn.view(n_rows, K) * mask[:, None])[..., None]).sum()
    tot.backward()
    ref = Wr.grad.clone(); ref[0] = 0
    assert torch.allclose(run(True).cpu(), ref, rtol=1e-4, atol=1e-4)


def test_plan_padmask_entry_point(ops):
    """recguru_embedding_bwd_plan_padmask: mask (and its count) derived inside the plan == pad mask kernel + plan."""
    g = torch.Generator().manual_seed(12)
    V1, n_rows, K, d = 5_002, 6_000, 5, 64
    src = torch.randint(0, 4, (n_rows,), generator=g)             # 25 % pads
    ids = [torch.randint(0, V1, (n_rows * gr,), generator=g).cuda() for gr in (1, 1, K)]
    X = torch.randn(n_rows, d, generator=g).cuda()
    coefs = [torch.randn(n_rows * gr, generator=g).cuda() for gr in (1, 1, K)]
    total = n_rows * (2 + K)
    out = []
    for fused in (False, True):
        mask = torch.full((n_rows,), 7.0, device="cuda")
        cnt = torch.zeros(3, dtype=torch.int32, device="cuda")
        ws = torch.zeros(ops.embedding_bwd_workspace_bytes(total, V1, d), dtype=torch.uint8, device="cuda")
        sites = [ops.Site(ids[i], X, coef=coefs[i], group=gr, live=mask) for i, gr in enumerate((1, 1, K))]
        if fused:
            ops.raw_embedding_bwd_plan(sites, V1, 0, ws, mask_ids=src.cuda(), pad_index=0, count3=cnt)
        else:
            ops.raw_pad_mask_count(src.cuda(), 0, mask, cnt)
            ops.raw_embedding_bwd_plan(sites, V1, 0, ws)
        dW = torch.empty(V1, d, device="cuda")
        ops.raw_embedding_bwd_apply(sites, V1, d, dW, ws, mode=ops.DW_ZERO_FILL)
        out.append((mask.cpu(), cnt.cpu(), dW.cpu()))
    assert torch.equal(out[0][0], (src != 0).float()) and torch.equal(out[1][0], out[0][0])
    assert out[1][1].tolist() == [0, 0, int((src != 0).sum())] == out[0][1].tolist()
    assert torch.equal(out[0][2], out[1][2])
