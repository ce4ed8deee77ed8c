"""GPU parity of the whole training-step hot path object (recguru_b200.hotpath.TrainHotPath):
gather+PE+mask -> sampled scores + loss -> dh -> one dense table gradient for all three lookup
sites, against the oracle run through torch autograd on the CPU (which is exactly what the
reference does: F.embedding / matmul / loss module / loss.backward(), train_auto.py:110-125).

Also: the padding-aware and two-stream variants must give the SAME bits as the plain one, the
plan/apply split the same bits as the fused entry point, and a CUDA-graph replay the same bits
as eager launches."""
import pytest
import torch

from oracle import ref_path as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hp_mod(built_lib):
    from recguru_b200 import hotpath, ops
    return hotpath, ops


def make_step(B, L, d, K, V, seed, pad_frac=0.6):
    g = torch.Generator().manual_seed(seed)
    V1 = V + 2
    W = torch.randn(V1, d, generator=g) * 0.3
    W[0] = 0
    dec_in = torch.randint(1, V + 1, (B, L), generator=g)
    dec_out = torch.randint(1, V + 1, (B, L), generator=g)
    n_items = torch.randint(1, V + 1, (B, L * K), generator=g)
    lens = torch.randint(1, max(2, int(L * (1 - pad_frac)) + 1), (B,), generator=g)
    pad = torch.arange(L).unsqueeze(0) < (L - lens).unsqueeze(1)          # left padding, like seq_padding
    dec_in[pad] = 0
    dec_out[pad] = 0
    h = torch.randn(B, L, d, generator=g) * 0.3
    g_x0 = torch.randn(B, L, d, generator=g) * 0.01
    pe = O.pe_table(max(L, 8), d)
    return W, pe, dec_in, dec_out, n_items, h, g_x0


def oracle_step(W, pe, dec_in, dec_out, n_items, h, g_x0, K, kind):
    """The reference's op sequence + autograd; the transformer between the edges is replaced by
    given h (its output) and g_x0 (the gradient it returns to its input)."""
    B, L = dec_in.shape
    Wp = W.clone().requires_grad_(True)
    hp = h.clone().requires_grad_(True)
    x0, mask = O.gather_pe_mask(Wp, dec_in, pe, 0)
    p, n = O.sampled_scores(hp, Wp, dec_out, n_items, L, K)
    loss = O.LOSSES[kind](p, n, mask.view(-1))
    # d(total)/dW = d loss/dW (scoring sites) + <g_x0, d x0/dW> (input site)
    (loss + (x0 * g_x0).sum()).backward()
    dW = Wp.grad.clone()
    dW[0] = 0                      # nn.Embedding(padding_idx=0) never accumulates into its pad row
    return x0.detach(), mask, loss.detach(), hp.grad, dW


@pytest.mark.parametrize("kind", ["bpr_sas", "bpr", "sce"])
@pytest.mark.parametrize("B,L,d,K,V", [(8, 20, 128, 5, 300), (5, 13, 64, 3, 50), (16, 50, 32, 5, 2000)])
def test_step_matches_reference_autograd(hp_mod, kind, B, L, d, K, V):
    hotpath, ops = hp_mod
    W, pe, dec_in, dec_out, n_items, h, g_x0 = make_step(B, L, d, K, V, seed=B * L + d)
    x0_ref, mask_ref, loss_ref, dh_ref, dW_ref = oracle_step(W, pe, dec_in, dec_out, n_items, h, g_x0, K, kind)
    c = lambda x: x.cuda()
    results = []
    variants = [(False, False, None), (True, False, None), (False, True, None), (True, True, None),
                (True, True, False), (True, False, False)]         # fuse_dh: None = default (on with skip_masked)
    for skip, overlap, fuse in variants:
        hp = hotpath.TrainHotPath(c(W), c(pe), B, L, K, kind, skip_masked=skip, overlap_plan=overlap, fuse_dh=fuse)
        assert hp.fuse_dh == (skip if fuse is None else fuse)
        loss, dh, dW = hp.step(c(dec_in), c(dec_out), c(n_items), c(h), c(g_x0))
        torch.cuda.synchronize()
        results.append((loss.clone(), dh.clone(), dW.clone()))
        assert torch.equal(hp.x0.cpu(), x0_ref) and torch.equal(hp.mask.cpu(), mask_ref)
        assert torch.allclose(loss.cpu(), loss_ref, rtol=1e-5, atol=1e-6)
        assert torch.allclose(dh.cpu(), dh_ref, rtol=1e-4, atol=1e-7)
        assert torch.allclose(dW.cpu(), dW_ref, rtol=1e-4, atol=1e-6)
    # every variant: gradients bit-identical to the plain one; the loss scalar too when the same
    # scoring kernel ran (the padding-aware kernel sums the same terms in another order)
    # (three scoring kernels: dense; padding-aware compacting per CTA; padding-aware over the plan's
    # live-row list -- the one a step with the fused dh launches, in every schedule)
    def kernel_of(skip, overlap, fuse):
        return 0 if not skip else (2 if fuse is not False else 1)
    first = {}
    for (skip, overlap, fuse), r in zip(variants, results):
        assert torch.equal(r[1], results[0][1]) and torch.equal(r[2], results[0][2])
        ref = first.setdefault(kernel_of(skip, overlap, fuse), r[0])
        assert torch.equal(r[0], ref)
        assert torch.allclose(r[0], results[0][0], rtol=1e-6, atol=0)


def test_plan_apply_equals_fused(hp_mod):
    hotpath, ops = hp_mod
    torch.manual_seed(5)
    V1, d, n, K = 700, 128, 6000, 4
    ids_a = torch.randint(0, V1, (n,)).cuda()
    ids_b = torch.randint(0, V1, (n * K,)).cuda()
    Xa, Xb = torch.randn(n, d).cuda(), torch.randn(n, d).cuda()
    live = (torch.rand(n) < 0.3).float().cuda()
    ca = live * torch.randn(n).cuda()
    cb = (live.unsqueeze(1) * torch.randn(n, K).cuda()).reshape(-1)
    fused = ops.raw_embedding_bwd([ops.Site(ids_a, Xa, ca), ops.Site(ids_b, Xb, cb, group=K)], V1, d, padding_idx=0)
    ws = torch.empty(ops.embedding_bwd_workspace_bytes(n * (1 + K), V1, d), dtype=torch.uint8, device="cuda")
    # plan from the liveness mask only (coefficients "not yet computed"), apply later
    ops.raw_embedding_bwd_plan([ops.Site(ids_a, None, live=live), ops.Site(ids_b, None, group=K, live=live)],
                               V1, 0, ws)
    out = torch.full((V1, d), 7.0, device="cuda")
    ops.raw_fill_zero(out)
    ops.raw_embedding_bwd_apply([ops.Site(ids_a, Xa, ca, live=live), ops.Site(ids_b, Xb, cb, group=K, live=live)],
                                V1, d, out, ws, mode=ops.DW_PREZEROED)
    assert torch.equal(out, fused)
    # the same plan by gathered index (distinct mask buffers => no shared-mask fast path) must give
    # the same bits as the plan by row (one shared mask)
    ws2 = torch.empty_like(ws)
    live_b = live.clone()
    ops.raw_embedding_bwd_plan([ops.Site(ids_a, None, live=live), ops.Site(ids_b, None, group=K, live=live_b)],
                               V1, 0, ws2)
    out_idx = torch.zeros(V1, d, device="cuda")
    ops.raw_embedding_bwd_apply([ops.Site(ids_a, Xa, ca, live=live), ops.Site(ids_b, Xb, cb, group=K, live=live_b)],
                                V1, d, out_idx, ws2, mode=ops.DW_PREZEROED)
    assert torch.equal(out_idx, fused)
    # accumulate mode on top of an existing gradient
    out2 = fused.clone()
    ops.raw_embedding_bwd_apply([ops.Site(ids_a, Xa, ca, live=live), ops.Site(ids_b, Xb, cb, group=K, live=live)],
                                V1, d, out2, ws, mode=ops.DW_ACCUMULATE)
    assert torch.allclose(out2, 2 * fused, rtol=1e-6, atol=1e-6)


def test_graph_replay_equals_eager(hp_mod):
    hotpath, ops = hp_mod
    B, L, d, K, V = 32, 50, 128, 5, 5000
    W, pe, dec_in, dec_out, n_items, h, g_x0 = [x.cuda() for x in make_step(B, L, d, K, V, seed=77)]
    hp = hotpath.TrainHotPath(W, pe, B, L, K, "bpr_sas")
    loss, dh, dW = hp.step(dec_in, dec_out, n_items, h, g_x0)
    torch.cuda.synchronize()
    ref = (loss.clone(), dh.clone(), dW.clone())
    graph = hp.capture_step(dec_in, dec_out, n_items, h, g_x0)
    # new data in the same (static) buffers: the graph must compute on what is there at replay
    W2, pe2, di2, do2, ni2, h2, g2 = [x.cuda() for x in make_step(B, L, d, K, V, seed=78)]
    for dst, src in ((dec_in, di2), (dec_out, do2), (n_items, ni2), (h, h2), (g_x0, g2)):
        dst.copy_(src)
    graph.replay()
    torch.cuda.synchronize()
    got2 = (hp.sbuf["aux"][0].clone(), hp.dh.clone(), hp.dW.clone())
    hp2 = hotpath.TrainHotPath(W, pe, B, L, K, "bpr_sas", overlap_plan=False)
    e2 = hp2.step(di2, do2, ni2, h2, g2)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(got2, e2))          # same kernels, same bits
    assert not torch.equal(got2[0], ref[0])


def test_edges_called_one_by_one_equal_step(hp_mod):
    """The caller-driven order of INTEGRATION.md (forward_in, plan_backward, [transformer], forward_loss,
    backward) gives the same bits as step(), whichever of forward_in / plan_backward comes first, with
    and without an upstream gradient scale tensor."""
    hotpath, ops = hp_mod
    B, L, d, K, V = 16, 40, 128, 5, 3000
    W, pe, dec_in, dec_out, n_items, h, g_x0 = [x.cuda() for x in make_step(B, L, d, K, V, seed=11)]
    ref = hotpath.TrainHotPath(W, pe, B, L, K, "bpr_sas").step(dec_in, dec_out, n_items, h, g_x0)
    ref = tuple(t.clone() for t in ref)
    one = torch.ones(1, device="cuda")
    for plan_first in (True, False, None):
        for go in (None, one):
            hp = hotpath.TrainHotPath(W, pe, B, L, K, "bpr_sas")
            for _ in range(2):                                       # twice: per-step state must reset
                if plan_first is True:
                    hp.plan_backward(dec_in, dec_out, n_items)
                x0 = hp.forward_in(dec_in)
                if plan_first is False:
                    hp.plan_backward(dec_in, dec_out, n_items)
                loss = hp.forward_loss(h, dec_out, n_items, grad_loss=go)
                dh, dW = hp.backward(h, dec_in, dec_out, n_items, g_x0, grad_loss=go)
                torch.cuda.synchronize()
                assert torch.equal(loss, ref[0]) and torch.equal(dh, ref[1]) and torch.equal(dW, ref[2]), (plan_first, go)
    # an upstream scale that differs between forward and backward falls back to the stand-alone dh kernel
    hp = hotpath.TrainHotPath(W, pe, B, L, K, "bpr_sas")
    hp.forward_in(dec_in)
    hp.forward_loss(h, dec_out, n_items)
    half = torch.full((1,), 0.5, device="cuda")
    dh, dW = hp.backward(h, dec_in, dec_out, n_items, g_x0, grad_loss=half)
    assert torch.allclose(dh, 0.5 * ref[1], rtol=1e-6, atol=0)
