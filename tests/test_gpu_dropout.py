"""SURVEY 8f rank 4 + the P1' row: the nn.Dropout behind PositionalEncoding.forward
(Transformer/transformer.py:104-106) fused into K1 with BIT parity against torch's own CUDA dropout
under a seeded generator (same kept elements, same generator offset afterwards), its backward, the
learned position table (PositionalEncodingM, :327-342) with its gradient, and the fused
``[:, -1, :]`` user-representation extraction (gan_training.py:152-162)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_path as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_lib):
    from recguru_b200 import ops as o
    return o


def bits(x):
    return x.contiguous().view(torch.int32)


@pytest.mark.parametrize("shape", [(6, 12, 32), (64, 50, 64), (1024, 200, 128), (3, 7, 4)])
@pytest.mark.parametrize("p", [0.5, 0.1, 0.9])
def test_dropout_bit_parity_with_torch_generator(ops, shape, p):
    dev = torch.device("cuda")
    x = torch.randn(*shape, device=dev)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    for warm in (0, 3):                                   # a fresh generator, and one whose offset has moved
        torch.manual_seed(1234 + warm)
        for _ in range(warm):
            torch.rand(1000, device=dev)
        ref = F.dropout(x, p, True)
        off_ref = gen.get_offset()
        after_ref = torch.rand(5, device=dev)             # the stream continues identically afterwards
        torch.manual_seed(1234 + warm)
        for _ in range(warm):
            torch.rand(1000, device=dev)
        seed, off, threads = ops.torch_dropout_state(x.numel(), dev)
        out = ops.raw_dropout(x, p, seed, off, threads)
        assert gen.get_offset() == off_ref, "generator offset differs from torch's after the call"
        assert torch.equal(bits(out), bits(ref)), f"{int((bits(out) != bits(ref)).sum())} elements differ"
        assert torch.equal(torch.rand(5, device=dev), after_ref)


def test_k1_fused_dropout_equals_torch_dropout_of_k1(ops):
    from recguru_b200 import synth
    B, L, d, V, p = 64, 50, 128, 3000, 0.5
    W = synth.item_table(V, d).cuda()
    pe = O.pe_table(64, d).cuda()
    tb = synth.train_batch(B, L, V, 5, seed=4).to("cuda")
    x0, mask = ops.raw_gather_pe_mask(W, tb.dec_in, pe, 0)
    torch.manual_seed(99)
    ref = F.dropout(x0, p, True)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    off_ref = gen.get_offset()
    torch.manual_seed(99)
    seed, off, threads = ops.torch_dropout_state(x0.numel(), x0.device)
    out, m2 = ops.raw_gather_pe_mask_dropout(W, tb.dec_in, pe, 0, p, seed, off, threads)
    assert gen.get_offset() == off_ref
    assert torch.equal(m2, mask)
    assert torch.equal(bits(out), bits(ref)), "fused K1+dropout differs from dropout(K1)"   # incl. the -0.0 of masked rows
    # with an explicit mask (cross-domain call sites)
    torch.manual_seed(99)
    seed, off, threads = ops.torch_dropout_state(x0.numel(), x0.device)
    out3, _ = ops.raw_gather_pe_mask_dropout(W, tb.dec_in, pe, -1, p, seed, off, threads, mask_in=mask)
    assert torch.equal(bits(out3), bits(ref))


def test_module_training_mode_matches_reference_ops_and_autograd(ops):
    """lookup_pe_mask / pos_emb(x, mask) in TRAINING mode == F.embedding -> +pe -> *mask -> nn.Dropout with the
    same seed, forward bits and gradients (the table backward goes through the dropout mask)."""
    from recguru_b200 import modules as M, synth
    B, L, d, V, p = 16, 20, 64, 500, 0.5
    W0 = synth.item_table(V, d)
    tb = synth.train_batch(B, L, V, 5, seed=8).to("cuda")
    g = synth.hidden_states(B, L, d, seed=3).cuda()
    pos = M.FusedPositionalEncoding(d, p, max_len=64).cuda().train()
    for fused_site in ("lookup", "module"):
        emb = M.FusedItemEmbedding(V + 2, d, padding_idx=0, _weight=W0.clone()).cuda()
        torch.manual_seed(7)
        if fused_site == "lookup":
            x, mask = emb.lookup_pe_mask(tb.dec_in, pos, pad_index=0)
        else:
            mask = (tb.dec_in != 0).float()
            x = pos(emb(tb.dec_in), mask)                       # the reference's two call sites, unchanged
        (x * g).sum().backward()
        Wr = W0.clone().cuda().requires_grad_(True)
        torch.manual_seed(7)
        xr = F.dropout((F.embedding(tb.dec_in, Wr, padding_idx=0) + pos.pe[0, :L]) * (tb.dec_in != 0).float()[..., None], p, True)
        (xr * g).sum().backward()
        assert torch.equal(bits(x.detach()), bits(xr.detach())), fused_site
        assert torch.allclose(emb.weight.grad, Wr.grad, rtol=1e-5, atol=1e-6), fused_site
    pos.eval()
    x_eval, _ = emb.lookup_pe_mask(tb.dec_in, pos, pad_index=0)
    assert torch.equal(x_eval, ops.raw_gather_pe_mask(emb.weight.detach(), tb.dec_in, pos.pe, 0)[0])


def test_learned_position_module_and_gradient(ops):
    """P1' (Transformer/transformer.py:327-342): x + pos_emb(arange(L)), mask, dropout -- forward bits and the
    gradients of BOTH tables against torch autograd; state_dict key as in the reference."""
    from recguru_b200 import modules as M, synth
    B, L, d, V = 12, 30, 64, 400
    W0 = synth.item_table(V, d)
    tb = synth.train_batch(B, L, V, 5, seed=9).to("cuda")
    g = synth.hidden_states(B, L, d, seed=5).cuda()
    pos = M.FusedPositionalEncodingM(d, 0.5, max_len=50).cuda().eval()
    assert list(pos.state_dict().keys()) == ["pos_emb.weight"]
    mask = (tb.dec_in != 0).float()

    def reference(Wr, Pr):
        idx = torch.arange(L, device="cuda").repeat(B).view(B, L)           # transformer.py:339
        return (F.embedding(tb.dec_in, Wr) + F.embedding(idx, Pr)) * mask.unsqueeze(2)

    for site in ("lookup", "module"):
        emb = M.FusedItemEmbedding(V + 2, d, padding_idx=None, _weight=W0.clone()).cuda()
        pos.pos_emb.weight.grad = None
        if site == "lookup":
            x, _ = emb.lookup_pe_mask(tb.dec_in, pos, pad_index=0)
        else:
            x = pos(emb(tb.dec_in), mask)
        (x * g).sum().backward()
        Wr = W0.clone().cuda().requires_grad_(True)
        Pr = pos.pos_emb.weight.detach().clone().requires_grad_(True)
        xr = reference(Wr, Pr)
        (xr * g).sum().backward()
        assert torch.equal(x.detach(), xr.detach()), site
        assert torch.allclose(pos.pos_emb.weight.grad, Pr.grad, rtol=1e-5, atol=1e-6), site
        assert float(pos.pos_emb.weight.grad[L:].abs().sum()) == 0.0
        assert torch.allclose(emb.weight.grad, Wr.grad, rtol=1e-5, atol=1e-6), site
    # the raw reduction is bit-reproducible and accumulates
    d1 = ops.raw_pe_grad(g, mask)
    d2 = ops.raw_pe_grad(g, mask)
    assert torch.equal(d1, d2)
    acc = d1.clone()
    ops.raw_pe_grad(g, mask, out=acc, accumulate=True)
    assert torch.allclose(acc, 2 * d1)
    # training mode with the fused dropout in front of a learned table
    pos.train()
    torch.manual_seed(3)
    emb = M.FusedItemEmbedding(V + 2, d, padding_idx=None, _weight=W0.clone()).cuda()
    x, _ = emb.lookup_pe_mask(tb.dec_in, pos, pad_index=0)
    torch.manual_seed(3)
    xr = F.dropout(reference(W0.cuda(), pos.pos_emb.weight.detach()), 0.5, True)
    assert torch.equal(bits(x.detach()), bits(xr))


def test_extract_last_user_reps(ops):
    from recguru_b200 import catalog
    h = torch.randn(257, 50, 128, device="cuda")
    u32, u16 = catalog.user_reps(h, bf16=True)
    assert torch.equal(u32, h[:, -1, :]) and torch.equal(u16, h[:, -1, :].to(torch.bfloat16))
    hr = h.clone().requires_grad_(True)
    u = torch.ops.recguru.extract_last(hr)
    (u * u32).sum().backward()
    ref = torch.zeros_like(h); ref[:, -1, :] = u32
    assert torch.equal(hr.grad, ref)
